"""Turn the raw profiler output of a gpurun call (gpurun_out/) into the committed summaries under profiles/.

    python scripts/summarize_profiles.py <tag>        # e.g. r1j

* gpurun_out/launches.csv  (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv of
  `bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline`) -> profiles/launches_<tag>_bench_step.csv (the rows of
  ONE training step: stem_pack to stem_pack), profiles/launches_<tag>_bench_step_summary.txt (per-kernel totals) and
  profiles/gemm_traffic_<tag>.json (DRAM bytes per tensor-core launch: bench.py's roofline.traffic).
* gpurun_out/prof_<tag>.ncu-rep (ncu --set full of scripts/prof_kernels.py) -> profiles/ncu_full_<tag>_summary.txt.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
FULL_METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
KERNEL_LABELS = ["conv3x3 64x64 64->64 (conv64_v2)", "conv3x3 32x32 128->128", "conv3x3 16x16 256->256", "conv3x3 8x8 512->512",
                 "wgrad3x3 64x64 64->64", "wgrad3x3 32x32 128->128", "wgrad3x3 16x16 256->256",
                 "wgrad3x3 8x8 512->512", "gemm (B*642)x257x515"]


def launch_list(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        print("no", path)
        return
    raw = open(path).readlines()
    lines = [l for l in raw if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()
    for row in rows:
        per.setdefault(row["ID"], {"name": row["Kernel Name"]})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    ids = list(per.keys())
    # one training step = the launches between two consecutive fused-Adam launches (N = 1: one Adam per step)
    adams = [i for i, k in enumerate(ids) if "adam_kernel" in per[k]["name"]]
    if len(adams) < 2:
        print("launch list too short to contain a full step:", len(ids), "launches,", len(adams), "adam_kernel")
        return
    a, b = adams[-2] + 1, adams[-1] + 1
    keep = set(ids[a:b])
    with open(os.path.join(PROF, "launches_%s_bench_step.csv" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv "
                "python bench.py --config 3 --steps 1 --warmup 1 --no-graph --quick\n")
        f.write("# one full training step (launch IDs %s..%s: Adam to Adam); times are cold-cache and serialised\n" % (ids[a], ids[b - 1]))
        f.write(lines[0])
        for l, row in zip(lines[1:], rows):
            if row["ID"] in keep:
                f.write(l)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i in range(a, b):
        v = per[ids[i]]
        n = v["name"].split("(")[0].replace("void ", "").replace("at::native::", "").replace("native::", "")[:64]
        x = agg[n]
        x[0] += 1
        x[1] += v.get("gpu__time_duration.sum", 0) / 1e3
        x[2] += v.get("dram__bytes_read.sum", 0)
        x[3] += v.get("dram__bytes_write.sum", 0)
    tot = sum(x[1] for x in agg.values())
    tc = [x for n, x in agg.items() if "gemm_tc_kernel" in n or "wgrad_bf16_kernel" in n or "conv64_" in n or "gemm_persist" in n]
    n_tc, us_tc, by = sum(x[0] for x in tc), sum(x[1] for x in tc), sum(x[2] + x[3] for x in tc)
    with open(os.path.join(PROF, "launches_%s_bench_step_summary.txt" % tag), "w") as f:
        f.write("# per-kernel totals of ONE training step (B=256, configs[2], bf16x3, eager launches, ncu serialises all streams) from "
                "profiles/launches_%s_bench_step.csv\n" % tag)
        f.write("# %d launches, %.1f us serialised under ncu\n" % (b - a, tot))
        f.write("%-66s %5s %10s %6s %10s %10s\n" % ("kernel", "n", "us", "share", "dram rd MB", "dram wr MB"))
        for n, x in sorted(agg.items(), key=lambda q: -q[1][1]):
            f.write("%-66s %5d %10.1f %5.1f%% %10.1f %10.1f\n" % (n, x[0], x[1], 100 * x[1] / tot, x[2] / 1e6, x[3] / 1e6))
        f.write("\n# tensor-core kernels (gemm_tc_kernel + gemm_persist_kernel + wgrad_bf16_kernel + conv64_*): %d launches, %.1f us = %.1f%% of the step, DRAM "
                "traffic %.1f MB per step = %.2f MB per launch\n" % (n_tc, us_tc, 100 * us_tc / tot, by / 1e6, by / 1e6 / n_tc))
    sys.path.insert(0, ROOT)
    import bench
    rec = {"source": "profiles/launches_%s_bench_step.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, one B=256 "
                     "configs[2] step)" % tag,
           "precision": "bf16x3", "kernel_source_sha256": bench.kernel_source_hash(),
           "tensor_core_launches_per_step": n_tc, "dram_bytes_per_step": by, "dram_bytes_per_launch": by / n_tc,
           "tensor_core_share_of_step_under_ncu": us_tc / tot, "step_us_under_ncu": tot, "launches_per_step": b - a}
    json.dump(rec, open(os.path.join(PROF, "gemm_traffic_%s.json" % tag), "w"), indent=1)
    json.dump(rec, open(os.path.join(PROF, "gemm_traffic_current.json"), "w"), indent=1)
    print("launch list: %d launches / step, %.1f us, tensor-core share %.1f%%, %.2f MB DRAM per tensor-core launch" % (
        b - a, tot, 100 * us_tc / tot, by / 1e6 / n_tc))


def full_capture(tag):
    rep = os.path.join(OUT, "prof_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        print("no", rep)
        return
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    with open(os.path.join(PROF, "ncu_full_%s_bf16x3_summary.txt" % tag), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|wgrad_bf16_kernel|conv64_|"
                "chamfer_bwd' -c 12, scripts/prof_kernels.py (PROF_B layer shapes in launch order: see the script's CASES), "
                "PROF_PASSES=2 PROF_WG_PASSES=2 (3xBF16)\n")
        f.write("# source: gpurun_out/prof_%s.ncu-rep (not committed); extracted with ncu -i ... --page raw --csv\n" % tag)
        for k, row in enumerate(data):
            f.write("%s   [%s]\n" % (row[idx["Kernel Name"]][:70], KERNEL_LABELS[k] if k < len(KERNEL_LABELS) else ""))
            for m in FULL_METRICS:
                if m in idx:
                    f.write("   %-78s %s %s\n" % (m, row[idx[m]], units[idx[m]]))
    print("full capture:", len(data), "kernels")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "rX"
    launch_list(tag)
    full_capture(tag)
