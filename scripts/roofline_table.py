"""Per-shape tensor-core roofline table of one training step from bench.py's --dump-launches file.

    python scripts/roofline_table.py profiles/tc_launches_<tag>.txt profiles/roofline_table_<tag>.md "<title suffix>"
"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else ""
    peak = 1392.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:  # noqa: BLE001
        pass
    agg = collections.OrderedDict()
    for line in open(src):
        if line.startswith("#") or not line.strip():
            continue
        parts = line.rsplit(None, 3)
        name, gf, ms = parts[0].strip(), float(parts[1]), float(parts[2])
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += gf; a[2] += ms
    tot_ms = sum(a[2] for a in agg.values())
    tot_gf = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# Per-shape tensor-core roofline of one training step (B=256, BASELINE configs[2], 3xBF16)%s\n\n" % title)
        f.write("Source: %s (CUDA events per launch, every launch alone on the GPU; `bench.py --dump-launches`).  Peak = "
                "MEASURED_PEAKS.json bf16_tflops_sustained = %.1f TFLOP/s.  `algorithmic` = 2*M*N*K fp32 FLOPs / time;\n"
                "`pipe` = executed tensor work (3 bf16 products per contraction) / peak.  Shapes that appear several times "
                "in a step are summed.\nconv_nhwc with c_out = 64 runs on conv64_v2_kernel, plain matrices with N = 128 j + "
                "(1..4) on the tail-column variant of gemm_tc_kernel, with one column tile on gemm_persist_kernel,\n"
                "everything else on gemm_tc_kernel / wgrad_bf16_kernel.\n\n" % (os.path.relpath(src, ROOT), peak))
        f.write("| launch shape | launches | GFLOP | ms | algorithmic TFLOP/s | pipe % of peak | share of tensor-core time |\n")
        f.write("|---|---|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
            tf = a[1] / a[2]
            f.write("| %s | %d | %.2f | %.4f | %.1f | %.1f | %.1f %% |\n" % (name, a[0], a[1], a[2], tf, 300.0 * tf / peak,
                                                                       100.0 * a[2] / tot_ms))
        tf = tot_gf / tot_ms
        f.write("| **all %d launches** | | %.1f | %.3f | %.1f | %.1f | 100 %% |\n" % (n, tot_gf, tot_ms, tf, 300.0 * tf / peak))
    print("wrote", dst, "-", n, "launches,", round(tot_ms, 3), "ms")


if __name__ == "__main__":
    main()
