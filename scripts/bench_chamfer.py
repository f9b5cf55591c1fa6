"""BASELINE.json configs[4]: Chamfer / contact kernel sweep on one B200 (roofline report).

For every (points per cloud, batch) it times the fused nearest-neighbour kernel (both Chamfer directions),
the Chamfer backward and the contact ray-parity kernel with CUDA events and reports, per SURVEY.md §8d:
  * pair evaluations / s and the share of the FP32 issue peak they represent (the honest bound: the fused
    kernel does 200-3300 flop per algorithmic byte, it cannot be HBM bound),
  * GB/s on ALGORITHMIC bytes and % of the measured HBM peak (the number BASELINE.json asks for),
  * the HBM bandwidth the reference's materialised (B,M,N) formulation would have needed for the same time.

    python scripts/bench_chamfer.py [--quick] [--out profiles/chamfer_sweep_r1.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import functional as Fb  # noqa: E402
from obman_train_b200.icosphere import icosphere  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "chamfer_sweep.json"))
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    sm_clk = peaks.get("sm_max_mhz", 1965.0) * 1e6
    fp32_instr_peak = 148 * 128 * sm_clk  # lane-instructions / s (4 schedulers x 32 lanes per SM)
    sizes = [600, 1024, 2048, 2500, 4096, 10000]
    batches = [32, 128, 512, 2048]
    if args.quick:
        sizes, batches = [600, 2500, 10000], [32, 512]
    rows = []
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in sizes:
        for b in batches:
            x = torch.randn(b, n, 3, device="cuda", generator=g) * 60
            y = torch.randn(b, n, 3, device="cuda", generator=g) * 60
            reps = max(2, min(50, int(2e10 / (2.0 * b * n * n))))
            ms = timeit(lambda: Fb.nearest_neighbours(x, y), reps)
            pairs = 2.0 * b * n * n
            algo_bytes = 12.0 * b * (n + n) + 8.0 * b * (n + n)
            xg = x.clone().requires_grad_(True)
            l1, l2 = Fb.chamfer(xg, y)
            loss = (l1 + l2).mean()
            ms_bwd = timeit(lambda: torch.autograd.grad(loss, xg, retain_graph=True), max(2, reps))
            bwd_bytes = 16.0 * b * (n + n) + 12.0 * b * n
            rows.append({
                "points": n, "batch": b, "fwd_ms": ms, "pairs_per_s": pairs / ms * 1e3,
                "fp32_issue_frac": pairs * 7.3 / (ms * 1e-3) / fp32_instr_peak,
                "algo_gbs": algo_bytes / ms / 1e6, "algo_frac_of_hbm": algo_bytes / ms / 1e6 / hbm,
                "reference_formulation_gbs_needed": 4.0 * b * n * n * 10 / ms / 1e6,
                "bwd_ms": ms_bwd, "bwd_algo_gbs": bwd_bytes / ms_bwd / 1e6,
                "bwd_frac_of_hbm": bwd_bytes / ms_bwd / 1e6 / hbm,
            })
            print(json.dumps(rows[-1]), flush=True)
            del x, y, xg
    # contact ray parity: 778 hand vertices vs icosphere meshes
    contact = []
    for sub, bs in ((3, [64, 256]), (4, [64, 256])):
        v, f = icosphere(sub)
        faces = torch.tensor(f, dtype=torch.int32, device="cuda")
        for b in bs:
            obj = torch.tensor(v, dtype=torch.float32, device="cuda").unsqueeze(0) * 40 + torch.randn(b, v.shape[0], 3, device="cuda", generator=g)
            pts = torch.randn(b, 778, 3, device="cuda", generator=g) * 35
            ms = timeit(lambda: Fb.mesh_exterior(pts, obj, faces), 20)
            tests = 778.0 * f.shape[0] * b
            contact.append({"faces": int(f.shape[0]), "batch": b, "ms": ms, "ray_tests_per_s": tests / ms * 1e3,
                            "fp32_issue_frac": tests * 30 / (ms * 1e-3) / fp32_instr_peak,
                            "reference_temporaries_gb": 8 * 4.0 * b * 778 * f.shape[0] * 3 / 1e9})
            print(json.dumps(contact[-1]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump({"hbm_peak_gbs": hbm, "peak_source": "measured" if peaks else "fallback",
                   "fp32_lane_instr_per_s_peak": fp32_instr_peak, "chamfer": rows, "contact_raycast": contact}, fh, indent=1)


if __name__ == "__main__":
    main()
