"""BASELINE.json configs[0]: single 256x256 image, hand-only forward (image_demo.py path, random-init weights) on the
host CPU - the reference's own CPU-runnable case.  Times the oracle port (pinned against the
reference's files by tests/test_oracle_vs_reference.py) with torch CPU; writes one JSON line.

    python scripts/config1_cpu.py [--threads N] [--reps K]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nets  # noqa: E402
from obman_train_b200.networks.handnet import HandNet  # noqa: E402

CFG = dict(resnet_version=18, mano_root="synthetic", mano_comps=30, mano_use_pca=True, mano_use_shape=True,
           mano_neurons=[1024, 256], mano_center_idx=0, mano_lambda_verts=0.167, no_loss=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    torch.manual_seed(0)
    model = HandNet(**{k: v for k, v in CFG.items() if k != "no_loss"}).eval()
    state = {k: v.detach() for k, v in model.state_dict().items()}
    tables = {s: {k: v.detach() for k, v in getattr(model.mano_branch, "mano_layer_" + s).named_buffers()
                  if k != "th_faces"} for s in ("right", "left")}
    g = torch.Generator().manual_seed(0)
    sample = {"images": torch.rand(1, 3, 256, 256, generator=g) - 0.5, "sides": ["left"], "root": "wrist",
              "joints3d": torch.ones(1, 21, 3)}
    times = []
    with torch.no_grad():
        for it in range(3 + args.reps):
            t0 = time.perf_counter()
            _, results, _ = nets.handnet_forward(state, CFG, sample, tables, None, None)
            if it >= 3:
                times.append(time.perf_counter() - t0)
    assert results["verts"].shape == (1, 778, 3) and results["joints"].shape == (1, 21, 3)
    times.sort()
    print(json.dumps({"config": "BASELINE.json configs[0]: B=1, 256x256, hand-only forward (ResNet-18 + ManoBranch + MANO), "
                                "oracle port on torch CPU", "threads": args.threads, "reps": args.reps,
                      "ms_median": 1e3 * times[len(times) // 2], "ms_min": 1e3 * times[0],
                      "images_per_s": 1.0 / times[len(times) // 2]}))


if __name__ == "__main__":
    main()
