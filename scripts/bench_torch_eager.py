"""Practical comparator: the reference ALGORITHM in eager PyTorch on one B200 - cuDNN / cuBLAS /
ATen, i.e. what the reference itself would run - on the bench workload (BASELINE configs[1], batch 64), fwd + bwd
+ torch.optim.Adam.  Uses oracle/nets.py on CUDA tensors (test infrastructure, not the product).  Two settings:
PyTorch's default (TF32 allowed for cuDNN convolutions) and strict fp32."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import nets  # noqa: E402
from obman_train_b200.networks.handnet import HandNet  # noqa: E402


def run(allow_tf32, steps=10, warmup=3):
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    wl = bench.Workload(int(os.environ.get("EAGER_CONFIG", "2")))
    batch = int(os.environ.get("EAGER_BATCH", str(wl.batch)))
    model = HandNet(**wl.cfg).eval()
    state = {k: v.detach().clone().cuda() for k, v in model.state_dict().items()}
    leaves = []
    for k, v in state.items():
        if v.is_floating_point() and "running_" not in k and "th_" not in k and ".fc." not in k:
            v.requires_grad_(True)
            leaves.append(v)
    opt = torch.optim.Adam(leaves, lr=1e-4)
    tables = {s: {k: v.detach().cuda() for k, v in getattr(model.mano_branch, "mano_layer_" + s).named_buffers()
                  if k != "th_faces"} for s in ("right", "left")}
    grid, faces = model.atlas_branch.test_verts.cuda(), model.atlas_branch.test_faces
    from obman_train_b200.assets import load_contacts
    zones = load_contacts()[1]
    sample = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in wl.sample(batch, 0).items()}

    def step():
        opt.zero_grad()
        total, _, _ = nets.handnet_forward(state, wl.cfg, sample, tables, grid, faces, zones)
        total.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"allow_tf32_cudnn": allow_tf32, "ms_per_step": ms, "images_per_s": batch / ms * 1e3, "batch": batch, "workload": wl.name}


if __name__ == "__main__":
    out = [run(True), run(False)]
    print(json.dumps({"comparator": "eager PyTorch (cuDNN/cuBLAS) reference algorithm, one B200", "runs": out}))
