"""Run the nearest-neighbour search and the ray / triangle parity kernel on seeded inputs (ties, a NaN sample, ragged
sizes) and save the raw outputs: tests/test_gpu_geometry.py runs this once per kernel variant (environment switches are
read once per process) and compares the files bit for bit.

    python scripts/dump_search_kernels.py out.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import functional as Fb  # noqa: E402
from obman_train_b200.icosphere import icosphere  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1234)
out = {}
for tag, (B, N, M) in {"chamfer": (5, 2562, 2500), "contact": (7, 778, 642), "small": (3, 37, 1029)}.items():
    x = torch.randn(B, N, 3, device="cuda", generator=g) * 40
    y = torch.randn(B, M, 3, device="cuda", generator=g) * 40
    y[:, 1::7] = y[:, 0:-1:7][:, : y[:, 1::7].shape[1]]          # duplicated candidates: ties must keep the lowest index
    x[0, : min(N, 50)] = y[0, : min(N, 50)]                          # exact hits (distance 0)
    if tag == "small":
        x[1, 5, 1] = float("nan")                                   # NaN query: its own distance is NaN
        y[2, 9, 0] = float("nan")                                   # NaN candidate: every distance of the sample is NaN
    minx, idxx, miny, idxy = Fb.nearest_neighbours(x, y)
    for k, v in (("minx", minx), ("idxx", idxx), ("miny", miny), ("idxy", idxy)):
        out[tag + "_" + k] = v.cpu().numpy()
verts, faces = icosphere(3)
verts = torch.as_tensor(verts, dtype=torch.float32, device="cuda")
faces_t = torch.as_tensor(faces, dtype=torch.int32, device="cuda").contiguous()
for tag, (B, P) in {"ray_hand": (9, 778), "ray_small": (4, 33)}.items():
    obj = (verts[None] * (30 + 20 * torch.rand(B, 1, 1, device="cuda", generator=g))
           + 3 * torch.randn(B, verts.shape[0], 3, device="cuda", generator=g)).contiguous()
    pts = (35 * torch.randn(B, P, 3, device="cuda", generator=g)).contiguous()
    exterior, hits = Fb.mesh_exterior(pts, obj, faces_t)
    out[tag + "_hits"] = hits.cpu().numpy()
torch.cuda.synchronize()
np.savez(sys.argv[1], **out)
print("dumped", len(out), "arrays")
