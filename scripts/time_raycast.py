"""Time the ray / triangle parity kernel at the workload size (778 hand vertices x 5120 object faces, B = 256)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200._lib import call, ptr, stream_ptr  # noqa: E402
from obman_train_b200.icosphere import icosphere  # noqa: E402

B = int(os.environ.get("PROF_B", "256"))
verts, faces = icosphere(4)
verts = torch.as_tensor(verts, dtype=torch.float32, device="cuda")
faces = torch.as_tensor(faces, dtype=torch.int32, device="cuda").contiguous()
g = torch.Generator(device="cuda").manual_seed(0)
obj = (verts[None] * (40 + 20 * torch.rand(B, 1, 1, device="cuda", generator=g))
       + 5 * torch.randn(B, verts.shape[0], 3, device="cuda", generator=g)).contiguous()
pts = (60 * torch.randn(B, 778, 3, device="cuda", generator=g)).contiguous()
hits = torch.empty(B, 778, dtype=torch.int32, device="cuda")
scratch = torch.empty(B, (faces.shape[0] + 1) // 2, 32, device="cuda")
fn = lambda: call("obman_raycast_hits", ptr(pts), ptr(obj), ptr(faces), B, 778, obj.shape[1], faces.shape[0], ptr(hits), ptr(scratch), stream_ptr())  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("raycast B=%d 778 x %d faces: %.4f ms  %.2f Gtests/s  packed=%s  hits checksum %d  inside %d" % (
    B, faces.shape[0], ms, B * 778.0 * faces.shape[0] / ms / 1e6, os.environ.get("OBMAN_RAYCAST_PACKED", "1") + "/stream=" + os.environ.get("OBMAN_RAYCAST_STREAM", "1"),
    int(hits.sum()), int((hits % 2 == 1).sum())))
