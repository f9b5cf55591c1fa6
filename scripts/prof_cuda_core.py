"""One launch each of the CUDA-core kernels that changed in round 2 (for `ncu --set full`): packed nearest-neighbour search,
streamed ray / triangle parity, the decoder's first-layer forward / backward kernels, max-pool backward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from obman_train_b200 import functional as Fb  # noqa: E402
from obman_train_b200.icosphere import icosphere  # noqa: E402
from obman_train_b200.networks.branches.atlasutils import PointGenCon  # noqa: E402

B = int(os.environ.get("PROF_B", "128"))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 2562, 3, device="cuda", generator=g) * 50
y = torch.randn(B, 2500, 3, device="cuda", generator=g) * 50
Fb.nearest_neighbours(x, y)
verts, faces = icosphere(4)
verts = torch.as_tensor(verts, dtype=torch.float32, device="cuda")
faces = torch.as_tensor(faces, dtype=torch.int32, device="cuda").contiguous()
obj = (verts[None] * 50 + 5 * torch.randn(B, verts.shape[0], 3, device="cuda", generator=g)).contiguous()
pts = (60 * torch.randn(B, 778, 3, device="cuda", generator=g)).contiguous()
Fb.mesh_exterior(pts, obj, faces)
dec = PointGenCon(bottleneck_size=515, out_factor=200).cuda().eval()
feat = torch.randn(B, 512, device="cuda", generator=g).requires_grad_(True)
out = dec.decode(feat, verts)
out.sum().backward()
torch.cuda.synchronize()
print("done")
