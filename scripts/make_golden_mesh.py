"""Generate tests/golden/mesh_regul_golden.npz by RUNNING THE REFERENCE's own code (container only):
``laplacianloss.cotangent`` + ``Laplacian.forward`` / ``Laplacian.backward`` called as plain methods (the legacy
autograd plumbing around them no longer runs on torch >= 1.5, the numerical body does) and ``atlasbranch.edge_loss``
with autograd.  Seeded inputs on the ico-2 sphere (162 vertices / 320 faces).

    python scripts/make_golden_mesh.py
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import icosphere, refhook  # noqa: E402
from obman_train_b200.manopth.synthetic import synthetic_mano_tables  # noqa: E402


def main():
    refhook.set_mano_tables(synthetic_mano_tables("right"), synthetic_mano_tables("left"))
    refhook.install()
    from mano_train.networks.branches import laplacianloss
    from mano_train.networks.branches.atlasbranch import edge_loss

    g = torch.Generator().manual_seed(4321)
    verts, faces = icosphere.icosphere(2)
    sphere = torch.tensor(verts, dtype=torch.float32)
    out = {"sphere_verts": verts, "faces": faces.astype(np.int64)}
    # deformed spheres in mm, the kind of mesh the decoder emits early in training
    V = (sphere.unsqueeze(0) * 40 + torch.randn(3, verts.shape[0], 3, generator=g) * 8).contiguous()
    lap = laplacianloss.Laplacian(faces[None].astype(np.int64), sphere)
    lx = lap.forward(V)
    loss = torch.norm(lx.view(-1, 3), p=2, dim=1).mean()
    # gradient of the mean row norm w.r.t. Lx, pushed through the reference's own backward (L g)
    norms = torch.norm(lx, dim=2, keepdim=True)
    g_lx = lx / norms / (lx.shape[0] * lx.shape[1])
    gV = lap.backward(g_lx)
    out.update(lap_V=V.numpy(), lap_Lx=lx.numpy(), lap_loss=np.float32(loss.item()), lap_gV=gV.numpy(),
               lap_cot=laplacianloss.cotangent(sphere.unsqueeze(0), torch.tensor(faces[None].astype(np.int64))).numpy())
    Ve = V.clone().requires_grad_(True)
    el = edge_loss(Ve, faces.astype(np.int64))
    el.backward()
    out.update(edge_loss=np.float32(el.item()), edge_gV=Ve.grad.numpy())
    path = os.path.join(ROOT, "tests", "golden", "mesh_regul_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", ()) for k, v in out.items()})


if __name__ == "__main__":
    main()
