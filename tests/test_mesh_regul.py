"""Mesh regularisers: cotangent-Laplacian loss and edge-length regulariser.
CPU part: the oracle restatement against golden vectors produced by the reference's own code
(scripts/make_golden_mesh.py), and the host-side table builders of the product.  GPU part: the CUDA kernels
(csrc/mesh_regul.cu, through the C ABI) against the fp64 oracle and the reference golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry, icosphere

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mesh_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "mesh_regul_golden.npz"))


def test_oracle_laplacian_matches_reference_golden(mesh_golden):
    g = mesh_golden
    L = geometry.laplacian_matrix(g["sphere_verts"], g["faces"])
    assert np.abs(L - L.T).max() < 1e-12 and np.abs(L.sum(1)).max() < 1e-12
    V = torch.tensor(g["lap_V"], dtype=torch.float64, requires_grad=True)
    loss, lx = geometry.laplacian_loss(V, L)
    loss.backward()
    # the reference multiplies in fp32 (scipy csr of fp32 cotangents): 1e-5 of the largest entry
    assert np.abs(lx.detach().numpy() - g["lap_Lx"]).max() < 1e-5 * np.abs(g["lap_Lx"]).max()
    assert abs(loss.item() - float(g["lap_loss"])) < 1e-6 * float(g["lap_loss"])
    assert np.abs(V.grad.numpy() - g["lap_gV"]).max() < 1e-5 * np.abs(g["lap_gV"]).max()


def test_oracle_edge_loss_matches_reference_golden(mesh_golden):
    g = mesh_golden
    V = torch.tensor(g["lap_V"], dtype=torch.float64, requires_grad=True)
    el = geometry.edge_loss(V, g["faces"])
    el.backward()
    assert abs(el.item() - float(g["edge_loss"])) < 1e-5 * float(g["edge_loss"])
    assert np.abs(V.grad.numpy() - g["edge_gV"]).max() < 1e-4 * np.abs(g["edge_gV"]).max()


@pytest.mark.parametrize("sub", [1, 3])
def test_product_tables_match_dense_laplacian(sub):
    from obman_train_b200.networks.branches.laplacianloss import laplacian_ell, vertex_face_table
    v, f = icosphere.icosphere(sub)
    L = geometry.laplacian_matrix(v, f)
    nbr, w = laplacian_ell(v, f)
    assert nbr.dtype == np.int32 and w.dtype == np.float32 and nbr.shape[1] <= 6
    D = np.zeros_like(L)
    np.add.at(D, (np.repeat(np.arange(len(v)), nbr.shape[1]), nbr.reshape(-1)), w.reshape(-1).astype(np.float64))
    off = L - np.diag(np.diag(L))
    assert np.abs(D - off).max() < 1e-6 * np.abs(off).max()
    vf = vertex_face_table(len(v), f)
    for i in (0, 5, len(v) - 1):
        inc = sorted(int(x) for x in vf[i] if x >= 0)
        assert inc == sorted(np.nonzero((f == i).any(1))[0].tolist())
    assert (vf >= 0).sum() == 3 * len(f)


@pytest.mark.reference
def test_oracle_laplacian_matches_reference_code_live(mano_tables_np):
    from oracle import refhook
    refhook.set_mano_tables(mano_tables_np["right"], mano_tables_np["left"])
    refhook.install()
    from mano_train.networks.branches import laplacianloss
    v, f = icosphere.icosphere(3)
    sphere = torch.tensor(v, dtype=torch.float32)
    V = sphere.unsqueeze(0) * 50 + torch.randn(2, len(v), 3, generator=torch.Generator().manual_seed(3)) * 5
    lap = laplacianloss.Laplacian(f[None].astype(np.int64), sphere)
    lx_ref = lap.forward(V)
    loss, lx = geometry.laplacian_loss(V.double(), geometry.laplacian_matrix(v, f))
    assert (lx_ref.double() - lx).abs().max() < 1e-5 * lx.abs().max()
    ref_loss = torch.norm(lx_ref.view(-1, 3), p=2, dim=1).mean()
    assert abs(ref_loss.item() - loss.item()) < 1e-5 * loss.item()


# ------------------------------------------------------------------------------------------------------------------
# CUDA kernels
# ------------------------------------------------------------------------------------------------------------------
def _cuda_tables(v, f):
    from obman_train_b200.networks.branches.laplacianloss import laplacian_ell, vertex_face_table
    nbr, w = laplacian_ell(v, f)
    return (torch.from_numpy(nbr).cuda(), torch.from_numpy(w).cuda(),
            torch.from_numpy(f.astype(np.int32)).cuda(), torch.from_numpy(vertex_face_table(len(v), f)).cuda())


@pytest.mark.gpu
def test_cuda_mesh_regularisers_match_reference_golden(mesh_golden):
    from obman_train_b200 import functional as Fb
    g = mesh_golden
    nbr, w, faces, vf = _cuda_tables(g["sphere_verts"], g["faces"])
    V = torch.tensor(g["lap_V"]).cuda().requires_grad_(True)
    loss, lx = Fb.laplacian_loss(V, nbr, w)
    loss.backward()
    assert np.abs(lx.cpu().numpy() - g["lap_Lx"]).max() < 1e-4 * np.abs(g["lap_Lx"]).max()
    assert abs(loss.item() - float(g["lap_loss"])) < 1e-4 * float(g["lap_loss"])
    assert np.abs(V.grad.cpu().numpy() - g["lap_gV"]).max() < 1e-4 * np.abs(g["lap_gV"]).max()
    V2 = torch.tensor(g["lap_V"]).cuda().requires_grad_(True)
    el = Fb.edge_loss(V2, faces, vf)
    el.backward()
    assert abs(el.item() - float(g["edge_loss"])) < 1e-4 * float(g["edge_loss"])
    assert np.abs(V2.grad.cpu().numpy() - g["edge_gV"]).max() < 1e-4 * np.abs(g["edge_gV"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("sub,B", [(3, 5), (4, 2), (0, 1)])
def test_cuda_mesh_regularisers_match_fp64_oracle(sub, B):
    from obman_train_b200 import functional as Fb
    v, f = icosphere.icosphere(sub)
    nbr, w, faces, vf = _cuda_tables(v, f)
    gen = torch.Generator().manual_seed(10 + sub)
    V = torch.tensor(v, dtype=torch.float32).unsqueeze(0) * 45 + torch.randn(B, len(v), 3, generator=gen) * 6
    # fp64 oracle with autograd; the upstream gradient is a non-trivial scalar to exercise gloss
    Vo = V.double().requires_grad_(True)
    lo, lxo = geometry.laplacian_loss(Vo, geometry.laplacian_matrix(v, f))
    eo = geometry.edge_loss(Vo, f)
    (0.3 * lo + 0.7 * eo).backward()
    Vc = V.cuda().requires_grad_(True)
    lc, lxc = Fb.laplacian_loss(Vc, nbr, w)
    ec = Fb.edge_loss(Vc, faces, vf)
    (0.3 * lc + 0.7 * ec).sum().backward()
    assert abs(lc.item() - lo.item()) < 1e-4 * abs(lo.item())
    assert abs(ec.item() - eo.item()) < 1e-4 * abs(eo.item())
    assert (lxc.cpu().double() - lxo).abs().max() < 1e-4 * lxo.abs().max()
    gd = Vc.grad.cpu().double() - Vo.grad
    assert gd.abs().max() < 1e-4 * Vo.grad.abs().max(), (gd.abs().max(), Vo.grad.abs().max())
    # bit-reproducible (no float atomics anywhere on this path)
    lc2, _ = Fb.laplacian_loss(Vc.detach(), nbr, w)
    assert torch.equal(lc2, lc.detach())


@pytest.mark.gpu
def test_cuda_laplacian_zero_rows_have_zero_gradient():
    """A flat (all-equal) mesh gives Lx = 0 everywhere: torch.norm's sub-gradient at 0 is 0, no NaN."""
    from obman_train_b200 import functional as Fb
    v, f = icosphere.icosphere(1)
    nbr, w, faces, vf = _cuda_tables(v, f)
    V = torch.full((2, len(v), 3), 7.0, device="cuda", requires_grad=True)
    loss, _ = Fb.laplacian_loss(V, nbr, w)
    el = Fb.edge_loss(V, faces, vf)
    (loss + el).sum().backward()
    assert loss.item() == 0.0 and el.item() == 0.0
    assert torch.isfinite(V.grad).all() and V.grad.abs().max().item() == 0.0
