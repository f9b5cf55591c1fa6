"""The C-ABI library builds, loads without a GPU driver and exports every symbol the header declares."""
import ctypes
import os
import re

from obman_train_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    protos = _lib.parse_header()
    assert len(protos) >= 10
    for name in protos:
        assert hasattr(lib, name), "missing export " + name
    with open(os.path.join(ROOT, "include", "obman_b200.h")) as f:
        declared = set(re.findall(r"\b(obman_\w+)\s*\(", f.read()))
    assert declared == set(protos), declared ^ set(protos)


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.obman_version() >= 100
    assert isinstance(lib.obman_get_last_error(), bytes)


def test_bad_arguments_are_rejected_before_any_launch():
    lib = _lib.load()
    rc = lib.obman_nn_fwd(None, None, 0, 0, 0, None, None, None, None, 3, None)
    assert rc == -1
    assert b"obman_nn_fwd" in lib.obman_get_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "obman_train_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)


def test_every_public_op_refuses_cpu_tensors_instead_of_falling_back():
    """There is no CPU path: each Python entry point of the product must raise on host tensors."""
    import numpy as np
    import pytest
    import torch
    from obman_train_b200 import dense, encoder, functional as Fb, mlp
    from obman_train_b200.manopth.manolayer import ManoLayer
    from obman_train_b200.networks.branches.atlasbranch import edge_loss
    from obman_train_b200.networks.branches.atlasutils import ChamferLoss
    from obman_train_b200.networks.branches.laplacianloss import LaplacianLoss
    from obman_train_b200.icosphere import icosphere
    x = torch.randn(2, 12, 3)
    y = torch.randn(2, 9, 3)
    verts, faces = icosphere(0)
    with pytest.raises(RuntimeError, match="CUDA"):
        Fb.chamfer(x, y)
    with pytest.raises(RuntimeError, match="CUDA"):
        ChamferLoss()(x, y)
    with pytest.raises(RuntimeError, match="CUDA"):
        Fb.nearest_neighbours(x, y)
    with pytest.raises(RuntimeError, match="CUDA"):
        Fb.mesh_exterior(x, x, torch.tensor(np.asarray(faces), dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA"):
        LaplacianLoss(faces, torch.tensor(verts, dtype=torch.float32))(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        edge_loss(x, faces)
    with pytest.raises(RuntimeError, match="CUDA"):
        dense.gemm(torch.randn(4, 32), torch.randn(8, 32), passes=3)
    with pytest.raises(RuntimeError):
        mlp.linear(torch.randn(4, 32), torch.randn(8, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        encoder.resnet18_features(torch.randn(1, 3, 64, 64), [])
    layer = ManoLayer(center_idx=0, ncomps=30, side="right", mano_root="synthetic")
    with pytest.raises(RuntimeError):
        layer(torch.zeros(2, 33), th_betas=torch.zeros(2, 10))
