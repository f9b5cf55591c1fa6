"""Import hook that executes the reference's own, unmodified files on CPU (oracle; container only).

``/root/reference`` is mounted read-only in the build container and does NOT exist on the GPU box, so
this module is used only (a) by tests that pin ``oracle.geometry`` / ``oracle.nets`` against the real
reference (skipped when the mount is absent), (b) by scripts/make_golden.py to generate the committed
fixtures under tests/golden/.  Nothing is copied: sources are read from where they lie and compiled in
memory with two byte-level substitutions:

  ``.cuda()`` -> ````            (no GPU in the container; several modules call it at import time)
  ``pretrained=True`` -> ``pretrained=False``   (no network for the ImageNet weights)

plus stub modules for the absent third-party packages: ``trimesh.creation.icosphere`` ->
``oracle.icosphere``; ``manopth.manolayer.ManoLayer`` -> ``oracle.mano`` on tables registered with
``set_mano_tables``; empty ``matplotlib`` / ``mpl_toolkits`` / ``mano_train.visualize.displaymano``.
"""
import importlib.abc
import importlib.util
import os
import sys
import types

import numpy as np
import torch

from . import icosphere as _ico
from . import mano as _mano

REF_ROOT = "/root/reference"
_PREFIXES = ("mano_train", "handobjectdatasets")
_MANO_TABLES = {}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mano_train"))


def set_mano_tables(right, left):
    """Register the (numpy) MANO tables the stubbed ManoLayer should use."""
    _MANO_TABLES["right"] = right
    _MANO_TABLES["left"] = left


class _RefLoader(importlib.abc.Loader):
    def __init__(self, path, is_pkg):
        self.path = path
        self.is_pkg = is_pkg

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        if self.is_pkg and not os.path.exists(self.path):
            return  # namespace-like package without __init__.py
        with open(self.path, "r") as f:
            src = f.read()
        src = src.replace(".cuda()", "").replace("pretrained=True", "pretrained=False")
        exec(compile(src, self.path, "exec"), module.__dict__)


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] not in _PREFIXES:
            return None
        rel = fullname.replace(".", "/")
        pkg_dir = os.path.join(REF_ROOT, rel)
        if os.path.isdir(pkg_dir):
            init = os.path.join(pkg_dir, "__init__.py")
            spec = importlib.util.spec_from_loader(fullname, _RefLoader(init, True), is_package=True)
            spec.submodule_search_locations = [pkg_dir]
            return spec
        if os.path.exists(pkg_dir + ".py"):
            return importlib.util.spec_from_loader(fullname, _RefLoader(pkg_dir + ".py", False))
        return None


class _StubManoLayer(torch.nn.Module):
    """Signature of manopth.manolayer.ManoLayer as used at manobranch.py:92-105,170-182."""

    def __init__(self, center_idx=None, flat_hand_mean=True, ncomps=6, side="right",
                 mano_root="mano/models", use_pca=True, root_rot_mode="axisang",
                 joint_rot_mode="axisang", robust_rot=False):
        super().__init__()
        t = _MANO_TABLES[side]
        self.side, self.center_idx, self.ncomps, self.use_pca = side, center_idx, ncomps, use_pca
        comps = np.asarray(t["hands_components"])
        mean = np.zeros(45) if flat_hand_mean else np.asarray(t["hands_mean"])
        f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)  # noqa: E731
        self.register_buffer("th_betas", f32(t["betas"]).view(1, 10))
        self.register_buffer("th_shapedirs", f32(t["shapedirs"]))
        self.register_buffer("th_posedirs", f32(t["posedirs"]))
        self.register_buffer("th_v_template", f32(t["v_template"]).unsqueeze(0))
        self.register_buffer("th_J_regressor", f32(t["J_regressor"]))
        self.register_buffer("th_weights", f32(t["weights"]))
        self.register_buffer("th_faces", torch.tensor(np.asarray(t["f"]).astype(np.int64)))
        self.register_buffer("th_hands_mean", f32(mean).unsqueeze(0))
        self.register_buffer("th_comps", f32(comps))
        self.register_buffer("th_selected_comps", f32(comps[:ncomps]))

    def forward(self, th_pose_coeffs, th_betas=torch.zeros(1), th_trans=torch.zeros(1),
                root_palm=torch.Tensor([0]), share_betas=torch.Tensor([0])):
        tables = {k: v for k, v in self.named_buffers()}
        return _mano.mano_forward(tables, th_pose_coeffs, th_betas, th_trans, bool(root_palm),
                                  self.side, self.center_idx, self.use_pca, self.ncomps)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Install the finder and the stubs (idempotent).  Also chdir-independent asset lookup: the
    reference opens "assets/contact_zones.pkl" relative to cwd, so callers should run
    ``with refhook.cwd():`` around calls that reach contactloss.py:263."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("/root/reference is not mounted here")
    sys.meta_path.insert(0, _RefFinder())

    class _Mesh(object):
        def __init__(self, v, f):
            self.vertices, self.faces = v, f

    creation = _stub("trimesh.creation",
                     icosphere=lambda subdivisions=3, **kw: _Mesh(*_ico.icosphere(subdivisions)))
    _stub("trimesh", creation=creation)
    manolayer = _stub("manopth.manolayer", ManoLayer=_StubManoLayer)
    _stub("manopth", manolayer=manolayer)
    if "matplotlib" not in sys.modules:
        pyplot = _stub("matplotlib.pyplot")
        _stub("matplotlib", pyplot=pyplot)
        art3d = _stub("mpl_toolkits.mplot3d.art3d", Poly3DCollection=object)
        mplot3d = _stub("mpl_toolkits.mplot3d", art3d=art3d)
        _stub("mpl_toolkits", mplot3d=mplot3d)
    _stub("mano_train.visualize.displaymano")
    _installed = True
    # batch_pairwise_dist's use_cuda=True default selects torch.cuda.LongTensor (contactloss.py:60-69)
    from mano_train.networks.branches import contactloss
    contactloss.batch_pairwise_dist.__defaults__ = (False,)


class cwd(object):
    def __enter__(self):
        self.prev = os.getcwd()
        os.chdir(REF_ROOT)

    def __exit__(self, *a):
        os.chdir(self.prev)
