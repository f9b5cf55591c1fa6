"""The licensed MANO_{LEFT,RIGHT}.pkl files are pickles that contain chumpy objects (chumpy is absent here and no longer
installs on current numpy).  ``manopth.manolayer.load_mano_pickle`` reads them through an Unpickler that maps every
``chumpy.*`` class onto a stub.  The real files cannot be shipped, so the test fabricates a pickle
with the same structure: a ``chumpy.ch.Ch`` instance whose state dict carries the array under ``x``, a scipy sparse
``J_regressor`` and plain ndarrays, written while a fake ``chumpy`` package is importable and read back after it is gone."""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from obman_train_b200.manopth.manolayer import ManoLayer, load_mano_pickle
from obman_train_b200.manopth.synthetic import synthetic_mano_tables


def _write_fake_mano(path, tables):
    pkg, mod = types.ModuleType("chumpy"), types.ModuleType("chumpy.ch")

    class Ch(object):  # pickles as GLOBAL chumpy.ch.Ch + BUILD(state dict), like a chumpy object
        def __init__(self, x):
            self.x = np.asarray(x)
            self._dirty_vars = set()
            self._itr = None

    Ch.__module__, Ch.__qualname__ = "chumpy.ch", "Ch"
    mod.Ch = Ch
    pkg.ch = mod
    sys.modules["chumpy"], sys.modules["chumpy.ch"] = pkg, mod
    try:
        data = {
            "v_template": tables["v_template"], "shapedirs": Ch(tables["shapedirs"]), "posedirs": tables["posedirs"],
            "J_regressor": sp.csc_matrix(tables["J_regressor"]), "weights": tables["weights"],
            "hands_mean": tables["hands_mean"], "hands_components": tables["hands_components"],
            "f": tables["f"].astype(np.uint32), "kintree_table": np.zeros((2, 16), dtype=np.int64), "bs_style": "lbs",
        }
        with open(path, "wb") as f:
            pickle.dump(data, f, protocol=2)
    finally:
        del sys.modules["chumpy"], sys.modules["chumpy.ch"]


@pytest.mark.parametrize("side", ["right", "left"])
def test_mano_pickle_with_chumpy_objects_loads_without_chumpy(tmp_path, side):
    tables = synthetic_mano_tables(side)
    fname = "MANO_RIGHT.pkl" if side == "right" else "MANO_LEFT.pkl"
    _write_fake_mano(str(tmp_path / fname), tables)
    assert "chumpy" not in sys.modules
    with pytest.raises(ModuleNotFoundError):
        with open(str(tmp_path / fname), "rb") as f:
            pickle.load(f)          # a plain unpickler needs chumpy ...
    got = load_mano_pickle(str(tmp_path / fname))   # ... the stub unpickler does not
    for key in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "hands_mean", "hands_components"):
        assert np.allclose(np.asarray(got[key], dtype=np.float64), np.asarray(tables[key], dtype=np.float64)), key
    assert np.array_equal(np.asarray(got["f"]).astype(np.int64), tables["f"].astype(np.int64))
    # the drop-in layer built from the file carries the reference's buffer names / shapes (checkpoint keys)
    layer = ManoLayer(center_idx=9, ncomps=30, side=side, mano_root=str(tmp_path), use_pca=True, flat_hand_mean=False)
    ref = ManoLayer(center_idx=9, ncomps=30, side=side, mano_root="synthetic", use_pca=True, flat_hand_mean=False)
    a, b = dict(layer.named_buffers()), dict(ref.named_buffers())
    assert set(a) == set(b) == {"th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor",
                                "th_weights", "th_faces", "th_hands_mean", "th_comps", "th_selected_comps"}
    for k in a:
        assert a[k].shape == b[k].shape and torch.allclose(a[k].double(), b[k].double()), k
    assert a["th_selected_comps"].shape == (30, 45) and a["th_faces"].dtype == torch.int64


def test_missing_mano_file_names_the_synthetic_escape_hatch(tmp_path):
    with pytest.raises(FileNotFoundError, match="synthetic"):
        ManoLayer(side="right", mano_root=str(tmp_path))
