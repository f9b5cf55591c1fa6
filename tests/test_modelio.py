"""Checkpoint interchange: reference-format checkpoints (DataParallel ``module.`` keys,
``torch.optim.Adam`` state) load strictly into the drop-in HandNet / FlatAdamTrainer, and back."""
import os
import warnings

import pytest
import torch

from obman_train_b200.modelutils import modelio
from obman_train_b200.trainer import FlatAdamTrainer
from tests.util import FULL_CFG


def _product_model(seed):
    from obman_train_b200.networks.handnet import HandNet
    torch.manual_seed(seed)
    return HandNet(**FULL_CFG).eval()


def _reference_style_checkpoint(model, path, steps=3, prefix="module."):
    """What traineval.py:374-384 writes: DataParallel keys + a real torch.optim.Adam state after `steps` updates."""
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    g = torch.Generator().manual_seed(5)
    for _ in range(steps):
        for n, p in model.named_parameters():
            # like the reference: ``fc`` never runs, so it never gets a gradient (and no optimizer state)
            p.grad = None if ".fc." in "." + n + "." else torch.randn(p.shape, generator=g) * 1e-2
        opt.step()
    sd = {prefix + k: v.clone() for k, v in model.state_dict().items()}
    torch.save({"epoch": 7, "network": "handnet", "state_dict": sd, "best_score": 0.25, "optimizer": opt.state_dict()},
               path)
    return opt


def test_reference_format_checkpoint_loads_strictly(tmp_path):
    src = _product_model(0)
    path = str(tmp_path / "checkpoint.pth.tar")
    opt = _reference_style_checkpoint(src, path)
    dst = _product_model(1)
    trainer = FlatAdamTrainer(dst, lr=3e-4)
    ptrs = [p.data_ptr() for p in trainer.params]
    epoch, best = modelio.load_checkpoint(dst, path, optimizer=trainer, strict=True)
    assert (epoch, best) == (7, 0.25)
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    # loaded in place: parameters are still views of the flat buffer (graph-captured addresses stay valid)
    assert ptrs == [p.data_ptr() for p in trainer.params]
    base = trainer.flat_p.untyped_storage().data_ptr()
    assert all(p.untyped_storage().data_ptr() == base for p in trainer.params)
    # optimizer: moments, step counter and hyper-parameters
    assert trainer.step_count == 3 and trainer.lr == pytest.approx(1e-4)
    ref_state = opt.state_dict()["state"]
    n_fc = sum(1 for n, _ in dst.named_parameters() if ".fc." in "." + n + ".")
    assert len(ref_state) == trainer.optim_len - n_fc  # fc parameters never had state
    for p, off, idx in zip(trainer.params, trainer.offsets, trainer.optim_index):
        n = p.numel()
        assert torch.equal(trainer.exp_avg[off:off + n].view_as(p), ref_state[idx]["exp_avg"])
        assert torch.equal(trainer.exp_avg_sq[off:off + n].view_as(p), ref_state[idx]["exp_avg_sq"])


def test_trainer_state_dict_resumes_under_torch_adam(tmp_path):
    src = _product_model(0)
    path = str(tmp_path / "checkpoint.pth.tar")
    opt = _reference_style_checkpoint(src, path)
    dst = _product_model(1)
    trainer = FlatAdamTrainer(dst)
    modelio.load_checkpoint(dst, path, optimizer=trainer)
    # write a checkpoint from the product side and resume a plain torch.optim.Adam from it (reference's traineval.py)
    os.makedirs(str(tmp_path / "exp"), exist_ok=True)
    modelio.save_checkpoint({"epoch": 8, "network": "handnet", "state_dict": dst.state_dict(), "best_score": 0.2,
                             "optimizer": trainer.state_dict()}, is_best=True, checkpoint=str(tmp_path / "exp"),
                            snapshot=4)
    for name in ("checkpoint.pth.tar", "checkpoint_8.pth.tar", "model_best.pth.tar"):
        assert os.path.isfile(str(tmp_path / "exp" / name)), name
    ckpt = torch.load(str(tmp_path / "exp" / "checkpoint.pth.tar"), weights_only=False)
    third = _product_model(2)
    opt3 = torch.optim.Adam([p for p in third.parameters() if p.requires_grad], lr=1.0)
    opt3.load_state_dict(ckpt["optimizer"])
    a, b = opt.state_dict(), opt3.state_dict()
    assert a["state"].keys() == b["state"].keys()
    for k in a["state"]:
        assert torch.equal(a["state"][k]["exp_avg"], b["state"][k]["exp_avg"])
        assert torch.equal(a["state"][k]["exp_avg_sq"], b["state"][k]["exp_avg_sq"])
        assert float(a["state"][k]["step"]) == float(b["state"][k]["step"])
    assert b["param_groups"][0]["lr"] == pytest.approx(1e-4)


def test_strict_mismatch_raises_and_missing_file_is_value_error(tmp_path):
    src = _product_model(0)
    path = str(tmp_path / "c.pth.tar")
    _reference_style_checkpoint(src, path, steps=1, prefix="")  # keys without the DataParallel prefix also load
    cfg = dict(FULL_CFG, atlas_separate_encoder=False)
    from obman_train_b200.networks.handnet import HandNet
    other = HandNet(**cfg).eval()
    with pytest.raises(RuntimeError, match="Unexpected key"):
        modelio.load_checkpoint(other, path, strict=True)
    modelio.load_checkpoint(other, path, strict=False)  # the reference's fallback (reload.py:100-108)
    with pytest.raises(ValueError, match="no checkpoint found"):
        modelio.load_checkpoint(other, str(tmp_path / "nope.tar"))
    # optimizer of a different parameter count: warning, not an exception (modelio.py:57-70)
    trainer = FlatAdamTrainer(other)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        modelio.load_checkpoint(other, path, optimizer=trainer, strict=False)
    assert any("load optimizer" in str(x.message) for x in w)


def test_load_atlas_remaps_encoder_and_checkpoint_averaging(tmp_path):
    from obman_train_b200.networks.handnet import HandNet
    shared = HandNet(**dict(FULL_CFG, atlas_separate_encoder=False)).eval()
    p1, p2 = str(tmp_path / "a.tar"), str(tmp_path / "b.tar")
    _reference_style_checkpoint(shared, p1, steps=0)
    sep = _product_model(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        modelio.load_checkpoint(sep, p1, strict=False, load_atlas=True)
    for k, v in shared.base_net.state_dict().items():
        assert torch.equal(sep.atlas_base_net.state_dict()[k], v), k
    # averaging (modelio.py:10-28)
    other = HandNet(**dict(FULL_CFG, atlas_separate_encoder=False)).eval()
    _reference_style_checkpoint(other, p2, steps=0)
    avg = HandNet(**dict(FULL_CFG, atlas_separate_encoder=False)).eval()
    epoch, _ = modelio.load_checkpoints(avg, [p1, p2])
    assert epoch == 7
    k = "base_net.layer1.0.conv1.weight"
    assert torch.allclose(avg.state_dict()[k], (shared.state_dict()[k] + other.state_dict()[k]) / 2)
    assert avg.state_dict()["base_net.bn1.num_batches_tracked"].dtype == torch.int64


@pytest.mark.reference
def test_reference_handnet_state_dict_loads_into_product(tmp_path, mano_tables_np):
    """The REAL reference HandNet (executed from /root/reference through the import hook), wrapped in DataParallel as
    traineval.py:130 does, saved as the reference saves it -> strict load into the product model."""
    from oracle import refhook
    refhook.set_mano_tables(mano_tables_np["right"], mano_tables_np["left"])
    refhook.install()
    from mano_train.networks.handnet import HandNet as RefHandNet
    torch.manual_seed(11)
    ref = RefHandNet(**{k: v for k, v in FULL_CFG.items()})
    wrapped = torch.nn.DataParallel(ref)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, ref.parameters()), lr=1e-4)
    path = str(tmp_path / "checkpoint.pth.tar")
    torch.save({"epoch": 1, "network": "handnet", "state_dict": wrapped.state_dict(), "best_score": 1.0,
                "optimizer": opt.state_dict()}, path)
    dst = _product_model(4)
    trainer = FlatAdamTrainer(dst)
    modelio.load_checkpoint(dst, path, optimizer=trainer, strict=True)
    own = dst.state_dict()
    for k, v in ref.state_dict().items():
        assert torch.equal(own[k], v), k
    assert trainer.optim_len == len(opt.state_dict()["param_groups"][0]["params"])


def test_reload_model_rebuilds_from_opts_and_checkpoint(tmp_path):
    """netscripts.reload.reload_model (reload.py:35-111): options missing from old ``opt.pkl`` files get the reference's
    defaults; a checkpoint with one key missing falls back to the non-strict load with a warning."""
    import pickle
    import numpy as np
    from obman_train_b200.netscripts import reload as rl
    opts = {"atlas_lambda_regul_edges": 0.0, "atlas_lambda": 0.167, "center_idx": 0, "hidden_neurons": [1024, 256],
            "use_shape": True, "mano_lambda_verts": 0.167, "atlas_predict_trans": True, "atlas_predict_scale": True,
            "atlas_final_lambda": 0.167, "mano_lambda_joints3d": 0.167}
    exp = tmp_path / "exp"
    exp.mkdir()
    with open(str(exp / "opt.pkl"), "wb") as f:
        pickle.dump(opts, f)
    assert rl.get_opts(str(exp / "checkpoint.pth.tar")) == opts and rl.get_opts(str(exp)) == opts
    src = rl.reload_model.__globals__["HandNet"](
        resnet_version=18, atlas_mesh=True, atlas_points_nb=642, atlas_lambda_regul_edges=0.0, atlas_lambda=0.167,
        atlas_final_lambda=0.167, atlas_predict_trans=True, atlas_predict_scale=True, atlas_ico_divisions=3,
        mano_root="synthetic", mano_center_idx=0, mano_comps=30, mano_neurons=[1024, 256], mano_use_shape=True,
        mano_use_pca=True, mano_lambda_verts=0.167, mano_lambda_joints3d=0.167)
    torch.manual_seed(9)
    for p in src.parameters():
        p.data.normal_()
    state = {"module." + k: v for k, v in src.state_dict().items()}
    torch.save({"epoch": 3, "state_dict": state, "best_score": 1.0}, str(exp / "checkpoint.pth.tar"))
    model = rl.reload_model(str(exp / "checkpoint.pth.tar"), opts, mano_root="synthetic", ico_divisions=3)
    assert not model.training
    for (k, a), (_, b) in zip(src.state_dict().items(), model.state_dict().items()):
        assert torch.equal(a, b), k
    # one parameter missing: strict load fails, the reference's non-strict fallback kicks in (reload.py:100-108)
    del state["module.atlas_branch.decode_scale.2.bias"]
    torch.save({"epoch": 3, "state_dict": state, "best_score": 1.0}, str(exp / "partial.pth.tar"))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        rl.reload_model(str(exp / "partial.pth.tar"), opts, mano_root="synthetic")
    assert any("trying without strict" in str(x.message) for x in w)
    # no_beta drops the shape regressor
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nb = rl.reload_model(str(exp / "checkpoint.pth.tar"), opts, mano_root="synthetic", no_beta=True)
    assert not any(k.startswith("mano_branch.shape_reg") for k in nb.state_dict())
    out = tmp_path / "m.obj"
    rl.save_obj(str(out), np.array([[0.0, 1.0, 2.0], [1, 0, 0], [0, 0, 1]]), np.array([[0, 1, 2]]))
    assert open(str(out)).read().splitlines() == ["v 0.000000 1.000000 2.000000", "v 1.000000 0.000000 0.000000",
                                                  "v 0.000000 0.000000 1.000000", "f 1 2 3"]
