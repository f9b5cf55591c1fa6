"""Checkpoint interchange with the reference.

Mirrors /root/reference/mano_train/modelutils/modelio.py: ``load_checkpoint`` (:31-84), ``load_checkpoints``
(:10-28, multi-checkpoint averaging) and ``save_checkpoint`` (:87-104) with the same arguments, return values,
warnings and error behaviour, so that ``release_models/*/checkpoint.pth.tar`` files written by the reference load
strictly into the drop-in ``HandNet`` and checkpoints written here resume under the reference's traineval.py.

Differences that the B200 path needs:
* the reference always wraps the model in ``torch.nn.DataParallel`` (keys carry a ``module.`` prefix,
  modelio.py:35-46); the drop-in model runs one process per GPU without a wrapper, so the prefix is added or
  stripped to match whatever ``model`` is;
* ``optimizer`` may be a ``FlatAdamTrainer`` (flat-buffer fused Adam): its ``load_state_dict`` /
  ``state_dict`` speak ``torch.optim.Adam``'s format;
* parameters are loaded IN PLACE (``copy_``), so the flat parameter buffer of a ``FlatAdamTrainer`` and the
  addresses baked into a captured CUDA graph stay valid;
* ``map_location="cpu"`` on ``torch.load`` (the reference loads to the device the checkpoint was saved from).
"""
import os
import shutil
import traceback
import warnings

import torch


def _read(path):
    try:
        return torch.load(path, map_location="cpu", weights_only=False)
    except TypeError:  # torch < 1.13
        return torch.load(path, map_location="cpu")


def _match_prefix(state_dict, model):
    """Rename checkpoint keys to the convention of ``model`` (with or without DataParallel's ``module.``)."""
    wants = any(k.startswith("module.") for k in model.state_dict().keys())
    out = {}
    for key, val in state_dict.items():
        has = key.startswith("module.")
        if wants and not has:
            key = "module." + key
        elif has and not wants:
            key = key[len("module."):]
        out[key] = val
    return out


def _to_atlas_encoder(state_dict):
    """``load_atlas`` (modelio.py:47-55): an AtlasNet-only checkpoint's encoder becomes the separate Atlas encoder."""
    return {(k.replace("base_net", "atlas_base_net") if "base_net" in k and "atlas_base_net" not in k else k): v
            for k, v in state_dict.items()}


def _load_in_place(model, state_dict, strict):
    """``model.load_state_dict`` semantics (same error text on mismatch) with in-place copies."""
    own = model.state_dict()
    missing = [k for k in own if k not in state_dict]
    unexpected = [k for k in state_dict if k not in own]
    errors = []
    for k, v in state_dict.items():
        if k in own and tuple(own[k].shape) != tuple(v.shape):
            errors.append("size mismatch for {}: copying a param with shape {} from checkpoint, the shape in "
                          "current model is {}.".format(k, tuple(v.shape), tuple(own[k].shape)))
    if strict:
        if unexpected:
            errors.insert(0, "Unexpected key(s) in state_dict: {}. ".format(", ".join('"%s"' % k for k in unexpected)))
        if missing:
            errors.insert(0, "Missing key(s) in state_dict: {}. ".format(", ".join('"%s"' % k for k in missing)))
    if errors:
        raise RuntimeError("Error(s) in loading state_dict for {}:\n\t{}".format(
            model.__class__.__name__, "\n\t".join(errors)))
    with torch.no_grad():
        for k, v in state_dict.items():
            if k in own:
                own[k].copy_(v.to(own[k].dtype))
    return missing, unexpected


def load_checkpoints(model, resume_paths, strict=True):
    """Average several checkpoints element-wise (integer buffers are taken from the last one)."""
    dicts, epochs = [], []
    for path in resume_paths:
        ckpt = _read(path)
        dicts.append(ckpt["state_dict"])
        epochs.append(ckpt["epoch"])
    mean = {}
    for key, last in dicts[-1].items():
        if not torch.is_floating_point(last):
            mean[key] = last
        else:
            mean[key] = torch.stack([d[key] for d in dicts]).mean(0)
    _load_in_place(model, _match_prefix(mean, model), strict)
    return max(epochs), None


def _best_score(checkpoint):
    """The score stored next to the weights: "best_auc" / the deprecated "best_acc" (with the reference's warning) /
    "best_score" (modelio.py:76-83)."""
    for key, deprecated in (("best_auc", False), ("best_acc", True), ("best_score", False)):
        if key in checkpoint:
            if deprecated:
                warnings.warn("Using deprecated best_acc instead of best_auc")
            return checkpoint[key]
    raise KeyError("checkpoint holds none of best_auc / best_acc / best_score")


def _restore_optimizer(optimizer, checkpoint, resume_path):
    """Optimizer state is best effort, as in the reference (modelio.py:60-73): a parameter-group mismatch is reported
    and training continues with a fresh optimizer state."""
    stored = checkpoint["optimizer"]
    absent = set(optimizer.state_dict().keys()) - set(stored.keys())
    if absent:
        warnings.warn("Missing keys in optimizer ! : {}".format(absent))
    try:
        optimizer.load_state_dict(stored)
    except ValueError:
        traceback.print_exc()
        warnings.warn("Couldn' load optimizer from {}".format(resume_path))


def load_checkpoint(model, resume_path, optimizer=None, strict=True, load_atlas=False):
    """-> (epoch, best score).  Same arguments, warnings and errors as modelio.py:31-84; the weights are copied IN PLACE
    into ``model`` whatever prefix convention (``module.`` or none) either side uses."""
    if not os.path.isfile(resume_path):
        raise ValueError("=> no checkpoint found at '{}'".format(resume_path))
    print("=> loading checkpoint '{}'".format(resume_path))
    checkpoint = _read(resume_path)
    weights = checkpoint["state_dict"]
    if load_atlas:
        weights = _to_atlas_encoder(weights)
    weights = _match_prefix(weights, model)
    absent = set(model.state_dict().keys()) - set(weights.keys())
    if absent:
        warnings.warn("Missing keys ! : {}".format(absent))
    _load_in_place(model, weights, strict)
    print("=> loaded checkpoint '{}' (epoch {})".format(resume_path, checkpoint["epoch"]))
    if optimizer is not None:
        _restore_optimizer(optimizer, checkpoint, resume_path)
    return checkpoint["epoch"], _best_score(checkpoint)


def save_checkpoint(state, is_best, checkpoint="checkpoint", filename="checkpoint.pth.tar", snapshot=None):
    """Write ``state`` (the dict traineval.py:374-384 builds: epoch, network, state_dict, best score, optimizer) to
    ``checkpoint/filename``; every ``snapshot``-th epoch also as ``checkpoint_<epoch>.pth.tar``, and as
    ``model_best.pth.tar`` when ``is_best`` (same files as modelio.py:87-104).  The main file is written to a temporary
    name first and renamed, so a job killed mid-write never leaves a truncated checkpoint behind."""
    target = os.path.join(checkpoint, filename)
    partial = target + ".tmp"
    torch.save(state, partial)
    os.replace(partial, target)
    copies = []
    if snapshot and state["epoch"] % snapshot == 0:
        copies.append("checkpoint_{}.pth.tar".format(state["epoch"]))
    if is_best:
        copies.append("model_best.pth.tar")
    for name in copies:
        shutil.copyfile(target, os.path.join(checkpoint, name))
