"""Rebuild a trained model from its experiment folder - drop-in for
/root/reference/mano_train/netscripts/reload.py:16-111 (``save_obj``, ``get_opts``, ``reload_model``), the entry point of
the reference's demos (image_demo.py:66-68, webcam_demo.py) and of netscripts/simulate.py.

Same arguments and defaults-for-missing-options as the reference; differences: the model is NOT wrapped in
``torch.nn.DataParallel`` (one process per GPU; ``modelio.load_checkpoint`` strips the ``module.`` prefix of the stored
keys), weights are loaded in place, and the module stays on the CPU until the caller moves it (``model.cuda()``) - the
reference's constructor calls ``.cuda()`` itself.  ``get_loader`` (dataset plumbing) is out of scope.
"""
import os
import pickle
import traceback
import warnings

from ..modelutils import modelio
from ..networks.handnet import HandNet

# options older checkpoints lack and the value the reference substitutes (reload.py:42-73)
_OPT_DEFAULTS = (("absolute_lambda", 0), ("atlas_predict_trans", False), ("atlas_lambda_laplacian", False),
                 ("atlas_residual", False), ("mano_lambda_joints3d", False), ("mano_lambda_joints2d", False),
                 ("mano_adapt_skeleton", False), ("contact_lambda", 0), ("collision_lambda", 0), ("mano_use_pca", True),
                 ("atlas_separate_encoder", False), ("atlas_final_lambda", 0), ("atlas_predict_scale", False))


# HandNet keyword <- key of the experiment's opt.pkl (reload.py:76-104); most share the name, three do not
_HANDNET_FROM_OPTS = tuple((k, k) for k in (
    "absolute_lambda", "atlas_lambda_regul_edges", "atlas_lambda_laplacian", "atlas_predict_trans",
    "atlas_predict_scale", "atlas_residual", "atlas_lambda", "atlas_final_lambda", "atlas_separate_encoder",
    "contact_lambda", "collision_lambda", "mano_adapt_skeleton", "mano_use_pca", "mano_lambda_verts",
    "mano_lambda_joints3d", "mano_lambda_joints2d")) + (("mano_center_idx", "center_idx"),
                                                        ("mano_neurons", "hidden_neurons"))
# what the reference hard-codes for every reloaded model
_FIXED = dict(resnet_version=18, atlas_mesh=True, atlas_points_nb=642, mano_comps=30)


def save_obj(filename, verticies, faces):
    """Wavefront .obj writer (reload.py:16-22); faces are 0-based on input, 1-based in the file."""
    with open(filename, "w") as fp:
        for v in verticies:
            fp.write("v %f %f %f\n" % (v[0], v[1], v[2]))
        for f in faces + 1:
            fp.write("f %d %d %d\n" % (f[0], f[1], f[2]))


def get_opts(resume_checkpoint):
    """``opt.pkl`` next to the checkpoint (reload.py:25-32); accepts the folder or the ``.tar`` file inside it."""
    if resume_checkpoint.endswith("tar"):
        resume_checkpoint = os.path.dirname(resume_checkpoint)
    with open(os.path.join(resume_checkpoint, "opt.pkl"), "rb") as p_f:
        return pickle.load(p_f)


def reload_model(model_path, checkpoint_opts, mano_root="misc/mano", ico_divisions=3, no_beta=False):
    checkpoint_opts = dict(checkpoint_opts)
    for key, default in _OPT_DEFAULTS:
        checkpoint_opts.setdefault(key, default)
    mano_use_shape = False if no_beta else checkpoint_opts["use_shape"]
    kwargs = {handnet_kw: checkpoint_opts[opt_key] for handnet_kw, opt_key in _HANDNET_FROM_OPTS}
    kwargs.update(_FIXED, atlas_ico_divisions=ico_divisions, mano_root=mano_root, mano_use_shape=mano_use_shape)
    model = HandNet(**kwargs)
    model.eval()
    try:
        modelio.load_checkpoint(model, resume_path=model_path, strict=True)
    except RuntimeError:
        traceback.print_exc()
        warnings.warn("Couldn' load model in strict mode, trying without strict")
        modelio.load_checkpoint(model, resume_path=model_path, strict=False)
    return model
