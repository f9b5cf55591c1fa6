"""Shared synthetic inputs / configs for the parity tests (seeded, mm scale)."""
import torch

from obman_train_b200.assets import load_contacts

FULL_CFG = dict(resnet_version=18, mano_root="synthetic", mano_comps=30, mano_use_shape=True,
                mano_neurons=[1024, 256], mano_center_idx=0, mano_lambda_verts=0.167,
                mano_lambda_joints3d=0.167, mano_lambda_shape=0.167, mano_lambda_pose_reg=0.167,
                atlas_lambda=0.167, atlas_final_lambda=0.167, atlas_predict_trans=True,
                atlas_predict_scale=True, atlas_trans_weight=0.167, atlas_scale_weight=0.167,
                atlas_separate_encoder=True, atlas_ico_divisions=2, atlas_lambda_regul_edges=0.1,
                contact_lambda=1, collision_lambda=1, contact_zones="zones", contact_mode="dist_tanh",
                collision_mode="dist_tanh", contact_thresh=10, collision_thresh=20)


def randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
            m.running_mean.data = torch.randn(m.running_mean.shape, generator=g) * 0.1
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)


def make_sample(B, H, seed, n_gt=600, sides=None):
    """Sample dict keyed by the plain strings of the query enums."""
    g = torch.Generator().manual_seed(seed)
    verts, _ = load_contacts()
    hand = torch.tensor(verts * 1000, dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
    return {
        "images": torch.rand(B, 3, H, H, generator=g) - 0.5,
        "sides": sides if sides is not None else ["right" if i % 2 == 0 else "left" for i in range(B)],
        "root": "wrist",
        "joints3d": torch.randn(B, 21, 3, generator=g) * 40,
        "verts3d": hand + torch.randn(B, 778, 3, generator=g) * 5,
        "objpoints3d": torch.randn(B, n_gt, 3, generator=g) * 40 + 30,
    }


def enum_sample(sample, device="cuda"):
    """Re-key a plain sample with the product's query enums and move tensors to ``device``."""
    from obman_train_b200.queries import TransQueries, BaseQueries
    out = {TransQueries.images: sample["images"], BaseQueries.sides: sample["sides"], "root": sample["root"]}
    for k, q in (("joints3d", TransQueries.joints3d), ("verts3d", TransQueries.verts3d),
                 ("objpoints3d", TransQueries.objpoints3d)):
        if k in sample:
            out[q] = sample[k]
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in out.items()}
