"""Synthetic MANO tables with the real template geometry.

The licensed ``MANO_{LEFT,RIGHT}.pkl`` files are not redistributable and are absent here
(SURVEY.md "three facts").  Benchmarks and tests therefore run on tables that have the real
template vertices/faces (from the packaged contact-zones fixture) and synthetic but structurally
valid blend-shape / regressor / skinning data: same shapes, dtypes and buffer names as the
buffers ``manopth.manolayer.ManoLayer`` registers.
"""
import numpy as np

from ..assets import template_mesh


def synthetic_mano_tables(side="right", seed=0):
    """Return a dict of numpy arrays keyed like the ManoLayer pickle:
    v_template (778,3), shapedirs (778,3,10), posedirs (778,3,135), J_regressor (16,778),
    weights (778,16), hands_mean (45,), hands_components (45,45), f (1538,3), betas (10,)."""
    rng = np.random.RandomState(seed + (0 if side == "right" else 1000))
    verts, faces = template_mesh()
    verts = verts.copy()
    faces = faces.copy()
    if side == "left":
        verts[:, 0] *= -1
        faces = faces[:, ::-1].copy()
    # 16 joint anchors by farthest-point sampling of the template (deterministic)
    anchors = [0]
    d = np.linalg.norm(verts - verts[0], axis=1)
    for _ in range(15):
        nxt = int(d.argmax())
        anchors.append(nxt)
        d = np.minimum(d, np.linalg.norm(verts - verts[nxt], axis=1))
    centres = verts[anchors]
    d2 = ((verts[None] - centres[:, None]) ** 2).sum(-1)  # (16,778)
    jreg = np.exp(-d2 / (2 * 0.01 ** 2))
    jreg /= jreg.sum(1, keepdims=True)
    joints = jreg @ verts
    d2w = ((verts[:, None] - joints[None]) ** 2).sum(-1)  # (778,16)
    w = np.exp(-d2w / (2 * 0.02 ** 2)) + 1e-6
    w /= w.sum(1, keepdims=True)
    q, _ = np.linalg.qr(rng.randn(45, 45))
    return {
        "v_template": verts,
        "shapedirs": rng.randn(778, 3, 10) * 0.002,
        "posedirs": rng.randn(778, 3, 135) * 0.0005,
        "J_regressor": jreg,
        "weights": w,
        "hands_mean": rng.randn(45) * 0.2,
        "hands_components": q,
        "f": faces.astype(np.int64),
        "betas": np.zeros(10),
    }
