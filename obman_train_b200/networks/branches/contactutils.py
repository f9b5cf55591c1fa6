"""Mirror of mano_train/networks/branches/contactutils.py (hot-path part): inside/outside test.

``batch_mesh_contains_points`` keeps the reference's signature
(/root/reference/mano_train/networks/branches/contactutils.py:62-66) but runs the fused ray-parity
kernel (csrc/contact.cu) instead of materialising (B, P*F, 3) temporaries.  ``load_contacts`` is the
fixture reader the reference keeps in handobjectdatasets/contactutils.py:8-45.
"""
import torch

from ... import functional as F_b200
from ...assets import load_contacts  # noqa: F401  (re-exported, same name as the reference)

RAY_DIRECTION = (0.4395064455, 0.617598629942, 0.652231566745)


def batch_mesh_contains_points(ray_origins, obj_triangles, direction=None):
    """ray_origins (B,P,3), obj_triangles (B,F,3,3) -> bool (B,P), True = exterior.

    The triangle-soup form of the reference API is kept; internally the soup is viewed as 3F vertices
    with trivial faces.  ``direction`` must be the reference's fixed default (it is compiled into the
    kernel); anything else raises.
    """
    if direction is not None:
        d = [float(x) for x in torch.as_tensor(direction).reshape(-1).tolist()]
        if max(abs(a - b) for a, b in zip(d, RAY_DIRECTION)) > 1e-6:
            raise ValueError("only the reference's fixed ray direction is supported")
    B, F = obj_triangles.shape[0], obj_triangles.shape[1]
    verts = obj_triangles.reshape(B, F * 3, 3)
    faces = torch.arange(F * 3, device=verts.device, dtype=torch.int32).view(F, 3)
    exterior, _ = F_b200.mesh_exterior(ray_origins, verts, faces)
    return exterior
