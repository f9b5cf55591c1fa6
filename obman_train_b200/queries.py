"""Sample-dict keys consumed by HandNet.forward: same enum members and values as
/root/reference/handobjectdatasets/queries.py:4-46.  When the reference package is
importable its own enums are re-exported, so that dicts built by the reference's data loaders index
identically; otherwise equivalent enums are defined here."""
try:  # pragma: no cover - only where the reference tree is on sys.path
    from handobjectdatasets.queries import BaseQueries, TransQueries  # noqa: F401
except Exception:  # noqa: BLE001
    from enum import Enum

    class BaseQueries(Enum):
        camintrs = "camintrs"
        depth = "depth"
        hand_poses = "hand_poses"
        hand_pcas = "hand_pcas"
        images = "images"
        joints2d = "joints2d"
        joints3d = "joints3d"
        meta = "meta"
        objpoints2d = "objpoints2d"
        objpoints3d = "objpoints3d"
        objverts3d = "objverts3d"
        objfaces = "objfaces"
        verts3d = "verts3d"
        sides = "sides"
        segms = "segms"
        manoidxs = "manoidxs"

    class TransQueries(Enum):
        camintrs = "camintrs"
        depth = "depth"
        images = "images"
        joints2d = "joints2d "
        joints3d = "joints3d"
        objfaces = "objfaces"
        objpoints2d = "objpoints2d"
        objpoints3d = "objpoints3d"
        objverts3d = "objverts3d"
        segms = "segms"
        verts3d = "verts3d"
        center3d = "center3d"
        affinetrans = "affinetrans"
        rotmat = "rotmat"
        sdf = "sdf"
        sdf_points = "sdf_points"
        mapvals = "mapvals"
        mapidxs = "mapidxs"
