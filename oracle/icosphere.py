"""Unit icosphere (oracle side; test infrastructure only).

Restates what ``trimesh.creation.icosphere(subdivisions=k)`` produces for the
reference's AtlasNet test mesh (/root/reference/mano_train/networks/branches/atlasbranch.py:64-70):
a regular icosahedron whose every triangle is split in four through its edge
midpoints ``k`` times, every vertex re-projected on the unit sphere.
Vertex/face ORDER is not part of the contract (the AtlasNet decoder is
point-wise; only grid<->faces consistency matters): 12->42->162->642->2562
vertices, 20->80->320->1280->5120 faces (F = 2V - 4).
"""
import numpy as np

_T = (1.0 + 5.0 ** 0.5) / 2.0

_SEED_VERTS = np.array(
    [
        (-1, _T, 0), (1, _T, 0), (-1, -_T, 0), (1, -_T, 0),
        (0, -1, _T), (0, 1, _T), (0, -1, -_T), (0, 1, -_T),
        (_T, 0, -1), (_T, 0, 1), (-_T, 0, -1), (-_T, 0, 1),
    ],
    dtype=np.float64,
)

_SEED_FACES = np.array(
    [
        (0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11),
        (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
        (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
        (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1),
    ],
    dtype=np.int64,
)


def icosphere(subdivisions=3):
    """Return (verts (V,3) float64 on the unit sphere, faces (F,3) int64)."""
    verts = _SEED_VERTS / np.linalg.norm(_SEED_VERTS, axis=1, keepdims=True)
    faces = _SEED_FACES.copy()
    for _ in range(int(subdivisions)):
        # unique undirected edges -> one new midpoint vertex each
        edges = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
        edges_sorted = np.sort(edges, axis=1)
        uniq, inverse = np.unique(edges_sorted, axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
        mids = verts[uniq].mean(axis=1)
        mid_idx = inverse.reshape(3, -1).T + len(verts)  # (F,3): ab, bc, ca
        a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
        ab, bc, ca = mid_idx[:, 0], mid_idx[:, 1], mid_idx[:, 2]
        faces = np.concatenate(
            [
                np.stack([a, ab, ca], 1),
                np.stack([ab, b, bc], 1),
                np.stack([ca, bc, c], 1),
                np.stack([ab, bc, ca], 1),
            ],
            axis=0,
        )
        verts = np.concatenate([verts, mids], axis=0)
        verts = verts / np.linalg.norm(verts, axis=1, keepdims=True)
    return verts, faces
