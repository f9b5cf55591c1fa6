"""Unit icosphere used as the AtlasNet test mesh (the reference takes it from
``trimesh.creation.icosphere``, /root/reference/mano_train/networks/branches/atlasbranch.py:64-70).

12 -> 42 -> 162 -> 642 -> 2562 vertices, 20 -> 80 -> 320 -> 1280 -> 5120 faces; vertex order is not part of
the contract (the decoder is point-wise), only grid <-> faces consistency is."""
import numpy as np


def icosphere(subdivisions=3):
    """Return (verts (V,3) float64 on the unit sphere, faces (F,3) int64)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
             (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.asarray(v, dtype=np.float64) / np.sqrt(1 + t * t) for v in verts]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
             (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
             (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(int(subdivisions)):
        cache = {}

        def midpoint(i, j):
            key = (i, j) if i < j else (j, i)
            if key not in cache:
                m = verts[i] + verts[j]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        new_faces = []
        for a, b, c in faces:
            ab, bc, ca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
            new_faces.extend([(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)])
        faces = new_faces
    return np.stack(verts), np.asarray(faces, dtype=np.int64)
