"""Cotangent-Laplacian regulariser on the predicted object mesh - drop-in for
/root/reference/mano_train/networks/branches/laplacianloss.py (``LaplacianLoss(faces, vertices)(verts)``).

The reference builds the block-diagonal (B*N x B*N) scipy matrix of the UNIT ICOSPHERE's cotangent Laplacian on the
first call (:100-131), multiplies on the host in numpy inside a legacy ``autograd.Function`` (:70-150) - which no longer
runs on torch >= 1.5 - and returns ``mean_i ||(L V)_i||`` (:36-41).  L depends only on the fixed sphere, so here it is
built once in float64 on the host as an ELL table (<= 7 entries per row) and the product, the row norms, the mean
and the backward ``L^T g = L g`` run in two small CUDA kernels (csrc/mesh_regul.cu).
"""
import numpy as np
import torch

from ... import functional as Fb


def cotangent_weights(verts, faces):
    """Per face (F,3): cot of the angle opposite to edges (v2,v3), (v3,v1), (v1,v2), divided by 2... following the
    reference's formula (:153-185): C = [l2^2+l3^2-l1^2, l1^2+l3^2-l2^2, l1^2+l2^2-l3^2] / (4 * 2*Area) with Heron's area."""
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    v1, v2, v3 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    l1 = np.linalg.norm(v2 - v3, axis=1)
    l2 = np.linalg.norm(v3 - v1, axis=1)
    l3 = np.linalg.norm(v1 - v2, axis=1)
    sp = (l1 + l2 + l3) * 0.5
    area2 = 2.0 * np.sqrt(sp * (sp - l1) * (sp - l2) * (sp - l3))
    cot = np.stack([l2 ** 2 + l3 ** 2 - l1 ** 2, l1 ** 2 + l3 ** 2 - l2 ** 2, l1 ** 2 + l2 ** 2 - l3 ** 2], 1)
    return cot / area2[:, None] / 4.0


def laplacian_ell(verts, faces):
    """ELL form of L = (C + C^T) - diag(rowsum) (:117-129): returns nbr (N,K) int32, w (N,K) float32 holding the
    off-diagonal entries; the diagonal equals -sum_k w[i,k] and is applied implicitly by the kernel."""
    f = np.asarray(faces, dtype=np.int64)
    n = int(np.asarray(verts).shape[0])
    cot = cotangent_weights(verts, f)
    rows = f[:, [1, 2, 0]].reshape(-1)
    cols = f[:, [2, 0, 1]].reshape(-1)
    vals = cot.reshape(-1)
    entries = {}
    for r, c, val in zip(rows.tolist(), cols.tolist(), vals.tolist()):
        entries[(r, c)] = entries.get((r, c), 0.0) + val
        entries[(c, r)] = entries.get((c, r), 0.0) + val
    per_row = [[] for _ in range(n)]
    for (r, c), val in sorted(entries.items()):
        per_row[r].append((c, val))
    k = max(1, max(len(x) for x in per_row))
    nbr = np.tile(np.arange(n, dtype=np.int32)[:, None], (1, k))
    w = np.zeros((n, k), dtype=np.float32)
    for i, lst in enumerate(per_row):
        for q, (c, val) in enumerate(lst):
            nbr[i, q] = c
            w[i, q] = val
    return nbr, w


def vertex_face_table(n_verts, faces):
    """(N,Kf) int32 table of the faces incident to each vertex, -1 padded (used by the edge-loss backward)."""
    f = np.asarray(faces, dtype=np.int64)
    per = [[] for _ in range(n_verts)]
    for fi, tri in enumerate(f.tolist()):
        for v in set(tri):
            per[v].append(fi)
    k = max(1, max(len(x) for x in per))
    out = -np.ones((n_verts, k), dtype=np.int32)
    for i, lst in enumerate(per):
        out[i, :len(lst)] = lst
    return out


class LaplacianLoss(object):
    """Encourages minimal mean curvature shapes (same constructor / call as the reference's class)."""

    def __init__(self, faces, vertices):
        verts = vertices.detach().cpu().numpy() if torch.is_tensor(vertices) else np.asarray(vertices)
        nbr, w = laplacian_ell(verts, np.asarray(faces))
        self._nbr_host, self._w_host = torch.from_numpy(nbr), torch.from_numpy(w)
        self._dev = {}
        self.Lx = None

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (self._nbr_host.to(device), self._w_host.to(device))
        return self._dev[key]

    def __call__(self, verts):
        nbr, w = self._tables(verts.device)
        loss, self.Lx = Fb.laplacian_loss(verts, nbr, w)
        return loss.squeeze(0)  # the reference returns a 0-dim tensor (torch.norm(...).mean())
