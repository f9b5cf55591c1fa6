"""Device-side scalar-loss stage: autograd nodes over csrc/loss_head.cu.

The reference assembles its loss with a dozen ATen reductions and Python-float multiplications per step
(ManoLoss.compute_loss manobranch.py:251-324, AtlasLoss.compute_loss atlasbranch.py:199-287, HandNet.forward
handnet.py:279-283,363-383).  Here the lambdas live in a small DEVICE vector (``LossWeights``), the mean-squared-error
terms of one branch are one launch (``sq_terms``), the GT object statistics one launch (``object_targets``) and the
weighted total one launch (``combine``).  Because the kernels read the lambdas from device memory when they run, a
captured CUDA graph follows ``HandNet.decay_regul`` and any other change of a weight without re-capture.
"""
import ctypes

import torch

from ._lib import call, ptr, stream_ptr

MAX_SQ_TERMS = 8
MAX_COMBINE_TERMS = 12
GROUPS = 4


def _ptrs(tensors):
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _ints(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def _floats(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def _chk(t, what):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("{}: expected a CUDA float32 tensor (obman_train_b200 has no CPU path)".format(what))
    return t if t.is_contiguous() else t.contiguous()


class LossWeights(object):
    """Named loss weights mirrored in a device vector.  ``w[name] = value`` updates the host value and, once the
    vector exists on a device, the device slot (a fill kernel: safe while replays of a captured step are in flight,
    and seen by the next replay).  ``None`` and ``0`` weights are stored as 0."""

    def __init__(self, names):
        self.names = list(names)
        self.slot = {n: i for i, n in enumerate(self.names)}
        self.host = [0.0] * len(self.names)
        self._dev = {}

    def __setitem__(self, name, value):
        v = 0.0 if value is None else float(value)
        i = self.slot[name]
        if self.host[i] != v:
            self.host[i] = v
            for vec in self._dev.values():
                vec[i:i + 1].fill_(v)

    def __getitem__(self, name):
        return self.host[self.slot[name]]

    def device(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = torch.tensor(self.host, dtype=torch.float32, device=device)
        return self._dev[key]


class _Workspace(object):
    """Per call site scratch of sq_terms (partial sums + the ticket counter, which must be zero before the first
    launch and is left zero by the kernel).  One per call site: a call site runs once per step on one stream."""

    def __init__(self):
        self._bufs = {}

    def get(self, device):
        key = str(device)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(32 * MAX_SQ_TERMS, device=device, dtype=torch.float32),
                               torch.zeros(1, device=device, dtype=torch.int32))
        return self._bufs[key]


class _SqTermsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, weights, workspace, *tensors):
        # spec: list of (has_b, col0, col1, slot); tensors: a_0, [b_0], a_1, [b_1], ...
        a, b = [], []
        it = iter(tensors)
        for has_b, _, _, _ in spec:
            a.append(_chk(next(it), "sq_terms input"))
            b.append(_chk(next(it).detach(), "sq_terms target") if has_b else None)
        rows, width = [], []
        for t, tb in zip(a, b):
            w = t.shape[-1] if t.dim() > 1 else 1
            width.append(w)
            rows.append(t.numel() // w)
            if tb is not None and tb.shape != t.shape:
                raise RuntimeError("sq_terms: target shape {} differs from prediction shape {}".format(
                    tuple(tb.shape), tuple(t.shape)))
        col0 = [s[1] for s in spec]
        col1 = [width[k] if s[2] is None else s[2] for k, s in enumerate(spec)]
        slot = [s[3] for s in spec]
        dev = a[0].device
        partial, ticket = workspace.get(dev)
        terms = torch.empty(len(spec), device=dev)
        wsum = torch.empty(1, device=dev)
        call("obman_sq_terms_fwd", _ptrs(a), _ptrs(b), _ints(rows), _ints(width), _ints(col0), _ints(col1),
             _ints(slot), len(spec), ptr(weights), ptr(partial), ptr(ticket), ptr(terms), ptr(wsum), stream_ptr())
        ctx.geom = (rows, width, col0, col1, slot)
        ctx.has_b = [s[0] for s in spec]
        ctx.weights = weights
        ctx.save_for_backward(*[t for pair in zip(a, b) for t in pair if t is not None])
        ctx.mark_non_differentiable(terms)
        return wsum, terms

    @staticmethod
    def backward(ctx, gwsum, _gterms):
        rows, width, col0, col1, slot = ctx.geom
        saved = iter(ctx.saved_tensors)
        a, b = [], []
        for has_b in ctx.has_b:
            a.append(next(saved))
            b.append(next(saved) if has_b else None)
        gwsum = _chk(gwsum, "sq_terms gradient")
        grads, ga = [], []
        pos = 3
        for k, has_b in enumerate(ctx.has_b):
            need = ctx.needs_input_grad[pos]
            ga.append(torch.empty_like(a[k]) if need else None)
            grads.append(ga[-1])
            pos += 1
            if has_b:
                grads.append(None)
                pos += 1
        if any(g is not None for g in ga):
            call("obman_sq_terms_bwd", _ptrs(a), _ptrs(b), _ptrs(ga), _ints(rows), _ints(width), _ints(col0),
                 _ints(col1), _ints(slot), len(a), ptr(ctx.weights), ptr(gwsum), stream_ptr())
        return (None, None, None) + tuple(grads)


def sq_terms(terms, weights_dev, workspace):
    """``terms``: list of (prediction, target or None, (col0, col1) or None, weight slot).  Each term is
    torch.nn.functional.mse_loss(prediction[..., col0:col1], target[..., col0:col1]) (target None = zeros).
    Returns (weighted sum (1,), values (K,)); only the weighted sum carries gradient."""
    if not 1 <= len(terms) <= MAX_SQ_TERMS:
        raise RuntimeError("sq_terms: between 1 and {} terms".format(MAX_SQ_TERMS))
    spec, flat = [], []
    for pred, target, cols, slot in terms:
        c0, c1 = (0, None) if cols is None else cols
        spec.append((target is not None, c0, c1, slot))
        flat.append(pred)
        if target is not None:
            flat.append(target)
    return _SqTermsFn.apply(spec, weights_dev, workspace, *flat)


class _CombineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, weights, *tensors):
        # spec: list of (n_tensors (1 | 2), scale, slot, group)
        p, q, lens = [], [], []
        it = iter(tensors)
        for n_t, _, _, _ in spec:
            t = _chk(next(it), "combine term")
            p.append(t)
            lens.append(t.numel())
            if n_t == 2:
                t2 = _chk(next(it), "combine term")
                if t2.numel() != t.numel():
                    raise RuntimeError("combine: the two vectors of a term must have the same length")
                q.append(t2)
            else:
                q.append(None)
        scale = [s[1] for s in spec]
        slot = [s[2] for s in spec]
        group = [s[3] for s in spec]
        dev = p[0].device
        total = torch.empty(1, device=dev)
        aux = torch.empty(GROUPS + len(spec), device=dev)
        call("obman_loss_combine_fwd", _ptrs(p), _ptrs(q), _ints(lens), _floats(scale), _ints(slot), _ints(group),
             len(spec), ptr(weights), ptr(total), ptr(aux), stream_ptr())
        ctx.meta = (scale, slot, [tuple(t.shape) for t in tensors], [s[0] for s in spec])
        ctx.weights = weights
        ctx.mark_non_differentiable(aux)
        return total, aux

    @staticmethod
    def backward(ctx, gtotal, _gaux):
        scale, slot, shapes, counts = ctx.meta
        gtotal = _chk(gtotal, "combine gradient")
        gterm = torch.empty(len(scale), device=gtotal.device)
        call("obman_loss_combine_bwd", _ints(slot), _floats(scale), len(scale), ptr(ctx.weights), ptr(gtotal),
             ptr(gterm), stream_ptr())
        grads, pos = [], 0
        for k, n_t in enumerate(counts):
            for _ in range(n_t):
                # one scalar per term, expanded (a stride-0 view, no kernel) to the shape of its vector
                g = gterm[k:k + 1]
                n = 1
                for d in shapes[pos]:
                    n *= d
                grads.append(g.view(shapes[pos]) if n == 1 else g.expand(shapes[pos]))
                pos += 1
        return (None, None) + tuple(grads)


def combine(terms, weights_dev):
    """``terms``: list of (tensor or (tensor, tensor), scale, weight slot, group).  value_k = scale * sum(tensor(s));
    returns (total (1,) = sum_k w[slot_k] * value_k, group sums (4,), values (K,)); gradient flows through total."""
    if not 1 <= len(terms) <= MAX_COMBINE_TERMS:
        raise RuntimeError("combine: between 1 and {} terms".format(MAX_COMBINE_TERMS))
    spec, flat = [], []
    for t, scale, slot, group in terms:
        ts = t if isinstance(t, (tuple, list)) else (t,)
        spec.append((len(ts), float(scale), int(slot), int(group)))
        flat.extend(ts)
    total, aux = _CombineFn.apply(spec, weights_dev, *flat)
    return total, aux[:GROUPS], aux[GROUPS:]


def object_targets(gt, want_centred=True):
    """gt (B,M,3) GT object points -> (centroid (B,3), scale (B,1) = max_i |gt_i - centroid|, centred (B,M,3));
    AtlasLoss.compute_loss's targets (atlasbranch.py:211-227).  No gradient (targets)."""
    gt = _chk(gt.detach(), "objpoints3d")
    B, M, _ = gt.shape
    centroid = torch.empty((B, 3), device=gt.device)
    scale = torch.empty((B, 1), device=gt.device)
    centred = torch.empty_like(gt) if want_centred else None
    call("obman_object_targets", ptr(gt), B, M, ptr(centroid), ptr(scale), ptr(centred), stream_ptr())
    return centroid, scale, centred


class _AffinePointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, scales, trans):
        verts = _chk(verts, "verts")
        B, N, _ = verts.shape
        s = None if scales is None else _chk(scales, "scales").reshape(B)
        t = None if trans is None else _chk(trans, "translations")
        out = torch.empty_like(verts)
        call("obman_affine_points_fwd", ptr(verts), ptr(s), ptr(t), B, N, ptr(out), stream_ptr())
        ctx.save_for_backward(verts, s if s is not None else verts.new_empty(0))
        ctx.has = (scales is not None, trans is not None, None if scales is None else tuple(scales.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        verts, s = ctx.saved_tensors
        has_s, has_t, s_shape = ctx.has
        B, N, _ = verts.shape
        g = _chk(g, "gradient")
        gv = torch.empty_like(verts) if ctx.needs_input_grad[0] else None
        gs = torch.empty(B, device=g.device) if (has_s and ctx.needs_input_grad[1]) else None
        gt = torch.empty((B, 3), device=g.device) if (has_t and ctx.needs_input_grad[2]) else None
        call("obman_affine_points_bwd", ptr(g), ptr(verts), ptr(s) if has_s else None, B, N, ptr(gv), ptr(gs), ptr(gt),
             stream_ptr())
        return gv, None if gs is None else gs.view(s_shape), gt


def affine_points(verts, scales=None, translations=None):
    """scales (B,1) * verts (B,N,3) + translations (B,3) per sample (atlasbranch.py:133-138) as one kernel per direction."""
    return _AffinePointsFn.apply(verts, scales, translations)
