"""Linear layers and the AtlasNet point decoder as autograd nodes over the tcgen05 GEMM kernels.

``linear``        nn.Linear (+ReLU) of ManoBranch / AtlasBranch heads (manobranch.py:57-67,82-85,124-147;
                  atlasbranch.py:44-61).
``point_decoder`` PointGenCon (atlasutils.py:42-75) on the AtlasBranch input cat(grid, feature)
                  (atlasbranch.py:117-131), BatchNorm1d in eval mode, with conv1 split algebraically:
                  conv1(x)[b,:,n] = W[:, :3] grid[n] + (W[:, 3:] feat_b + bias), which removes the
                  (B,515,N) concatenated tensor and 61 % of the decoder MACs.
"""
import torch

from . import dense, streams
from ._lib import call, ptr, stream_ptr

BN_EPS = 1e-5


def _r32(n):
    return (n + 31) // 32 * 32


def _zeros(*shape):
    return torch.zeros(shape, device="cuda", dtype=torch.float32)


def _empty(*shape):
    return torch.empty(shape, device="cuda", dtype=torch.float32)


def _pad_cols(t, width):
    """Copy (M,N) into a zero-padded (M,width) buffer (no copy when already that wide and contiguous)."""
    if t.shape[1] == width and t.is_contiguous():
        return t
    out = _zeros(t.shape[0], width)
    out[:, :t.shape[1]] = t
    return out


def pad_scale_mask(src, width, alpha=1.0, mask=None, cols=None):
    """(rows, width) buffer = alpha * src[:, :cols] where mask > 0 (mask None: everywhere), zero-padded - one launch."""
    rows = src.shape[0]
    C = src.shape[1] if cols is None else cols
    out = _empty(rows, width)
    call("obman_pad_scale_mask", ptr(src), src.stride(0), ptr(mask), 0 if mask is None else mask.stride(0), rows, C,
         float(alpha), ptr(out), width, stream_ptr())
    return out


def packed_transpose(w, n=None, k=None):
    """(K, r32(N)) packed bf16 hi|lo transpose of w (N,K): the B operand of a Linear layer's data gradient."""
    N = w.shape[0] if n is None else n
    K = w.shape[1] if k is None else k
    out = _empty(K, _r32(N))
    call("obman_pack_bf16_t", ptr(w), w.stride(0), N, K, ptr(out), out.stride(0), stream_ptr())
    return out


def colsum(t, C=None):
    C = t.shape[1] if C is None else C
    out = _empty(C)
    call("obman_colsum", ptr(t), t.shape[0], C, t.stride(0), ptr(out), stream_ptr())
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x = x.contiguous()
        w = weight.detach().contiguous()
        if x.shape[1] % 4:
            raise RuntimeError("linear: in_features must be a multiple of 4 for the TMA path")
        pf = dense.PASSES[dense.get_precision()["fwd"]]
        y = dense.gemm(x, w, bias=None if bias is None else bias.detach().contiguous(), relu=relu, passes=pf)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x, w, y = ctx.saved_tensors
        pb = dense.PASSES[dense.get_precision()["bwd"]]
        pw = dense.PASSES[dense.get_precision()["wgrad"]]
        N, K = w.shape
        # ReLU mask, contiguity and the zero padding of the columns to a multiple of 32 in one launch
        if not (g.is_cuda and g.dtype == torch.float32):
            raise RuntimeError("linear backward: expected a CUDA float32 gradient")
        if g.stride(-1) != 1:
            g = g.contiguous()
        gp = pad_scale_mask(g, _r32(N), mask=y if ctx.relu else None)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if pb == dense.BF16X3:
                gx = dense.gemm(gp, packed_transpose(w), passes=pb, n=K, k=N, packed=True)
            else:
                wt = _zeros(K, _r32(N))
                wt[:, :N] = w.t()
                gx = dense.gemm(gp, wt, passes=pb, n=K, k=N)
        if ctx.needs_input_grad[1]:
            xk = x if x.shape[1] == _r32(K) else pad_scale_mask(x, _r32(K))
            gw = dense.wgrad_matrix(gp, xk, passes=pw)[:N, :K].contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = colsum(gp, N)
        return gx, gw, gb, None


def linear(x, weight, bias=None, relu=False):
    return _LinearFn.apply(x, weight, bias, relu)


class _Layer(object):
    """Folded 1x1-conv (+BN eval) layer of the point decoder: y = relu?(x Wf^T + shift)."""

    def __init__(self, w, cbias, bn, packed=False, plain=False):
        self.w = w.detach().reshape(w.shape[0], w.shape[1]).contiguous()
        self.cbias = cbias.detach().contiguous()
        self.O, self.I = self.w.shape
        self.bn = None if bn is None else [t.detach().contiguous() for t in bn]
        gamma, beta, mean, var = self.bn if self.bn is not None else (None, None, None, None)
        self.packed = packed
        self.shift, self.scale, self.rstd = _empty(self.O), _empty(self.O), _empty(self.O)
        if plain:
            # BatchNorm-folded weights in plain fp32, no data-gradient layout (the layer is evaluated outside the GEMMs)
            self.ld = (self.I + 3) // 4 * 4
            self.wf = _empty(self.O, self.ld)
            self.wf_lo = self.wft = self.wft_lo = None
            call("obman_fold_conv", ptr(self.w), ptr(self.cbias), ptr(gamma), ptr(beta), ptr(mean), ptr(var),
                 BN_EPS, self.O, self.I, 1, 1, self.ld, 0, 0, ptr(self.wf), None, None, None,
                 ptr(self.shift), ptr(self.scale), ptr(self.rstd), stream_ptr())
            self.fw = self.bw = {}
            return
        if packed:
            # folded weights in the packed bf16 hi|lo layout (3xBF16): wf (O, r32(I)), wft (I, r32(O))
            self.ld = _r32(self.I)
            self.wf, self.wf_lo = _empty(self.O, self.ld), None
            self.wft, self.wft_lo = _zeros(self.I, _r32(self.O)), None
            call("obman_fold_conv", ptr(self.w), ptr(self.cbias), ptr(gamma), ptr(beta), ptr(mean), ptr(var),
                 BN_EPS, self.O, self.I, 1, 1, self.ld, 0, 1, ptr(self.wf), None, ptr(self.wft), None,
                 ptr(self.shift), ptr(self.scale), ptr(self.rstd), stream_ptr())
            self.fw = {"packed": True}
            self.bw = {"packed": True}
            return
        self.ld = (self.I + 3) // 4 * 4
        # folded weights, pre-split into tf32 hi / lo parts: wf + wf_lo = scale * w (fprop), wft + wft_lo (dgrad)
        self.wf, self.wf_lo = _empty(self.O, self.ld), _empty(self.O, self.ld)
        self.wft, self.wft_lo = _zeros(self.I, _r32(self.O)), _zeros(self.I, _r32(self.O))   # (I, O padded)
        wft_tmp, wft_tmp_lo = _empty(self.I, self.O), _empty(self.I, self.O)
        call("obman_fold_conv", ptr(self.w), ptr(self.cbias), ptr(gamma), ptr(beta), ptr(mean), ptr(var),
             BN_EPS, self.O, self.I, 1, 1, self.ld, 0, 0, ptr(self.wf), ptr(self.wf_lo), ptr(wft_tmp),
             ptr(wft_tmp_lo), ptr(self.shift), ptr(self.scale), ptr(self.rstd), stream_ptr())
        self.wft[:, :self.O] = wft_tmp
        self.wft_lo[:, :self.O] = wft_tmp_lo
        self.fw = {"w_lo": self.wf_lo}
        self.bw = {"w_lo": self.wft_lo}

    def finish(self, dwraw, gsum):
        """dwraw (>=O rows, row stride ld_raw, first I columns valid) -> (gw, gcbias, ggamma, gbeta)."""
        gw = _empty(self.O, self.I)
        gcb = _empty(self.O)
        gg = gbt = None
        mean = None
        if self.bn is not None:
            gg, gbt = _empty(self.O), _empty(self.O)
            mean = self.bn[2]
        call("obman_bn_wgrad_finish", ptr(dwraw), dwraw.stride(0), ptr(self.w), ptr(self.cbias),
             ptr(self.scale), ptr(self.rstd), ptr(mean), ptr(gsum), self.O, self.I, 1, 1, dwraw.stride(0), 0,
             ptr(gw), ptr(gg), ptr(gbt), ptr(gcb), stream_ptr())
        return gw, gcb, gg, gbt


class _PointDecoderFn(torch.autograd.Function):
    """verts (B,N,3) = out_factor * PointGenCon(cat(grid, feat repeated)), BatchNorm in eval mode.

    args: feat (B,F), grid (N,3) shared or (B,N,3) per sample, out_factor, then
    conv1.w, conv1.b, conv2.w, conv2.b, conv3.w, conv3.b, conv4.w, conv4.b,
    bn1 (g,b,m,v), bn2 (g,b,m,v), bn3 (g,b,m,v)."""

    @staticmethod
    def forward(ctx, feat, grid, out_factor, *p):
        pf = dense.PASSES[dense.get_precision()["fwd"]]
        st = stream_ptr()
        feat = feat.contiguous()
        grid = grid.detach().contiguous()
        B, Fdim = feat.shape
        N = grid.shape[-2]
        per_sample = grid.dim() == 3
        pk = pf == dense.BF16X3
        l1 = _Layer(p[0], p[1], p[8:12], plain=True)   # conv1 is split algebraically below: fp32 folded weights
        l2 = _Layer(p[2], p[3], p[12:16], packed=pk)
        l3 = _Layer(p[4], p[5], p[16:20], packed=pk)
        l4 = _Layer(p[6], p[7], None, packed=pk)
        C1, C2, C3 = l1.O, l2.O, l3.O
        # layer 1: grid part (K = 3, evaluated inside the layer-1 kernel) + feature part (B x F GEMM) -> relu(G + F)
        wfeat = l1.wf[:, 3:3 + Fdim]                         # (C1,F) view of the folded conv1 weights
        if pf == dense.BF16X3:
            wpk = _empty(C1, _r32(Fdim))
            call("obman_pack_bf16", ptr(wfeat), l1.wf.stride(0), C1, Fdim, ptr(wpk), wpk.stride(0), st)
            Fb = dense.gemm(feat, wpk, bias=l1.shift, passes=pf, n=C1, k=Fdim, packed=True)        # (B,C1)
        else:
            Fb = dense.gemm(feat, wfeat.contiguous(), bias=l1.shift, passes=pf)
        ld1, ld2 = _r32(C1), _r32(C2)
        h1 = _empty(B * N, ld1)
        wg4 = pad_scale_mask(l1.wf, 4, cols=3)               # (C1, 4): the grid columns, one float4 per channel
        call("obman_pointmlp_l1_fwd", ptr(grid), N * 3 if per_sample else 0, ptr(wg4), ptr(Fb), B, N,
             C1, ld1, 1, ptr(h1), st)
        # padding columns of h2 (and of the gradient buffers below) are never read as data: GEMM A operands are fetched
        # through tensor maps whose K extent is the true channel count (TMA zero-fills beyond it), and in the
        # weight-gradient GEMMs a padding channel only produces a padding row / column of the raw gradient
        h2 = _empty(B * N, ld2)
        dense.gemm(h1, l2.wf, out=h2, bias=l2.shift, relu=True, passes=pf, n=C2, k=C1, **l2.fw)
        h3 = dense.gemm(h2, l3.wf, bias=l3.shift, relu=True, passes=pf, n=C3, k=C2, **l3.fw)
        y = dense.gemm(h3, l4.wf, bias=l4.shift * out_factor, alpha=out_factor, passes=pf, n=3, k=C3, **l4.fw)
        ctx.layers = (l1, l2, l3, l4)
        ctx.param_shapes = [tuple(t.shape) for t in p]
        ctx.saved = (feat, grid, h1, h2, h3, B, N, per_sample, out_factor)
        return y.view(B, N, 3)

    @staticmethod
    def backward(ctx, gy):
        pb = dense.PASSES[dense.get_precision()["bwd"]]
        pw = dense.PASSES[dense.get_precision()["wgrad"]]
        st = stream_ptr()
        l1, l2, l3, l4 = ctx.layers
        feat, grid, h1, h2, h3, B, N, per_sample, out_factor = ctx.saved
        C1, C2, C3 = l1.O, l2.O, l3.O
        M = B * N
        Fdim = feat.shape[1]
        # data-gradient chain on the current stream, weight gradients / BN finishes on the auxiliary stream
        keep = []

        def side_layer(layer, g, h, C):
            keep.extend((g, h))
            streams.fork()
            with streams.on_aux():
                if pw == dense.BF16X3:
                    gsum = _empty(g.shape[1])       # over the padded width; finish reads the first C entries
                    return layer.finish(dense.wgrad_matrix(g, h, passes=pw, dy_colsum=gsum), gsum)
                return layer.finish(dense.wgrad_matrix(g, h, passes=pw), colsum(g, C))

        if not (gy.is_cuda and gy.dtype == torch.float32):
            raise RuntimeError("point_decoder backward: expected a CUDA float32 gradient")
        g4 = pad_scale_mask(gy.reshape(M, 3) if gy.is_contiguous() else gy.contiguous().view(M, 3), 32, alpha=out_factor)
        gw4, gb4, _, _ = side_layer(l4, g4, h3, 3)
        g3 = dense.gemm(g4, l4.wft, mask_src=h3, passes=pb, n=C3, k=3, **l4.bw)       # (M,C3)
        gw3, gb3, gg3, gbt3 = side_layer(l3, g3, h2, C3)
        g2 = _empty(M, h2.shape[1])
        dense.gemm(g3, l3.wft, out=g2, mask_src=h2, passes=pb, n=C2, k=C3, **l3.bw)
        gw2, gb2, gg2, gbt2 = side_layer(l2, g2, h1, C2)
        g1 = _empty(M, h1.shape[1])
        dense.gemm(g2, l2.wft, out=g1, mask_src=h1, passes=pb, n=C1, k=C2, **l2.bw)
        # raw weight gradient of conv1 = [grid part (C1,3) | feature part (C1,F)], written into one (C1, 3+F) buffer
        gFc = _empty(B, C1)
        dw1 = _empty(C1, 3 + Fdim)
        call("obman_pointmlp_l1_bwd", ptr(g1), ptr(grid), N * 3 if per_sample else 0, B, N, C1, g1.shape[1], ptr(gFc),
             ptr(_empty(B, 3, C1)), ptr(dw1), dw1.stride(0), st)
        gF = pad_scale_mask(gFc, _r32(C1))
        featp = feat if Fdim == _r32(Fdim) else pad_scale_mask(feat, _r32(Fdim))
        dw1[:, 3:] = dense.wgrad_matrix(gF, featp, passes=pw)[:C1, :Fdim]
        gw1, gb1, gg1, gbt1 = l1.finish(dw1, colsum(gF, C1))
        gfeat = None
        if ctx.needs_input_grad[0]:
            wfeat = l1.wf[:, 3:3 + Fdim]
            if pb == dense.BF16X3:
                wft = _empty(Fdim, _r32(C1))
                call("obman_pack_bf16_t", ptr(wfeat), l1.wf.stride(0), C1, Fdim, ptr(wft), wft.stride(0), st)
                gfeat = dense.gemm(gF, wft, passes=pb, n=Fdim, k=C1, packed=True)
            else:
                wft = _zeros(Fdim, _r32(C1))
                wft[:, :C1] = wfeat.t()
                gfeat = dense.gemm(gF, wft, passes=pb, n=Fdim, k=C1)
        streams.join()
        del keep
        p = ctx.param_shapes
        return (gfeat, None, None,
                gw1.view(p[0]), gb1, gw2.view(p[2]), gb2, gw3.view(p[4]), gb3, gw4.view(p[6]), gb4,
                gg1, gbt1, None, None, gg2, gbt2, None, None, gg3, gbt3, None, None)


class _BatchStatBN(object):
    """BatchNorm1d with batch statistics on a (rows, ld) activation whose first C columns are real (csrc/bn_train.cu works
    on the padded width; padded channels get gamma = beta = 0 and come out as zeros)."""

    def __init__(self, gamma, beta, rmean, rvar, ld, momentum):
        from . import _lib
        self.C, self.ld = gamma.shape[0], ld
        self.gamma = pad_scale_mask(gamma.detach().view(1, -1), ld).view(-1)
        self.beta = pad_scale_mask(beta.detach().view(1, -1), ld).view(-1)
        self.rmean_mod, self.rvar_mod = rmean, rvar
        self.rmean = pad_scale_mask(rmean.detach().view(1, -1), ld).view(-1)
        self.rvar = pad_scale_mask(rvar.detach().view(1, -1), ld).view(-1)
        self.momentum = momentum
        self.mean, self.rstd, self.scale, self.shift = _empty(ld), _empty(ld), _empty(ld), _empty(ld)
        self._chunks = _lib.load().obman_bn_chunks

    def _partial(self, rows):
        return _empty(max(1, self._chunks(int(rows), int(self.ld))) * self.ld * 2)

    def forward(self, z, relu=True):
        rows = z.shape[0]
        call("obman_bn_stats", ptr(z), rows, self.ld, self.ld, ptr(self.gamma), ptr(self.beta), BN_EPS,
             float(self.momentum), ptr(self._partial(rows)), ptr(self.mean), ptr(self.rstd), ptr(self.scale),
             ptr(self.shift), ptr(self.rmean), ptr(self.rvar), stream_ptr())
        with torch.no_grad():   # the module's running statistics (state-dict entries) follow the padded copies
            self.rmean_mod.copy_(self.rmean[:self.C])
            self.rvar_mod.copy_(self.rvar[:self.C])
        y = torch.empty_like(z)
        call("obman_bn_apply_fwd", ptr(z), rows, self.ld, self.ld, ptr(self.scale), ptr(self.shift), None, int(relu),
             ptr(y), stream_ptr())
        return y

    def backward(self, g, y, z):
        """-> (dz (rows, ld), d gamma (C,), d beta (C,)) for g = gradient w.r.t. y = relu(bn(z))."""
        rows = z.shape[0]
        sum_g, sum_gz = _empty(self.ld), _empty(self.ld)
        dz = torch.empty_like(z)
        call("obman_bn_bwd", ptr(g), ptr(y), ptr(z), rows, self.ld, self.ld, ptr(self.mean), ptr(self.rstd),
             ptr(self.scale), ptr(self._partial(rows)), ptr(sum_g), ptr(sum_gz), ptr(dz), None, stream_ptr())
        return dz, sum_gz[:self.C].contiguous(), sum_g[:self.C].contiguous()


class _PointDecoderTrainFn(torch.autograd.Function):
    """PointGenCon with its three BatchNorm1d layers in TRAINING mode (batch statistics over all B*N points,
    running statistics updated): atlasutils.py:65-75 under model.train().  Same argument list as _PointDecoderFn plus
    the three momenta.  3xBF16 path only; not the benchmarked configuration (the README recipe freezes BatchNorm)."""

    @staticmethod
    def forward(ctx, feat, grid, out_factor, momenta, *p):
        if dense.get_precision()["fwd"] != "bf16x3":
            raise RuntimeError("point_decoder: batch-statistics BatchNorm runs on the 3xBF16 path only")
        pf = dense.BF16X3
        st = stream_ptr()
        feat = feat.contiguous()
        grid = grid.detach().contiguous()
        B, Fdim = feat.shape
        N = grid.shape[-2]
        per_sample = grid.dim() == 3
        # plain (unfolded) weights: layer 1 in fp32 (split algebraically), layers 2-4 packed for the tensor cores
        l1 = _Layer(p[0], p[1], None, plain=True)
        l2 = _Layer(p[2], p[3], None, packed=True)
        l3 = _Layer(p[4], p[5], None, packed=True)
        l4 = _Layer(p[6], p[7], None, packed=True)
        C1, C2, C3 = l1.O, l2.O, l3.O
        ld1, ld2, ld3 = _r32(C1), _r32(C2), _r32(C3)
        bn1 = _BatchStatBN(p[8], p[9], p[10], p[11], ld1, momenta[0])
        bn2 = _BatchStatBN(p[12], p[13], p[14], p[15], ld2, momenta[1])
        bn3 = _BatchStatBN(p[16], p[17], p[18], p[19], ld3, momenta[2])
        wfeat = l1.wf[:, 3:3 + Fdim]
        wpk = _empty(C1, _r32(Fdim))
        call("obman_pack_bf16", ptr(wfeat), l1.wf.stride(0), C1, Fdim, ptr(wpk), wpk.stride(0), st)
        Fb = dense.gemm(feat, wpk, bias=l1.shift, passes=pf, n=C1, k=Fdim, packed=True)     # (B,C1) incl. the conv bias
        z1 = _empty(B * N, ld1)     # the layer-1 kernel writes the padding columns as zeros
        wg4 = pad_scale_mask(l1.wf, 4, cols=3)
        call("obman_pointmlp_l1_fwd", ptr(grid), N * 3 if per_sample else 0, ptr(wg4), ptr(Fb), B, N,
             C1, ld1, 0, ptr(z1), st)
        y1 = bn1.forward(z1)
        z2 = _zeros(B * N, ld2)
        dense.gemm(y1, l2.wf, out=z2, bias=l2.shift, passes=pf, n=C2, k=C1, packed=True)
        y2 = bn2.forward(z2)
        z3 = _zeros(B * N, ld3)
        dense.gemm(y2, l3.wf, out=z3, bias=l3.shift, passes=pf, n=C3, k=C2, packed=True)
        y3 = bn3.forward(z3)
        y = dense.gemm(y3, l4.wf, bias=l4.shift * out_factor, alpha=out_factor, passes=pf, n=3, k=C3, packed=True)
        ctx.layers, ctx.bns = (l1, l2, l3, l4), (bn1, bn2, bn3)
        ctx.param_shapes = [tuple(t.shape) for t in p]
        ctx.saved = (feat, grid, z1, y1, z2, y2, z3, y3, B, N, per_sample, out_factor)
        return y.view(B, N, 3)

    @staticmethod
    def backward(ctx, gy):
        pb = pw = dense.BF16X3
        st = stream_ptr()
        l1, l2, l3, l4 = ctx.layers
        bn1, bn2, bn3 = ctx.bns
        feat, grid, z1, y1, z2, y2, z3, y3, B, N, per_sample, out_factor = ctx.saved
        C1, C2, C3 = l1.O, l2.O, l3.O
        M = B * N
        Fdim = feat.shape[1]

        def conv_grads(layer, dz, x, C):
            """weight (O, I) and bias gradients of a plain 1x1 convolution from its output gradient dz and input x."""
            gsum = _empty(dz.shape[1])
            dw = dense.wgrad_matrix(dz, x, passes=pw, dy_colsum=gsum)
            return dw[:layer.O, :layer.I].contiguous(), gsum[:C].contiguous()

        g4 = pad_scale_mask(gy.reshape(M, 3) if gy.is_contiguous() else gy.contiguous().view(M, 3), 32, alpha=out_factor)
        gw4, gb4 = conv_grads(l4, g4, y3, 3)
        gy3 = _zeros(M, z3.shape[1])
        dense.gemm(g4, l4.wft, out=gy3, passes=pb, n=C3, k=3, packed=True)
        dz3, gg3, gbt3 = bn3.backward(gy3, y3, z3)
        gw3, gb3 = conv_grads(l3, dz3, y2, C3)
        gy2 = _zeros(M, z2.shape[1])
        dense.gemm(dz3, l3.wft, out=gy2, passes=pb, n=C2, k=C3, packed=True)
        dz2, gg2, gbt2 = bn2.backward(gy2, y2, z2)
        gw2, gb2 = conv_grads(l2, dz2, y1, C2)
        gy1 = _zeros(M, z1.shape[1])
        dense.gemm(dz2, l2.wft, out=gy1, passes=pb, n=C1, k=C2, packed=True)
        dz1, gg1, gbt1 = bn1.backward(gy1, y1, z1)
        gFc = _empty(B, C1)
        dw1 = _empty(C1, 3 + Fdim)
        call("obman_pointmlp_l1_bwd", ptr(dz1), ptr(grid), N * 3 if per_sample else 0, B, N, C1, dz1.shape[1], ptr(gFc),
             ptr(_empty(B, 3, C1)), ptr(dw1), dw1.stride(0), st)
        gF = pad_scale_mask(gFc, _r32(C1))
        featp = feat if Fdim == _r32(Fdim) else pad_scale_mask(feat, _r32(Fdim))
        dw1[:, 3:] = dense.wgrad_matrix(gF, featp, passes=pw)[:C1, :Fdim]
        gb1 = colsum(gF, C1)
        gfeat = None
        if ctx.needs_input_grad[0]:
            wfeat = l1.wf[:, 3:3 + Fdim]
            wft = _empty(Fdim, _r32(C1))
            call("obman_pack_bf16_t", ptr(wfeat), l1.wf.stride(0), C1, Fdim, ptr(wft), wft.stride(0), st)
            gfeat = dense.gemm(gF, wft, passes=pb, n=Fdim, k=C1, packed=True)
        ps = ctx.param_shapes
        return (gfeat, None, None, None,
                dw1.view(ps[0]), gb1, gw2.view(ps[2]), gb2, gw3.view(ps[4]), gb3, gw4.view(ps[6]), gb4,
                gg1, gbt1, None, None, gg2, gbt2, None, None, gg3, gbt3, None, None)


def point_decoder_train(feat, grid, out_factor, conv_params, bn_params, momenta):
    """point_decoder with the BatchNorm1d layers in training mode (batch statistics)."""
    return _PointDecoderTrainFn.apply(feat, grid, out_factor, list(momenta), *(list(conv_params) + list(bn_params)))


def point_decoder(feat, grid, out_factor, conv_params, bn_params):
    """conv_params: [w1,b1,w2,b2,w3,b3,w4,b4]; bn_params: [g1,b1,m1,v1, g2,..., g3,...]."""
    return _PointDecoderFn.apply(feat, grid, out_factor, *(list(conv_params) + list(bn_params)))
