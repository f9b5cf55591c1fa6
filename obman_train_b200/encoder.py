"""ResNet-18 image encoder as ONE hand-scheduled autograd node over the tcgen05 convolution kernels.

Forward and backward of mano_train/networks/bases/resnet.py:154-188 (conv7x7/2 -> BN -> ReLU -> maxpool ->
4x2 BasicBlocks -> spatial mean) with BatchNorm in eval mode (the README recipe trains with
--freeze_batchnorm, i.e. fixed running statistics and trainable gamma/beta):

* activations are NHWC; BN scale is folded into the weights once per step, BN shift / residual add /
  ReLU run in the convolution epilogue, so no elementwise pass ever touches HBM;
* the backward pass fuses the ReLU mask and the residual-branch add into the dgrad epilogues and derives
  d(gamma), d(beta) from the raw weight gradient (sum_k W*dWraw), so no pre-BN tensor is saved.
"""
import torch

from . import dense, streams
from ._lib import call, ptr, stream_ptr

BN_EPS = 1e-5
DEBUG = None  # set to a dict to capture intermediate gradients (diagnostics only)

# Gradient sink (set by FlatAdamTrainer for the duration of a step): an object with
#   view(param_data_ptr) -> tensor view of the flat gradient buffer shaped like that parameter, or None
#   stage_done(first_ptr, last_ptr)  called once per encoder when the gradients of layer3 + layer4 (94 % of the
#                                    encoder's parameters) are complete on the WGRAD stream
# With a sink the weight-gradient kernels write straight into the flat buffer (no per-parameter gradient tensors,
# no gather copy) and the trainer can start the all-reduce of that range under the rest of the backward pass.
_grad_sink = None


def set_grad_sink(sink):
    global _grad_sink
    prev, _grad_sink = _grad_sink, sink
    return prev

# (name, out_ch, in_ch, ksize, stride) of the 20 conv+BN units in state-dict order
def resnet18_units():
    units = [("conv1", "bn1", 64, 3, 7, 2)]
    inp = 64
    for li, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
        for bi in range(2):
            s = stride if bi == 0 else 1
            p = "layer{}.{}.".format(li, bi)
            units.append((p + "conv1", p + "bn1", planes, inp, 3, s))
            units.append((p + "conv2", p + "bn2", planes, planes, 3, 1))
            if bi == 0 and (s != 1 or inp != planes):
                units.append((p + "downsample.0", p + "downsample.1", planes, inp, 1, s))
            inp = planes
    return units


def _empty(*shape):
    return torch.empty(shape, device="cuda", dtype=torch.float32)


def stem_view(xs):
    """xs (B, Ho, Wo + 4, 16) from obman_stem_pack -> geometry of the overlapping (B, Ho, Wo, 64) view the stem
    convolution reads: pixel stride 16 floats, so pixel j of the view covers unpadded pixels j-2 .. j+1."""
    B, Ho, Wp, C = xs.shape
    return (B, Ho, Wp - 4, 64, Ho * Wp * C, Wp * C, C)


class _NullCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _Unit(object):
    """Per-step folded weights of one conv+BN unit."""

    def __init__(self, w, gamma, beta, mean, var, ksize, stride, stem=False, need_dgrad=True, packed=False, fold=True):
        self.w, self.gamma, self.mean = w, gamma, mean
        self.beta, self.var = beta, var
        self.beta_ptr = beta.data_ptr()
        self.O, self.I = w.shape[0], w.shape[1]
        self.k, self.stride, self.stem = ksize, stride, stem
        self.Ip = 64 if stem else self.I
        self.slots = 4 if stem else ksize * ksize
        # weights are folded AND pre-split for the A-in-tensor-memory kernels: packed bf16 hi|lo (3xBF16) or
        # tf32 hi / lo arrays (3xTF32)
        self.packed = packed
        self.wf = _empty(self.O, self.slots * self.Ip)
        self.wf_lo = None if packed else _empty(self.O, self.slots * self.Ip)
        self.wft = self.wft_lo = None
        if need_dgrad and not stem:
            self.wft = _empty(self.I, self.slots * self.O)
            self.wft_lo = None if packed else _empty(self.I, self.slots * self.O)
        self.shift, self.scale, self.rstd = _empty(self.O), _empty(self.O), _empty(self.O)
        if fold:
            call("obman_fold_conv", ptr(w), None, ptr(gamma), ptr(beta), ptr(mean), ptr(var), BN_EPS, self.O, self.I,
                 ksize, ksize, self.Ip, int(stem), int(packed), ptr(self.wf), ptr(self.wf_lo), ptr(self.wft),
                 ptr(self.wft_lo), ptr(self.shift), ptr(self.scale), ptr(self.rstd), stream_ptr())
        if stem:
            self.taps = ([-2, -1, 0, 1], [0, 0, 0, 0], [0] * 4, list(range(4)))
            self.in_step = 1
        else:
            dh, dw, phase, slot, step = dense.fprop_taps(ksize, stride, ksize // 2)
            self.taps = (dh, dw, phase, slot)
            self.in_step = step

    def fprop(self, x, h_out, w_out, addend=None, relu=True, passes=3):
        out = _empty(x.shape[0], h_out, w_out, self.O)
        dense.conv_nhwc(x, self.wf, self.O, self.taps, self.in_step, out, h_out, w_out, bias=self.shift,
                        addend=addend, relu=relu, passes=passes, w_slots=self.slots,
                        algo_k=147 if self.stem else None, w_lo=self.wf_lo,
                        x_geom=stem_view(x) if self.stem else None)
        return out

    def dgrad(self, g, h_in, w_in, addend=None, mask_src=None, passes=3, phase00_only=False, addend_phase00=False):
        """g (B,Ho,Wo,O) -> gx (B,h_in,w_in,I) = conv^T(g) [+ addend] [masked by mask_src > 0].
        A stride-2 convolution is four launches that write the four interleaved output phases; each is a small
        GEMM (a quarter of the pixels, 1-4 taps) that cannot fill 148 SMs on its own, so two of them are issued on
        the CHAIN auxiliary stream (disjoint outputs, same inputs).

        A 1x1 stride-2 convolution (the downsample branch) reaches output phase (0, 0) only.  ``phase00_only``: the
        caller promises to read nothing but that phase, so the other three quarters of ``gx`` stay uninitialised
        instead of being zero-filled; ``addend_phase00`` is the consuming side: ``addend`` holds data in phase (0, 0)
        only and is treated as zero elsewhere (three of the four phase launches then skip the addend read)."""
        B = g.shape[0]
        s, k = self.stride, self.k
        # a 1x1 stride-2 convolution reaches one of the four output phases only: one contiguous memset instead of
        # three strided fills
        sparse = s > 1 and k == 1 and addend is None
        gx = torch.zeros((B, h_in, w_in, self.I), device="cuda", dtype=torch.float32) if sparse and not phase00_only \
            else _empty(B, h_in, w_in, self.I)
        two_lanes = s > 1 and k > 1 and streams.enabled()
        if two_lanes:
            streams.fork(streams.CHAIN)
        for ph in range(s):
            for pw in range(s):
                dh, dw, slot = dense.dgrad_taps(k, s, k // 2, (ph, pw))
                off = (ph * w_in + pw) * self.I
                strides = (h_in * w_in * self.I, s * w_in * self.I, s * self.I)
                if not dh:  # this output phase receives no contribution from the convolution
                    if addend is not None:
                        view = gx[:, ph::s, pw::s]
                        src = addend[:, ph::s, pw::s]
                        view.copy_(src if mask_src is None else src * (mask_src[:, ph::s, pw::s] > 0))
                    continue  # (without addend: gx was allocated zero-filled, see above)
                lane = streams.on_aux(streams.CHAIN) if (two_lanes and ph == 1) else _NullCtx()
                add = None if (addend_phase00 and s > 1 and (ph, pw) != (0, 0)) else addend
                with lane:
                    dense.conv_nhwc(g, self.wft, self.I, (dh, dw, None, slot), 1, gx, h_in // s, w_in // s,
                                    out_strides=strides, out_offset=off, addend=add, mask_src=mask_src,
                                    passes=passes, w_slots=k * k, w_lo=self.wft_lo)
        if two_lanes:
            streams.join(streams.CHAIN)
        return gx

    def wgrad(self, g, x, passes=3, colsum_out=None):
        """Raw weight gradient; ``colsum_out`` (O,) additionally receives the per-channel sum of ``g`` (3xBF16 path:
        accumulated by the threads that split ``g`` for the tensor cores, no separate pass over the gradient)."""
        dwraw = _empty(self.O, self.slots * self.Ip)
        dense.wgrad_nhwc(g, x, self.taps, self.in_step, dwraw, passes=passes,
                         algo_k=147 if self.stem else None, x_geom=stem_view(x) if self.stem else None,
                         dy_colsum=colsum_out)
        return dwraw

    def finish(self, dwraw, gbeta_sum):
        """-> (gw, ggamma, gbeta); entries are None where the gradient went straight into the sink's flat buffer."""
        sink = _grad_sink
        views = [None, None, None]
        if sink is not None:
            views = [sink.view(self.w.data_ptr()), sink.view(self.gamma.data_ptr()), sink.view(self.beta_ptr)]
        direct = all(v is not None for v in views)
        gw = views[0] if direct else torch.empty_like(self.w)
        ggamma = views[1] if direct else torch.empty_like(self.gamma)
        gbeta = views[2] if direct else torch.empty_like(self.gamma)
        call("obman_bn_wgrad_finish", ptr(dwraw), dwraw.stride(0), ptr(self.w), None, ptr(self.scale),
             ptr(self.rstd), ptr(self.mean), ptr(gbeta_sum), self.O, self.I, self.k, self.k, self.Ip,
             int(self.stem), ptr(gw), ptr(ggamma), ptr(gbeta), None, stream_ptr())
        if direct:
            sink.wrote(self.w.data_ptr(), self.gamma.data_ptr(), self.beta_ptr)
            return None, None, None
        return gw, ggamma, gbeta


def fold_units(units):
    """BatchNorm folding + bf16 hi|lo packing of all units in ONE launch (obman_fold_conv_batch; packed layout only)."""
    import ctypes

    def ptrs(ts):
        return (ctypes.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])

    def ints(vs):
        return (ctypes.c_int * len(vs))(*[int(v) for v in vs])

    call("obman_fold_conv_batch", len(units), ptrs([u.w for u in units]), ptrs([u.gamma for u in units]),
         ptrs([u.beta for u in units]), ptrs([u.mean for u in units]), ptrs([u.var for u in units]), BN_EPS,
         ints([u.O for u in units]), ints([u.I for u in units]), ints([u.k for u in units]),
         ints([u.Ip for u in units]), ints([int(u.stem) for u in units]), ptrs([u.wf for u in units]),
         ptrs([u.wft for u in units]), ptrs([u.shift for u in units]), ptrs([u.scale for u in units]),
         ptrs([u.rstd for u in units]), stream_ptr())


def colsum(x2d_rows, C, t):
    out = _empty(C)
    call("obman_colsum", ptr(t), int(x2d_rows), int(C), int(C), ptr(out), stream_ptr())
    return out


class _EncoderFn(torch.autograd.Function):
    """features (B,512) = resnet18(images (B,3,H,W)); params = 5 tensors per unit (w, gamma, beta, mean, var)."""

    @staticmethod
    def forward(ctx, images, *params):
        if not (images.is_cuda and images.dtype == torch.float32):
            raise RuntimeError("encoder: expected CUDA float32 images (no CPU path)")
        images = images.contiguous()
        B, _, H, W = images.shape
        if H % 32 or W % 32:
            raise RuntimeError("encoder: image height/width must be multiples of 32")
        pf = dense.PASSES[dense.get_precision()["fwd"]]
        specs = resnet18_units()
        packed = pf == dense.BF16X3
        unit_params = [[p.detach().contiguous() for p in params[5 * i:5 * i + 5]] for i in range(len(specs))]

        def make_unit(i, fold=True):
            _, _, O, I, k, s = specs[i]
            w, gamma, beta, mean, var = unit_params[i]
            return _Unit(w, gamma, beta, mean, var, k, s, stem=(i == 0), packed=packed, fold=fold)

        st = stream_ptr()
        xs = _empty(B, H // 2, W // 2 + 4, 16)
        if packed:
            # BN folding + bf16 packing of all 20 units: one launch on the auxiliary stream, underneath the stem pack
            streams.fork()
            with streams.on_aux():
                units = [make_unit(i, fold=False) for i in range(len(specs))]
                fold_units(units)
            call("obman_stem_pack", ptr(images), B, H, W, ptr(xs), st)
            streams.join()
        else:
            units = [make_unit(0)]
            # 3xTF32 layouts: one folding launch per unit; units 1..19 on the auxiliary stream underneath the HBM-bound
            # front of the network (stem pack, stem convolution, max-pool)
            streams.fork()
            with streams.on_aux():
                units.extend(make_unit(i) for i in range(1, len(specs)))
            call("obman_stem_pack", ptr(images), B, H, W, ptr(xs), st)
        c1 = units[0].fprop(xs, H // 2, W // 2, relu=True, passes=pf)
        if DEBUG is not None:
            DEBUG["act_c1"] = c1
        hp, wp = H // 4, W // 4
        p = _empty(B, hp, wp, 64)
        pidx = torch.empty((B, hp, wp, 64), device="cuda", dtype=torch.uint8)
        call("obman_maxpool_fwd", ptr(c1), B, H // 2, W // 2, 64, ptr(p), ptr(pidx), st)
        if not packed:
            streams.join()
        if DEBUG is not None:
            DEBUG["pool_idx"] = pidx
        x, h, w_ = p, hp, wp
        blocks = []
        ui = 1
        for li in range(4):
            for bi in range(2):
                u1, u2 = units[ui], units[ui + 1]
                ud = None
                nxt = ui + 2
                if nxt < len(specs) and "downsample" in specs[nxt][0] and specs[nxt][0].startswith("layer{}.{}.".format(li + 1, bi)):
                    ud = units[nxt]
                    nxt += 1
                ho, wo = h // u1.stride, w_ // u1.stride
                a = u1.fprop(x, ho, wo, relu=True, passes=pf)
                r = ud.fprop(x, ho, wo, relu=False, passes=pf) if ud is not None else x
                out = u2.fprop(a, ho, wo, addend=r, relu=True, passes=pf)
                blocks.append((u1, u2, ud, x, a, out, h, w_))
                if DEBUG is not None:
                    DEBUG["act_a%d" % (len(blocks) - 1)] = a
                    DEBUG["act_out%d" % (len(blocks) - 1)] = out
                x, h, w_ = out, ho, wo
                ui = nxt
        feats = _empty(B, 512)
        call("obman_meanpool_fwd", ptr(x), B, h * w_, 512, ptr(feats), st)
        ctx.units, ctx.blocks = units, blocks
        ctx.saved = (xs, pidx, p, H, W)
        return feats

    @staticmethod
    def backward(ctx, gfeat):
        units, blocks = ctx.units, ctx.blocks
        xs, pidx, p, H, W = ctx.saved
        pb = dense.PASSES[dense.get_precision()["bwd"]]      # data gradients
        pw = dense.PASSES[dense.get_precision()["wgrad"]]    # weight gradients
        st = stream_ptr()
        gfeat = gfeat.contiguous()
        B = gfeat.shape[0]
        last = blocks[-1][5]
        hl, wl = last.shape[1], last.shape[2]
        g2 = torch.empty_like(last)
        call("obman_meanpool_bwd", ptr(gfeat), ptr(last), B, hl * wl, 512, ptr(g2), st)
        grads = {}
        # Critical path (current stream): the chain of data gradients.  Off the path (auxiliary stream): per unit the
        # column sum of the incoming gradient (d beta), the weight-gradient GEMM and its BatchNorm finish.  ``keep``
        # pins every tensor that crosses the two streams until they are joined (see streams.py).
        keep = []

        def side_unit(u, g, x, rows, gb=None):
            keep.extend((g, x))
            streams.fork()  # g was produced on the current stream
            with streams.on_aux():
                if gb is None and pw == dense.BF16X3:
                    gb = _empty(u.O)
                    dwraw = u.wgrad(g, x, pw, colsum_out=gb)   # d beta comes out of the weight-gradient kernel
                else:
                    if gb is None:
                        gb = colsum(rows, u.O, g)
                    dwraw = u.wgrad(g, x, pw)
                grads[id(u)] = u.finish(dwraw, gb)
            return gb

        for bidx, (u1, u2, ud, x, a, out, h, w_) in reversed(list(enumerate(blocks))):
            ho, wo = out.shape[1], out.shape[2]
            if DEBUG is not None:
                DEBUG["g_out_%d" % bidx] = g2.clone()
            rows = B * ho * wo
            gb2 = side_unit(u2, g2, a, rows)
            if ud is not None:
                side_unit(ud, g2, x, rows, gb=gb2)
            if ud is not None:
                # the downsample branch's data gradient only needs g2: second lane, next to conv2's dgrad
                keep.append(g2)
                streams.fork(streams.CHAIN)
                with streams.on_aux(streams.CHAIN):
                    gres = ud.dgrad(g2, h, w_, passes=pb, phase00_only=True)
            else:
                gres = g2
            g1 = u2.dgrad(g2, ho, wo, mask_src=a, passes=pb)
            if DEBUG is not None:
                DEBUG["g_a_%d" % bidx] = g1.clone()
            side_unit(u1, g1, x, rows)
            if ud is not None:
                streams.join(streams.CHAIN)
            keep.append(gres)
            g2 = u1.dgrad(g1, h, w_, addend=gres, mask_src=x, passes=pb, addend_phase00=ud is not None)
            if bidx == 4 and _grad_sink is not None:
                # layer4 and layer3 are done on the WGRAD stream: their flat-buffer range can go on the wire now
                first = blocks[4][0]                      # layer3.0.conv1
                last = blocks[7][1]                       # layer4.1.conv2 (bn2 is its last parameter)
                _grad_sink.stage_done(first.w.data_ptr(), last.beta_ptr, streams.aux_stream() if streams.enabled() else None)
        # g2 is now the gradient w.r.t. the max-pool output (already masked by p > 0)
        gc1 = _empty(B, H // 2, W // 2, 64)
        call("obman_maxpool_bwd", ptr(g2), ptr(pidx), B, H // 2, W // 2, 64, ptr(gc1), st)
        if DEBUG is not None:
            DEBUG["g_p"] = g2.clone()
            DEBUG["g_c1"] = gc1.clone()
        u0 = units[0]
        if pw == dense.BF16X3:
            gb0 = _empty(64)
            grads[id(u0)] = u0.finish(u0.wgrad(gc1, xs, pw, colsum_out=gb0), gb0)
        else:
            grads[id(u0)] = u0.finish(u0.wgrad(gc1, xs, pw), colsum(B * (H // 2) * (W // 2), 64, gc1))
        streams.join()
        del keep
        outs = [None]
        for u in units:
            gw, gg, gbt = grads[id(u)]
            outs.extend([gw, gg, gbt, None, None])
        return tuple(outs)


class _TrainUnit(_Unit):
    """conv + BatchNorm with BATCH statistics (model.train(), i.e. training without --freeze_batchnorm): the raw
    weights go to the tensor-core kernels (bf16 hi|lo packed, no folding), the statistics / normalisation / ReLU and
    their gradients are the HBM-bound kernels of csrc/bn_train.cu."""

    def __init__(self, w, gamma, beta, rmean, rvar, ksize, stride, stem=False, momentum=0.1, update_stats=True):
        super(_TrainUnit, self).__init__(w, gamma, beta, rmean, rvar, ksize, stride, stem=stem, packed=True, fold=False)
        self.rmean, self.rvar = rmean, rvar
        self.momentum, self.update_stats = momentum, update_stats
        self.mean_b = _empty(self.O)
        # gamma = NULL: plain (unscaled) weights in the packed layouts, shift = 0
        call("obman_fold_conv", ptr(w), None, None, None, None, None, BN_EPS, self.O, self.I, ksize, ksize, self.Ip,
             int(stem), 1, ptr(self.wf), None, ptr(self.wft), None, ptr(self.shift), ptr(self.scale), ptr(self.rstd),
             stream_ptr())

    def _partial(self, rows):
        chunks = _lib_load().obman_bn_chunks(int(rows), int(self.O))
        return _empty(max(1, chunks) * self.O * 2)

    def forward(self, x, h_out, w_out, addend=None, relu=True, passes=dense.BF16X3):
        """-> (z raw convolution output, y = [relu](bn(z) [+ addend])), both (B, h_out, w_out, O)."""
        z = _empty(x.shape[0], h_out, w_out, self.O)
        dense.conv_nhwc(x, self.wf, self.O, self.taps, self.in_step, z, h_out, w_out, passes=passes, w_slots=self.slots,
                        algo_k=147 if self.stem else None, x_geom=stem_view(x) if self.stem else None)
        rows = z.shape[0] * h_out * w_out
        upd = self.update_stats
        call("obman_bn_stats", ptr(z), rows, self.O, self.O, ptr(self.gamma), ptr(self.beta), BN_EPS,
             float(self.momentum), ptr(self._partial(rows)), ptr(self.mean_b), ptr(self.rstd), ptr(self.scale),
             ptr(self.shift), ptr(self.rmean) if upd else None, ptr(self.rvar) if upd else None, stream_ptr())
        y = torch.empty_like(z)
        call("obman_bn_apply_fwd", ptr(z), rows, self.O, self.O, ptr(self.scale), ptr(self.shift), ptr(addend),
             int(relu), ptr(y), stream_ptr())
        return z, y

    def bn_backward(self, g, mask_src, z, want_masked=False):
        """g = gradient w.r.t. the unit's output (after the ReLU whose mask is ``mask_src > 0``; None = no ReLU) ->
        (dz, d beta, d gamma, g masked or None)."""
        rows = z.shape[0] * z.shape[1] * z.shape[2]
        sum_g, sum_gz = _empty(self.O), _empty(self.O)
        dz = torch.empty_like(z)
        gm = torch.empty_like(z) if want_masked else None
        call("obman_bn_bwd", ptr(g), ptr(mask_src), ptr(z), rows, self.O, self.O, ptr(self.mean_b), ptr(self.rstd),
             ptr(self.scale), ptr(self._partial(rows)), ptr(sum_g), ptr(sum_gz), ptr(dz), ptr(gm), stream_ptr())
        return dz, sum_g, sum_gz, gm

    def weight_grad(self, dz, x, passes=dense.BF16X3):
        """(O, I, k, k) weight gradient from the raw stacked-tap layout (bn_wgrad_finish with unit scale: a re-layout)."""
        dwraw = self.wgrad(dz, x, passes)
        gw = torch.empty_like(self.w)
        ones = torch.ones(self.O, device=dz.device)   # (train mode is not the benchmarked path)
        call("obman_bn_wgrad_finish", ptr(dwraw), dwraw.stride(0), ptr(self.w), None, ptr(ones), ptr(ones), None,
             ptr(ones), self.O, self.I, self.k, self.k, self.Ip, int(self.stem), ptr(gw), None, None, None, stream_ptr())
        return gw


def _lib_load():
    from . import _lib
    return _lib.load()


class _EncoderTrainFn(torch.autograd.Function):
    """resnet18(images) with BatchNorm in TRAINING mode (batch statistics, running statistics updated with the modules'
    momentum): bases/resnet.py:154-188 under model.train().  params = 5 tensors per unit (w, gamma, beta,
    running_mean, running_var); ``momenta`` = one float per unit."""

    @staticmethod
    def forward(ctx, images, momenta, *params):
        if not (images.is_cuda and images.dtype == torch.float32):
            raise RuntimeError("encoder: expected CUDA float32 images (no CPU path)")
        if dense.get_precision()["fwd"] != "bf16x3":
            raise RuntimeError("encoder: batch-statistics BatchNorm runs on the 3xBF16 path only")
        images = images.contiguous()
        B, _, H, W = images.shape
        if H % 32 or W % 32:
            raise RuntimeError("encoder: image height/width must be multiples of 32")
        specs = resnet18_units()
        st = stream_ptr()
        units = []
        for i, (_, _, O, I, k, s) in enumerate(specs):
            w, gamma, beta, rm, rv = [p.detach() for p in params[5 * i:5 * i + 5]]
            units.append(_TrainUnit(w.contiguous(), gamma.contiguous(), beta.contiguous(), rm, rv, k, s, stem=(i == 0),
                                    momentum=momenta[i]))
        xs = _empty(B, H // 2, W // 2 + 4, 16)
        call("obman_stem_pack", ptr(images), B, H, W, ptr(xs), st)
        z0, c1 = units[0].forward(xs, H // 2, W // 2)
        hp, wp = H // 4, W // 4
        p = _empty(B, hp, wp, 64)
        pidx = torch.empty((B, hp, wp, 64), device="cuda", dtype=torch.uint8)
        call("obman_maxpool_fwd", ptr(c1), B, H // 2, W // 2, 64, ptr(p), ptr(pidx), st)
        x, h, w_ = p, hp, wp
        blocks = []
        ui = 1
        for li in range(4):
            for bi in range(2):
                u1, u2 = units[ui], units[ui + 1]
                ud = None
                nxt = ui + 2
                if nxt < len(specs) and "downsample" in specs[nxt][0] and specs[nxt][0].startswith("layer{}.{}.".format(li + 1, bi)):
                    ud = units[nxt]
                    nxt += 1
                ho, wo = h // u1.stride, w_ // u1.stride
                z1, a = u1.forward(x, ho, wo)
                zd, r = ud.forward(x, ho, wo, relu=False) if ud is not None else (None, x)
                z2, out = u2.forward(a, ho, wo, addend=r)
                blocks.append((u1, u2, ud, x, z1, a, zd, z2, out, h, w_))
                x, h, w_ = out, ho, wo
                ui = nxt
        feats = _empty(B, 512)
        call("obman_meanpool_fwd", ptr(x), B, h * w_, 512, ptr(feats), st)
        ctx.units, ctx.blocks = units, blocks
        ctx.saved = (xs, z0, c1, pidx, H, W)
        return feats

    @staticmethod
    def backward(ctx, gfeat):
        units, blocks = ctx.units, ctx.blocks
        xs, z0, c1, pidx, H, W = ctx.saved
        st = stream_ptr()
        gfeat = gfeat.contiguous()
        B = gfeat.shape[0]
        last = blocks[-1][8]
        g2 = torch.empty_like(last)
        call("obman_meanpool_bwd", ptr(gfeat), ptr(last), B, last.shape[1] * last.shape[2], 512, ptr(g2), st)
        grads = {}
        for (u1, u2, ud, x, z1, a, zd, z2, out, h, w_) in reversed(blocks):
            ho, wo = out.shape[1], out.shape[2]
            # out = relu(bn2(conv2(a)) + r): the ReLU mask comes from ``out``; its masked gradient also feeds the residual
            dz2, gb2, gg2, gmask = u2.bn_backward(g2, out, z2, want_masked=True)
            grads[id(u2)] = (u2.weight_grad(dz2, a), gg2, gb2)
            ga = u2.dgrad(dz2, ho, wo, passes=dense.BF16X3)
            if ud is not None:
                dzd, gbd, ggd, _ = ud.bn_backward(gmask, None, zd)
                grads[id(ud)] = (ud.weight_grad(dzd, x), ggd, gbd)
                gres = ud.dgrad(dzd, h, w_, passes=dense.BF16X3)
            else:
                gres = gmask
            dz1, gb1, gg1, _ = u1.bn_backward(ga, a, z1)
            grads[id(u1)] = (u1.weight_grad(dz1, x), gg1, gb1)
            g2 = u1.dgrad(dz1, h, w_, addend=gres, passes=dense.BF16X3)
        gc1 = _empty(B, H // 2, W // 2, 64)
        call("obman_maxpool_bwd", ptr(g2), ptr(pidx), B, H // 2, W // 2, 64, ptr(gc1), st)
        u0 = units[0]
        dz0, gb0, gg0, _ = u0.bn_backward(gc1, c1, z0)
        grads[id(u0)] = (u0.weight_grad(dz0, xs), gg0, gb0)
        outs = [None, None]
        for u in units:
            gw, gg, gbt = grads[id(u)]
            outs.extend([gw, gg, gbt, None, None])
        return tuple(outs)


def resnet18_features_train(images, params, momenta):
    """Batch-statistics BatchNorm (the module is in training mode)."""
    return _EncoderTrainFn.apply(images, list(momenta), *params)


def resnet18_features(images, params):
    return _EncoderFn.apply(images, *params)
