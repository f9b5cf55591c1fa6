"""Low-level Python entry points of the tcgen05 dense kernels (obman_gemm / obman_conv_nhwc /
obman_wgrad_nhwc) and the tap tables that express fprop / dgrad / wgrad of a strided, padded
convolution as shifted-box implicit GEMMs.  Activations are NHWC, weights are (C_out, KH*KW*C_in).
"""
import ctypes

import torch

from ._lib import call, ptr, stream_ptr

PASSES = {"tf32": 1, "bf16x3": 2, "tf32x3": 3}
BF16X3 = 2
_precision = {"fwd": "bf16x3", "bwd": "bf16x3", "wgrad": "bf16x3"}


def set_precision(fwd="bf16x3", bwd="bf16x3", wgrad=None):
    """Per direction (fwd = fprop, bwd = data gradients, wgrad = weight gradients): 'bf16x3' (fp32 operands
    split into bf16 hi + lo, three products at the bf16 rate, ~2^-17 relative error per product), 'tf32x3'
    (3xTF32 split) or 'tf32' (single pass).  fwd and bwd share one weight layout: both or neither 'bf16x3'."""
    assert fwd in PASSES and bwd in PASSES
    if (fwd == "bf16x3") != (bwd == "bf16x3"):
        raise ValueError("set_precision: 'bf16x3' must be chosen for both fwd and bwd or for neither")
    if wgrad is None:
        wgrad = bwd
    assert wgrad in PASSES
    _precision["fwd"] = fwd
    _precision["bwd"] = bwd
    _precision["wgrad"] = wgrad


def get_precision():
    return dict(_precision)


# ---- optional per-launch profiling (bench.py roofline): CUDA events on the launching stream ------------------
_prof = None


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """Returns algorithmic TFLOP/s of all tensor-core launches since profile_begin (sum flops / sum time)."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    global last_profile
    last_profile = [(tag, f, e0.elapsed_time(e1)) for f, e0, e1, tag in rec]
    ms = sum(t for _, _, t in last_profile)
    flops = sum(f for _, f, _ in last_profile)
    steps = max(1, getattr(profile_end, "steps", 1))
    return {"tflops": flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, "ms_total": ms, "launches": len(rec),
            "ms_per_step": ms / steps, "launches_per_step": len(rec) / steps, "flops": flops}


last_profile = []


def _tc_call(flops, name, *args, tag=""):
    if _prof is None:
        return call(name, *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    _prof.append((float(flops), e0, e1, name[6:] + " " + tag))


def _ints(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def _chk(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError("{}: expected a contiguous CUDA float32 tensor".format(name))
    return t


def split_tf32(w):
    """(hi, lo) with hi = tf32-rounded w and lo = w - hi: pre-split weights for the A-in-TMEM 3xTF32 path."""
    _chk(w, "w")
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    call("obman_split_tf32", ptr(w), w.numel(), ptr(hi), ptr(lo), stream_ptr())
    return hi, lo


def r32(n):
    return (int(n) + 31) // 32 * 32


def pack_bf16(w, k=None):
    """(rows, r32(K)) buffer holding w[:, :K] in the packed bf16 hi|lo layout of the 3xBF16 kernels."""
    _chk(w, "w")
    rows = w.shape[0]
    K = w.shape[1] if k is None else k
    out = torch.empty((rows, r32(K)), device=w.device, dtype=torch.float32)
    call("obman_pack_bf16", ptr(w), w.stride(0), rows, K, ptr(out), out.stride(0), stream_ptr())
    return out


def gemm(a, w, out=None, bias=None, addend=None, mask_src=None, alpha=1.0, relu=False,
         accumulate=False, passes=3, n=None, k=None, w_lo=None, packed=False):
    """out[M,N] = epilogue(alpha * a[M,:K] @ w[:N,:K]^T).  ``a`` / ``w`` may have padded leading
    dimensions (row stride multiple of 4 floats); ``n`` / ``k`` give the logical sizes.  ``w_lo``: residual
    of pre-split weights (``w`` is then the tf32-rounded part), selects the A-in-TMEM kernel.  With
    passes == BF16X3 ``w`` must be (or, unless ``packed``, is converted here to) the pack_bf16 layout."""
    _chk(a, "a"); _chk(w, "w")
    M = a.shape[0]
    K = k if k is not None else a.shape[1]
    N = n if n is not None else w.shape[0]
    if passes == BF16X3 and not packed:
        w = pack_bf16(w, K)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    _tc_call(2.0 * M * N * K, "obman_gemm", ptr(a), a.stride(0), ptr(w), ptr(w_lo) if passes == 3 else None,
             w.stride(0), M, N, K, ptr(out), out.stride(0),
             ptr(bias), ptr(addend), ptr(mask_src), float(alpha), int(relu), int(accumulate), int(passes),
             stream_ptr(), tag="M%d N%d K%d" % (M, N, K))
    return out


# ---- tap tables -------------------------------------------------------------------------------------
def fprop_taps(ksize, stride, pad):
    """Taps of y[h,w] = sum_{kh,kw} x[h*stride + kh - pad, w*stride + kw - pad] * W[kh,kw] in the
    phase-view coordinates of obman_conv_nhwc: returns (dh, dw, phase, wslot, in_step)."""
    dh, dw, phase, slot = [], [], [], []
    for kh in range(ksize):
        for kw in range(ksize):
            oh, ow = kh - pad, kw - pad
            if stride == 1:
                dh.append(oh); dw.append(ow); phase.append(0)
            else:
                ph, pw = oh % 2, ow % 2
                dh.append((oh - ph) // 2); dw.append((ow - pw) // 2); phase.append(ph * 2 + pw)
            slot.append(kh * ksize + kw)
    return dh, dw, phase, slot, stride


def dgrad_taps(ksize, stride, pad, out_phase=(0, 0)):
    """Taps of dx[h,w] = sum dy[(h + pad - kh)/stride, (w + pad - kw)/stride] * W[kh,kw] for the output
    pixels h = stride*i + out_phase[0], w = stride*j + out_phase[1]: (dh, dw, wslot) on the dy grid."""
    dh, dw, slot = [], [], []
    for kh in range(ksize):
        for kw in range(ksize):
            th, tw = out_phase[0] + pad - kh, out_phase[1] + pad - kw
            if th % stride or tw % stride:
                continue
            dh.append(th // stride); dw.append(tw // stride); slot.append(kh * ksize + kw)
    return dh, dw, slot


def conv_nhwc(x, w, c_out, taps, in_step, out, h_out, w_out, out_strides=None, out_offset=0,
              bias=None, addend=None, mask_src=None, relu=False, passes=3, w_slots=None, algo_k=None,
              w_lo=None, x_geom=None):
    """Raw obman_conv_nhwc call.  x (N,H,W,C) contiguous; w (c_out, slots*C); taps = (dh, dw, phase, slot).
    ``out`` is any tensor whose storage receives element (n,h,w,c) at out_offset + n*sN + h*sH + w*sW + c.
    ``x_geom`` = (N, H, W, C, sN, sH, sW): read ``x``'s storage as that strided (possibly overlapping) NHWC view."""
    _chk(x, "x"); _chk(w, "w")
    if x_geom is None:
        n_img, h_in, w_in, c_in = x.shape
        xs = (0, 0, 0)
    else:
        n_img, h_in, w_in, c_in = x_geom[:4]
        xs = tuple(int(v) for v in x_geom[4:7])
    dh, dw, phase, slot = taps
    if w_slots is None:
        w_slots = w.shape[1] // c_in
    if out_strides is None:
        out_strides = (h_out * w_out * c_out, w_out * c_out, c_out)
    esz = 4

    def off(t):
        return None if t is None else t.data_ptr() + out_offset * esz

    k_eff = algo_k if algo_k is not None else len(dh) * c_in
    _tc_call(2.0 * n_img * h_out * w_out * c_out * k_eff,
             "obman_conv_nhwc", ptr(x), n_img, h_in, w_in, c_in, int(in_step), xs[0], xs[1], xs[2], ptr(w),
             ptr(w_lo) if passes == 3 else None, int(c_out),
         int(w_slots), len(dh), _ints(dh), _ints(dw), _ints(phase) if phase is not None else None,
         _ints(slot), off(out), int(h_out), int(w_out), int(out_strides[0]), int(out_strides[1]),
         int(out_strides[2]), ptr(bias), off(addend), off(mask_src), int(relu), int(passes),
         stream_ptr(), tag="n%d %dx%d c%d->%d taps%d" % (n_img, h_out, w_out, c_in, c_out, len(dh)))
    return out


def wgrad_nhwc(dy, x, taps, in_step, dw_out, passes=3, algo_k=None, x_geom=None, dy_colsum=None):
    """dw_out (c_out, num_taps*c_in) = sum over pixels dy (N,Ho,Wo,c_out) x shifted x (N,H,W,c_in); the tap
    order is the weight-slot order (taps = (dh, dw, phase, slot) with slot == position).  ``x_geom`` as in conv_nhwc.
    ``dy_colsum`` (c_out,), 3xBF16 only: also receives the per-channel sum of dy (fused into the operand split)."""
    _chk(dy, "dy"); _chk(x, "x"); _chk(dw_out, "dw")
    n_img, h_out, w_out, c_out = dy.shape
    if x_geom is None:
        _, h_in, w_in, c_in = x.shape
        xs = (0, 0, 0)
    else:
        _, h_in, w_in, c_in = x_geom[:4]
        xs = tuple(int(v) for v in x_geom[4:7])
    dh, dw, phase, slot = taps
    if list(slot) != list(range(len(dh))):
        raise RuntimeError("wgrad_nhwc: taps must be listed in weight-slot order")
    k_eff = algo_k if algo_k is not None else len(dh) * c_in
    _tc_call(2.0 * n_img * h_out * w_out * c_out * k_eff,
             "obman_wgrad_nhwc", ptr(dy), n_img, h_out, w_out, c_out, ptr(x), h_in, w_in, c_in,
             int(in_step), xs[0], xs[1], xs[2], len(dh), _ints(dh), _ints(dw),
             _ints(phase) if phase is not None else None, ptr(dw_out), ptr(dy_colsum), int(passes), stream_ptr(),
             tag="n%d %dx%d c%d->%d taps%d" % (n_img, h_out, w_out, c_in, c_out, len(dh)))
    return dw_out


def wgrad_matrix(dy, x, dw_out=None, passes=3, dy_colsum=None):
    """dw[N,K] = dy[M,N]^T @ x[M,K]; N and K (the padded row lengths) multiples of 32."""
    M, N = dy.shape
    K = x.shape[1]
    if dw_out is None:
        dw_out = torch.empty((N, K), device=dy.device, dtype=torch.float32)
    return wgrad_nhwc(dy.view(1, 1, M, N), x.view(1, 1, M, K), ([0], [0], None, [0]), 1, dw_out, passes,
                      dy_colsum=dy_colsum)
