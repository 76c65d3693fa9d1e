"""Torch-tensor wrappers over the C ABI (include/vnqa_b200.h).  PyTorch is used for device memory and the
current stream only; all arithmetic happens in libvnqa_b200.so."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import ConvArgs
from ._lib import check as _check

IMG_PIX = 130
_KERNELS_PER_CALL = {"bn_backward": 3, "vnqa_device_check": 0, "conv_wgrad": 1}   # entry points that launch != 1 kernel


def check(rc, what):
    _check(rc, what)
    _count(_KERNELS_PER_CALL.get(what, 1))
    if TRACE is not None:       # diagnostics only (tools/hangwatch.py): one event per launch on the launching stream
        st = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(st)
        TRACE.append((what, ev, st.cuda_stream))

# ---- measurement hooks (bench.py): kernel-launch counter and optional per-launch CUDA-event timing of the
# tensor-core kernels on the stream they are launched on.  Off by default; they never change what is computed.
LAUNCHES = 0
PROFILE = None      # list of (tag, flops, start_event, end_event) when enabled
TRACE = None        # collections.deque of (entry point, completion event, stream) when enabled (tools/hangwatch.py)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


class _timed:
    def __init__(self, tag, flops):
        self.tag, self.flops = tag, flops

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.tag, self.flops, self.e0, self.e1))


def _p(t):
    if t is None:
        return None
    assert t.is_cuda, "videonavqa_b200 ops need CUDA tensors (no CPU fallback)"
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    assert t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t


def round_up(x, m):
    return (x + m - 1) // m * m


def device_check():
    check(_lib.lib().vnqa_device_check(), "vnqa_device_check")


# ------------------------------------------------------------------------------------------ layout / packing
def act_dtype(act):
    """torch dtype of the tensor-core operand / trunk activation storage (include/vnqa_b200.h VNQA_ACT_*)."""
    return torch.float16 if act else torch.float32


def frames_to_nhwc(v_input, img_of, n_img, act=0):
    """(B, C, 10, 13, T) fp32 -> frame-image-major NHWC [n_img][130][C].  Two memory layouts are read in place: the
    contiguous one (T innermost) and the frame-blocked view the reference's drivers build from per-frame stem outputs
    (every (b, t) a contiguous (C, 10, 13) block); anything else is made contiguous first."""
    B, C, H, W, T = v_input.shape
    assert (H, W) == (10, 13) and v_input.dtype == torch.float32
    x = torch.empty(n_img, IMG_PIX, C, device=v_input.device, dtype=act_dtype(act))
    sb, sc, sh, sw, st = v_input.stride()
    if not v_input.is_contiguous() and (sc, sh, sw) == (IMG_PIX, W, 1) and (B == 1 or sb >= C * IMG_PIX) \
            and (T == 1 or st >= C * IMG_PIX):
        with _timed("hbm.k0_blocked", float(n_img * IMG_PIX * C * (4 + x.element_size()))):
            check(_lib.lib().vnqa_frames_blocked_to_nhwc(_p(v_input), _p(x), _p(img_of), B, C, T, sb, st, act,
                                                         _stream()), "frames_blocked_to_nhwc")
        return x
    v_input = v_input.contiguous()
    with _timed("hbm.k0", float(n_img * IMG_PIX * C * (4 + x.element_size()))):
        check(_lib.lib().vnqa_frames_to_nhwc(_p(v_input), _p(x), _p(img_of), B, C, T, act, _stream()), "frames_to_nhwc")
    return x


def to_half(x, scale=1.0):
    """fp32 -> fp16 (times a power-of-two scale): tensors entering the fp16 domain."""
    out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    check(_lib.lib().vnqa_to_half(_p(_f32c(x)), _p(out), x.numel(), float(scale), _stream()), "to_half")
    return out


def pack_conv_weight(w, transpose_flip=False, act=0):
    """(Cout, Cin, kh, kw) -> [taps][rows_pad][K] rounded to the operand type; transpose_flip builds the dgrad operand."""
    cout, cin = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3]
    rows = cin if transpose_flip else cout
    K = cout if transpose_flip else cin
    rows_pad = round_up(rows, 128)
    out = torch.empty(taps, rows_pad, K, device=w.device, dtype=act_dtype(act))
    check(_lib.lib().vnqa_pack_conv_weight(_p(_f32c(w)), _p(out), cout, cin, taps, rows_pad, int(transpose_flip), act,
                                           _stream()), "pack_conv_weight")
    return out


def pack_matrix(w, transpose=False, act=0):
    """Linear weight (N, K) as a 1-tap 'convolution' operand: [rows_pad][K] (or its transpose [K_pad][N])."""
    n, k = w.shape
    return pack_conv_weight(w.view(n, k, 1, 1), transpose_flip=transpose, act=act)[0]


def permute_fc_weight(w, C, inverse=False, n_pad=None, out=None):
    n = w.shape[0]
    n_pad = n if inverse else (n_pad or round_up(n, 128))
    if out is None:
        out = torch.empty(n if inverse else n_pad, C * IMG_PIX, device=w.device, dtype=torch.float32)
    assert out.is_contiguous() and out.numel() == (n if inverse else n_pad) * C * IMG_PIX
    check(_lib.lib().vnqa_permute_fc_weight(_p(_f32c(w)), _p(out), n, n_pad, C, int(inverse), _stream()),
          "permute_fc_weight")
    return out


# ------------------------------------------------------------------------------------------ tensor-core GEMMs
def conv_gemm(w_packed, x, *, taps, cin, cout, n_img, epi, out, flat=False, x_ld=0, splits=1, no_round=False,
              bias=None, scale=None, shift=None, img_frame=None, stat_sum=None, stat_sq=None, film=None,
              film_row=None, film_ld=0, alpha_off=0, res=None, z_out=None, add=None, dbias=None, act=0, out_f32=False,
              grad_unscale=1.0, img_hw=None, pool=False):
    """img_hw: (H, W) of the conv's images, None = the trunk's 10 x 13 (the frozen stem passes 40 x 52 / 20 x 26)."""
    assert x.dtype == act_dtype(act) and w_packed.dtype == act_dtype(act), (x.dtype, w_packed.dtype, act)
    a = ConvArgs()
    a.act_dtype, a.out_f32, a.grad_unscale = act, int(out_f32), float(grad_unscale)
    a.img_h, a.img_w = img_hw if img_hw is not None else (0, 0)
    a.pool = int(pool)
    a.w_packed = w_packed.data_ptr()
    a.x = x.data_ptr()
    a.taps, a.cin, a.cout, a.rows_pad, a.n_img = taps, cin, cout, round_up(cout, 128), n_img
    a.flat, a.x_ld, a.splits, a.epi, a.no_round = int(flat), x_ld, splits, epi, int(no_round)
    a.out = out.data_ptr()
    for name, t in (("bias", bias), ("scale", scale), ("shift", shift), ("img_frame", img_frame),
                    ("stat_sum", stat_sum), ("stat_sq", stat_sq), ("film", film), ("film_row", film_row),
                    ("res", res), ("z_out", z_out), ("add", add), ("dbias", dbias)):
        setattr(a, name, t.data_ptr() if t is not None else None)
    a.film_ld, a.alpha_off = film_ld, alpha_off
    rows = n_img if flat else n_img * (IMG_PIX if img_hw is None else img_hw[0] * img_hw[1])
    tag = f"conv_gemm.{'flat' if flat else str(taps) + 'tap'}.epi{epi}.k{cin * splits}.n{cout}"
    if img_hw is not None:
        tag += f".{img_hw[0]}x{img_hw[1]}"
    with _timed(tag, 2.0 * rows * cout * cin * splits * taps):
        check(_lib.lib().vnqa_conv_gemm(ctypes.byref(a), _stream()), "conv_gemm")
    return out


def nchw_to_nhwc_affine(x, scale=None, shift=None, act=0, out=None):
    """x (N, C, H, W) fp32 -> [N][H*W][C] in the operand type, times scale[c] plus shift[c] (csrc/stem.cu)."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    N, C, H, W = x.shape
    if out is None:
        out = torch.empty(N, H * W, C, device=x.device, dtype=act_dtype(act))
    with _timed("hbm.stem_in", x.numel() * 4 + out.numel() * out.element_size()):
        check(_lib.lib().vnqa_nchw_to_nhwc_affine(_p(x), _p(out), _p(scale), _p(shift), N, C, H * W, act, _stream()),
              "nchw_to_nhwc_affine")
    return out


def nhwc_to_nchw(x, hw, out=None):
    """x [N][H*W][C] (fp32 or half) -> (N, C, H, W) fp32."""
    assert x.is_contiguous() and x.dim() == 3 and x.dtype in (torch.float32, torch.float16)
    N, P, C = x.shape
    assert P == hw[0] * hw[1]
    if out is None:
        out = torch.empty(N, C, hw[0], hw[1], device=x.device, dtype=torch.float32)
    check(_lib.lib().vnqa_nhwc_to_nchw(_p(x), _p(out), N, C, P, int(x.dtype == torch.float16), _stream()),
          "nhwc_to_nchw")
    return out


_WS_BYTES = {}


def conv_wgrad(dy, x, dw, *, n_img, cout, cin, taps, flat=False, act=0, grad_unscale=1.0):
    """dw (fp32, PyTorch layout) is overwritten; deterministic reduction order (csrc/wgrad.cu)."""
    assert dy.dtype == act_dtype(act) and x.dtype == act_dtype(act) and dw.dtype == torch.float32 and dw.is_contiguous()
    rows = n_img if flat else n_img * IMG_PIX
    tag = f"conv_wgrad.{'flat' if flat else str(taps) + 'tap'}.k{cin}.n{cout}"
    key = (cout, cin, taps)
    if key not in _WS_BYTES:
        zb = ctypes.c_int64(0)
        total = _lib.lib().vnqa_conv_wgrad_ws_bytes(cout, cin, taps, ctypes.byref(zb))
        if total < 0:
            raise _lib.VnqaError(f"vnqa_conv_wgrad_ws_bytes rejected ({cout}, {cin}, {taps})")
        _WS_BYTES[key] = (int(total), int(zb.value))
    total, zero = _WS_BYTES[key]
    ws = torch.empty(total, device=dw.device, dtype=torch.uint8)
    ws[:zero].zero_()                       # turn counters (a few KB); the partial-sum area needs no initialisation
    with _timed(tag, 2.0 * rows * cout * cin * taps):
        check(_lib.lib().vnqa_conv_wgrad(_p(dy), _p(x), _p(dw), _p(ws), n_img, cout, cin, taps, int(flat), act,
                                         float(grad_unscale), _stream()), "conv_wgrad")
    return dw


# ------------------------------------------------------------------------------------------ small fp32 ops
def sgemm(a, b, *, M, N, K, a_strides, b_strides, out=None, bias=None, relu=False, accumulate=False, ksplit=1):
    """out[M][N] = A B^T (+bias, relu); A(m,k) = a[m*rs + k*cs], B(n,k) likewise."""
    if out is None:
        out = (torch.zeros if ksplit > 1 else torch.empty)(M, N, device=a.device, dtype=torch.float32)
    check(_lib.lib().vnqa_sgemm(_p(a), _p(b), _p(out), _p(bias), M, N, K, a_strides[0], a_strides[1], b_strides[0],
                                b_strides[1], out.stride(0), int(relu), int(accumulate), ksplit, _stream()), "sgemm")
    return out


def linear(x, w, bias=None, relu=False):
    """x (M,K) @ w (N,K)^T + bias."""
    M, K = x.shape
    N = w.shape[0]
    return sgemm(x, w, M=M, N=N, K=K, a_strides=(x.stride(0), x.stride(1)), b_strides=(w.stride(0), w.stride(1)),
                 bias=bias, relu=relu)


def linear_dgrad(dy, w):
    """dy (M,N) @ w (N,K) -> (M,K)."""
    M, N = dy.shape
    K = w.shape[1]
    return sgemm(dy, w, M=M, N=K, K=N, a_strides=(dy.stride(0), dy.stride(1)), b_strides=(w.stride(1), w.stride(0)))


def linear_wgrad(dy, x, ksplit=1, out=None):
    """dy (M,N)^T @ x (M,K) -> (N,K); with `out` the product is accumulated onto it."""
    M, N = dy.shape
    K = x.shape[1]
    return sgemm(dy, x, M=N, N=K, K=M, a_strides=(dy.stride(1), dy.stride(0)), b_strides=(x.stride(1), x.stride(0)),
                 ksplit=ksplit, out=out, accumulate=out is not None)


def colsum(x, out=None):
    rows, n = x.shape
    acc = out is not None
    if out is None:
        out = torch.empty(n, device=x.device, dtype=torch.float32)
    check(_lib.lib().vnqa_colsum(_p(x), rows, n, x.stride(0), _p(out), int(acc), _stream()), "colsum")
    return out


def bn_finalize_train(stat_sum, stat_sq, frame_ct, n_frames, gamma, beta, running_mean, running_var, momentum, eps):
    C = gamma.numel()
    dev = gamma.device
    scale, shift, mean, rstd = (torch.empty(n_frames, C, device=dev, dtype=torch.float32) for _ in range(4))
    check(_lib.lib().vnqa_bn_finalize_train(_p(stat_sum), _p(stat_sq), _p(frame_ct), n_frames, C, _p(gamma), _p(beta),
                                            _p(running_mean), _p(running_var), momentum, eps, _p(scale), _p(shift),
                                            _p(mean), _p(rstd), _stream()), "bn_finalize_train")
    return scale, shift, mean, rstd


def bn_eval_affine(gamma, beta, running_mean, running_var, eps):
    C = gamma.numel()
    scale = torch.empty(C, device=gamma.device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    check(_lib.lib().vnqa_bn_eval_affine(_p(gamma), _p(beta), _p(running_mean), _p(running_var), eps, C, _p(scale),
                                         _p(shift), _stream()), "bn_eval_affine")
    return scale, shift


def bn_apply(y, img_frame, scale, shift, act=0):
    n_img, _, C = y.shape
    x = torch.empty(y.shape, device=y.device, dtype=act_dtype(act))
    with _timed("hbm.bn_apply", float(y.numel() * (4 + x.element_size()))):
        check(_lib.lib().vnqa_bn_apply(_p(y), _p(x), _p(img_frame), _p(scale), _p(shift), n_img, C, act, _stream()),
              "bn_apply")
    return x


def bn_backward(dx, y, img_frame, frame_ct, n_frames, mean, rstd, gamma, dbias, dgamma, dbeta, act=0,
                grad_unscale=1.0):
    n_img, _, C = y.shape
    assert dx.dtype == act_dtype(act) and y.dtype == torch.float32
    s1 = torch.empty(n_frames, C, device=y.device, dtype=torch.float32)
    s2 = torch.empty_like(s1)
    dy = torch.empty(y.shape, device=y.device, dtype=act_dtype(act))
    check(_lib.lib().vnqa_bn_backward(_p(dx), _p(y), _p(img_frame), _p(frame_ct), n_frames, _p(mean), _p(rstd),
                                      _p(gamma), n_img, C, _p(s1), _p(s2), _p(dy), _p(dbias), _p(dgamma), _p(dbeta),
                                      act, float(grad_unscale), _stream()), "bn_backward")
    return dy


def film_backward(dout, z, film, film_row, alpha_off, dfilm, dbias, mask_relu=True, act=0, grad_unscale=1.0):
    n_img, _, C = z.shape
    assert dout.dtype == act_dtype(act) and z.dtype == act_dtype(act)
    dz = torch.empty_like(z)
    with _timed("hbm.film_bwd", 3.0 * z.numel() * z.element_size()):
        check(_lib.lib().vnqa_film_backward(_p(dout), _p(z), _p(film), _p(film_row), film.stride(0), alpha_off, n_img,
                                            C, _p(dz), _p(dfilm), _p(dbias), int(mask_relu), act, float(grad_unscale),
                                            _stream()), "film_backward")
    return dz


def splitk_reduce(partial, splits, rows, N, bias, dst_row, out):
    check(_lib.lib().vnqa_splitk_reduce(_p(partial), splits, rows, N, _p(bias), _p(dst_row), _p(out), _stream()),
          "splitk_reduce")
    return out


def gather_rows(src, src_row, rows, N, round_tf32=False):
    out = torch.empty(rows, N, device=src.device, dtype=torch.float32)
    check(_lib.lib().vnqa_gather_rows(_p(src), _p(src_row), rows, N, _p(out), int(round_tf32), _stream()), "gather_rows")
    return out


def gp_maxpool(y, img_of, B, T, tail):
    pooled = torch.empty(B, tail * IMG_PIX, device=y.device, dtype=torch.float32)
    arg = torch.empty(B, tail * IMG_PIX, device=y.device, dtype=torch.int32)
    check(_lib.lib().vnqa_gp_maxpool(_p(y), _p(img_of), B, T, tail, _p(pooled), _p(arg), _stream()), "gp_maxpool")
    return pooled, arg


def gp_maxpool_backward(dpooled, pooled, arg, B, tail, n_img):
    dy = torch.zeros(n_img, IMG_PIX, tail, device=pooled.device, dtype=torch.float32)
    check(_lib.lib().vnqa_gp_maxpool_backward(_p(dpooled), _p(pooled), _p(arg), B, tail, _p(dy), _stream()),
          "gp_maxpool_backward")
    return dy


# ------------------------------------------------------------------------------------------ question encoder
def embed_gates(q_input, q_lens32, embed, w_ih, b_ih, b_hh):
    B, L = q_input.shape
    E = embed.shape[1]
    H = w_ih.shape[0] // 4
    gx = torch.zeros(B, L, 4 * H, device=embed.device, dtype=torch.float32)
    check(_lib.lib().vnqa_embed_gates(_p(q_input), _p(q_lens32), _p(embed), _p(w_ih), _p(b_ih), _p(b_hh), B, L, E, H,
                                      _p(gx), _stream()), "embed_gates")
    return gx


def lstm_chain_forward(gx, q_lens32, whh_t, h_state, c_state, n_pass, save=False, want_states=False):
    B, L, H4 = gx.shape
    H = H4 // 4
    dev = gx.device
    h_pass = torch.empty(n_pass, B, H, device=dev, dtype=torch.float32)
    states = torch.zeros(n_pass, B, L, H, device=dev, dtype=torch.float32) if want_states else None
    sg = sc = sh = None
    if save:
        sg = torch.zeros(n_pass, B, L, H4, device=dev, dtype=torch.float32)
        sc = torch.zeros(n_pass, B, L, H, device=dev, dtype=torch.float32)
        sh = torch.zeros(n_pass, B, L, H, device=dev, dtype=torch.float32)
    check(_lib.lib().vnqa_lstm_chain_forward(_p(gx), _p(q_lens32), _p(whh_t), _p(h_state), _p(c_state), n_pass, B, L, H,
                                             _p(h_pass), _p(sg), _p(sc), _p(sh), _p(states), _stream()),
          "lstm_chain_forward")
    if want_states:
        return h_pass, sg, sc, sh, states
    return h_pass, sg, sc, sh


def lstm_chain_backward(dh_pass, q_lens32, whh, save_gates, save_c, c0, dh_steps=None):
    n_pass, B, L, H4 = save_gates.shape
    H = H4 // 4
    dgates = torch.zeros_like(save_gates)
    dgx = torch.empty(B, L, H4, device=whh.device, dtype=torch.float32)
    check(_lib.lib().vnqa_lstm_chain_backward(_p(dh_pass), _p(q_lens32), _p(whh), _p(save_gates), _p(save_c), _p(c0),
                                              n_pass, B, L, H, _p(dgates), _p(dgx), _p(dh_steps), _stream()),
          "lstm_chain_backward")
    return dgates, dgx


def embed_gather(q_input, q_lens32, embed):
    B, L = q_input.shape
    E = embed.shape[1]
    x = torch.empty(B, L, E, device=embed.device, dtype=torch.float32)
    check(_lib.lib().vnqa_embed_gather(_p(q_input), _p(q_lens32), _p(embed), B, L, E, _p(x), _stream()), "embed_gather")
    return x


def embed_backward(q_input, q_lens32, dx, dembed):
    B, L = q_input.shape
    check(_lib.lib().vnqa_embed_backward(_p(q_input), _p(q_lens32), _p(dx), B, L, dx.shape[-1], _p(dembed), _stream()),
          "embed_backward")
    return dembed


# ------------------------------------------------------------------------------------------ multi-hop decoder
def tmh_hops_forward(states, h_last, img_frame, img_clip, W, NB, enc_g, enc_b, w_att, b_att):
    n_pass, B, L, H = states.shape
    n_img = img_frame.numel()
    dev = states.device
    hvec = torch.empty(NB, n_img, H, device=dev, dtype=torch.float32)
    enc_out = torch.empty(n_img, H, device=dev, dtype=torch.float32)
    enc_stat = torch.empty(n_img, 2, device=dev, dtype=torch.float32)
    coefs = torch.empty(NB, n_img, W, device=dev, dtype=torch.float32)
    check(_lib.lib().vnqa_tmh_hops_forward(_p(states), _p(h_last), _p(img_frame), _p(img_clip), n_img, B, L, H, W, NB,
                                           _p(enc_g), _p(enc_b), _p(w_att), _p(b_att), _p(hvec), _p(enc_out),
                                           _p(enc_stat), _p(coefs), _stream()), "tmh_hops_forward")
    return hvec, enc_out, enc_stat, coefs


def tmh_hops_backward(states, h_last, img_frame, img_clip, q_lens32, W, NB, enc_g, w_att, hvec, enc_out, enc_stat,
                      coefs, dhvec, dw_att, db_att, denc_g, denc_b):
    n_pass, B, L, H = states.shape
    n_img = img_frame.numel()
    dstates = torch.zeros_like(states)
    check(_lib.lib().vnqa_tmh_hops_backward(_p(states), _p(h_last), _p(img_frame), _p(img_clip), _p(q_lens32), n_img, B,
                                            L, H, W, NB, _p(enc_g), _p(w_att), _p(hvec), _p(enc_out), _p(enc_stat),
                                            _p(coefs), _p(dhvec), _p(dstates), _p(dw_att), _p(db_att), _p(denc_g),
                                            _p(denc_b), _stream()), "tmh_hops_backward")
    return dstates


def ln_slice_forward(cond, gamma, beta, film_row, film, off, width):
    rows, N = cond.shape
    stat = torch.empty(rows, 2, device=cond.device, dtype=torch.float32)
    check(_lib.lib().vnqa_ln_slice_forward(_p(cond), rows, N, _p(gamma), _p(beta), _p(film_row), film.stride(0), off,
                                           width, _p(film), _p(stat), _stream()), "ln_slice_forward")
    return stat


def ln_slice_backward(cond, gamma, film_row, dfilm, off, width, stat, dgamma, dbeta):
    rows, N = cond.shape
    dcond = torch.empty_like(cond)
    check(_lib.lib().vnqa_ln_slice_backward(_p(cond), rows, N, _p(gamma), _p(film_row), dfilm.stride(0), off, width,
                                            _p(dfilm), _p(stat), _p(dcond), _p(dgamma), _p(dbeta), _stream()),
          "ln_slice_backward")
    return dcond


# ------------------------------------------------------------------------------------------ attention tail
def attn_tail_forward(feats, v_lens32, n_proc, w1, b1, wh, bh, wih_t, whh_t, b_ih, b_hh, save=False):
    B, T, A = feats.shape
    dev = feats.device
    hs = torch.empty(B, T, A, device=dev, dtype=torch.float32)
    coefs = ctxt = gates = cs = None
    if save:
        coefs = torch.empty(B, T, T, device=dev, dtype=torch.float32)
        ctxt = torch.empty(B, T, A, device=dev, dtype=torch.float32)
        gates = torch.empty(B, T, 4 * A, device=dev, dtype=torch.float32)
        cs = torch.empty(B, T, A, device=dev, dtype=torch.float32)
    check(_lib.lib().vnqa_attn_tail_forward(_p(feats), _p(v_lens32), n_proc, B, T, A, _p(w1), _p(b1), _p(wh), _p(bh),
                                            _p(wih_t), _p(whh_t), _p(b_ih), _p(b_hh), _p(hs), _p(coefs), _p(ctxt),
                                            _p(gates), _p(cs), _stream()), "attn_tail_forward")
    return hs, coefs, ctxt, gates, cs


def attn_tail_backward(feats, v_lens32, n_proc, w1, wh, wih, whh, hs, coefs, gates, cs, dhs, dw1, db1, dwh, dbh):
    B, T, A = feats.shape
    dfeats = torch.empty_like(feats)
    dgates = torch.empty(B, T, 4 * A, device=feats.device, dtype=torch.float32)
    check(_lib.lib().vnqa_attn_tail_backward(_p(feats), _p(v_lens32), n_proc, B, T, A, _p(w1), _p(wh), _p(wih), _p(whh),
                                             _p(hs), _p(coefs), _p(gates), _p(cs), _p(dhs), _p(dfeats), _p(dgates),
                                             _p(dw1), _p(db1), _p(dwh), _p(dbh), _stream()), "attn_tail_backward")
    return dfeats, dgates
