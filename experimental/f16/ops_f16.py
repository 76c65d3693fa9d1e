"""Torch-tensor wrappers over the CANDIDATE fp16 C ABI (experimental/f16/vnqa_f16.h).  Not imported by the package;
not yet run on hardware.  Mirrors videonavqa_b200/ops.py: PyTorch supplies device memory and the current stream only.
"""
from __future__ import annotations

import ctypes
import os
import sys
from ctypes import POINTER, c_float, c_int32, c_int64, c_void_p

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from videonavqa_b200 import _lib as _main          # noqa: E402
from videonavqa_b200 import ops as _ops            # noqa: E402  (launch counter / timing hooks are shared)

LIB_PATH = os.path.join(_ROOT, "build", "libvnqa_f16_candidate.so")
IMG_PIX = 130
HALF = torch.float16


class ConvArgsH(ctypes.Structure):
    """Mirror of vnqa_conv_args_h."""
    _fields_ = [
        ("w_packed", c_void_p), ("x", c_void_p),
        ("taps", c_int32), ("cin", c_int32), ("cout", c_int32), ("rows_pad", c_int32), ("n_img", c_int32),
        ("flat", c_int32), ("x_ld", c_int32), ("splits", c_int32), ("epi", c_int32),
        ("out", c_void_p), ("bias", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
        ("img_frame", c_void_p), ("stat_sum", c_void_p), ("stat_sq", c_void_p),
        ("film", c_void_p), ("film_row", c_void_p), ("film_ld", c_int32), ("alpha_off", c_int32),
        ("res", c_void_p), ("z_out", c_void_p), ("add", c_void_p), ("dbias", c_void_p),
        ("grad_unscale", c_float), ("out_f32", c_int32),
    ]


P, I32, I64, F32 = c_void_p, c_int32, c_int64, c_float
SIGNATURES = {
    "vnqa_pack_conv_weight_h": [P, P, I32, I32, I32, I32, I32, P],
    "vnqa_to_half": [P, P, I64, F32, P],
    "vnqa_conv_gemm_h": [POINTER(ConvArgsH), P],
    "vnqa_conv_wgrad_h": [P, P, P, P, I32, I32, I32, I32, I32, F32, P],
    "vnqa_frames_to_nhwc_h": [P, P, P, I32, I32, I32, P],
    "vnqa_bn_apply_h": [P, P, P, P, P, I32, I32, P],
    "vnqa_bn_backward_h": [P, P, P, P, I32, P, P, P, I32, I32, P, P, P, P, P, P, F32, P],
    "vnqa_film_backward_h": [P, P, P, P, I32, I32, I32, I32, P, P, P, I32, F32, P],
}
_handle = None


def lib():
    global _handle
    if _handle is None:
        _main.lib()      # the candidate library resolves set_error / the tensor-map encoder from the main one
        if not os.path.exists(LIB_PATH):
            raise _main.VnqaError(f"{LIB_PATH} is missing: make -C experimental/f16")
        h = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(h, name)
            fn.argtypes = argtypes
            fn.restype = c_int32
        _handle = h
    return _handle


_p, _stream, check, round_up, _timed = _ops._p, _ops._stream, _ops.check, _ops.round_up, _ops._timed


def _h16c(t):
    assert t.dtype == HALF and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t


def frames_to_nhwc(v_input, img_of, n_img):
    B, C, H, W, T = v_input.shape
    assert (H, W) == (10, 13)
    x = torch.empty(n_img, IMG_PIX, C, device=v_input.device, dtype=HALF)
    check(lib().vnqa_frames_to_nhwc_h(_p(_ops._f32c(v_input)), _p(x), _p(img_of), B, C, T, _stream()), "frames_to_nhwc_h")
    return x


def pack_conv_weight(w, transpose_flip=False):
    """(Cout, Cin, kh, kw) fp32 -> fp16 [taps][rows_pad][K]; transpose_flip builds the dgrad operand."""
    cout, cin = w.shape[0], w.shape[1]
    taps = w.shape[2] * w.shape[3]
    rows = cin if transpose_flip else cout
    K = cout if transpose_flip else cin
    rows_pad = round_up(rows, 128)
    out = torch.empty(taps, rows_pad, K, device=w.device, dtype=HALF)
    check(lib().vnqa_pack_conv_weight_h(_p(_ops._f32c(w)), _p(out), cout, cin, taps, rows_pad, int(transpose_flip),
                                        _stream()), "pack_conv_weight_h")
    return out


def pack_matrix(w, transpose=False):
    n, k = w.shape
    return pack_conv_weight(w.view(n, k, 1, 1), transpose_flip=transpose)[0]


def to_half(t, scale=1.0):
    out = torch.empty(t.shape, device=t.device, dtype=HALF)
    check(lib().vnqa_to_half(_p(_ops._f32c(t)), _p(out), t.numel(), float(scale), _stream()), "to_half")
    return out


def conv_gemm(w_packed, x, *, taps, cin, cout, n_img, epi, out, flat=False, x_ld=0, splits=1, bias=None, scale=None,
              shift=None, img_frame=None, stat_sum=None, stat_sq=None, film=None, film_row=None, film_ld=0,
              alpha_off=0, res=None, z_out=None, add=None, dbias=None, grad_unscale=1.0, out_f32=False):
    _h16c(w_packed); _h16c(x)
    a = ConvArgsH()
    a.w_packed = w_packed.data_ptr()
    a.x = x.data_ptr()
    a.taps, a.cin, a.cout, a.rows_pad, a.n_img = taps, cin, cout, round_up(cout, 128), n_img
    a.flat, a.x_ld, a.splits, a.epi = int(flat), x_ld, splits, epi
    a.out = out.data_ptr()
    for name, t in (("bias", bias), ("scale", scale), ("shift", shift), ("img_frame", img_frame),
                    ("stat_sum", stat_sum), ("stat_sq", stat_sq), ("film", film), ("film_row", film_row),
                    ("res", res), ("z_out", z_out), ("add", add), ("dbias", dbias)):
        setattr(a, name, t.data_ptr() if t is not None else None)
    a.film_ld, a.alpha_off = film_ld, alpha_off
    a.grad_unscale, a.out_f32 = float(grad_unscale), int(out_f32)
    f32_out = out_f32 or epi in (_main.EPI_BIAS_RELU_STATS, _main.EPI_PARTIAL)
    assert out.dtype == (torch.float32 if f32_out else HALF), (epi, out.dtype)
    for t in (res, z_out, add):
        assert t is None or t.dtype == HALF
    rows = n_img if flat else n_img * IMG_PIX
    tag = f"conv_gemm.{'flat' if flat else str(taps) + 'tap'}.epi{epi}.k{cin * splits}.n{cout}"
    with _timed(tag, 2.0 * rows * cout * cin * splits * taps):
        check(lib().vnqa_conv_gemm_h(ctypes.byref(a), _stream()), "conv_gemm")
    return out


def conv_wgrad(dy, x, dw, *, n_img, cout, cin, taps, flat=False, grad_unscale=1.0):
    _h16c(dy); _h16c(x)
    rows = n_img if flat else n_img * IMG_PIX
    tag = f"conv_wgrad.{'flat' if flat else str(taps) + 'tap'}.k{cin}.n{cout}"
    ws = torch.zeros(taps, cout, cin, device=dw.device, dtype=torch.float32) if taps != 1 else None
    with _timed(tag, 2.0 * rows * cout * cin * taps):
        check(lib().vnqa_conv_wgrad_h(_p(dy), _p(x), _p(dw), _p(ws), n_img, cout, cin, taps, int(flat),
                                      float(grad_unscale), _stream()), "conv_wgrad")
    return dw


def bn_apply(y, img_frame, scale, shift):
    n_img, _, C = y.shape
    x = torch.empty(y.shape, device=y.device, dtype=HALF)
    with _timed("hbm.bn_apply", y.numel() * 6.0):
        check(lib().vnqa_bn_apply_h(_p(_ops._f32c(y)), _p(x), _p(img_frame), _p(scale), _p(shift), n_img, C, _stream()),
              "bn_apply")
    return x


def bn_backward(dx, y, img_frame, frame_ct, n_frames, mean, rstd, gamma, dbias, dgamma, dbeta, grad_unscale=1.0):
    n_img, _, C = y.shape
    s1 = torch.empty(n_frames, C, device=y.device, dtype=torch.float32)
    s2 = torch.empty_like(s1)
    dy = torch.empty(y.shape, device=y.device, dtype=HALF)
    check(lib().vnqa_bn_backward_h(_p(_h16c(dx)), _p(y), _p(img_frame), _p(frame_ct), n_frames, _p(mean), _p(rstd),
                                   _p(gamma), n_img, C, _p(s1), _p(s2), _p(dy), _p(dbias), _p(dgamma), _p(dbeta),
                                   float(grad_unscale), _stream()), "bn_backward")
    return dy


def film_backward(dout, z, film, film_row, alpha_off, dfilm, dbias, mask_relu=True, grad_unscale=1.0):
    n_img, _, C = z.shape
    dz = torch.empty_like(z)
    with _timed("hbm.film_bwd", 3.0 * z.numel() * 2):
        check(lib().vnqa_film_backward_h(_p(_h16c(dout)), _p(_h16c(z)), _p(film), _p(film_row), film.stride(0), alpha_off,
                                         n_img, C, _p(dz), _p(dfilm), _p(dbias), int(mask_relu), float(grad_unscale),
                                         _stream()), "film_backward")
    return dz
