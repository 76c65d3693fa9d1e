"""One launch of every hot kernel at the benchmark shape (2240 frame images, C = 1024), for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'conv_gemm_kernel|wgrad_kernel|film_bwd|bn_|frames_to' \
        -o gpurun_out/r02_prof python tools/ncu_targets.py [--operands fp16|tf32]

Random operands; nothing is checked here (parity lives in tests/), the point is one representative launch per kernel.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videonavqa_b200 import ops  # noqa: E402
from videonavqa_b200._lib import EPI_BIAS_RELU, EPI_DGRAD_RES, EPI_FILM, EPI_PLAIN  # noqa: E402


def stem(a):
    """ncu -k regex:'conv_gemm_kernel|nchw_to_nhwc|nhwc_to_nchw' ... python tools/ncu_targets.py --stem"""
    from videonavqa_b200.models import ObjDetectCNN
    torch.manual_seed(0)
    m = ObjDetectCNN(27, num_filters=512, tail_hidden_dim=1024, tail_dropout_p=0, logits=True,
                     pretrained_features=True).cuda().eval().set_operands(a.operands)
    x = torch.randn(64, 128, 40, 52, device="cuda").clamp_min_(0)
    m(x)                                       # packs the weights (not a target)
    torch.cuda.synchronize()
    for _ in range(a.reps):
        m(x)
        torch.cuda.synchronize()
    print("ncu stem target done", a.operands)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--operands", choices=["fp16", "tf32"], default="fp16")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--stem", action="store_true", help="one call of the frozen stem (64 frames) instead of the trunk kernels")
    a = ap.parse_args()
    if a.stem:
        return stem(a)
    act = 1 if a.operands == "fp16" else 0
    dt = ops.act_dtype(act)
    dev = torch.device("cuda")
    n_img, C, B, T = 2240, 1024, 64, 35
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)     # noqa: E731
    x = rnd(n_img, 130, C).clamp_min_(0).to(dt)
    dy = (rnd(n_img, 130, C) * 0.05).to(dt)
    w3 = ops.pack_conv_weight(rnd(C, C, 3, 3) * 0.01, act=act)
    w3t = ops.pack_conv_weight(rnd(C, C, 3, 3) * 0.01, True, act)
    w1 = ops.pack_conv_weight(rnd(C, C, 1, 1) * 0.03, act=act)
    bias = rnd(C) * 0.1
    film = rnd(T * B, 2 * C * 5).clamp_min_(0)
    film_row = torch.arange(n_img, device=dev, dtype=torch.int32)
    img_frame = (torch.arange(n_img, device=dev, dtype=torch.int32) // B)
    v = rnd(B, 512, 10, 13, T).clamp_min_(0)
    img_of = torch.arange(B * T, device=dev, dtype=torch.int32).view(T, B).t().contiguous().view(-1)
    out, z = torch.empty_like(x), torch.empty_like(x)
    dw3 = torch.empty(C, C, 3, 3, device=dev)
    dw1 = torch.empty(C, C, 1, 1, device=dev)
    dfilm = torch.zeros_like(film)
    db = torch.zeros(C, device=dev)
    y0 = rnd(n_img, 130, C).clamp_min_(0)
    scale, shift = rnd(T, C).abs() + 0.5, rnd(T, C)
    for _ in range(a.reps):
        ops.frames_to_nhwc(v, img_of, n_img, act)
        ops.conv_gemm(w1, x, taps=1, cin=C, cout=C, n_img=n_img, epi=EPI_BIAS_RELU, out=out, bias=bias, act=act)
        ops.conv_gemm(w3, x, taps=9, cin=C, cout=C, n_img=n_img, epi=EPI_FILM, out=out, bias=bias, film=film,
                      film_row=film_row, film_ld=film.stride(0), alpha_off=0, res=x, z_out=z, act=act)
        ops.conv_gemm(w3t, dy, taps=9, cin=C, cout=C, n_img=n_img, epi=EPI_DGRAD_RES, out=out, add=dy, res=x, dbias=db,
                      act=act)
        ops.conv_gemm(w1, dy, taps=1, cin=C, cout=C, n_img=n_img, epi=EPI_PLAIN, out=out, act=act)
        ops.conv_wgrad(dy, x, dw3, n_img=n_img, cout=C, cin=C, taps=9, act=act)
        ops.conv_wgrad(dy, x, dw1, n_img=n_img, cout=C, cin=C, taps=1, act=act)
        ops.film_backward(dy, z, film, film_row, 0, dfilm, db, act=act)
        ops.bn_apply(y0, img_frame, scale, shift, act)
        torch.cuda.synchronize()
    print("ncu targets done", a.operands)


if __name__ == "__main__":
    main()
