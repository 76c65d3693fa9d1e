"""GPU parity of the frozen stem through the C ABI: the layout kernels of csrc/stem.cu, the
40 x 52 / 20 x 26 geometries and the BIAS / BIAS_AFFINE_RELU(+pool) epilogues of csrc/conv_gemm.cu, and the drop-in
``ObjDetectCNN`` (models/obj_detector.py:69-86) against (i) the golden vector produced by the unmodified reference
module, (ii) the CPU oracle that emulates the operand rounding, (iii) strict-fp32 cuDNN at the drivers' full size
(64 images of 128 x 40 x 52, 512 filters, eval/utils.py:42-51).

Tolerances.  Operands carry a 10-bit mantissa (fp16 / tf32), accumulation is fp32: a single convolution on operands
that are already representable differs from strict fp32 only by the accumulation order (<= 5e-5 of the output scale).
The whole stem stores five intermediate maps in operand precision; measured against the reference's fp64 run the CPU
emulation of exactly that rounding is 7e-4 in relative L2 (oracle/stem_oracle.py), so the end-to-end bounds are
rel-L2 <= 2e-3 and max <= 4e-3 of the output scale.  Against the emulating oracle the bound is only a little tighter,
<= 8e-4 / 1.5e-3 (measured 3.8e-4 / 6.5e-4): a per-layer accumulation difference of 3e-5 flips about 4 % of the stored
intermediates to the neighbouring 10-bit value (7e-4 apart), 1.4e-4 per stored map, five maps.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import film_oracle as O
from oracle import stem_oracle as S

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _errs(got, ref):
    d = got.double() - ref.double()
    return float(d.abs().max() / ref.double().abs().max()), float(d.norm() / ref.double().norm())


def _operand(x, act):
    """(kernel operand, same values as fp32)"""
    if act:
        h = x.half()
        return h, h.float()
    x = O.tf32_round(x)
    return x, x


@pytest.mark.parametrize("act", [1, 0], ids=["fp16", "tf32"])
def test_layout_kernels(act):
    from videonavqa_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    for (N, C, H, W) in [(3, 128, 40, 52), (2, 96, 8, 52), (1, 34, 3, 7)]:
        x = torch.randn(N, C, H, W, device="cuda", generator=g)
        s = torch.randn(C, device="cuda", generator=g)
        t = torch.randn(C, device="cuda", generator=g)
        got = ops.nchw_to_nhwc_affine(x, s, t, act=act)
        assert got.shape == (N, H * W, C) and got.dtype == (torch.float16 if act else torch.float32)
        exact = (x.double() * s.double().view(1, C, 1, 1) + t.double().view(1, C, 1, 1)).permute(0, 2, 3, 1).reshape(N, H * W, C)
        want = exact.float().half().float() if act else O.tf32_round(exact.float())
        d = (got.float() - want).abs()
        ulp = want.abs().clamp_min(2.0 ** -14) * 2.0 ** -10          # one unit of a 10-bit mantissa
        assert bool((d <= ulp).all()), float((d / ulp).max())         # at most a double-rounding tie
        assert float((d == 0).float().mean()) > 0.999
        plain = ops.nchw_to_nhwc_affine(x, None, None, act=act)
        wantp = x.permute(0, 2, 3, 1).reshape(N, H * W, C)
        assert torch.equal(plain.float(), wantp.half().float() if act else O.tf32_round(wantp))
        # and back: a pure permutation, exact for fp32 and for half inputs
        back = ops.nhwc_to_nchw(plain, (H, W))
        assert back.dtype == torch.float32 and torch.equal(back, plain.float().view(N, H, W, C).permute(0, 3, 1, 2))
        if not act:
            assert torch.equal(ops.nhwc_to_nchw(wantp.contiguous(), (H, W)), x)


GEOM_CASES = [
    # N, cin, cout, H, W
    (64, 128, 512, 40, 52),      # conv11 at the drivers' size (CTA pairs)
    (7, 512, 512, 20, 26),       # conv21/22 geometry, N not a multiple of the 4-image tile
    (5, 256, 256, 40, 52),       # 2-image tiles, odd N: the last tile's second image is TMA zero fill
    (3, 64, 32, 16, 52),         # single-CTA kernels (cout < 256), one partial m-tile
    (9, 128, 128, 10, 13),       # the trunk's own geometry through the new epilogues
]


@pytest.mark.parametrize("act", [1, 0], ids=["fp16", "tf32"])
@pytest.mark.parametrize("case", GEOM_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_geometries_and_stem_epilogues(case, act):
    """conv3x3 over H x W images with EPI_BIAS, EPI_BIAS_AFFINE_RELU and the fused 2 x 2 max-pool, full tensors,
    against strict-fp32 cuDNN on the same (representable) operands."""
    from videonavqa_b200 import _lib, ops
    N, cin, cout, H, W = case
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + cin + H)
    xk, xf = _operand(torch.randn(N, H * W, cin, device="cuda", generator=g).clamp_min_(0), act)
    wk, wf = _operand(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5, act)
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1
    scale = torch.randn(cout, device="cuda", generator=g)              # negative scales included
    shift = torch.randn(cout, device="cuda", generator=g) * 0.2
    wp = ops.pack_conv_weight(wf, act=act)
    x_nchw = xf.view(N, H, W, cin).permute(0, 3, 1, 2)
    ref = F.conv2d(x_nchw, wf, bias, padding=1)                        # (N, cout, H, W)
    hw = None if (H, W) == (10, 13) else (H, W)

    def nhwc(t):
        return t.permute(0, 2, 3, 1).reshape(N, -1, cout)

    out = torch.full((N, H * W, cout), float("nan"), device="cuda", dtype=xk.dtype)
    ops.conv_gemm(wp, xk, taps=9, cin=cin, cout=cout, n_img=N, epi=_lib.EPI_BIAS, out=out, bias=bias, act=act,
                  img_hw=hw, no_round=True)
    emax, el2 = _errs(out.float(), nhwc(ref))
    bound = (1.2e-3, 4e-4) if act else (5e-5, 3e-5)     # fp16 OUTPUT rounding (2^-11 relative) dominates for act = 1
    assert emax <= bound[0] and el2 <= bound[1], ("BIAS", emax, el2)

    want = F.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    outf = torch.full((N, H * W, cout), float("nan"), device="cuda", dtype=torch.float32)
    ops.conv_gemm(wp, xk, taps=9, cin=cin, cout=cout, n_img=N, epi=_lib.EPI_BIAS_AFFINE_RELU, out=outf, bias=bias,
                  scale=scale, shift=shift, act=act, img_hw=hw, out_f32=bool(act), no_round=True)
    emax, el2 = _errs(outf, nhwc(want))
    assert emax <= 5e-5 and el2 <= 3e-5, ("AFFINE_RELU", emax, el2)

    if W in (26, 52):
        wantp = F.max_pool2d(want, 2, 2)
        outp = torch.full((N, (H // 2) * (W // 2), cout), float("nan"), device="cuda", dtype=torch.float32)
        ops.conv_gemm(wp, xk, taps=9, cin=cin, cout=cout, n_img=N, epi=_lib.EPI_BIAS_AFFINE_RELU, out=outp, bias=bias,
                      scale=scale, shift=shift, act=act, img_hw=hw, pool=True, out_f32=bool(act), no_round=True)
        emax, el2 = _errs(outp, nhwc(wantp))
        assert emax <= 5e-5 and el2 <= 3e-5, ("POOL", emax, el2)


def _module(filters, operands, P=None):
    from videonavqa_b200.models import ObjDetectCNN
    m = ObjDetectCNN(nb_classes=27, num_filters=filters, tail_hidden_dim=64, tail_dropout_p=0, logits=True,
                     pretrained_features=True)
    if P is not None:
        r = m.load_state_dict(P, strict=False)
        assert not r.unexpected_keys
    return m.cuda().eval().set_operands(operands)


@pytest.mark.parametrize("operands", ["fp16", "tf32"])
def test_stem_matches_reference_golden(operands, golden_dir):
    g = np.load(golden_dir + "/stem_small.npz")
    n, h, w, filters, seed = (int(v) for v in g["case"])
    P, x = S.synthetic_stem_case(n, h, w, filters, seed)
    m = _module(filters, operands, P)
    y = m(x.cuda())
    assert y.shape == g["y"].shape and y.dtype == torch.float32 and not y.requires_grad
    emax, el2 = _errs(y.cpu(), torch.from_numpy(g["y64"]))
    assert emax <= 4e-3 and el2 <= 2e-3, ("vs the reference module (fp64)", emax, el2)
    emu = S.stem_forward(P, x, emulate=operands)
    emax, el2 = _errs(y.cpu(), emu)
    assert emax <= 1.5e-3 and el2 <= 8e-4, ("vs the emulating oracle", emax, el2)
    # NHWC fast path: same values in the kernels' own layout / operand type
    z = m.features(x.cuda(), nhwc=True)
    assert z.shape == (n, (h // 4) * (w // 4), filters) and z.dtype == m.precision.dtype
    zz = z.float().view(n, h // 4, w // 4, filters).permute(0, 3, 1, 2)
    assert _errs(zz, y)[0] <= 1e-3


@pytest.mark.parametrize("operands", ["fp16", "tf32"])
def test_stem_full_size_vs_strict_fp32(operands):
    """The drivers' configuration: 512 filters, 64 frames of (128, 40, 52) per call (eval/utils.py:42-51,
    q_and_v_eval.py:104-108), trained-looking BatchNorm statistics; reference = the same network in strict-fp32
    PyTorch on the GPU (cuDNN, TF32 off)."""
    torch.manual_seed(5)
    m = _module(512, operands)
    with torch.no_grad():
        for bn in (m.bn_input, m.bn1, m.bn2, m.bn3):
            bn.weight.normal_(0.8, 0.4)
            bn.bias.normal_(0.0, 0.3)
            bn.running_mean.normal_(0.0, 0.4)
            bn.running_var.uniform_(0.5, 2.0)
        for c in (m.conv11, m.conv12, m.conv21, m.conv22, m.conv31, m.conv32):
            c.bias.normal_(0.0, 0.1)
    m._weight_cache.invalidate()
    x = torch.randn(64, 128, 40, 52, device="cuda").clamp_min_(0) * 2
    y = m(x)
    P = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = S.stem_forward(P, x)
    assert y.shape == (64, 512, 10, 13)
    emax, el2 = _errs(y, ref)
    assert emax <= 4e-3 and el2 <= 2e-3, (emax, el2)
    # one call per frame, as the drivers do, is bit-identical to itself (no atomics on this path)
    assert torch.equal(m(x), y)
    # smaller, odd batch through the same weights (the drivers' last test batch is zero-padded, but val/test loops
    # may feed any N): rows must not depend on their batch neighbours
    y5 = m(x[:5].contiguous())
    assert torch.equal(y5, y[:5])


def test_stem_refuses_what_it_does_not_implement():
    from videonavqa_b200.models import ObjDetectCNN
    m = ObjDetectCNN(27, num_filters=32, pretrained_features=True).cuda()
    x = torch.zeros(1, 128, 40, 52, device="cuda")
    with pytest.raises(NotImplementedError):
        m.train()(x)
    with pytest.raises(NotImplementedError):
        ObjDetectCNN(27, num_filters=32, pretrained_features=False).cuda().eval()(x)
    with pytest.raises(ValueError):
        m.eval()(torch.zeros(1, 128, 40, 50, device="cuda"))
    with pytest.raises(Exception):
        m.eval()(torch.zeros(1, 128, 40, 52))            # CPU tensor: no fallback
