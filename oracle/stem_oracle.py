"""CPU oracle of the frozen stem - TEST INFRASTRUCTURE ONLY.

Restatement of ``ObjDetectCNN.forward`` with ``pretrained_features=True`` in ``eval()`` mode
(models/obj_detector.py:69-86), the way the drivers run it (eval/utils.py:42-51, eval/q_and_v_eval.py:104-110), as plain
functional PyTorch on the CPU.  Pinned against the unmodified reference module by ``oracle/gen_golden_stem.py`` ->
``tests/golden/stem_small.npz`` (tests/test_oracle_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this module; the product path (videonavqa_b200.models.obj_detector) never does.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5     # nn.BatchNorm2d default, obj_detector.py:21,25,30,35


def bn_eval(x, P, name):
    """BatchNorm2d in eval mode: (x - running_mean) / sqrt(running_var + eps) * weight + bias."""
    sh = (1, -1, 1, 1)
    inv = torch.rsqrt(P[name + ".running_var"].to(x.dtype) + EPS)
    return (x - P[name + ".running_mean"].to(x.dtype).view(sh)) * (inv * P[name + ".weight"].to(x.dtype)).view(sh) \
        + P[name + ".bias"].to(x.dtype).view(sh)


def conv(x, P, name):
    return F.conv2d(x, P[name + ".weight"].to(x.dtype), P[name + ".bias"].to(x.dtype), padding=1)


def stem_forward(P, inputs, dtype=torch.float32, emulate=None):
    """inputs (N, 128, H, W) -> (N, F, H/4, W/4).  P: state_dict-keyed tensors.

    emulate: None, or 'fp16' / 'tf32' - round every tensor that the CUDA path stores in operand precision (the
    BatchNorm'ed input, the packed weights, each convolution output that feeds the next convolution) the way the
    kernels do, so that a comparison isolates the accumulation order from the documented operand rounding."""
    rnd = _rounder(emulate)
    Pw = dict(P)
    if emulate:
        for k in list(Pw):
            if k.startswith("conv") and k.endswith(".weight"):
                Pw[k] = rnd(Pw[k].float())
    x = inputs.to(dtype)
    h = rnd(bn_eval(x, Pw, "bn_input"))                                   # obj_detector.py:70
    h = conv(rnd(conv(h, Pw, "conv11")), Pw, "conv12")                    # :72
    h = rnd(F.max_pool2d(F.relu(bn_eval(h, Pw, "bn1")), 2, 2))            # :73-75
    h = conv(rnd(conv(h, Pw, "conv21")), Pw, "conv22")                    # :77
    h = rnd(F.max_pool2d(F.relu(bn_eval(h, Pw, "bn2")), 2, 2))            # :78-80
    h = conv(rnd(conv(h, Pw, "conv31")), Pw, "conv32")                    # :82
    return F.relu(bn_eval(h, Pw, "bn3"))                                  # :83-86 (pretrained_features -> return)


def _rounder(emulate):
    if emulate is None:
        return lambda t: t
    if emulate == "fp16":
        return lambda t: t.to(torch.float16).to(t.dtype)
    if emulate == "tf32":
        def r(t):
            # round to nearest (ties away, as cvt.rna.tf32.f32) onto a 10-bit mantissa
            a = t.float().contiguous().view(torch.int32)
            a = (a + 0x1000) & ~0x1FFF
            return a.view(torch.float32).to(t.dtype)
        return r
    raise ValueError(emulate)


def synthetic_stem_case(n=3, h=16, w=52, filters=32, seed=21):
    """Seeded input + parameters with NON-trivial BatchNorm statistics and affines (negative gammas included: the
    affine must be applied before the max-pool) and non-zero biases.  Returns (P, inputs)."""
    rng = np.random.default_rng(seed)

    def t(*shape, scale=1.0, shift=0.0):
        return torch.from_numpy((rng.standard_normal(shape) * scale + shift).astype(np.float32))

    P = {}
    for name, c in (("bn_input", 128), ("bn1", filters), ("bn2", filters), ("bn3", filters)):
        P[name + ".weight"] = t(c, scale=0.5, shift=0.8)
        P[name + ".bias"] = t(c, scale=0.3)
        P[name + ".running_mean"] = t(c, scale=0.4)
        P[name + ".running_var"] = torch.from_numpy(rng.uniform(0.5, 2.0, c).astype(np.float32))
    cin = 128
    for name in ("conv11", "conv12", "conv21", "conv22", "conv31", "conv32"):
        P[name + ".weight"] = t(filters, cin, 3, 3, scale=(2.0 / (9 * cin)) ** 0.5)
        P[name + ".bias"] = t(filters, scale=0.1)
        cin = filters
    inputs = torch.from_numpy(np.maximum(rng.standard_normal((n, 128, h, w)), 0.0).astype(np.float32) * 2.0)
    return P, inputs
