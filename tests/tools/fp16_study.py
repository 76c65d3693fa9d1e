"""CPU study for the next round: would fp16 operand storage (same 10-bit mantissa as tf32, twice the tcgen05 rate,
half the HBM bytes) keep the logits inside the 1e-3 tolerance, and how small do the backward signals get?

Uses the oracle's straight-through rounding hooks with the rounding function swapped (test/tool code, not product).
    python tests/tools/fp16_study.py [--clips 2] [--frames 4]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import film_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--C", type=int, default=1024)
    ap.add_argument("--nb", type=int, default=5)
    a = ap.parse_args()
    B = a.clips
    p = O.init_params("attn", c_in=512, C=a.C, nb=a.nb, hidden=128, at_hidden=128, T=35, classes=70, vocab=134, seed=3)
    v, q, _, q_lens, labels = O.synthetic_batch(B, seed=5)
    v_lens = torch.tensor(sorted([max(1, a.frames - i) for i in range(B)], reverse=True))

    stats = []

    def run(mode, grad_scale=None):
        orig_q = O._q
        if grad_scale is not None:
            # also round the backward signal that crosses each rounded activation, scaled by 2^k as a loss-scale would
            def q_hook(t):
                r = orig_q(t)
                if r.requires_grad and t.dim() == 4:
                    def hook(g):
                        nz = g[g != 0].abs()
                        stats.append((tuple(g.shape), float(nz.max()) if nz.numel() else 0.0,
                                      float((nz * grad_scale < 6.1e-5).float().mean()) if nz.numel() else 0.0,
                                      float((nz * grad_scale < 6e-8).float().mean()) if nz.numel() else 0.0))
                        return (g * grad_scale).half().float() / grad_scale
                    r.register_hook(hook)
                return r
            O._q = q_hook
        if mode == "fp32":
            ctx = None
        else:
            ctx = O.emulate_tf32()
            if mode in ("fp16", "fp16x"):
                pass
            if mode == "fp16":
                O._ROUND = lambda t: t.half().float()
            elif mode == "bf16":
                O._ROUND = lambda t: t.bfloat16().float()
        orig = O._ROUND
        if ctx:
            ctx.__enter__()
        try:
            lo, loss, g, _ = O.loss_and_grads("attn", {k: t.clone() for k, t in p.items()}, O.new_bn_state(a.C),
                                              O.new_film_state(B, 128), v, q, v_lens, q_lens, labels, training=True)
        finally:
            O._q = orig_q
            if ctx:
                ctx.__exit__(None, None, None)
        return lo.detach(), g

    base_round = O._ROUND
    ref, gref = run("fp32")
    for mode in ("tf32", "fp16", "bf16"):
        O._ROUND = base_round
        lo, g = run(mode)
        O._ROUND = base_round
        err = float((lo - ref).abs().max() / ref.abs().max())
        gl2 = {k: float((g[k] - gref[k]).norm() / gref[k].norm().clamp_min(1e-30))
               for k in ("conv_init.weight", "film_pipeline.0.weight", "film_pipeline.4.weight", "out_linear.weight")
               if k in g}
        print(f"{mode}: logits normalised max err {err:.3e}; grad rel-L2 {gl2}")


    # backward signal range: fraction of non-zero dy elements that are sub-normal / flushed in fp16 for a loss scale
    for k in (0, 8, 12):
        del stats[:]
        O._ROUND = lambda t: t.half().float()
        lo, g = run("fp16x", grad_scale=float(2 ** k))
        O._ROUND = base_round
        gl2 = {n: float((g[n] - gref[n]).norm() / gref[n].norm().clamp_min(1e-30))
               for n in ("conv_init.weight", "film_pipeline.0.weight", "film_pipeline.4.weight")}
        mx = max(s_[1] for s_ in stats)
        sub = max(s_[2] for s_ in stats)
        flushed = max(s_[3] for s_ in stats)
        print(f"fp16 fwd + fp16 dy with loss scale 2^{k}: max|dy| {mx:.3e}, worst layer: {sub:.3f} sub-normal, "
              f"{flushed:.4f} flushed; grad rel-L2 vs fp32 {gl2}")


if __name__ == "__main__":
    main()
