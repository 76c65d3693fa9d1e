"""PyTorch-eager run of the reference ALGORITHM on the same B200 ("the real beat-this bar").

/root/reference does not exist on the GPU box, so this times the oracle's functional restatement (same per-frame loop,
same nn.functional calls -> cuDNN / cuBLAS) with device tensors, TF32 off and on.  Checker / measurement tooling only;
written after the round-1 GPU budget was spent, not yet run.
    python tests/tools/eager_gpu_baseline.py [--batch 64] [--steps 3]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import film_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B = a.batch
    p = {k: t.to(dev) for k, t in O.init_params("attn", seed=1).items()}
    v, q, v_lens, q_lens, labels = [t.to(dev) for t in O.synthetic_batch(B, seed=0)]
    out = {}
    for tag, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True

        def step():
            lo, loss, g, _ = O.loss_and_grads("attn", p, {k: t.to(dev) for k, t in O.new_bn_state(1024).items()},
                                              tuple(t.to(dev) for t in O.new_film_state(B, 128)), v, q, v_lens, q_lens,
                                              labels, training=True)
            return loss
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[tag] = {"ms_per_step": round(ms, 2), "clips_per_s": round(B / ms * 1e3, 2)}
    print(json.dumps({"what": "PyTorch eager (oracle code on cuda:0), FiLM-attn fwd+bwd, no optimizer", "batch": B, **out}))


if __name__ == "__main__":
    main()
