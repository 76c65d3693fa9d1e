"""Helpers shared by the oracle and GPU parity tests: load a golden case written by
oracle/gen_golden.py and rebuild its inputs."""
import hashlib
import os

import numpy as np
import torch

from oracle.gen_golden import CASES
from oracle.film_oracle import synthetic_batch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    kind, kw, bkw, _, _ = CASES[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    v, _, _, _, _ = synthetic_batch(kw["batch_size"], c_in=kw["num_input_channels"], vocab=kw["vocab_size"],
                                    classes=kw["nb_classes"], **bkw)
    sha = np.frombuffer(hashlib.sha256(v.numpy().tobytes()).digest(), dtype=np.uint8)
    assert (sha == z["v_sha256"]).all(), "numpy RNG drift: regenerate tests/golden with oracle/gen_golden.py"
    params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    grads = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g.")}
    batch = dict(v=v, q=torch.from_numpy(z["q_input"]), v_lens=torch.from_numpy(z["v_lens"]),
                 q_lens=torch.from_numpy(z["q_lens"]), labels=torch.from_numpy(z["labels"]))
    return kind, kw, params, grads, batch, z


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def norm_max_err(a, b):
    """max|a-b| / max|b| - the logits tolerance metric of SURVEY.md §8(d)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
