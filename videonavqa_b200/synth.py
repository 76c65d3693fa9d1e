"""Seeded synthetic batches of the shape the reference feeds the hot path:
post-ReLU stem features (B, C_in, 10, 13, T), zero-padded token ids in [1, vocab), question lengths U[5, 30],
video lengths all T or the ragged distribution ceil(clip(N(75, 27), 12, 140) / 4) sorted descending.
numpy RNG, so every host regenerates the same bits.  (The test oracle carries an identical generator; they are
compared in tests/test_host_cpu.py.)"""
import numpy as np
import torch


def synthetic_batch(B, c_in=512, T=35, Lq=56, vocab=134, classes=70, seed=0, ragged=False, q_len_range=(5, 30)):
    rng = np.random.default_rng(seed)
    v = np.maximum(rng.standard_normal((B, c_in, 10, 13, T), dtype=np.float32), 0.0)
    q_lens = rng.integers(q_len_range[0], q_len_range[1] + 1, size=B)
    q = rng.integers(1, vocab, size=(B, Lq))
    q[np.arange(Lq)[None, :] >= q_lens[:, None]] = 0
    if ragged:
        raw = np.clip(rng.normal(75, 27, size=B), 12, 140)
        v_lens = np.minimum(np.ceil(raw / 4).astype(np.int64), T)
        v_lens = -np.sort(-v_lens)
    else:
        v_lens = np.full(B, T, dtype=np.int64)
    labels = rng.integers(0, classes, size=B)
    return (torch.from_numpy(v), torch.from_numpy(q.astype(np.int64)), torch.from_numpy(v_lens.astype(np.int64)),
            torch.from_numpy(q_lens.astype(np.int64)), torch.from_numpy(labels.astype(np.int64)))
