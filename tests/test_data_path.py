"""Data path, first slice: uint8 staging + device-side conversion against the UNMODIFIED
reference ``VNQADataset.__getitem__`` (eval/dataset.py:57-93), whose output on the committed clip
``tests/golden/data_clip.mp4`` is pinned by hashes in ``tests/golden/data_small.npz`` (oracle/gen_golden_data.py).
Bit-exact: the path is byte movement plus one correctly rounded division."""
import hashlib
import os
import random

import numpy as np
import pytest
import torch

from oracle import data_oracle as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).digest()


def _golden_clip():
    from videonavqa_b200 import data
    g = np.load(os.path.join(GOLD, "data_small.npz"))
    path = os.path.join(GOLD, "data_clip.mp4")
    assert hashlib.sha256(open(path, "rb").read()).digest() == g["clip_sha256"].tobytes()
    frames = data.decode_clip_uint8(path)
    assert frames.shape == (int(g["n_raw"]), 160, 208, 3) and frames.dtype == np.uint8
    return g, frames


def test_oracle_and_subsampling_match_the_reference_dataset():
    from videonavqa_b200 import data
    g, frames = _golden_clip()
    random.seed(7)                                            # the seed the golden generator gave the reference
    idx = data.subsample_indices(frames.shape[0])
    assert idx == g["indices"].tolist() and len(idx) == int(g["v_len"])
    v64, v_len = D.clip_tensor(frames, idx)
    assert v_len == int(g["v_len"]) and v64.dtype == torch.float64 and tuple(v64.shape) == (3, 160, 208, 35)
    assert _sha(v64) == g["sha256_f64"].tobytes()
    v32 = v64.float()
    assert _sha(v32) == g["sha256_f32"].tobytes()
    for i, val in zip(g["sample_idx"], g["sample_val"]):
        assert float(v32[tuple(i)]) == float(val)
    # a different generator object with the same interface (e.g. one per DataLoader worker)
    r = random.Random(7)
    assert data.subsample_indices(frames.shape[0], rng=r) == idx
    assert data.subsample_indices(0) == [] and data.subsample_indices(1) == [0]
    assert all(4 * k <= i < 4 * k + 4 for k, i in enumerate(data.subsample_indices(140)))


def test_stager_host_logic():
    from videonavqa_b200 import _lib, data
    g, frames = _golden_clip()
    st = data.VideoBatchStager(2, "cpu")
    assert st.h2d_bytes == 2 * 35 * 160 * 208 * 3
    st.host.fill_(7)                                          # stale content of a previous batch
    n = st.put(0, frames, indices=g["indices"].tolist())
    assert n == 10 and np.array_equal(st.host[0, 3].numpy(), frames[int(g["indices"][3])])
    assert int(st.host[0, 10:].max()) == 0                    # zero frames behind the clip's end (dataset.py:79-80)
    with pytest.raises(ValueError):
        st.put(1, frames.astype(np.float32))
    with pytest.raises(ValueError):
        st.put(1, np.zeros((200, 160, 208, 3), np.uint8), indices=list(range(36)))
    with pytest.raises(_lib.VnqaError):
        st.to_device()


@pytest.mark.gpu
def test_device_conversion_is_bit_identical_to_the_reference():
    from videonavqa_b200 import data
    g, frames = _golden_clip()
    rng = np.random.RandomState(1)
    full = rng.randint(0, 256, size=(140, 160, 208, 3)).astype(np.uint8)      # 140 raw frames -> all 35 slots used
    st = data.VideoBatchStager(3, "cuda")
    lens = [st.put(0, frames, indices=g["indices"].tolist()),
            st.put(1, full, rng=random.Random(3)),
            st.put(2, full[:6], indices=[1, 5])]
    assert lens == [10, 35, 2]
    out = st.to_device()
    torch.cuda.synchronize()
    assert out.shape == (3, 3, 160, 208, 35) and out.dtype == torch.float32
    o = out.cpu()
    assert _sha(o[0]) == g["sha256_f32"].tobytes()            # the reference's own output, bit for bit
    idx1 = data.subsample_indices(140, rng=random.Random(3))
    assert torch.equal(o[1], D.clip_tensor(full, idx1)[0].float())
    assert torch.equal(o[2], D.clip_tensor(full[:6], [1, 5])[0].float())
    # second batch through the same stager: stale frames of longer clips must not survive
    st.put(1, full[:6], indices=[0])
    o2 = st.to_device().cpu()
    assert torch.equal(o2[1], D.clip_tensor(full[:6], [0])[0].float()) and torch.equal(o2[0], o[0])
    # any other frame count / frame size goes through the generic kernel
    st2 = data.VideoBatchStager(2, "cuda", max_frames=20, height=24, width=30)
    small = rng.randint(0, 256, size=(50, 24, 30, 3)).astype(np.uint8)
    n0 = st2.put(0, small, rng=random.Random(9))
    n1 = st2.put(1, small[:3], indices=[2])
    o3 = st2.to_device().cpu()
    assert (n0, n1) == (13, 1) and o3.shape == (2, 3, 24, 30, 20)
    assert torch.equal(o3[0], D.clip_tensor(small, data.subsample_indices(50, rng=random.Random(9)), max_frames=20)[0].float())
    assert torch.equal(o3[1], D.clip_tensor(small[:3], [2], max_frames=20)[0].float())
