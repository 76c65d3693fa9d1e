"""Property tests of the host-side logic (hypothesis, CPU): the frontier-driven gradient reducer, the ragged-batch
plan against the reference's frame loop, the flat parameter buffers and the frame subsampling."""
import random

import torch
from hypothesis import given, settings, strategies as st

from videonavqa_b200 import dp
from videonavqa_b200.optim import FlatBuffers


class _Recorder:
    """Stands in for the NCCL collective: records the [a, b) slices handed over, in order."""

    def __init__(self):
        self.slices, self.joins = [], 0

    def all_reduce(self, t, a=0, b=None, join=True):
        self.slices.append((a, t.numel() if b is None else b))

    def join(self):
        self.joins += 1


class _Mod(torch.nn.Module):
    def __init__(self, sizes):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(n)) for n in sizes])


@settings(max_examples=60, deadline=None)
@given(sizes=st.lists(st.integers(1, 5000), min_size=1, max_size=12), seed=st.integers(0, 10 ** 6),
       min_kb=st.sampled_from([0, 1, 4, 64]), overlap=st.booleans(), drop=st.integers(0, 3))
def test_flat_reducer_sends_every_gradient_exactly_once_and_only_when_final(sizes, seed, min_kb, overlap, drop):
    """Whatever the order in which the backward pass finishes the gradients (and even if some parameters get none):
    the slices handed to the collective are disjoint, cover the whole bucket, and a slice is only sent once every
    parameter inside it - and every parameter behind it - is final."""
    m = _Mod(sizes)
    flat = FlatBuffers(m.parameters())
    rec = _Recorder()
    old = dp._COLL[0]
    dp._COLL[0] = rec
    try:
        red = dp.FlatReducer(m, flat, overlap=overlap, min_bytes=min_kb << 10)
        names = [n for n, _ in m.named_parameters()]
        order = list(range(len(names)))
        random.Random(seed).shuffle(order)
        order = order[:len(order) - min(drop, len(order) - 1)]          # some parameters never report a gradient
        done = set()
        for i in order:
            before = len(rec.slices)
            red(names[i], flat.grad_views[i])
            done.add(i)
            for a, b in rec.slices[before:]:
                inside = [k for k in range(len(sizes)) if flat.offsets[k] < b and flat.offsets[k] + sizes[k] > a]
                behind = [k for k in range(len(sizes)) if flat.offsets[k] >= b]
                assert all(k in done for k in inside + behind)
        red("not_a_registered_parameter", None)                          # conv1x1_layers: ignored
        assert red.wait() == flat.total * 4 and rec.joins == 1
        if not overlap:
            assert rec.slices == [(0, flat.total)]
        # disjoint cover of [0, total), handed over from the top of the bucket downwards
        assert sorted(rec.slices) == sorted(rec.slices, key=lambda s: s[0])
        cover = sorted(rec.slices)
        assert cover[0][0] == 0 and cover[-1][1] == flat.total
        assert all(x[1] == y[0] for x, y in zip(cover, cover[1:]))
        assert [s[0] for s in rec.slices] == sorted((s[0] for s in rec.slices), reverse=True)
        # the reducer is reusable: a second backward starts from a clean frontier
        for i in range(len(names)):
            red(names[i], flat.grad_views[i])
        red.wait()
        assert rec.joins == 2
    finally:
        dp._COLL[0] = old


@settings(max_examples=40, deadline=None)
@given(sizes=st.lists(st.integers(1, 300), min_size=1, max_size=10))
def test_flat_buffers_keep_parameters_as_aligned_views(sizes):
    m = _Mod(sizes)
    vals = [torch.randn(n) for n in sizes]
    with torch.no_grad():
        for p, v in zip(m.parameters(), vals):
            p.copy_(v)
    flat = FlatBuffers(m.parameters())
    assert flat.is_intact() and flat.total % 4 == 0
    for p, v, o, g in zip(m.parameters(), vals, flat.offsets, flat.grad_views):
        assert o % 4 == 0 and torch.equal(p.detach(), v) and g.shape == p.shape
        assert p.data_ptr() == flat.param.data_ptr() + 4 * o and g.data_ptr() == flat.grad.data_ptr() + 4 * o
    flat.param.fill_(3.0)                       # the parameters ARE the buffer
    assert all(float(p.detach().min()) == 3.0 for p in m.parameters())
    m.ps[0].data = torch.zeros(sizes[0])        # something re-homed a parameter
    assert not flat.is_intact()


@settings(max_examples=60, deadline=None)
@given(lens=st.lists(st.integers(1, 35), min_size=1, max_size=9))
def test_plan_matches_the_reference_frame_loop_for_any_lengths(lens):
    """engine.Plan against the reference's bookkeeping (film_attn_pt_stem.py:190-210, 245-251): frame i is processed
    for the first ct_i = #{b : v_len[b] > i} clips (clips sorted by descending length), frames beyond the longest clip
    are not processed at all."""
    from videonavqa_b200.engine import Plan
    v_lens = sorted(lens, reverse=True)
    B, T = len(v_lens), 35
    plan = Plan(v_lens, B, T, torch.device("cpu"))
    img_of = plan.img_of.view(B, T)
    img = 0
    for i in range(v_lens[0]):
        ct = sum(1 for v in v_lens if v > i)
        assert img_of[:ct, i].tolist() == list(range(img, img + ct))       # frame-major, clip order inside a frame
        assert (img_of[ct:, i] == -1).all()
        assert plan.img_frame[img:img + ct].tolist() == [i] * ct
        img += ct
    assert (img_of[:, v_lens[0]:] == -1).all()
    assert plan.n_img == img == sum(v_lens) and plan.n_frames == v_lens[0]
    assert plan.frame_ct[:plan.n_frames].tolist() == [sum(1 for v in v_lens if v > i) for i in range(v_lens[0])]


@settings(max_examples=100, deadline=None)
@given(n=st.integers(0, 400), seed=st.integers(0, 10 ** 6))
def test_subsample_indices_properties(n, seed):
    """eval/dataset.py:81-86: one frame out of every group of 4, in order, ceil(n / 4) frames in total."""
    from videonavqa_b200 import data
    idx = data.subsample_indices(n, rng=random.Random(seed))
    assert len(idx) == (n + 3) // 4
    assert all(4 * k <= i < min(4 * k + 4, n) for k, i in enumerate(idx))
    r = random.Random(seed)                      # the very call sequence of the reference
    assert idx == [r.randint(i, min(i + 4, n) - 1) for i in range(0, n, 4)]
