"""CUDA-graph replay of the train step against eager launches - ``-m gpu``.

Two identical modules take the same three different batches (tokens, question lengths, frames and labels change; the
clip-length pattern is the graph's identity): one through ``GraphedTrainStep.__call__`` (graph replays only), one
through the eager loop body of eval/q_and_v_eval.py:119-139.  Losses, every parameter, both Adam moments, the BatchNorm
running statistics and the optimizer's step count must agree.  The only non-bit-exact pieces of the step are the
atomic accumulations of the BatchNorm sums (fp64) and bias gradients, and Adam turns last-bit differences of near-zero
gradient elements into +-lr differences of single weights (m / sqrt(v) is a sign function there - the same holds
between two eager runs).  The comparison is therefore made on ONE step from an identical state: loss, the whole
gradient bucket, both moments and the BatchNorm statistics to 1e-5; further replays are checked through their losses.
"""
import pytest
import torch

from golden_util import load_case

pytestmark = pytest.mark.gpu


def _build(kind, kw, params, capturable):
    from test_gpu_modules import build_module
    from videonavqa_b200.optim import FusedClipAdam
    m = build_module(kind, kw, params, operands="fp16").train()
    opt = FusedClipAdam(m.parameters(), lr=1e-3, max_norm=1.0, capturable=capturable).attach(m)
    return m, opt


@pytest.mark.parametrize("case", ["attn_small", "gp_small"])
def test_graph_replay_matches_eager_steps(case):
    from videonavqa_b200.graph import GraphedTrainStep
    kind, kw, params, _, b, _ = load_case(case)
    B = kw["batch_size"]
    g = torch.Generator().manual_seed(5)
    batches = []
    for i in range(3):
        q_lens = torch.randint(1, b["q"].shape[1] + 1, (B,), generator=g)
        q = torch.randint(1, kw["vocab_size"], tuple(b["q"].shape), generator=g)
        q[torch.arange(q.shape[1])[None, :] >= q_lens[:, None]] = 0
        v = (b["v"] * (1.0 + 0.1 * i)).cuda()
        labels = torch.randint(0, kw["nb_classes"], (B,), generator=g)
        batches.append((v, q.cuda(), q_lens.cuda(), labels.cuda()))
    v_lens = b["v_lens"].clone()                      # CPU tensor: the graph's identity

    me, oe = _build(kind, kw, params, capturable=False)
    mg, og = _build(kind, kw, params, capturable=True)
    loss_fn = torch.nn.CrossEntropyLoss(reduction="sum")

    def eager(m, o, v, q, ql, y):
        m.init_hidden()
        loss = loss_fn(m(v, q, v_lens, ql), y)
        loss.backward()
        o.step()
        o.zero_grad()
        return loss

    # both modules first train one step EAGERLY on the default stream, and the graphed one keeps its loss tensor (and
    # with it the autograd graph) alive - the situation of a training script that switches to graph replay later
    stale = [eager(me, oe, *batches[0]), eager(mg, og, *batches[0])]
    assert abs(float(stale[0].detach()) - float(stale[1].detach())) <= 1e-5 * abs(float(stale[0].detach()))
    before = [p.detach().clone() for p in mg.parameters()]
    n_tracked = int(mg.bn_init.num_batches_tracked)
    step = GraphedTrainStep(mg, og, *batches[0][:2], v_lens, *batches[0][2:], loss_fn=loss_fn)
    # capturing (and its eager warm-up steps) must leave the training state untouched
    assert all(torch.equal(p, q) for p, q in zip(mg.parameters(), before))
    assert int(og._step_dev.item()) == 1 and int(mg.bn_init.num_batches_tracked) == n_tracked

    # identical starting state (the two eager first steps already differ in single weights: Adam turns last-bit
    # differences of near-zero gradient elements - atomics - into +-lr), then ONE step each on a new batch
    with torch.no_grad():
        og.flat.param.copy_(oe.flat.param); og._m.copy_(oe._m); og._v.copy_(oe._v)
        for (_, a), (_, c) in zip(me.named_buffers(), mg.named_buffers()):
            c.copy_(a)
    # (conv1x1_layers are never optimised - quirk Q1 - and identical in both modules by construction)
    le = float(eager(me, oe, *batches[1]).detach())
    lg = float(step(*batches[1]))
    assert abs(le - lg) <= 1e-5 * abs(le), (le, lg)

    def rel(a, c):
        return float((a.double() - c.double()).norm() / a.double().norm().clamp_min(1e-30))
    assert rel(oe.flat.grad, og.flat.grad) <= 1e-5          # the gradient bucket of the step (kept until the next backward)
    assert rel(oe._m, og._m) <= 1e-5 and rel(oe._v, og._v) <= 1e-5
    assert torch.allclose(me.bn_init.running_mean, mg.bn_init.running_mean, rtol=1e-5, atol=1e-7)
    assert torch.allclose(me.bn_init.running_var, mg.bn_init.running_var, rtol=1e-5, atol=1e-7)
    for ca, cc in zip(me.conv1x1_layers, mg.conv1x1_layers):          # Q1: unregistered, but they do receive .grad
        assert cc.weight.grad is not None and rel(ca.weight.grad, cc.weight.grad) <= 1e-5

    # further replays really see new inputs, and the bookkeeping that no Python runs for stays right
    losses = [(float(eager(me, oe, *bt).detach()), float(step(*bt))) for bt in (batches[2], batches[0])]
    assert all(abs(a - c) <= 2e-3 * abs(a) for a, c in losses), losses
    assert len({round(c, 3) for _, c in losses} | {round(lg, 3)}) == 3
    assert int(me.bn_init.num_batches_tracked) == int(mg.bn_init.num_batches_tracked)
    assert float(og.state_dict()["state"][0]["step"]) == 4.0 == float(oe.state_dict()["state"][0]["step"])


def test_capturable_adam_matches_host_step_count():
    """vnqa_clip_adam_dev (bias corrections formed on the device from the device-resident step count) against the
    host-scalar path, 5 steps on random gradients: identical bits (same double arithmetic, same rounding)."""
    from videonavqa_b200.optim import FusedClipAdam
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in (1000, 37, 4096)]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a = FusedClipAdam(ps, lr=1e-3, max_norm=1.0)
    c = FusedClipAdam(qs, lr=1e-3, max_norm=1.0, capturable=True)
    for _ in range(5):
        for p, q in zip(ps, qs):
            p.grad = torch.randn_like(p)
            q.grad = p.grad.clone()
        a.step()
        c.step()
    for p, q in zip(ps, qs):
        assert torch.equal(p, q)
    assert float(c.state_dict()["state"][2]["step"]) == 5.0
