"""CUDA-graph replay of the train step against eager launches - ``-m gpu``.

Two identical modules take the same three different batches (tokens, question lengths, frames and labels change; the
clip-length pattern is the graph's identity): one through ``GraphedTrainStep.__call__`` (graph replays only), one
through the eager loop body of eval/q_and_v_eval.py:119-139.  Losses, every parameter, both Adam moments, the BatchNorm
running statistics and the optimizer's step count must agree; the only non-bit-exact pieces of the step are the
atomic accumulations of the BatchNorm sums (fp64) and bias gradients, hence 1e-5 instead of equality.
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
    loss_fn = torch.nn.CrossEntropyLoss(reduction="sum")
    eager_losses = []
    for v, q, ql, y in batches:
        me.init_hidden()
        loss = loss_fn(me(v, q, v_lens, ql), y)
        loss.backward()
        oe.step()
        oe.zero_grad()
        eager_losses.append(float(loss))

    mg, og = _build(kind, kw, params, capturable=True)
    before = [p.detach().clone() for p in mg.parameters()]
    step = GraphedTrainStep(mg, og, *batches[0][:2], v_lens, *batches[0][2:], loss_fn=loss_fn)
    # capturing (and its eager warm-up steps) must leave the training state untouched
    assert all(torch.equal(p, q) for p, q in zip(mg.parameters(), before))
    assert int(og._step_dev.item()) == 0 and int(mg.bn_init.num_batches_tracked) == 0
    graph_losses = [float(step(v, q, ql, y)) for v, q, ql, y in batches]

    for a, c in zip(eager_losses, graph_losses):
        assert abs(a - c) <= 1e-5 * max(1.0, abs(a)), (eager_losses, graph_losses)
    assert len({round(x, 4) for x in graph_losses}) == 3                 # the replays really saw three different batches
    for (n, p), q in zip(me.named_parameters(), mg.parameters()):
        assert float((p - q).abs().max()) <= 1e-5 * max(1e-3, float(p.abs().max())), n
    assert float((oe._m - og._m).abs().max()) <= 1e-5 * float(oe._m.abs().max())
    assert float((oe._v - og._v).abs().max()) <= 1e-5 * float(oe._v.abs().max())
    assert torch.allclose(me.bn_init.running_mean, mg.bn_init.running_mean, rtol=1e-5, atol=1e-7)
    assert torch.allclose(me.bn_init.running_var, mg.bn_init.running_var, rtol=1e-5, atol=1e-7)
    assert int(me.bn_init.num_batches_tracked) == int(mg.bn_init.num_batches_tracked)
    sd = og.state_dict()
    assert float(sd["state"][0]["step"]) == 3.0 == float(oe.state_dict()["state"][0]["step"])


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
