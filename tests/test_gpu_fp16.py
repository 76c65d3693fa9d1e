"""The fp16-operand path (engine.Precision('fp16'): kind::f16 MMAs, half trunk activations, 2^12 loss scale on the
trunk's backward signals) through the same gates as the tf32 path - ``-m gpu``.

  * the four golden cases written by the unmodified reference: logits <= 1e-3 (normalised max error), gradients
    against the oracle with fp16 operand rounding emulated at the same points;
  * eval.sh size (C = 1024, 5 blocks): B = 4 ragged train step against the fp32 CPU oracle, B = 64 x 35 train step
    against the oracle run in strict fp32 on the GPU (logits, argmax, every parameter gradient);
  * range: the un-normalised residual stream after 5 ResBlocks and the loss-scaled backward signals stay far inside
    the fp16 range (max 65504), no inf / nan anywhere.
"""
import pytest
import torch

from golden_util import load_case, norm_max_err, rel_l2
from oracle import film_oracle as O
from test_gpu_benchmark_size import CFG, _assert_rows, _grad_table, _module, _named
from test_gpu_modules import all_named_params, build_module

pytestmark = pytest.mark.gpu

CASES = ["attn_small", "attn_h128", "gp_small", "tmh_small"]


@pytest.mark.parametrize("name", CASES)
def test_golden_cases_fp16_operands(name):
    kind, kw, params, grads_ref, b, z = load_case(name)
    m = build_module(kind, kw, params).train().set_operands("fp16")
    m.init_hidden()
    lo = m(b["v"].cuda(), b["q"].cuda(), b["v_lens"].cuda(), b["q_lens"].cuda())
    loss = torch.nn.functional.cross_entropy(lo, b["labels"].cuda(), reduction="sum")
    loss.backward()
    ltol = 2e-3 if name == "tmh_small" else 1e-3        # same allowance as the tf32 path (tests/test_gpu_modules.py)
    e_ref = norm_max_err(lo.detach().cpu(), torch.from_numpy(z["logits_train"]))
    C = kw["num_res_block_channels"]
    with O.emulate_fp16():
        lo_emu, _, g_emu, _ = O.loss_and_grads(kind, params, O.new_bn_state(C),
                                               O.new_film_state(kw["batch_size"], kw["hidden_size"]), b["v"], b["q"],
                                               b["v_lens"], b["q_lens"], b["labels"], training=True)
    e_emu = norm_max_err(lo.detach().cpu(), lo_emu)
    named = all_named_params(m)
    worst, worst_k = 0.0, None
    for k, gref in grads_ref.items():
        g = named[k].grad
        assert g is not None and torch.isfinite(g).all(), k
        if float(gref.abs().max()) < 1e-6:
            assert float(g.abs().max()) < 1e-6, k
            continue
        e = rel_l2(g.cpu(), g_emu[k])
        if e > worst:
            worst, worst_k = e, k
        assert rel_l2(g.cpu(), gref) <= 1e-1, (k, rel_l2(g.cpu(), gref))
    print(f"{name} fp16: logits vs reference {e_ref:.2e}, vs fp16-emulating oracle {e_emu:.2e}, worst gradient rel-L2 vs "
          f"emulating oracle {worst:.2e} [{worst_k}]")
    assert e_ref <= ltol and e_emu <= 5e-4
    assert worst <= (2e-2 if kind in ("gp", "tmh") else 8e-3), (worst_k, worst)
    # eval forward through the same operands (a fresh module: the train step above moved the BatchNorm statistics)
    m = build_module(kind, kw, params).eval().set_operands("fp16")
    m.init_hidden()
    with torch.no_grad():
        le = m(b["v"].cuda(), b["q"].cuda(), b["v_lens"].cuda(), b["q_lens"].cuda()).cpu()
    assert norm_max_err(le, torch.from_numpy(z["logits_eval"])) <= ltol


def test_fp16_train_step_eval_sh_size_ragged_vs_cpu_oracle():
    B = 4
    m = _module(B, 11, True).set_operands("fp16")
    v, q, _, q_lens, labels = O.synthetic_batch(B, seed=12)
    v_lens = torch.tensor([35, 20, 9, 3])
    params = {k: p.detach().cpu() for k, p in _named(m).items()}
    m.init_hidden()
    lo = m(v.cuda(), q.cuda(), v_lens.cuda(), q_lens.cuda())
    loss = torch.nn.functional.cross_entropy(lo, labels.cuda(), reduction="sum")
    loss.backward()
    lo_ref, loss_ref, g_ref, _ = O.loss_and_grads("attn", params, O.new_bn_state(1024), O.new_film_state(B, 128), v, q,
                                                  v_lens, q_lens, labels, training=True)
    with O.emulate_fp16():
        lo_emu, _, g_emu, _ = O.loss_and_grads("attn", params, O.new_bn_state(1024), O.new_film_state(B, 128), v, q,
                                               v_lens, q_lens, labels, training=True)
    e_ref, e_emu = norm_max_err(lo.detach().cpu(), lo_ref), norm_max_err(lo.detach().cpu(), lo_emu)
    print(f"fp16 logits: vs fp32 oracle {e_ref:.2e}, vs fp16-emulating oracle {e_emu:.2e}")
    assert e_ref <= 1e-3 and e_emu <= 5e-4
    assert abs(loss.item() - float(loss_ref)) <= 1e-3 * abs(float(loss_ref))
    rows = _grad_table(_named(m), (g_emu, g_ref))
    print("parameter gradient rel-L2 (fp16 operands)  vs fp16-oracle   vs fp32-oracle")
    for k, a, b in rows:
        print(f"  {k:38s} {a:10.2e}    {b:10.2e}")
    _assert_rows(rows)


def test_fp16_b64_train_step_vs_strict_fp32_oracle_on_gpu():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        B = 64
        m = _module(B, 15, True).set_operands("fp16")
        dev = torch.device("cuda")
        v, q, v_lens, q_lens, labels = (t.to(dev) for t in O.synthetic_batch(B, seed=16))
        m.init_hidden()
        lo = m(v, q, v_lens, q_lens)
        loss = torch.nn.functional.cross_entropy(lo, labels, reduction="sum")
        loss.backward()
        params = {k: p.detach() for k, p in _named(m).items()}
        bn = {k: t.to(dev) for k, t in O.new_bn_state(1024).items()}
        st = tuple(t.to(dev) for t in O.new_film_state(B, 128))
        lo_ref, loss_ref, g_ref, _ = O.loss_and_grads("attn", params, bn, st, v, q, v_lens, q_lens, labels, training=True)
        e_ref = norm_max_err(lo.detach(), lo_ref)
        agree = float((lo.argmax(1) == lo_ref.argmax(1)).float().mean())
        print(f"fp16 B=64 train logits vs fp32 {e_ref:.2e}, argmax agreement {agree:.3f}")
        assert e_ref <= 1e-3 and agree == 1.0
        assert abs(loss.item() - float(loss_ref)) <= 1e-3 * abs(float(loss_ref))
        rows = _grad_table(_named(m), (g_ref, g_ref))
        print("parameter gradient rel-L2 (fp16 operands, B=64) vs fp32-oracle")
        for k, a, _ in rows:
            print(f"  {k:38s} {a:10.2e}")
        _assert_rows([(k, 0.0, b) for k, _, b in rows])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_fp16_range_of_residual_stream_and_backward_signals():
    """Overflow / underflow margins at the benchmark shape: the residual stream is never normalised after bn_init, so
    it grows block by block; the backward signals carry the 2^12 loss scale."""
    from videonavqa_b200 import engine
    B = 16
    m = _module(B, 17, True).set_operands("fp16")
    dev = torch.device("cuda")
    v, q, v_lens, q_lens, labels = (t.to(dev) for t in O.synthetic_batch(B, seed=18))
    v = v * 4.0                                             # a hot stem: 4x the synthetic feature scale
    names, tensors = m._collect_params()
    P = dict(zip(names, tensors))
    m.init_hidden()
    logits, ctx = m._forward_impl(P, v, q, v_lens, q_lens, save=True)
    peak = max(float(x.float().abs().max()) for x in ctx["xs"] + ctx["zs"] + ctx["ress"])
    print(f"fp16 trunk: largest |activation| over x / z / res of 5 blocks = {peak:.1f} (fp16 max 65504)")
    assert torch.isfinite(logits).all() and peak < 65504 / 8
    lg = logits.detach().requires_grad_(True)
    dlogits = torch.autograd.grad(torch.nn.functional.cross_entropy(lg, labels, reduction="sum"), lg)[0]
    grads = engine.attn_backward(P, m._cache, ctx, dlogits, ctx["q_in"], ctx["q_lens32"])
    for k, g in grads.items():
        assert torch.isfinite(g).all(), k
    assert float(grads["conv_init.weight"].abs().max()) > 0
