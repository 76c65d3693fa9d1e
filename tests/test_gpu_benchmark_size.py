"""Module-level parity at the eval.sh / benchmark shape (C_in 512, C 1024, 5 ResBlocks, hidden 128) - ``-m gpu``.

  * train step, B = 4 ragged clips: logits, loss and EVERY parameter gradient against the CPU oracle, both plain fp32
    (the reference's arithmetic) and with the product's tf32 operand rounding emulated;
  * B = 64 x 35 frames (2240 frame images, the shape bench.py times): eval logits of a handful of clips against the
    CPU oracle run on just those clips (clips are independent in eval mode), and a full train step (logits + every
    gradient) against the oracle code executed in strict fp32 on the same GPU (cuDNN/cuBLAS with TF32 disabled).

Tolerances: logits max|d|/max|logit| <= 1e-3 (measured 1.2e-4 .. 1.4e-4).  Gradients: §8(d) asks
for ||dg||/||g|| <= 1e-3 per parameter; that is below the conditioning of the reference itself at this size
(test_reference_gradients_move_more_than_1e3_under_a_1e6_input_perturbation): a relative perturbation delta of the
forward activations flips the ReLU / FiLM-ReLU mask of the ~delta fraction of elements that sit within delta of zero,
each flip is an O(1) change of that gradient element, i.e. ~sqrt(delta) in relative L2 per layer, growing with depth
in the backward direction.  fp32 accumulation-order noise (1e-5) alone gives ~3e-3 at the last block and ~2e-2 at
conv_init, tf32 operand rounding (2e-4) gives 2e-3 .. 4.5e-2.  Measured table in DESIGN.md §5; asserted bounds:
GRAD_TOL_EMU against the oracle with the same tf32 rounding points, GRAD_TOL_FP32 against plain fp32, and TAIL_TOL for
the parameters behind the last ReLU (no mask between them and the loss).  The kernels themselves are pinned tightly
on identical inputs in tests/test_gpu_fullsize.py (wgrad 3e-5, dgrad 2e-4 incl. the tf32 rounding of its output).
"""
import pytest
import torch

from golden_util import norm_max_err, rel_l2
from oracle import film_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(q_embedding_size=128, nb_classes=70, num_input_channels=512, num_res_block_channels=1024,
           num_res_blocks=5, hidden_size=128, at_hidden_size=128, max_num_frames=35, vocab_size=134)
GRAD_TOL_EMU = 4e-2           # trunk parameters vs the tf32-emulating oracle (mask flips from accumulation-order noise)
GRAD_TOL_FP32 = 8e-2          # trunk parameters vs plain fp32 (mask flips from tf32 operand rounding)
TAIL_TOL = 5e-4               # parameters after the last ReLU of the trunk: no mask flips, plain arithmetic error
TAIL = ("fc_embed_attn.", "lstm_attn.", "out_linear.")
# mathematically zero gradients: quirk Q4, and fc_attn_1.bias whenever the longest clip has all T frames (the score
# bias shifts every unmasked logit of the softmax over frames by the same amount)
ZERO_GRAD = ("fc_hidden_attn.weight", "fc_hidden_attn.bias", "fc_attn_1.bias")


def _module(B, seed, train):
    from videonavqa_b200.models import FiLMAttnPretrainedStem
    torch.manual_seed(seed)
    m = FiLMAttnPretrainedStem(batch_size=B, **CFG).cuda().set_operands("tf32")   # tests/test_gpu_fp16.py re-sets fp16
    # non-trivial biases / BatchNorm affine so that their gradients are exercised
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias") and "lstm" not in n and "film_layer.0" not in n:
                p.uniform_(-0.05, 0.05)
        m.bn_init.weight.uniform_(0.8, 1.2)
    return m.train(train)


def _named(m):
    named = dict(m.named_parameters())
    for k, conv in enumerate(m.conv1x1_layers):
        named[f"conv1x1_layers.{k}.weight"] = conv.weight
        named[f"conv1x1_layers.{k}.bias"] = conv.bias
    return named


def _grad_table(named, refs):
    """rows: (name, rel-L2 vs each reference gradient dict)."""
    rows = []
    for k in sorted(refs[0]):
        g = named[k].grad
        assert g is not None, f"no gradient for {k}"
        if k in ZERO_GRAD:
            assert float(g.abs().max()) < 2e-5 and float(refs[-1][k].abs().max()) < 2e-5, k
            continue
        rows.append((k,) + tuple(rel_l2(g.detach().to(r[k].device), r[k]) for r in refs))
    return rows


def test_train_step_eval_sh_size_ragged_vs_cpu_oracle():
    B = 4
    m = _module(B, 11, True)
    v, q, _, q_lens, labels = O.synthetic_batch(B, seed=12)
    v_lens = torch.tensor([35, 20, 9, 3])
    params = {k: p.detach().cpu() for k, p in _named(m).items()}
    m.init_hidden()
    lo = m(v.cuda(), q.cuda(), v_lens.cuda(), q_lens.cuda())
    loss = torch.nn.functional.cross_entropy(lo, labels.cuda(), reduction="sum")
    loss.backward()
    lo_ref, loss_ref, g_ref, _ = O.loss_and_grads("attn", params, O.new_bn_state(1024), O.new_film_state(B, 128), v, q,
                                                  v_lens, q_lens, labels, training=True)
    with O.emulate_tf32():
        lo_emu, _, g_emu, _ = O.loss_and_grads("attn", params, O.new_bn_state(1024), O.new_film_state(B, 128), v, q,
                                               v_lens, q_lens, labels, training=True)
    e_ref, e_emu = norm_max_err(lo.detach().cpu(), lo_ref), norm_max_err(lo.detach().cpu(), lo_emu)
    print(f"logits: vs fp32 oracle {e_ref:.2e}, vs tf32-emulating oracle {e_emu:.2e}")
    assert e_ref <= 1e-3 and e_emu <= 3e-4
    assert abs(loss.item() - float(loss_ref)) <= 1e-3 * abs(float(loss_ref))
    rows = _grad_table(_named(m), (g_emu, g_ref))
    print("parameter gradient rel-L2                  vs tf32-oracle   vs fp32-oracle")
    for k, a, b in rows:
        print(f"  {k:38s} {a:10.2e}    {b:10.2e}")
    _assert_rows(rows)


def _assert_rows(rows):
    for k, a, b in rows:
        if k.startswith(TAIL):
            assert a <= TAIL_TOL and b <= TAIL_TOL, (k, a, b)
        else:
            assert a <= GRAD_TOL_EMU, (k, a)
            assert b <= GRAD_TOL_FP32, (k, b)


def test_reference_gradients_move_more_than_1e3_under_a_1e6_input_perturbation():
    """Conditioning of the REFERENCE arithmetic at the eval.sh size, measured with the oracle alone in strict fp32 on
    the GPU: scaling the input video by (1 + 1e-6 * noise) - far below any operand rounding - moves the conv-stack
    gradients by more than 1e-3 in relative L2 (ReLU-mask flips), while the logits move by ~1e-6.  This is why the
    per-parameter gradient bounds above are what they are; it involves no product code."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = torch.device("cuda")
        B = 8
        p = {k: t.to(dev) for k, t in O.init_params("attn", seed=3).items()}
        v, q, v_lens, q_lens, labels = (t.to(dev) for t in O.synthetic_batch(B, seed=4))
        g = torch.Generator(device="cuda").manual_seed(5)
        v2 = v * (1.0 + 1e-6 * torch.randn(v.shape, device=dev, generator=g))

        def run(vv):
            bn = {k: t.to(dev) for k, t in O.new_bn_state(1024).items()}
            st = tuple(t.to(dev) for t in O.new_film_state(B, 128))
            lo, _, gr, _ = O.loss_and_grads("attn", p, bn, st, vv, q, v_lens, q_lens, labels, training=True)
            return lo, gr
        lo_a, g_a = run(v)
        lo_b, g_b = run(v2)
        e_logits = norm_max_err(lo_b, lo_a)
        e_init = rel_l2(g_b["conv_init.weight"], g_a["conv_init.weight"])
        e_last = rel_l2(g_b["film_pipeline.4.weight"], g_a["film_pipeline.4.weight"])
        print(f"reference under a 1e-6 input perturbation: logits {e_logits:.2e}, conv_init.weight grad {e_init:.2e}, "
              f"film_pipeline.4.weight grad {e_last:.2e}")
        assert e_logits < 1e-4
        assert e_init > 1e-3, "the reference's own conv_init gradient is better conditioned than claimed"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_b64_eval_logits_of_sampled_clips_vs_cpu_oracle():
    """The exact shape bench.py times (2240 frame images, 2275 persistent tiles per conv launch)."""
    B = 64
    m = _module(B, 13, False)
    v, q, v_lens, q_lens, _ = O.synthetic_batch(B, seed=14)
    with torch.no_grad():
        m.init_hidden()
        lo = m(v.cuda(), q.cuda(), v_lens.cuda(), q_lens.cuda()).cpu()
    pick = [0, 17, 40, 63]
    params = {k: p.detach().cpu() for k, p in _named(m).items()}
    with torch.no_grad():
        ref, _ = O.film_attn_forward(params, O.new_bn_state(1024), O.new_film_state(len(pick), 128), v[pick], q[pick],
                                     v_lens[pick], q_lens[pick], training=False)
    err = float((lo[pick] - ref).abs().max() / ref.abs().max())
    print(f"B=64 eval logits, clips {pick}: normalised max error {err:.2e}")
    assert err <= 1e-3
    assert (lo[pick].argmax(1) == ref.argmax(1)).all()


def test_b64_train_step_vs_strict_fp32_oracle_on_gpu():
    """Full benchmark-size train step: logits, loss and every parameter gradient against the oracle's functional code
    executed on the same GPU in strict fp32 (checker only; the product path is the module under test)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        B = 64
        m = _module(B, 15, True)
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
        bn = {k: t.to(dev) for k, t in O.new_bn_state(1024).items()}
        with O.emulate_tf32():
            lo_emu, _, g_emu, _ = O.loss_and_grads("attn", params, bn, st, v, q, v_lens, q_lens, labels, training=True)
        e_ref, e_emu = norm_max_err(lo.detach(), lo_ref), norm_max_err(lo.detach(), lo_emu)
        agree = float((lo.argmax(1) == lo_ref.argmax(1)).float().mean())
        print(f"B=64 train logits: vs fp32 {e_ref:.2e}, vs tf32-emulating {e_emu:.2e}, argmax agreement {agree:.3f}")
        assert e_ref <= 1e-3 and e_emu <= 3e-4
        assert abs(loss.item() - float(loss_ref)) <= 1e-3 * abs(float(loss_ref))
        rows = _grad_table(_named(m), (g_emu, g_ref))
        print("parameter gradient rel-L2 (B=64)           vs tf32-oracle   vs fp32-oracle")
        for k, a, b in rows:
            print(f"  {k:38s} {a:10.2e}    {b:10.2e}")
        _assert_rows(rows)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("operands", ["fp16", "tf32"])
def test_b64_clip_permutation_equivariance(operands):
    """Size-independent property at the benchmark shape: in eval mode every clip is processed independently
    (tests/test_oracle_golden.py pins that for the oracle), so permuting the clips of a batch of equal-length clips
    permutes the logits - through 2240 frame images, the frame-major image order, the question-length sort and the
    attention tail.  Images land in different tiles and different CTAs after the permutation; the accumulation order
    inside a tile does not depend on the image's position, so the agreement is to the last bit."""
    B = 64
    m = _module(B, 21, False).set_operands(operands)
    v, q, v_lens, q_lens, _ = O.synthetic_batch(B, seed=22)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
    vc = v.cuda()
    with torch.no_grad():
        m.init_hidden()
        a = m(vc, q.cuda(), v_lens, q_lens)
        m.init_hidden()
        b = m(vc[perm.cuda()].contiguous(), q[perm].cuda(), v_lens[perm], q_lens[perm])
    assert torch.equal(a[perm.cuda()], b)
    assert not torch.equal(a, b)
