"""GPU checks of the fused global-norm clip + Adam (videonavqa_b200.optim, csrc/optim.cu) against the reference's own
two calls, ``clip_grad_norm_(params, 1.0)`` + ``torch.optim.Adam.step()`` (eval/q_and_v_eval.py:137-139, 333), and of
the flat gradient bucket written by the backward kernels.  ``-m gpu``."""
import copy

import pytest
import torch

from oracle import film_oracle as O

pytestmark = pytest.mark.gpu


def _small_model(seed=0):
    from videonavqa_b200.models import FiLMAttnPretrainedStem
    torch.manual_seed(seed)
    return FiLMAttnPretrainedStem(batch_size=4, q_embedding_size=32, nb_classes=5, num_input_channels=64,
                                  num_res_block_channels=128, num_res_blocks=2, hidden_size=64, at_hidden_size=64,
                                  max_num_frames=35, vocab_size=30).cuda().train()


def test_fused_clip_adam_matches_torch_on_random_gradients():
    """Five steps on random tensors of awkward sizes, gradient norms on both sides of the clipping threshold."""
    from videonavqa_b200.optim import FusedClipAdam
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(7, 13), (1,), (128, 33, 3, 3), (1000,), (5, 5, 5)]
    ref_p = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=1e-3)
    ours = FusedClipAdam(our_p, lr=1e-3, max_norm=1.0)
    for step in range(5):
        scale = [10.0, 1e-3, 3.0, 1e-4, 0.5][step]       # clipped, not clipped, ...
        grads = [torch.randn(s, device="cuda", generator=g) * scale for s in shapes]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        ref.step()
        ours.step()
        assert abs(float(ours.grad_norm) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        for p, q in zip(ref_p, our_p):
            err = float((p.detach() - q.detach()).abs().max() / p.detach().abs().max())
            assert err <= 1e-6, (step, tuple(p.shape), err)
    # optimizer state: standard entries, interchangeable with torch.optim.Adam in both directions
    sd = ours.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    for i, p in enumerate(ref_p):
        assert torch.allclose(sd["state"][i]["exp_avg"], ref.state[p]["exp_avg"], rtol=1e-5, atol=1e-9)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], ref.state[p]["exp_avg_sq"], rtol=1e-5, atol=1e-12)
    fresh = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in our_p], lr=1e-3)
    fresh.load_state_dict(sd)                                 # our checkpoint loads into the reference's optimizer
    back = FusedClipAdam([torch.nn.Parameter(p.detach().clone()) for p in our_p], lr=1e-3)
    back.load_state_dict(ref.state_dict())                    # and the reference's checkpoint loads into ours
    assert back._steps == 5
    assert torch.allclose(back.state[back.flat.params[2]]["exp_avg"], ref.state[ref_p[2]]["exp_avg"])


def test_flat_bucket_training_matches_plain_autograd_path():
    """The same model trained 3 steps (a) the reference's way - autograd-owned gradients, clip_grad_norm_, Adam - and
    (b) with FusedClipAdam attached (gradients written by the backward kernels into the flat bucket, parameters views
    of one flat buffer, packed-weight caches invalidated by the optimizer): same losses and parameters."""
    from videonavqa_b200.optim import FusedClipAdam
    ma = _small_model(1)
    mb = copy.deepcopy(ma)
    for ca, cb in zip(ma.conv1x1_layers, mb.conv1x1_layers):      # deepcopy keeps the plain list (quirk Q1)
        assert ca is not cb and torch.equal(ca.weight, cb.weight)
    v, q, _, q_lens, labels = O.synthetic_batch(4, c_in=64, vocab=30, classes=5, seed=2, Lq=16, q_len_range=(3, 16))
    v_lens = torch.tensor([35, 22, 22, 6])
    batch = [t.cuda() for t in (v, q, v_lens, q_lens)]
    labels = labels.cuda()
    oa = torch.optim.Adam(ma.parameters(), lr=1e-3)
    ob = FusedClipAdam(mb.parameters(), lr=1e-3, max_norm=1.0).attach(mb)
    assert ob.flat.is_intact()
    for step in range(3):
        la = torch.nn.functional.cross_entropy(_fwd(ma, batch), labels, reduction="sum")
        oa.zero_grad(set_to_none=True)
        la.backward()
        torch.nn.utils.clip_grad_norm_(ma.parameters(), 1.0)
        oa.step()
        lb = torch.nn.functional.cross_entropy(_fwd(mb, batch), labels, reduction="sum")
        ob.zero_grad(set_to_none=True)
        lb.backward()
        # p.grad is a view of the bucket (no copy)
        p0 = next(mb.parameters())
        assert p0.grad.data_ptr() == ob.flat.grad_views[0].data_ptr()
        ob.step()
        assert abs(la.item() - lb.item()) <= 2e-4 * abs(la.item()), (step, la.item(), lb.item())
    # Adam divides by sqrt(v): in the first steps an element's update is ~lr * sign(g), so elements whose gradient is
    # rounding noise (bias / statistic reductions use atomics) may move in opposite directions - bound the relative L2
    # and the worst element by 2 * lr * steps
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        d = (pa.detach() - pb.detach()).double()
        rel = float(d.norm() / pa.detach().double().norm().clamp_min(1e-12))
        if n not in ("fc_attn_1.bias", "fc_hidden_attn.weight", "fc_hidden_attn.bias"):   # mathematically zero gradients
            assert rel <= 2e-2, (n, rel)      # zero-initialised biases ARE their first three updates
        assert float(d.abs().max()) <= 2 * 1e-3 * 3 + 1e-6, (n, float(d.abs().max()))
    # state_dict still has the reference layout and loads into a fresh module
    sd = mb.state_dict()
    mc = _small_model(3)
    mc.load_state_dict(sd)
    assert torch.equal(mc.conv_init.weight, mb.conv_init.weight)


def _fwd(m, batch):
    m.init_hidden()
    return m(*batch)
