"""GPU A/B check of the fp16 candidate path (run on a B200 box; NOT part of the test suite yet).

    make -C experimental/f16 && python tests/tools/check_model_f16.py [--full] [--bench]

1. the four golden cases: logits vs the committed reference outputs (<= 1e-3), gradients vs the oracle
   with fp16 operand rounding emulated (oracle = checker only);
2. --full: eval.sh-size logits (C = 1024, 5 blocks, 4 ragged clips) vs the fp32 oracle;
3. --bench: training-step time of the fp16 path next to the shipped tf32 path, batch 64.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "experimental", "f16"))

import engine_f16                                            # noqa: E402
from oracle import film_oracle as O                          # noqa: E402  (checker)
from golden_util import load_case, norm_max_err, rel_l2       # noqa: E402
from test_gpu_modules import all_named_params, build_module  # noqa: E402
from videonavqa_b200.models import FiLMAttnPretrainedStem    # noqa: E402
from videonavqa_b200.synth import synthetic_batch            # noqa: E402


def golden(name):
    kind, kw, params, grads_ref, b, z = load_case(name)
    m = engine_f16.install(build_module(kind, kw, params).train())
    m.init_hidden()
    lo = m(b["v"].cuda(), b["q"].cuda(), b["v_lens"].cuda(), b["q_lens"].cuda())
    loss = torch.nn.functional.cross_entropy(lo, b["labels"].cuda(), reduction="sum")
    loss.backward()
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
        g = named[k].grad.cpu()
        if float(gref.abs().max()) < 1e-6:
            continue
        e = rel_l2(g, g_emu[k])
        if e > worst:
            worst, worst_k = e, k
    gtol = 2e-2 if kind in ("gp", "tmh") else 2e-3      # max-pool argmax ties re-route whole gradient elements
    ltol = 2e-3 if name == "tmh_small" else 1e-3        # same allowance as tests/test_gpu_modules.py
    ok = e_ref <= ltol and e_emu <= 3e-4 and worst <= gtol
    print(f"{name}: logits vs reference {e_ref:.3e} (<= 1e-3), vs fp16-emulating oracle {e_emu:.3e} (<= 3e-4), "
          f"worst gradient rel-L2 vs emulating oracle {worst:.3e} [{worst_k}] (<= 2e-3)  {'ok' if ok else 'FAIL'}")
    return ok


def full():
    B = 4
    kw = dict(batch_size=B, q_embedding_size=128, nb_classes=70, num_input_channels=512, num_res_block_channels=1024,
              num_res_blocks=5, hidden_size=128, at_hidden_size=128, max_num_frames=35, vocab_size=134)
    torch.manual_seed(5)
    m = FiLMAttnPretrainedStem(**kw).cuda().eval()
    v, q, _, q_lens, labels = synthetic_batch(B, seed=9)
    v_lens = torch.tensor([6, 5, 3, 2])
    params = {k: t.detach().cpu() for k, t in all_named_params(m).items()}
    with torch.no_grad():
        m.init_hidden()
        lo32 = m(v.cuda(), q.cuda(), v_lens.cuda(), q_lens.cuda()).cpu()
        engine_f16.install(m)
        m.init_hidden()
        lo16 = m(v.cuda(), q.cuda(), v_lens.cuda(), q_lens.cuda()).cpu()
    bn = O.new_bn_state(1024)
    ref, _ = O.film_attn_forward(params, bn, O.new_film_state(B, 128), v, q, v_lens, q_lens, training=False)
    e32, e16 = norm_max_err(lo32, ref), norm_max_err(lo16, ref)
    print(f"eval.sh-size eval logits vs fp32 oracle: shipped tf32 path {e32:.3e}, fp16 candidate {e16:.3e} (<= 1e-3)")
    return e16 <= 1e-3


def bench():
    B = 64
    kw = dict(batch_size=B, q_embedding_size=128, nb_classes=70, num_input_channels=512, num_res_block_channels=1024,
              num_res_blocks=5, hidden_size=128, at_hidden_size=128, max_num_frames=35, vocab_size=134)
    v, q, v_lens, q_lens, labels = [t.cuda() for t in synthetic_batch(B, seed=0)]
    for tag in ("tf32 (shipped)", "fp16 (candidate)"):
        torch.manual_seed(1)
        m = FiLMAttnPretrainedStem(**kw).cuda().train()
        if tag.startswith("fp16"):
            engine_f16.install(m)
        opt = torch.optim.Adam(m.parameters(), lr=1e-4)

        def step():
            m.init_hidden()
            loss = torch.nn.functional.cross_entropy(m(v, q, v_lens, q_lens), labels, reduction="sum")
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step(); opt.zero_grad(set_to_none=True)
            return loss
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            last = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{tag}: {ms:.2f} ms/step, {B / ms * 1e3:.1f} clips/s, loss {float(last):.4f}")
        del m, opt
        torch.cuda.empty_cache()
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--bench", action="store_true")
    a = ap.parse_args()
    t0 = time.time()
    ok = all([golden(n) for n in ("attn_small", "attn_h128", "gp_small", "tmh_small")])
    if a.full:
        ok = full() and ok
    if a.bench:
        bench()
    print(f"{'ALL OK' if ok else 'FAILED'} ({time.time() - t0:.0f} s)")
    sys.exit(0 if ok else 1)
