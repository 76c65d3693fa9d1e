"""Print parity numbers of the CUDA modules against the golden vectors (reference outputs) - run on the B200 box.
    python tools/gpu_parity_report.py [case ...]
"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import load_case, rel_l2, norm_max_err
from videonavqa_b200.models import (FiLMAttnPretrainedStem, FiLMGlobalPoolingPretrainedStem,
                                    TimeMultiHopFiLMPretrainedStem)

CLS = {"attn": FiLMAttnPretrainedStem, "gp": FiLMGlobalPoolingPretrainedStem, "tmh": TimeMultiHopFiLMPretrainedStem}


def build(kind, kw, params):
    m = CLS[kind](**kw)
    sd = {k: v for k, v in params.items() if not k.startswith("conv1x1_layers")}
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith("bn_init.") for k in missing.missing_keys), missing
    m = m.cuda()
    with torch.no_grad():
        for k, conv in enumerate(m.conv1x1_layers):
            conv.weight.copy_(params[f"conv1x1_layers.{k}.weight"])
            conv.bias.copy_(params[f"conv1x1_layers.{k}.bias"])
    return m


def main():
    names = sys.argv[1:] or ["attn_small", "attn_h128", "gp_small", "tmh_small"]
    for name in names:
        kind, kw, params, grads, b, z = load_case(name)
        m = build(kind, kw, params)
        v, q, vl, ql, lab = (b[k].cuda() for k in ("v", "q", "v_lens", "q_lens", "labels"))
        m.eval(); m.init_hidden()
        with torch.no_grad():
            lo = m(v, q, vl, ql)
        torch.cuda.synchronize()
        print(f"[{name}] eval logits  normalised max err {norm_max_err(lo.cpu(), torch.from_numpy(z['logits_eval'])):.3e}"
              f"  film_h err {norm_max_err(m.film_hidden[0][0].cpu(), torch.from_numpy(z['film_h_eval'])):.3e}")
        m.train(); m.init_hidden()
        lo = m(v, q, vl, ql)
        loss = torch.nn.functional.cross_entropy(lo, lab, reduction="sum")
        loss.backward()
        torch.cuda.synchronize()
        print(f"[{name}] train logits normalised max err {norm_max_err(lo.detach().cpu(), torch.from_numpy(z['logits_train'])):.3e}"
              f"  loss {loss.item():.6f} vs {float(z['loss_train']):.6f}")
        print(f"[{name}] bn running_mean err {norm_max_err(m.bn_init.running_mean.cpu(), torch.from_numpy(z['bn_running_mean'])):.3e}"
              f" running_var err {norm_max_err(m.bn_init.running_var.cpu(), torch.from_numpy(z['bn_running_var'])):.3e}"
              f" tracked {int(m.bn_init.num_batches_tracked)} vs {int(z['bn_num_batches_tracked'])}")
        named = dict(m.named_parameters())
        for k, conv in enumerate(m.conv1x1_layers):
            named[f"conv1x1_layers.{k}.weight"] = conv.weight
            named[f"conv1x1_layers.{k}.bias"] = conv.bias
        for k, gref in grads.items():
            g = named[k].grad
            if g is None:
                print(f"   grad {k:32s} MISSING"); continue
            g = g.cpu()
            if float(gref.abs().max()) < 1e-6:
                print(f"   grad {k:32s} ~zero ref; max |g| {float(g.abs().max()):.3e}")
            else:
                print(f"   grad {k:32s} rel-l2 {rel_l2(g, gref):.3e}   normalised max {norm_max_err(g, gref):.3e}")


if __name__ == "__main__":
    main()
