"""Pin the CPU oracle (oracle/film_oracle.py) against the golden vectors produced by the
unmodified reference modules (oracle/gen_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import film_oracle as O
from golden_util import load_case, rel_l2, norm_max_err

CASE_NAMES = ["attn_small", "attn_h128", "gp_small", "tmh_small"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_eval_logits(name):
    kind, kw, params, _, b, z = load_case(name)
    C = kw["num_res_block_channels"]
    st = O.new_film_state(kw["batch_size"], kw["hidden_size"])
    with torch.no_grad():
        logits, st = O.FORWARDS[kind](params, O.new_bn_state(C), st, b["v"], b["q"], b["v_lens"],
                                      b["q_lens"], training=False)
    assert norm_max_err(logits, torch.from_numpy(z["logits_eval"])) < 2e-5
    # the reference keeps film_hidden in q_len-sorted order (models/film_attn_pt_stem.py:153-160)
    perm = b["q_lens"].sort(0, descending=True)[1]
    assert norm_max_err(st[0][perm], torch.from_numpy(z["film_h_eval"])) < 2e-5
    assert norm_max_err(st[1][perm], torch.from_numpy(z["film_c_eval"])) < 2e-5


@pytest.mark.parametrize("name", CASE_NAMES)
def test_train_logits_grads_bn(name):
    kind, kw, params, grads, b, z = load_case(name)
    C = kw["num_res_block_channels"]
    bn = O.new_bn_state(C)
    st = O.new_film_state(kw["batch_size"], kw["hidden_size"])
    logits, loss, g, st = O.loss_and_grads(kind, params, bn, st, b["v"], b["q"], b["v_lens"], b["q_lens"],
                                           b["labels"], training=True)
    assert norm_max_err(logits, torch.from_numpy(z["logits_train"])) < 2e-5
    assert abs(float(loss) - float(z["loss_train"])) < 1e-4 * max(1.0, abs(float(z["loss_train"])))
    # BatchNorm running statistics are updated once per processed frame (Q6)
    assert int(bn["num_batches_tracked"]) == int(z["bn_num_batches_tracked"]) == int(b["v_lens"][0])
    assert norm_max_err(bn["running_mean"], torch.from_numpy(z["bn_running_mean"])) < 1e-5
    assert norm_max_err(bn["running_var"], torch.from_numpy(z["bn_running_var"])) < 1e-5
    perm = b["q_lens"].sort(0, descending=True)[1]
    assert norm_max_err(st[0][perm], torch.from_numpy(z["film_h_train"])) < 2e-5
    for k, gref in grads.items():
        assert k in g, k
        if float(gref.abs().max()) < 1e-6:
            # mathematically zero gradients (Q4: fc_hidden_attn cancels in the softmax; fc_attn_1.bias
            # when no frame is left unprocessed): rounding noise in both implementations
            assert float(g[k].abs().max()) < 1e-5, k
            continue
        assert rel_l2(g[k], gref) < 2e-4, (k, rel_l2(g[k], gref))


def test_state_dict_key_order_attn():
    """Registered-parameter layout of the CUDA branch."""
    _, _, params, _, _, z = load_case("attn_small")
    keys = [str(k) for k in z["state_dict_keys"]]
    # the golden file was produced on a CPU-only host: film_layer is absent there (Q2)
    assert keys[0] == "embed.weight" and "conv_init.weight" in keys and "out_linear.bias" in keys
    assert not any(k.startswith("conv1x1_layers") for k in keys)      # Q1
    assert "conv1x1_layers.0.weight" in params


def test_q5_unmasked_tail_frames():
    """Frames >= v_lens[0] get logit 0, not the mask (Q5): with v_lens[0]=30 the last five frames
    still receive attention weight."""
    kind, kw, params, _, b, _ = load_case("attn_small")
    assert int(b["v_lens"][0]) == 30
    cts = O.effective_batch_sizes(b["v_lens"], 35)
    assert len(cts) == 30 and cts[0] == 3 and cts[-1] == 1


def test_q3_state_carried_across_frames():
    kind, kw, params, _, b, _ = load_case("attn_small")
    st0 = O.new_film_state(kw["batch_size"], kw["hidden_size"])
    f0, st1 = O.film_values_lstm(params, "film_layer.", b["q"], b["q_lens"], st0)
    f1, _ = O.film_values_lstm(params, "film_layer.", b["q"], b["q_lens"], st1)
    assert float((f0 - f1).abs().max()) > 1e-6


@pytest.mark.parametrize("name", ["attn_small", "gp_small"])
def test_operand_rounding_emulation_stays_near_the_reference(name):
    """emulate_tf32 / emulate_fp16 (the checker modes used by the GPU gradient tests) only perturb the fp32 result at
    the 10-bit-mantissa level, and they are distinct roundings."""
    kind, kw, params, _, b, z = load_case(name)
    C = kw["num_res_block_channels"]
    outs = {}
    for tag, ctx in (("tf32", O.emulate_tf32()), ("fp16", O.emulate_fp16())):
        with ctx:
            lo, _, _, _ = O.loss_and_grads(kind, params, O.new_bn_state(C),
                                           O.new_film_state(kw["batch_size"], kw["hidden_size"]), b["v"], b["q"],
                                           b["v_lens"], b["q_lens"], b["labels"], training=True)
        outs[tag] = lo
        assert norm_max_err(lo, torch.from_numpy(z["logits_train"])) <= 2e-3, tag
    assert not torch.equal(outs["tf32"], outs["fp16"])
    # leaving the context restores the plain fp32 oracle
    lo, _, _, _ = O.loss_and_grads(kind, params, O.new_bn_state(C), O.new_film_state(kw["batch_size"], kw["hidden_size"]),
                                   b["v"], b["q"], b["v_lens"], b["q_lens"], b["labels"], training=True)
    assert norm_max_err(lo, torch.from_numpy(z["logits_train"])) <= 2e-5


def test_stem_oracle_matches_the_reference_module(golden_dir):
    """oracle/stem_oracle.py against the output of the UNMODIFIED reference ObjDetectCNN (oracle/gen_golden_stem.py),
    fp32 and fp64; the operand-rounding emulation stays within the bound the GPU tests use."""
    import hashlib
    import numpy as np
    from oracle import stem_oracle as S
    g = np.load(os.path.join(golden_dir, "stem_small.npz"))
    P, x = S.synthetic_stem_case(*[int(v) for v in g["case"]])
    assert hashlib.sha256(x.numpy().tobytes()).digest() == g["x_sha256"].tobytes(), "numpy RNG drift"
    for k, v in P.items():
        assert np.array_equal(v.numpy(), g["p." + k]), k
    y = S.stem_forward(P, x)
    scale = float(np.abs(g["y"]).max())
    assert float((y - torch.from_numpy(g["y"])).abs().max()) <= 5e-5 * scale          # fp32: summation order only
    y64 = S.stem_forward(P, x, dtype=torch.float64)
    assert float((y64 - torch.from_numpy(g["y64"])).abs().max()) <= 1e-10 * scale
    for emulate in ("fp16", "tf32"):
        ye = S.stem_forward(P, x, emulate=emulate).double()
        rel = float((ye - y64).norm() / y64.norm())
        assert 1e-5 < rel <= 1.5e-3, (emulate, rel)


@pytest.mark.parametrize("name", ["attn_small", "gp_small", "tmh_small"])
def test_oracle_size_independent_properties(name):
    """Properties of the path that hold at any size (and that the full-size GPU tests rely on): in eval mode a clip's
    logits depend on nothing but that clip (BatchNorm uses running statistics), frames behind v_len and tokens behind
    q_len are never read, and - for the models whose question attention is masked by packing (attn, gp) - the question
    padding length does not matter."""
    kind, kw, params, _, b, _ = load_case(name)
    B, H = kw["batch_size"], kw["hidden_size"]
    C = kw["num_res_block_channels"]
    fwd = O.FORWARDS[kind]

    def run(v, q, v_lens, q_lens):
        with torch.no_grad():
            return fwd(params, O.new_bn_state(C), O.new_film_state(v.shape[0], H), v, q, v_lens, q_lens,
                       training=False)[0]
    base = run(b["v"], b["q"], b["v_lens"], b["q_lens"])
    # (1) garbage in the frames behind each clip's end and in the tokens behind each question's end changes nothing
    v2, q2 = b["v"].clone(), b["q"].clone()
    for i in range(B):
        v2[i, ..., int(b["v_lens"][i]):] = 123.0
        q2[i, int(b["q_lens"][i]):] = 1 + (i % (kw["vocab_size"] - 1))
    if kind != "tmh":            # the multi-hop decoder attends over pad positions too (quirk Q10)
        assert torch.equal(run(v2, q2, b["v_lens"], b["q_lens"]), base)
    else:
        assert torch.equal(run(v2, b["q"], b["v_lens"], b["q_lens"]), base)
    # (2) clip independence: replace the other clips' content (same lengths: the effective batch per frame is part of
    # the reference's bookkeeping) and clip 0's logits stay put
    v3, q3 = b["v"].clone(), b["q"].clone()
    v3[1:] = torch.rand_like(v3[1:])
    q3[1:] = torch.where(q3[1:] > 0, torch.randint_like(q3[1:], 1, kw["vocab_size"]), q3[1:])
    out3 = run(v3, q3, b["v_lens"], b["q_lens"])
    assert torch.allclose(out3[0], base[0], rtol=0, atol=2e-6 * float(base.abs().max()))
    assert not torch.allclose(out3[1], base[1], atol=1e-3 * float(base.abs().max()))
