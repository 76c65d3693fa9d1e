"""CPU oracle for the VideoNavQA FiLM hot path.  TEST INFRASTRUCTURE ONLY.

A plain, functional PyTorch-CPU restatement of the reference algorithm (no nn.Module from
the reference, no CUDA).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file; the product path in
``videonavqa_b200/`` never does (it raises if the CUDA library is missing).

Pinning: the reference repository ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference modules themselves, generated in the build
container by ``oracle/gen_golden.py`` (imports ``/root/reference/models``) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks logits, gradients, BatchNorm
running statistics and the carried question-LSTM state.

Every function cites the reference lines it follows (paths relative to the reference root).

Parameters are passed as a flat ``dict[str, Tensor]`` keyed like the reference
``state_dict`` (CUDA branch) plus the unregistered 1x1 convolutions under
``conv1x1_layers.<k>.weight|bias``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

NEG_MASK = float(-(1 << 31))  # models/film_attn_pt_stem.py:251

# --------------------------------------------------------------------------------------
# Optional emulation of the product's operand rounding.  The CUDA path feeds its tensor-core MMAs with
# tf32 operands (10-bit mantissa, round-to-nearest applied once by the producer of each tensor: packed
# weights, the NHWC input, the BatchNorm output, every ResBlock activation).  With ``emulate_tf32()`` active the
# oracle rounds at exactly those points (straight-through in backward), so that ReLU masks agree and a gradient
# comparison measures implementation errors rather than mask flips.  Default (off) = the plain fp32 reference.
# --------------------------------------------------------------------------------------
_EMULATE_TF32 = False


class emulate_tf32:
    def __enter__(self):
        global _EMULATE_TF32
        self._old, _EMULATE_TF32 = _EMULATE_TF32, True

    def __exit__(self, *a):
        global _EMULATE_TF32
        _EMULATE_TF32 = self._old


def tf32_round(t):
    """cvt.rna.tf32.f32: round to nearest (ties away) to a 10-bit mantissa."""
    i = t.detach().contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    return r


def fp16_round(t):
    """cvt.rn.f16.f32 and back: the same 10-bit mantissa as tf32 (round to nearest even), fp16 exponent range."""
    return t.detach().half().float()


_ROUND = tf32_round


class emulate_fp16(emulate_tf32):
    """Same rounding points as emulate_tf32(), fp16 storage (the kind::f16 candidate path, experimental/f16)."""

    def __enter__(self):
        global _ROUND
        super().__enter__()
        self._old_round, _ROUND = _ROUND, fp16_round

    def __exit__(self, *a):
        global _ROUND
        _ROUND = self._old_round
        super().__exit__(*a)


def _q(t):
    if not _EMULATE_TF32 or t.dtype != torch.float32:
        return t
    return t + (_ROUND(t) - t).detach()


# --------------------------------------------------------------------------------------
# Question encoder: embedding + packed LSTM with state carried across frames
# --------------------------------------------------------------------------------------
def lstm_cell(x_gates, h, c, w_hh, b_hh):
    """One LSTM step, gate order i,f,g,o (torch.nn.LSTM / LSTMCell semantics, used at
    models/film_attn_pt_stem.py:160,293).  ``x_gates`` already holds W_ih x + b_ih."""
    gates = x_gates + h @ w_hh.t() + b_hh
    i, f, g, o = gates.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def packed_lstm_pass(x_emb, q_lens, h0, c0, w_ih, w_hh, b_ih, b_hh):
    """One pass of the question LSTM over a batch of zero-padded questions with
    pack_padded_sequence semantics (models/film_attn_pt_stem.py:153-171): clip b advances
    only while t < q_lens[b]; the returned state is the state at each clip's own last token.
    The sort / un-sort of the reference (:153-165) is a pure permutation and is elided.

    x_emb: (B, L, E).  Returns (outputs (B, Lmax, H) zero-padded, h_last (B,H), c_last (B,H)).
    """
    B, L, _ = x_emb.shape
    lmax = int(q_lens.max())
    xg = x_emb @ w_ih.t() + b_ih  # (B, L, 4H)
    h, c = h0, c0
    outs = []
    for t in range(lmax):
        h2, c2 = lstm_cell(xg[:, t], h, c, w_hh, b_hh)
        m = (q_lens > t).to(x_emb.dtype).unsqueeze(1)
        h = m * h2 + (1 - m) * h
        c = m * c2 + (1 - m) * c
        outs.append(m * h2)  # pad_packed_sequence pads with zeros (:161)
    return torch.stack(outs, dim=1), h, c


def film_values_lstm(p, prefix, q_input, q_lens, state):
    """compute_film_values, LSTM branch (models/film_attn_pt_stem.py:144-181,
    models/film_global_pooling_pt_stem.py:137-173).  ``state`` = (h, c) each (B, H) carried
    from the previous frame (quirk Q3); returns (film (B, 2*C*NB) >= 0, new state)."""
    x = F.embedding(q_input, p["embed.weight"])
    _, h, c = packed_lstm_pass(
        x, q_lens, state[0], state[1],
        p[prefix + "0.weight_ih_l0"], p[prefix + "0.weight_hh_l0"],
        p[prefix + "0.bias_ih_l0"], p[prefix + "0.bias_hh_l0"])
    film = F.relu(F.linear(h, p[prefix + "1.weight"], p[prefix + "1.bias"]))
    return film, (h, c)


# --------------------------------------------------------------------------------------
# Per-frame FiLM ResBlock stack
# --------------------------------------------------------------------------------------
def effective_batch_sizes(v_lens, num_frames):
    """Effective batch per frame for clips sorted by descending length
    (models/film_attn_pt_stem.py:201-208): ct_i = #{b : v_lens[b] > i}; the loop stops at the
    first frame with no clip left."""
    cts = []
    for i in range(num_frames):
        ct = int((v_lens > i).sum())
        if ct == 0:
            break
        cts.append(ct)
    return cts


def conv_init_bn(p, bn_state, x, training):
    """bn_init(relu(conv_init(x))) (models/film_attn_pt_stem.py:211): conv3x3 pad 1, ReLU, THEN
    BatchNorm2d (eps 1e-5, momentum 0.1).  ``bn_state`` = dict(running_mean, running_var,
    num_batches_tracked) updated in place in training mode (quirk Q6)."""
    y = F.relu(F.conv2d(_q(x), _q(p["conv_init.weight"]), p["conv_init.bias"], padding=1))
    out = F.batch_norm(y, bn_state["running_mean"], bn_state["running_var"],
                       p["bn_init.weight"], p["bn_init.bias"],
                       training=training, momentum=0.1, eps=1e-5)
    if training:
        bn_state["num_batches_tracked"] += 1
    return _q(out)


def film_res_block(p, k, x, alphas, betas):
    """One FiLM ResBlock (models/film_attn_pt_stem.py:217-241):
    res = relu(conv1x1_k(x)); z = conv3x3_k(res); out = relu(alpha*z + beta) + res."""
    res = _q(F.relu(F.conv2d(x, _q(p[f"conv1x1_layers.{k}.weight"]), p[f"conv1x1_layers.{k}.bias"])))
    z = F.conv2d(res, _q(p[f"film_pipeline.{k}.weight"]), p[f"film_pipeline.{k}.bias"], padding=1)
    return _q(F.relu(alphas[:, :, None, None] * z + betas[:, :, None, None]) + res)


def film_stack(p, x, film, num_blocks):
    """The block loop with the FiLM slice bookkeeping (models/film_attn_pt_stem.py:215-241)."""
    C = p["film_pipeline.0.weight"].shape[0]
    start = 0
    for k in range(num_blocks):
        x = film_res_block(p, k, x, film[:, start:start + C], film[:, start + C:start + 2 * C])
        start += 2 * C
    return x


def _num_blocks(p):
    k = 0
    while f"film_pipeline.{k}.weight" in p:
        k += 1
    return k


# --------------------------------------------------------------------------------------
# FiLMAttnPretrainedStem.forward
# --------------------------------------------------------------------------------------
def film_attn_forward(p, bn_state, film_state, v_input, q_input, v_lens, q_lens, training):
    """models/film_attn_pt_stem.py:188-301.  Returns (logits (B, classes), new film_state).

    v_input (B, Cin, 10, 13, T) fp32, clips sorted by descending v_lens; film_state = (h, c)
    each (B, hidden) (the reference keeps a leading layer axis of 1)."""
    B, T = v_input.shape[0], v_input.shape[-1]
    nb = _num_blocks(p)
    A = p["fc_attn_1.weight"].shape[1]
    cts = effective_batch_sizes(v_lens, T)
    feats = v_input.new_zeros(B, T, A)                         # :245-256 zero padding
    masks = v_input.new_zeros(B, T, 1)                         # :194
    for i, ct in enumerate(cts):
        x = conv_init_bn(p, bn_state, v_input[:ct, :, :, :, i], training)   # :210-211
        film, film_state = film_values_lstm(p, "film_layer.", q_input, q_lens, film_state)  # :213
        x = film_stack(p, x, film[:ct], nb)                    # :215-241
        f = F.linear(x.reshape(ct, -1), _q(p["fc_embed_attn.weight"]), p["fc_embed_attn.bias"])  # :244
        feats[:ct, i] = f
        masks[ct:, i, 0] = NEG_MASK                            # :251 (processed frames only, Q5)
    # frame scores on valid (b, t) only, zero elsewhere (:268-281)
    scores = v_input.new_zeros(B, T, 1)
    for i, ct in enumerate(cts):
        scores[:ct, i] = F.linear(feats[:ct, i], p["fc_attn_1.weight"], p["fc_attn_1.bias"])
    h = v_input.new_zeros(B, A)
    c = v_input.new_zeros(B, A)
    hs = []
    for _ in range(T):                                         # :283-295
        v_i = F.linear(h, p["fc_hidden_attn.weight"], p["fc_hidden_attn.bias"]).view(B, 1, 1)
        coefs = torch.softmax(v_i + scores + masks, dim=1)     # :288
        ctxt = torch.bmm(coefs.permute(0, 2, 1), feats).view(B, A)   # :290
        xg = F.linear(ctxt, p["lstm_attn.weight_ih"], p["lstm_attn.bias_ih"])
        h, c = lstm_cell(xg, h, c, p["lstm_attn.weight_hh"], p["lstm_attn.bias_hh"])   # :293
        hs.append(h)
    hs = torch.stack(hs, dim=1).reshape(B, T * A)              # :298
    return F.linear(hs, p["out_linear.weight"], p["out_linear.bias"]), film_state   # :301


# --------------------------------------------------------------------------------------
# FiLMGlobalPoolingPretrainedStem.forward
# --------------------------------------------------------------------------------------
def gp_tail(p, x):
    """relu(c1x1_tail(x)) flattened channel-major (models/film_global_pooling_pt_stem.py:228-232)."""
    y = F.relu(F.conv2d(x, _q(p["c1x1_tail.weight"]), p["c1x1_tail.bias"]))
    return y.reshape(y.shape[0], -1)


def film_gp_forward(p, bn_state, film_state, v_input, q_input, v_lens, q_lens, training):
    """models/film_global_pooling_pt_stem.py:180-238.  Returns (logits, new film_state)."""
    B, T = v_input.shape[0], v_input.shape[-1]
    nb = _num_blocks(p)
    cts = effective_batch_sizes(v_lens, T)
    per_frame = []
    for i, ct in enumerate(cts):
        x = conv_init_bn(p, bn_state, v_input[:ct, :, :, :, i], training)
        film, film_state = film_values_lstm(p, "film_layer.", q_input, q_lens, film_state)
        x = film_stack(p, x, film[:ct], nb)
        y = gp_tail(p, x)
        per_frame.append(F.pad(y, (0, 0, 0, B - ct)))           # :230-232 zero rows
    pooled = torch.stack(per_frame, dim=0).max(dim=0)[0]        # :235
    return F.linear(pooled, p["out_linear.weight"], p["out_linear.bias"]), film_state   # :238


# --------------------------------------------------------------------------------------
# TimeMultiHopFiLMPretrainedStem.forward
# --------------------------------------------------------------------------------------
def tmh_forward(p, bn_state, film_state, v_input, q_input, v_lens, q_lens, training):
    """models/time_multi_hop_pt_stem.py:191-250 with compute_film_encoding (:124-158) and
    decode_to_film_values (:165-184).  Word attention is unmasked and the full-width
    fc_attn_out + LayerNorm is recomputed per block (quirk Q10)."""
    B, T = v_input.shape[0], v_input.shape[-1]
    nb = _num_blocks(p)
    C = p["film_pipeline.0.weight"].shape[0]
    H = p["q_encoder.weight_hh_l0"].shape[1]
    cts = effective_batch_sizes(v_lens, T)
    per_frame = []
    for i, ct in enumerate(cts):
        x = conv_init_bn(p, bn_state, v_input[:ct, :, :, :, i], training)
        emb = F.embedding(q_input, p["embed.weight"], padding_idx=0)
        states, hl, cl = packed_lstm_pass(
            emb, q_lens, film_state[0], film_state[1],
            p["q_encoder.weight_ih_l0"], p["q_encoder.weight_hh_l0"],
            p["q_encoder.bias_ih_l0"], p["q_encoder.bias_hh_l0"])
        film_state = (hl, cl)
        enc = F.layer_norm(hl[:ct], (H,), p["encoder_norm.weight"], p["encoder_norm.bias"], 1e-5)  # :148
        W = states.shape[1]
        hctx = enc.view(ct, 1, H).repeat(1, W, 1)               # :157-158
        rnn = states[:ct]
        start = 0
        for k in range(nb):
            res = _q(F.relu(F.conv2d(x, _q(p[f"conv1x1_layers.{k}.weight"]), p[f"conv1x1_layers.{k}.bias"])))
            z = F.conv2d(res, _q(p[f"film_pipeline.{k}.weight"]), p[f"film_pipeline.{k}.bias"], padding=1)
            e = hctx * rnn                                      # :170
            coefs = torch.softmax(F.linear(e, p["fc_hidden_attn.weight"], p["fc_hidden_attn.bias"]), dim=1)
            hvec = torch.bmm(coefs.view(ct, 1, W), e).view(ct, H)   # :175-176
            cond = F.linear(hvec, p["fc_attn_out.weight"], p["fc_attn_out.bias"])   # :179
            hctx = hvec.view(ct, 1, H).repeat(1, W, 1)          # :180-181
            film = F.layer_norm(cond, (cond.shape[1],), p["decoder_norm.weight"], p["decoder_norm.bias"], 1e-5)
            a, b = film[:, start:start + C], film[:, start + C:start + 2 * C]
            start += 2 * C
            x = _q(F.relu(a[:, :, None, None] * z + b[:, :, None, None]) + res)
        y = gp_tail(p, x)
        per_frame.append(F.pad(y, (0, 0, 0, B - ct)))
    pooled = torch.stack(per_frame, dim=0).max(dim=0)[0]
    return F.linear(pooled, p["out_linear.weight"], p["out_linear.bias"]), film_state


FORWARDS = {"attn": film_attn_forward, "gp": film_gp_forward, "tmh": tmh_forward}


# --------------------------------------------------------------------------------------
# Synthetic inputs; numpy RNG so that the GPU box regenerates them bit-exactly
# --------------------------------------------------------------------------------------
def synthetic_batch(B, c_in=512, T=35, Lq=56, vocab=134, classes=70, seed=0, ragged=False,
                    q_len_range=(5, 30)):
    import numpy as np
    rng = np.random.default_rng(seed)
    v = np.maximum(rng.standard_normal((B, c_in, 10, 13, T), dtype=np.float32), 0.0)
    q_lens = rng.integers(q_len_range[0], q_len_range[1] + 1, size=B)
    q = rng.integers(1, vocab, size=(B, Lq))
    q[np.arange(Lq)[None, :] >= q_lens[:, None]] = 0
    if ragged:
        raw = np.clip(rng.normal(75, 27, size=B), 12, 140)
        v_lens = np.minimum(np.ceil(raw / 4).astype(np.int64), T)
        v_lens = -np.sort(-v_lens)
    else:
        v_lens = np.full(B, T, dtype=np.int64)
    labels = rng.integers(0, classes, size=B)
    return (torch.from_numpy(v), torch.from_numpy(q.astype(np.int64)),
            torch.from_numpy(v_lens.astype(np.int64)), torch.from_numpy(q_lens.astype(np.int64)),
            torch.from_numpy(labels.astype(np.int64)))


def init_params(kind, c_in=512, C=1024, nb=5, hidden=128, at_hidden=128, embed=128, vocab=134,
                classes=70, tail=32, T=35, seed=1):
    """Random parameters with the reference's initialisation scheme (weights_init,
    models/film_attn_pt_stem.py:111-127): Xavier-uniform weights / zero bias for Linear and
    Conv2d, Xavier W_ih + orthogonal W_hh + forget-gate bias 1 for the LSTM.  Used for
    benchmarks and property tests; parity tests load the reference's own weights from
    tests/golden instead."""
    g = torch.Generator().manual_seed(seed)

    def xavier(*shape):
        fan_out, fan_in = shape[0], shape[1]
        rf = 1
        for s in shape[2:]:
            rf *= s
        a = (6.0 / ((fan_in + fan_out) * rf)) ** 0.5
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    p = {}
    p["embed.weight"] = torch.randn(vocab, embed, generator=g)
    if kind != "attn":
        p["embed.weight"][0] = 0
    p["conv_init.weight"] = xavier(C, c_in, 3, 3)
    p["conv_init.bias"] = torch.zeros(C)
    p["bn_init.weight"] = torch.ones(C)
    p["bn_init.bias"] = torch.zeros(C)
    enc = "film_layer.0." if kind != "tmh" else "q_encoder."
    p[enc + "weight_ih_l0"] = xavier(4 * hidden, embed)
    q, _ = torch.linalg.qr(torch.randn(4 * hidden, hidden, generator=g))
    p[enc + "weight_hh_l0"] = q.contiguous()
    p[enc + "bias_ih_l0"] = torch.zeros(4 * hidden)
    bhh = (torch.rand(4 * hidden, generator=g) * 2 - 1) / hidden ** 0.5
    bhh[hidden:2 * hidden] = 1.0
    p[enc + "bias_hh_l0"] = bhh
    if kind != "tmh":
        p["film_layer.1.weight"] = xavier(2 * C * nb, hidden)
        p["film_layer.1.bias"] = torch.zeros(2 * C * nb)
    else:
        p["encoder_norm.weight"] = torch.ones(hidden)
        p["encoder_norm.bias"] = torch.zeros(hidden)
        p["fc_hidden_attn.weight"] = xavier(1, hidden)
        p["fc_hidden_attn.bias"] = torch.zeros(1)
        p["fc_attn_out.weight"] = xavier(2 * C * nb, hidden)
        p["fc_attn_out.bias"] = torch.zeros(2 * C * nb)
        p["decoder_norm.weight"] = torch.ones(2 * C * nb)
        p["decoder_norm.bias"] = torch.zeros(2 * C * nb)
    for k in range(nb):
        p[f"film_pipeline.{k}.weight"] = xavier(C, C, 3, 3)
        p[f"film_pipeline.{k}.bias"] = torch.zeros(C)
        p[f"conv1x1_layers.{k}.weight"] = xavier(C, C, 1, 1)
        p[f"conv1x1_layers.{k}.bias"] = torch.zeros(C)
    if kind == "attn":
        p["fc_embed_attn.weight"] = xavier(at_hidden, 130 * C)
        p["fc_embed_attn.bias"] = torch.zeros(at_hidden)
        p["fc_attn_1.weight"] = xavier(1, at_hidden)
        p["fc_attn_1.bias"] = torch.zeros(1)
        p["fc_hidden_attn.weight"] = xavier(1, at_hidden)
        p["fc_hidden_attn.bias"] = torch.zeros(1)
        k_ = 1.0 / at_hidden ** 0.5
        for n, shape in (("weight_ih", (4 * at_hidden, at_hidden)), ("weight_hh", (4 * at_hidden, at_hidden)),
                         ("bias_ih", (4 * at_hidden,)), ("bias_hh", (4 * at_hidden,))):
            p["lstm_attn." + n] = (torch.rand(*shape, generator=g) * 2 - 1) * k_
        p["out_linear.weight"] = xavier(classes, T * at_hidden)
        p["out_linear.bias"] = torch.zeros(classes)
    else:
        p["c1x1_tail.weight"] = xavier(tail, C, 1, 1)
        p["c1x1_tail.bias"] = torch.zeros(tail)
        p["out_linear.weight"] = xavier(classes, 130 * tail)
        p["out_linear.bias"] = torch.zeros(classes)
    return p


def new_bn_state(C, dtype=torch.float32):
    return {"running_mean": torch.zeros(C, dtype=dtype), "running_var": torch.ones(C, dtype=dtype),
            "num_batches_tracked": torch.zeros((), dtype=torch.int64)}


def new_film_state(B, hidden, dtype=torch.float32):
    """init_hidden (models/film_attn_pt_stem.py:133-138)."""
    return (torch.zeros(B, hidden, dtype=dtype), torch.zeros(B, hidden, dtype=dtype))


def loss_and_grads(kind, p, bn_state, film_state, v, q, v_lens, q_lens, labels, training=True):
    """CrossEntropyLoss(reduction='sum') + backward (eval/q_and_v_eval.py:124,136; eval.sh:15).
    Returns (logits, loss, {name: grad}, new film_state)."""
    leaves = {k: t.detach().clone().requires_grad_(t.is_floating_point()) for k, t in p.items()}
    logits, st = FORWARDS[kind](leaves, bn_state, film_state, v, q, v_lens, q_lens, training)
    loss = F.cross_entropy(logits, labels, reduction="sum")
    loss.backward()
    grads = {k: t.grad for k, t in leaves.items() if t.grad is not None}
    return logits.detach(), loss.detach(), grads, (st[0].detach(), st[1].detach())
