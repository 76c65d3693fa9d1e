"""Forward / backward orchestration of the FiLM hot path on top of the C-ABI kernels.

The reference runs a Python loop over the 35 frames with ~25 launches each
(models/film_attn_pt_stem.py:201-251).  Here every layer runs ONCE over all n_img = sum(v_lens) frame images
("frame-batched"): frames only couple through the BatchNorm running statistics (sequential EMA, reproduced by
vnqa_bn_finalize_train) and through the video-independent question-LSTM chain (vnqa_lstm_chain_forward), so
this is exactly the same arithmetic per image.

Everything here is host-side plumbing (buffer allocation through PyTorch's caching allocator, launch order);
no tensor arithmetic is done with torch ops on the hot path.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from ._lib import (EPI_BIAS_RELU, EPI_BIAS_RELU_AFFINE, EPI_BIAS_RELU_STATS, EPI_DGRAD_RES, EPI_FILM, EPI_PARTIAL,
                   EPI_PLAIN)

IMG_PIX = 130
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class _Nvtx:
    """NVTX range per op group (SURVEY.md §5 tracing), on when VNQA_NVTX=1: shows the stages of a step in Nsight
    Systems / Compute timelines; a no-op otherwise."""

    on = os.environ.get("VNQA_NVTX") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Nvtx.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _Nvtx.on:
            torch.cuda.nvtx.range_pop()


class Precision:
    """Operand / trunk-activation storage of the tensor-core kernels (include/vnqa_b200.h VNQA_ACT_*).

    ``tf32``: fp32 storage holding tf32-representable values (kind::tf32 MMA).
    ``fp16``: half storage - the same 10-bit mantissa - for the NHWC input, the BatchNorm output, every ResBlock tensor
    and the saved pre-FiLM activation (kind::f16 MMA: twice the rate, half the bytes).  BatchNorm input and statistics,
    FiLM parameters, split-K partials, both LSTMs, the attention tail and every parameter gradient stay fp32.  Backward
    signals inside the trunk are stored multiplied by ``loss_scale`` (a power of two; they reach 1e-7): d feats is
    scaled when it is converted to fp16, every kernel that leaves the fp16 domain multiplies by 1 / loss_scale."""

    def __init__(self, operands="tf32", loss_scale=4096.0):
        if operands not in ("tf32", "fp16"):
            raise ValueError(f"operands must be 'tf32' or 'fp16' (got {operands!r})")
        self.name = operands
        self.fp16 = operands == "fp16"
        self.act = 1 if self.fp16 else 0
        self.dtype = torch.float16 if self.fp16 else torch.float32
        self.loss_scale = float(loss_scale) if self.fp16 else 1.0
        self.unscale = 1.0 / self.loss_scale
        self.bk = 64 if self.fp16 else 32           # reduction elements per MMA stage row
        self.key = "h." if self.fp16 else ""        # prefix of the packed-weight cache keys


TF32 = Precision("tf32")


class Plan:
    """Ragged-batch bookkeeping (models/film_attn_pt_stem.py:190-210, 245-251) computed once per forward on the
    host from v_lens (ONE device->host read instead of the reference's >= 35)."""

    def __init__(self, v_lens_host, B, T, device):
        v = np.asarray(v_lens_host, dtype=np.int64)
        if v.shape != (B,):
            raise ValueError(f"v_lens must have shape ({B},), got {v.shape}")
        if (np.diff(v) > 0).any():
            raise ValueError("clips must be sorted by descending length (models/film_attn_pt_stem.py:184-186)")
        if v.min() < 1 or v.max() > T:
            raise ValueError(f"v_lens must lie in [1, {T}]")
        self.B, self.T = B, T
        self.n_frames = int(v[0])
        t_idx = np.arange(self.n_frames)
        frame_ct = (v[None, :] > t_idx[:, None]).sum(1).astype(np.int32)          # ct_i
        frame_off = np.concatenate([[0], np.cumsum(frame_ct)[:-1]]).astype(np.int64)
        self.n_img = int(frame_ct.sum())
        img_of = np.full((B, T), -1, dtype=np.int32)
        img_frame = np.empty(self.n_img, dtype=np.int32)
        img_clip = np.empty(self.n_img, dtype=np.int32)
        for t in range(self.n_frames):
            ct = int(frame_ct[t])
            img_of[:ct, t] = frame_off[t] + np.arange(ct)
            img_frame[frame_off[t]:frame_off[t] + ct] = t
            img_clip[frame_off[t]:frame_off[t] + ct] = np.arange(ct)
        film_row = img_frame * B + img_clip       # row of the (n_frames*B, 2*C*NB) FiLM matrix
        feat_row = img_clip * T + img_frame       # row of the (B*T, A) attention feature matrix
        packed = np.concatenate([img_of.reshape(-1), img_frame, film_row, feat_row, frame_ct,
                                 v.astype(np.int32), img_clip]).astype(np.int32)
        dev = torch.from_numpy(packed)
        if torch.device(device).type == "cuda":
            dev = dev.pin_memory().to(device, non_blocking=True)
        o = 0
        self.img_of = dev[o:o + B * T]; o += B * T
        self.img_frame = dev[o:o + self.n_img]; o += self.n_img
        self.film_row = dev[o:o + self.n_img]; o += self.n_img
        self.feat_row = dev[o:o + self.n_img]; o += self.n_img
        self.frame_ct = dev[o:o + self.n_frames]; o += self.n_frames
        self.v_lens32 = dev[o:o + B]; o += B
        self.img_clip = dev[o:o + self.n_img]
        self.frame_ct_host = frame_ct


_PLANS = {}


def get_plan(v_lens_host, B, T, device):
    """Plans are pure functions of (v_lens, B, T): keep the last few (their index arrays are read-only device tensors),
    so a loop over same-shaped batches neither rebuilds them nor copies them to the device again."""
    dev = torch.device(device)
    key = (tuple(int(x) for x in v_lens_host), B, T, dev.type, dev.index if dev.index is not None else
           (torch.cuda.current_device() if dev.type == "cuda" else -1))
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) >= 16:
            _PLANS.pop(next(iter(_PLANS)))
        plan = _PLANS[key] = Plan(v_lens_host, B, T, device)
    return plan


class GradSink(dict):
    """name -> gradient.  ``buffer(name, like)`` hands out the tensor a gradient is written into: a view of the flat
    gradient bucket when one is installed (videonavqa_b200.optim.FlatBuffers - zeroed once at the start of backward,
    all-reduced and consumed by the optimizer in place), else a fresh zero tensor.  ``sink[name] = t`` publishes a
    finished gradient and calls ``hook(name, t)`` so that a data-parallel reducer can start its all-reduce while the
    rest of the backward pass is still running (videonavqa_b200.dp)."""

    def __init__(self, hook=None, bucket=None):
        super().__init__()
        self._hook = hook
        self._bucket = bucket

    def buffer(self, name, like):
        v = self._bucket.get(name) if self._bucket is not None else None
        return v if v is not None else torch.zeros_like(like)

    def __setitem__(self, name, tensor):
        super().__setitem__(name, tensor)
        if self._hook is not None:
            self._hook(name, tensor)


class WeightCache:
    """Kernel-friendly repacks of the PyTorch-layout parameters; derived data, rebuilt when a parameter's version
    counter or storage changes (load_state_dict, optimizer step, .cuda()) or after ``invalidate()`` - needed when the
    weights are written through ``.data`` or raw pointers, which do not bump the version counter (broadcasts, the
    fused optimizer)."""

    def __init__(self):
        self._c = {}
        self._gen = {}          # data_ptr -> generation, bumped by invalidate(tensors)
        self._all = 0

    def get(self, key, tensor, build):
        ptr = tensor.data_ptr()
        tag = (ptr, tensor._version, tuple(tensor.shape), self._all, self._gen.get(ptr, 0))
        hit = self._c.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, build(tensor))
            self._c[key] = hit
        return hit[1]

    def invalidate(self, tensors=None):
        if tensors is None:
            self._all += 1
            return
        for t in tensors:
            self._gen[t.data_ptr()] = self._gen.get(t.data_ptr(), 0) + 1


def _choose_splits(n_img, k_total, bk=32):
    """split-K factor for the fc_embed_attn GEMM: smallest divisor of 130 that fills the SMs (K per split must be a
    multiple of the MMA stage depth); falls back to the largest legal one."""
    n_tiles = (n_img + 207) // 208
    legal = [s for s in (1, 2, 5, 10, 13, 26, 65, 130) if k_total % s == 0 and (k_total // s) % bk == 0]
    for s in legal:
        if n_tiles * s >= 140:
            return s
    return legal[-1]


# ======================================================================================== trunk forward
def lstm_generator_forward(P, cache, plan, q_input, q_lens32, film_h, film_c, save, nb, C):
    """FiLM generator of the attention / global-pooling models (compute_film_values, film_attn_pt_stem.py:144-181):
    chained question LSTM -> Linear -> ReLU.  Returns (film matrix, generator context)."""
    enc, lin = "film_layer.0.", "film_layer.1."
    B = plan.B
    gx = ops.embed_gates(q_input, q_lens32, P["embed.weight"], P[enc + "weight_ih_l0"], P[enc + "bias_ih_l0"],
                         P[enc + "bias_hh_l0"])
    whh_t = cache.get(enc + "whh_t", P[enc + "weight_hh_l0"], lambda w: w.t().contiguous())
    c0 = film_c.clone() if save else None
    h_pass, sg, sc, sh = ops.lstm_chain_forward(gx, q_lens32, whh_t, film_h, film_c, plan.n_frames, save=save)
    H = h_pass.shape[-1]
    film = ops.linear(h_pass.view(plan.n_frames * B, H), P[lin + "weight"], P[lin + "bias"], relu=True)
    return film, dict(h_pass=h_pass, gx=gx, sg=sg, sc=sc, sh=sh, c0=c0, enc_prefix=enc, film_lin_prefix=lin)


def lstm_generator_backward(P, cache, ctx, dfilm, q_input, q_lens32, sink):
    """Backward of lstm_generator_forward (the ReLU mask was already applied by film_backward)."""
    plan = ctx["plan"]
    enc, lin = ctx["enc_prefix"], ctx["film_lin_prefix"]
    h_pass = ctx["h_pass"]
    H = h_pass.shape[-1]
    hp2 = h_pass.view(plan.n_frames * plan.B, H)
    sink[lin + "weight"] = ops.linear_wgrad(dfilm, hp2, out=sink.buffer(lin + "weight", P[lin + "weight"]))
    sink[lin + "bias"] = ops.colsum(dfilm, out=sink.buffer(lin + "bias", P[lin + "bias"]))
    dh_pass = ops.linear_dgrad(dfilm, P[lin + "weight"])
    question_backward(P, ctx, dh_pass, q_input, q_lens32, sink, enc)


def trunk_forward(P, bn, cache, plan, v_input, q_input, q_lens32, film_h, film_c, training, save,
                  gen_fwd=lstm_generator_forward, prec=TF32):
    """conv_init+ReLU+BN, question encoder + FiLM generator, FiLM ResBlocks.  Returns (x_last, ctx)."""
    dev = v_input.device
    B, c_in = v_input.shape[0], v_input.shape[1]
    C = P["conv_init.weight"].shape[0]
    nb = 0
    while f"film_pipeline.{nb}.weight" in P:
        nb += 1
    n_img = plan.n_img
    ctx = {"plan": plan, "C": C, "nb": nb, "c_in": c_in, "prec": prec}
    act, pk = prec.act, prec.key

    # K0: frame-image-major NHWC
    with _Nvtx("vnqa.K0.frames_to_nhwc"):
        x0 = ops.frames_to_nhwc(v_input, plan.img_of, n_img, act)

    # ---- question encoder on a side stream (independent of the video until the first FiLM epilogue)
    main = torch.cuda.current_stream()
    side = _side_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side), _Nvtx("vnqa.K6.film_generator.fwd"):
        film, gctx = gen_fwd(P, cache, plan, q_input, q_lens32, film_h, film_c, save, nb, C)
        film.record_stream(main)        # allocated on the side stream, consumed on the main stream
        for t in gctx.values():
            if torch.is_tensor(t):
                t.record_stream(main)
    film_ready = torch.cuda.Event()
    film_ready.record(side)

    # ---- conv_init -> ReLU -> BatchNorm (:211)
    pack = lambda w: ops.pack_conv_weight(w, act=act)     # noqa: E731
    w_init = cache.get(pk + "conv_init", P["conv_init.weight"], pack)
    if training:
        y0 = torch.empty(n_img, IMG_PIX, C, device=dev, dtype=torch.float32)
        stat = torch.zeros(2, plan.n_frames, C, device=dev, dtype=torch.float64)
        ops.conv_gemm(w_init, x0, taps=9, cin=c_in, cout=C, n_img=n_img, epi=EPI_BIAS_RELU_STATS, out=y0, act=act,
                      bias=P["conv_init.bias"], img_frame=plan.img_frame, stat_sum=stat[0], stat_sq=stat[1])
        scale, shift, mean, rstd = ops.bn_finalize_train(stat[0], stat[1], plan.frame_ct, plan.n_frames,
                                                         P["bn_init.weight"], P["bn_init.bias"], bn["running_mean"],
                                                         bn["running_var"], BN_MOMENTUM, BN_EPS)
        bn["num_batches_tracked"] += plan.n_frames
        x = ops.bn_apply(y0, plan.img_frame, scale, shift, act)
        ctx.update(y0=y0, mean=mean, rstd=rstd)
    else:
        scale, shift = ops.bn_eval_affine(P["bn_init.weight"], P["bn_init.bias"], bn["running_mean"],
                                          bn["running_var"], BN_EPS)
        x = torch.empty(n_img, IMG_PIX, C, device=dev, dtype=prec.dtype)
        ops.conv_gemm(w_init, x0, taps=9, cin=c_in, cout=C, n_img=n_img, epi=EPI_BIAS_RELU_AFFINE, out=x, act=act,
                      bias=P["conv_init.bias"], scale=scale, shift=shift)

    # ---- FiLM ResBlocks (:217-241)
    xs, ress, zs = [x], [], []
    for k in range(nb):
        if _Nvtx.on:
            torch.cuda.nvtx.mark(f"vnqa.resblock.{k}.fwd")
        w1 = cache.get(f"{pk}c1x1.{k}", P[f"conv1x1_layers.{k}.weight"], pack)
        w3 = cache.get(f"{pk}c3x3.{k}", P[f"film_pipeline.{k}.weight"], pack)
        res = torch.empty_like(x)
        ops.conv_gemm(w1, x, taps=1, cin=C, cout=C, n_img=n_img, epi=EPI_BIAS_RELU, out=res, act=act,
                      bias=P[f"conv1x1_layers.{k}.bias"])
        if k == 0:
            main.wait_event(film_ready)
        z = torch.empty_like(x) if save else None
        out = torch.empty_like(x)
        ops.conv_gemm(w3, res, taps=9, cin=C, cout=C, n_img=n_img, epi=EPI_FILM, out=out, act=act,
                      bias=P[f"film_pipeline.{k}.bias"], film=film, film_row=plan.film_row, film_ld=film.stride(0),
                      alpha_off=2 * k * C, res=res, z_out=z)
        x = out
        if save:
            xs.append(x); ress.append(res); zs.append(z)
    if nb == 0:
        main.wait_event(film_ready)
    if save:
        ctx.update(x0=x0, xs=xs, ress=ress, zs=zs, film=film)
        ctx.update(gctx)
    return x, ctx


_SIDE = {}
SIDE_BWD = os.environ.get("VNQA_SIDE_BWD", "1") == "1"


def _side_stream(dev):
    if os.environ.get("VNQA_NO_SIDE_STREAM") == "1":     # bring-up switch: everything on the caller's stream
        return torch.cuda.current_stream()
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


# ======================================================================================== trunk backward
def trunk_backward(P, cache, ctx, dx_last, q_input, q_lens32, grads, gen_bwd=lstm_generator_backward, mask_relu=True):
    """Backward of trunk_forward given d(loss)/d(x_last); fills `grads` (name -> tensor).  mask_relu: the FiLM
    generator ends in a ReLU (attention / global-pooling models)."""
    plan, C, nb, c_in = ctx["plan"], ctx["C"], ctx["nb"], ctx["c_in"]
    prec = ctx["prec"]
    act, pk, un = prec.act, prec.key, prec.unscale      # dx_last and every trunk signal carry prec.loss_scale
    n_img, B = plan.n_img, plan.B
    dev = dx_last.device
    film = ctx["film"]
    dfilm = torch.zeros_like(film)
    dout = dx_last
    main = torch.cuda.current_stream()
    side = _side_stream(dev)
    gen_done = None

    def generator_backward():
        """FiLM generator backward (+ chained-LSTM BPTT + embedding), as soon as dfilm is complete (after block 0's
        FiLM backward).  Latency-bound (64 CTAs for ~2.4 ms) and independent of the remaining convolution backward,
        so it runs on the side stream.  The overlap is not free - the persistent tensor-core kernels use a static tile
        schedule, so clusters that find their SMs taken start their tile list late (block 0's dgrad: 3.3 -> up to
        6.9 ms) - but it still beats running it in line (measured, fp16: 80.1 vs 81.1 ms per step).
        VNQA_SIDE_BWD=0 runs it in line on the compute stream."""
        nonlocal gen_done
        if not SIDE_BWD:
            with _Nvtx("vnqa.K6.film_generator.bwd"):
                local = GradSink(None, grads._bucket if isinstance(grads, GradSink) else None)
                gen_bwd(P, cache, ctx, dfilm, q_input, q_lens32, local)
            gen_done = (None, local)
            return
        side.wait_stream(main)
        with torch.cuda.stream(side), _Nvtx("vnqa.K6.film_generator.bwd"):
            local = GradSink(None, grads._bucket if isinstance(grads, GradSink) else None)
            gen_bwd(P, cache, ctx, dfilm, q_input, q_lens32, local)
            for t in local.values():
                t.record_stream(main)
            dfilm.record_stream(side)
            gen_done = (torch.cuda.Event(), local)
            gen_done[0].record(side)

    if nb == 0:
        generator_backward()
    for k in reversed(range(nb)):
        if _Nvtx.on:
            torch.cuda.nvtx.mark(f"vnqa.resblock.{k}.bwd")
        res, z, x_in = ctx["ress"][k], ctx["zs"][k], ctx["xs"][k]
        db3 = grads.buffer(f"film_pipeline.{k}.bias", P[f"film_pipeline.{k}.bias"])
        dz = ops.film_backward(dout, z, film, plan.film_row, 2 * k * C, dfilm, db3, mask_relu=mask_relu, act=act,
                               grad_unscale=un)
        if k == 0:
            generator_backward()
        grads[f"film_pipeline.{k}.bias"] = db3
        dw3 = grads.buffer(f"film_pipeline.{k}.weight", P[f"film_pipeline.{k}.weight"])
        ops.conv_wgrad(dz, res, dw3, n_img=n_img, cout=C, cin=C, taps=9, act=act, grad_unscale=un)
        grads[f"film_pipeline.{k}.weight"] = dw3
        w3t = cache.get(f"{pk}c3x3t.{k}", P[f"film_pipeline.{k}.weight"], lambda w: ops.pack_conv_weight(w, True, act))
        db1 = grads.buffer(f"conv1x1_layers.{k}.bias", P[f"conv1x1_layers.{k}.bias"])
        dpre = torch.empty_like(res)
        ops.conv_gemm(w3t, dz, taps=9, cin=C, cout=C, n_img=n_img, epi=EPI_DGRAD_RES, out=dpre, add=dout, res=res,
                      dbias=db1, act=act, grad_unscale=un)
        del dz
        grads[f"conv1x1_layers.{k}.bias"] = db1
        dw1 = grads.buffer(f"conv1x1_layers.{k}.weight", P[f"conv1x1_layers.{k}.weight"])
        ops.conv_wgrad(dpre, x_in, dw1, n_img=n_img, cout=C, cin=C, taps=1, act=act, grad_unscale=un)
        grads[f"conv1x1_layers.{k}.weight"] = dw1
        w1t = cache.get(f"{pk}c1x1t.{k}", P[f"conv1x1_layers.{k}.weight"], lambda w: ops.pack_conv_weight(w, True, act))
        dx = torch.empty_like(res)
        ops.conv_gemm(w1t, dpre, taps=1, cin=C, cout=C, n_img=n_img, epi=EPI_PLAIN, out=dx, act=act)
        del dpre
        dout = dx
    # ---- BatchNorm + ReLU + conv_init
    db0 = grads.buffer("conv_init.bias", P["conv_init.bias"])
    dgamma = grads.buffer("bn_init.weight", P["bn_init.weight"])
    dbeta = grads.buffer("bn_init.bias", P["bn_init.bias"])
    dy0 = ops.bn_backward(dout, ctx["y0"], plan.img_frame, plan.frame_ct, plan.n_frames, ctx["mean"], ctx["rstd"],
                          P["bn_init.weight"], db0, dgamma, dbeta, act=act, grad_unscale=un)
    grads["conv_init.bias"] = db0
    grads["bn_init.weight"] = dgamma
    grads["bn_init.bias"] = dbeta
    dw0 = grads.buffer("conv_init.weight", P["conv_init.weight"])
    ops.conv_wgrad(dy0, ctx["x0"], dw0, n_img=n_img, cout=C, cin=c_in, taps=9, act=act, grad_unscale=un)
    grads["conv_init.weight"] = dw0
    # ---- join the side stream; its gradients are published (and all-reduced, if a hook is set) from the main stream
    if gen_done[0] is not None:
        main.wait_event(gen_done[0])
    for n_, t_ in gen_done[1].items():
        grads[n_] = t_


def question_backward(P, ctx, dh_pass, q_input, q_lens32, grads, enc, dh_steps=None):
    """BPTT through the chained question LSTM + embedding (autograd of compute_film_values, :144-171).
    dh_pass: gradient at the state after each pass; dh_steps: gradient at every step output (multi-hop model)."""
    sg, sc, sh = ctx["sg"], ctx["sc"], ctx["sh"]
    n_pass, B, L, H4 = sg.shape
    H = H4 // 4
    dgates, dgx = ops.lstm_chain_backward(dh_pass, q_lens32, P[enc + "weight_hh_l0"], sg, sc, ctx["c0"],
                                          dh_steps=dh_steps)
    steps = n_pass * B * L
    grads[enc + "weight_hh_l0"] = ops.linear_wgrad(dgates.view(steps, H4), sh.view(steps, H), ksplit=64,
                                                   out=grads.buffer(enc + "weight_hh_l0", P[enc + "weight_hh_l0"]))
    dgx2 = dgx.view(B * L, H4)
    xemb = ops.embed_gather(q_input, q_lens32, P["embed.weight"])
    E = xemb.shape[-1]
    grads[enc + "weight_ih_l0"] = ops.linear_wgrad(dgx2, xemb.view(B * L, E), ksplit=8,
                                                   out=grads.buffer(enc + "weight_ih_l0", P[enc + "weight_ih_l0"]))
    bsum = ops.colsum(dgx2, out=grads.buffer(enc + "bias_ih_l0", P[enc + "bias_ih_l0"]))
    grads[enc + "bias_ih_l0"] = bsum
    grads[enc + "bias_hh_l0"] = grads.buffer(enc + "bias_hh_l0", P[enc + "bias_hh_l0"]).copy_(bsum)
    dxemb = ops.linear_dgrad(dgx2, P[enc + "weight_ih_l0"])
    dembed = grads.buffer("embed.weight", P["embed.weight"])
    ops.embed_backward(q_input, q_lens32, dxemb, dembed)
    grads["embed.weight"] = dembed


# ======================================================================================== attention model
def attn_forward(P, bn, cache, v_input, q_input, v_lens_host, q_lens32, film_h, film_c, T, training, save, prec=TF32):
    """FiLMAttnPretrainedStem.forward (models/film_attn_pt_stem.py:188-301)."""
    dev = v_input.device
    B = v_input.shape[0]
    plan = get_plan(v_lens_host, B, T, dev)
    x, ctx = trunk_forward(P, bn, cache, plan, v_input, q_input, q_lens32, film_h, film_c, training, save, prec=prec)
    C, n_img = ctx["C"], plan.n_img
    A = P["fc_attn_1.weight"].shape[1]
    # fc_embed_attn over the NHWC-flattened map (:244): split-K tensor-core GEMM + reduction that scatters each
    # frame image to its (clip, frame) row of the zero-padded feature matrix (:245-256)
    wfc = cache.get(prec.key + "fc_embed", P["fc_embed_attn.weight"], lambda w: _fc_operand(w, C, prec))
    K = IMG_PIX * C
    splits = _choose_splits(n_img, K, prec.bk)
    partial = torch.empty(splits, n_img, A, device=dev, dtype=torch.float32)
    ops.conv_gemm(wfc, x, taps=1, cin=K // splits, cout=A, n_img=n_img, epi=EPI_PARTIAL, out=partial, flat=True,
                  x_ld=K, splits=splits, act=prec.act)
    feats = torch.zeros(B, T, A, device=dev, dtype=torch.float32)
    ops.splitk_reduce(partial, splits, n_img, A, P["fc_embed_attn.bias"], plan.feat_row, feats)
    wih_t = cache.get("tail_wih_t", P["lstm_attn.weight_ih"], lambda w: w.t().contiguous())
    whh_t = cache.get("tail_whh_t", P["lstm_attn.weight_hh"], lambda w: w.t().contiguous())
    hs, coefs, cx, gates, cs = ops.attn_tail_forward(
        feats, plan.v_lens32, plan.n_frames, P["fc_attn_1.weight"], P["fc_attn_1.bias"], P["fc_hidden_attn.weight"],
        P["fc_hidden_attn.bias"], wih_t, whh_t, P["lstm_attn.bias_ih"], P["lstm_attn.bias_hh"], save=save)
    logits = ops.linear(hs.view(B, T * A), P["out_linear.weight"], P["out_linear.bias"])
    if save:
        ctx.update(feats=feats, hs=hs, coefs=coefs, cx=cx, gates=gates, cs=cs, A=A, T=T, x_last=x)
    return logits, ctx


def _fc_operand(w, C, prec):
    """fc_embed_attn weight (A, C*130) permuted to the NHWC column order, in the operand type."""
    wp = ops.permute_fc_weight(w, C)
    return ops.to_half(wp) if prec.fp16 else wp


def attn_backward(P, cache, ctx, dlogits, q_input, q_lens32, grad_hook=None, grad_bucket=None):
    plan, C, A, T = ctx["plan"], ctx["C"], ctx["A"], ctx["T"]
    prec = ctx["prec"]
    B, n_img = plan.B, plan.n_img
    dev = dlogits.device
    grads = GradSink(grad_hook, grad_bucket)
    buf = lambda n: grads.buffer(n, P[n])     # noqa: E731
    hs = ctx["hs"]
    hs2 = hs.view(B, T * A)
    dlogits = dlogits.contiguous()
    grads["out_linear.weight"] = ops.linear_wgrad(dlogits, hs2, out=buf("out_linear.weight"))
    grads["out_linear.bias"] = ops.colsum(dlogits, out=buf("out_linear.bias"))
    dhs = ops.linear_dgrad(dlogits, P["out_linear.weight"]).view(B, T, A)
    dw1, db1 = buf("fc_attn_1.weight"), buf("fc_attn_1.bias")
    dwh, dbh = buf("fc_hidden_attn.weight"), buf("fc_hidden_attn.bias")
    dfeats, dgates = ops.attn_tail_backward(ctx["feats"], plan.v_lens32, plan.n_frames, P["fc_attn_1.weight"],
                                            P["fc_hidden_attn.weight"], P["lstm_attn.weight_ih"],
                                            P["lstm_attn.weight_hh"], hs, ctx["coefs"], ctx["gates"], ctx["cs"], dhs,
                                            dw1, db1, dwh, dbh)
    for n_, t_ in (("fc_attn_1.weight", dw1), ("fc_attn_1.bias", db1), ("fc_hidden_attn.weight", dwh),
                   ("fc_hidden_attn.bias", dbh)):
        grads[n_] = t_
    dg2 = dgates.view(B * T, 4 * A)
    grads["lstm_attn.weight_ih"] = ops.linear_wgrad(dg2, ctx["cx"].view(B * T, A), out=buf("lstm_attn.weight_ih"))
    hprev = torch.zeros_like(hs)
    hprev[:, 1:].copy_(hs[:, :-1])               # data movement only: h_{i-1} with h_{-1} = 0
    grads["lstm_attn.weight_hh"] = ops.linear_wgrad(dg2, hprev.view(B * T, A), out=buf("lstm_attn.weight_hh"))
    bsum = ops.colsum(dg2, out=buf("lstm_attn.bias_ih"))
    grads["lstm_attn.bias_ih"] = bsum
    grads["lstm_attn.bias_hh"] = buf("lstm_attn.bias_hh").copy_(bsum)
    # fc_embed_attn backward on the valid (clip, frame) rows
    df_rows = ops.gather_rows(dfeats.view(B * T, A), plan.feat_row, n_img, A, round_tf32=not prec.fp16)
    grads["fc_embed_attn.bias"] = ops.colsum(df_rows, out=buf("fc_embed_attn.bias"))
    if prec.fp16:
        df_rows = ops.to_half(df_rows, prec.loss_scale)      # d feats enters the fp16 domain scaled
    K = IMG_PIX * C
    x_last = ctx["x_last"].view(n_img, K)
    dwp = torch.empty(A, K, device=dev, dtype=torch.float32)
    ops.conv_wgrad(df_rows, x_last, dwp, n_img=n_img, cout=A, cin=K, taps=1, flat=True, act=prec.act,
                   grad_unscale=prec.unscale)
    grads["fc_embed_attn.weight"] = ops.permute_fc_weight(dwp, C, inverse=True, out=buf("fc_embed_attn.weight"))
    wfc_t = cache.get(prec.key + "fc_embed_t", P["fc_embed_attn.weight"],
                      lambda w: ops.pack_matrix(ops.permute_fc_weight(w, C, n_pad=w.shape[0]), transpose=True,
                                                act=prec.act))
    dx_last = torch.empty(n_img, IMG_PIX, C, device=dev, dtype=prec.dtype)
    ops.conv_gemm(wfc_t, df_rows, taps=1, cin=A, cout=K, n_img=n_img, epi=EPI_PLAIN, out=dx_last, flat=True, x_ld=A,
                  act=prec.act)
    trunk_backward(P, cache, ctx, dx_last, q_input, q_lens32, grads)
    return grads


# ======================================================================================== global-pooling model
def gp_forward(P, bn, cache, v_input, q_input, v_lens_host, q_lens32, film_h, film_c, training, save, prec=TF32):
    """FiLMGlobalPoolingPretrainedStem.forward (models/film_global_pooling_pt_stem.py:180-238)."""
    dev = v_input.device
    B, T = v_input.shape[0], v_input.shape[-1]
    plan = get_plan(v_lens_host, B, T, dev)
    x, ctx = trunk_forward(P, bn, cache, plan, v_input, q_input, q_lens32, film_h, film_c, training, save, prec=prec)
    return _gp_tail_forward(P, cache, plan, x, ctx, B, T, save)


def _gp_tail_forward(P, cache, plan, x, ctx, B, T, save):
    """relu(c1x1_tail) -> max over frames -> Linear (models/film_global_pooling_pt_stem.py:228-238)."""
    dev = x.device
    C, n_img = ctx["C"], plan.n_img
    tail = P["c1x1_tail.weight"].shape[0]
    prec = ctx["prec"]
    wt = cache.get(prec.key + "c1x1_tail", P["c1x1_tail.weight"], lambda w: ops.pack_conv_weight(w, act=prec.act))
    y = torch.empty(n_img, IMG_PIX, tail, device=dev, dtype=torch.float32)     # the tail map leaves the MMA domain: fp32
    ops.conv_gemm(wt, x, taps=1, cin=C, cout=tail, n_img=n_img, epi=EPI_BIAS_RELU, out=y, bias=P["c1x1_tail.bias"],
                  no_round=True, act=prec.act, out_f32=prec.fp16)
    pooled, arg = ops.gp_maxpool(y, plan.img_of, B, T, tail)
    logits = ops.linear(pooled, P["out_linear.weight"], P["out_linear.bias"])
    if save:
        ctx.update(pooled=pooled, arg=arg, tail=tail, x_last=x)
    return logits, ctx


def gp_backward(P, cache, ctx, dlogits, q_input, q_lens32, grad_hook=None, grad_bucket=None):
    grads = GradSink(grad_hook, grad_bucket)
    dx_last = _gp_tail_backward(P, cache, ctx, dlogits, grads)
    trunk_backward(P, cache, ctx, dx_last, q_input, q_lens32, grads)
    return grads


def _gp_tail_backward(P, cache, ctx, dlogits, grads):
    plan, C, tail = ctx["plan"], ctx["C"], ctx["tail"]
    B, n_img = plan.B, plan.n_img
    dlogits = dlogits.contiguous()
    pooled = ctx["pooled"]
    grads["out_linear.weight"] = ops.linear_wgrad(dlogits, pooled, out=grads.buffer("out_linear.weight", P["out_linear.weight"]))
    grads["out_linear.bias"] = ops.colsum(dlogits, out=grads.buffer("out_linear.bias", P["out_linear.bias"]))
    dpooled = ops.linear_dgrad(dlogits, P["out_linear.weight"])
    dy = ops.gp_maxpool_backward(dpooled, pooled, ctx["arg"], B, tail, n_img)     # through max + ReLU
    dy2 = dy.view(n_img * IMG_PIX, tail)
    grads["c1x1_tail.bias"] = ops.colsum(dy2, out=grads.buffer("c1x1_tail.bias", P["c1x1_tail.bias"]))
    x_last = ctx["x_last"]
    prec = ctx["prec"]
    dwt = grads.buffer("c1x1_tail.weight", P["c1x1_tail.weight"])
    if not prec.fp16:
        # tail conv gradients: tail is tiny (<= 64 channels) -> fp32 SGEMMs
        ops.linear_wgrad(dy2, x_last.view(n_img * IMG_PIX, C), ksplit=32, out=dwt.view(tail, C))
        grads["c1x1_tail.weight"] = dwt
        return ops.linear_dgrad(dy2, P["c1x1_tail.weight"].view(tail, C)).view(n_img, IMG_PIX, C)
    # fp16 trunk: x_last is half, so the tail gradients run on the tensor cores; dy enters the fp16 domain scaled
    dy_h = ops.to_half(dy, prec.loss_scale)
    ops.conv_wgrad(dy_h, x_last, dwt, n_img=n_img, cout=tail, cin=C, taps=1, act=1, grad_unscale=prec.unscale)
    grads["c1x1_tail.weight"] = dwt
    wtt = cache.get("h.c1x1_tail_t", P["c1x1_tail.weight"], lambda w: ops.pack_conv_weight(w, True, 1))
    dx_last = torch.empty(n_img, IMG_PIX, C, device=dy.device, dtype=torch.float16)
    ops.conv_gemm(wtt, dy_h, taps=1, cin=tail, cout=C, n_img=n_img, epi=EPI_PLAIN, out=dx_last, act=1)
    return dx_last


# ======================================================================================== multi-hop model
def tmh_generator_forward(P, cache, plan, q_input, q_lens32, film_h, film_c, save, nb, C, max_q_len=None):
    """compute_film_encoding + decode_to_film_values of TimeMultiHopFiLMPretrainedStem
    (models/time_multi_hop_pt_stem.py:124-184) for every frame image at once."""
    enc = "q_encoder."
    B = plan.B
    gx = ops.embed_gates(q_input, q_lens32, P["embed.weight"], P[enc + "weight_ih_l0"], P[enc + "bias_ih_l0"],
                         P[enc + "bias_hh_l0"])
    whh_t = cache.get(enc + "whh_t", P[enc + "weight_hh_l0"], lambda w: w.t().contiguous())
    c0 = film_c.clone() if save else None
    h_pass, sg, sc, sh, states = ops.lstm_chain_forward(gx, q_lens32, whh_t, film_h, film_c, plan.n_frames, save=save,
                                                        want_states=True)
    W = max_q_len                            # pad_packed_sequence pads to the longest question of the batch (:136)
    hvec, enc_out, enc_stat, coefs = ops.tmh_hops_forward(
        states, h_pass, plan.img_frame, plan.img_clip, W, nb, P["encoder_norm.weight"], P["encoder_norm.bias"],
        P["fc_hidden_attn.weight"], P["fc_hidden_attn.bias"])
    N = 2 * C * nb
    film = torch.zeros(plan.n_frames * B, N, device=gx.device, dtype=torch.float32)
    conds, stats = [], []
    for k in range(nb):
        cond = ops.linear(hvec[k], P["fc_attn_out.weight"], P["fc_attn_out.bias"])        # full width, every hop (Q10)
        stat = ops.ln_slice_forward(cond, P["decoder_norm.weight"], P["decoder_norm.bias"], plan.film_row, film,
                                    2 * k * C, 2 * C)
        if save:
            conds.append(cond); stats.append(stat)
    g = dict(h_pass=h_pass, gx=gx, sg=sg, sc=sc, sh=sh, c0=c0, enc_prefix=enc)
    if save:
        g.update(states=states, hvec=hvec, enc_out=enc_out, enc_stat=enc_stat, coefs=coefs, W=W,
                 conds=torch.stack(conds), ln_stats=torch.stack(stats))
    return film, g


def tmh_generator_backward(P, cache, ctx, dfilm, q_input, q_lens32, sink):
    plan, C, nb = ctx["plan"], ctx["C"], ctx["nb"]
    dev = dfilm.device
    hvec, W = ctx["hvec"], ctx["W"]
    H = hvec.shape[-1]
    N = 2 * C * nb
    dgam = sink.buffer("decoder_norm.weight", P["decoder_norm.weight"])
    dbet = sink.buffer("decoder_norm.bias", P["decoder_norm.bias"])
    dw_out = sink.buffer("fc_attn_out.weight", P["fc_attn_out.weight"])
    db_out = sink.buffer("fc_attn_out.bias", P["fc_attn_out.bias"])
    dhvec = torch.empty_like(hvec)
    for k in range(nb):
        dcond = ops.ln_slice_backward(ctx["conds"][k], P["decoder_norm.weight"], plan.film_row, dfilm, 2 * k * C, 2 * C,
                                      ctx["ln_stats"][k], dgam, dbet)
        ops.linear_wgrad(dcond, hvec[k], out=dw_out)         # accumulated over the hops
        ops.colsum(dcond, out=db_out)
        dhvec[k].copy_(ops.linear_dgrad(dcond, P["fc_attn_out.weight"]))
    sink["decoder_norm.weight"] = dgam
    sink["decoder_norm.bias"] = dbet
    sink["fc_attn_out.weight"] = dw_out
    sink["fc_attn_out.bias"] = db_out
    dw_att = sink.buffer("fc_hidden_attn.weight", P["fc_hidden_attn.weight"])
    db_att = sink.buffer("fc_hidden_attn.bias", P["fc_hidden_attn.bias"])
    deg = sink.buffer("encoder_norm.weight", P["encoder_norm.weight"])
    deb = sink.buffer("encoder_norm.bias", P["encoder_norm.bias"])
    dstates = ops.tmh_hops_backward(ctx["states"], ctx["h_pass"], plan.img_frame, plan.img_clip, q_lens32, W, nb,
                                    P["encoder_norm.weight"], P["fc_hidden_attn.weight"], hvec, ctx["enc_out"],
                                    ctx["enc_stat"], ctx["coefs"], dhvec, dw_att, db_att, deg, deb)
    sink["fc_hidden_attn.weight"] = dw_att
    sink["fc_hidden_attn.bias"] = db_att
    sink["encoder_norm.weight"] = deg
    sink["encoder_norm.bias"] = deb
    question_backward(P, ctx, None, q_input, q_lens32, sink, ctx["enc_prefix"], dh_steps=dstates)


def tmh_forward(P, bn, cache, v_input, q_input, v_lens_host, q_lens_host, q_lens32, film_h, film_c, training, save,
                prec=TF32):
    """TimeMultiHopFiLMPretrainedStem.forward (models/time_multi_hop_pt_stem.py:191-250): conv trunk with the
    multi-hop FiLM generator, then the global-pooling tail."""
    dev = v_input.device
    B, T = v_input.shape[0], v_input.shape[-1]
    plan = get_plan(v_lens_host, B, T, dev)
    max_q_len = int(max(q_lens_host))
    x, ctx = trunk_forward(P, bn, cache, plan, v_input, q_input, q_lens32, film_h, film_c, training, save,
                           gen_fwd=lambda *a: tmh_generator_forward(*a, max_q_len=max_q_len), prec=prec)
    return _gp_tail_forward(P, cache, plan, x, ctx, B, T, save)


def tmh_backward(P, cache, ctx, dlogits, q_input, q_lens32, grad_hook=None, grad_bucket=None):
    grads = GradSink(grad_hook, grad_bucket)
    dx_last = _gp_tail_backward(P, cache, ctx, dlogits, grads)
    trunk_backward(P, cache, ctx, dx_last, q_input, q_lens32, grads, gen_bwd=tmh_generator_backward, mask_relu=False)
    return grads
