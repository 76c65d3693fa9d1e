"""Shared machinery of the drop-in modules: parameter containers with the reference's registration order and
initialisation, the autograd bridge into videonavqa_b200.engine, and the carried question-LSTM state."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import engine, ops

# Operand storage of the tensor-core kernels when neither VNQA_OPERANDS nor set_operands() says otherwise.  fp16 has the
# same 10-bit mantissa as tf32 and passes the same gates (tests/test_gpu_fp16.py: logits 1.4e-4 from the fp32 reference
# at the benchmark shape, 100 % argmax agreement over 4096 clips, the same gradient table) at twice the MMA rate.
DEFAULT_OPERANDS = "fp16"


def reference_init(module):
    """The reference's weights_init (models/film_attn_pt_stem.py:111-127): Xavier-uniform weight + zero bias
    for every Linear / Conv2d; LSTM: Xavier W_ih, orthogonal W_hh, zero b_ih, forget-gate slice of the biases = 1."""
    if isinstance(module, (nn.Linear, nn.Conv2d)):
        nn.init.xavier_uniform_(module.weight.data)
        module.bias.data.zero_()
    if isinstance(module, nn.LSTM):
        nn.init.xavier_uniform_(module.weight_ih_l0)
        nn.init.orthogonal_(module.weight_hh_l0)
        hidden = module.hidden_size
        with torch.no_grad():
            module.bias_hh_l0[hidden:2 * hidden] = 1.0
            module.bias_ih_l0.zero_()


class _FilmFunction(torch.autograd.Function):
    """One autograd node for the whole hot path: forward and backward are hand-written CUDA pipelines."""

    @staticmethod
    def forward(ctx, mod, need_grad, v_input, q_input, v_lens, q_lens, *params):
        ctx.set_materialize_grads(False)
        P = dict(zip(mod._param_names, params))            # a trailing autograd anchor (bucket mode) is not a parameter
        logits, saved = mod._forward_impl(P, v_input, q_input, v_lens, q_lens, save=need_grad)
        ctx.mod, ctx.saved, ctx.P = mod, saved, P
        ctx.n_inputs = 6 + len(params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        mod = ctx.mod
        if dlogits is None:
            return (None,) * ctx.n_inputs
        bucket = getattr(mod, "_grad_bucket", None)
        if bucket:
            # flat-bucket mode (videonavqa_b200.optim.FlatBuffers): one memset, the kernels write every registered
            # parameter's gradient into its slice, the reducer / optimizer consume the bucket in place
            mod._grad_flat.zero_()
        grads = mod._backward_impl(ctx.P, ctx.saved, dlogits)
        hook = getattr(mod, "_grad_hook", None)
        if hook is not None and hasattr(hook, "wait"):
            hook.wait()     # in-flight gradient all-reduces (videonavqa_b200.dp) must land before autograd reads them
        ctx.saved = None
        out = []
        for n in mod._param_names:
            g = grads.get(n)
            if bucket and g is not None:
                # p.grad IS the bucket slice (no copy by AccumulateGrad).  Overwrite semantics: the reference zeroes
                # the gradients every iteration (eval/q_and_v_eval.py:139); accumulation across backward calls is
                # not supported in this mode.  The unregistered conv1x1_layers (Q1: never optimised, but they do
                # receive .grad) get theirs assigned the same way.  Autograd sees none of the parameters in this mode
                # (forward() hands detached aliases + a fresh anchor leaf to this node): their AccumulateGrad nodes
                # remember the stream they were first used on and make the engine synchronise with it after every
                # backward, which is illegal while a CUDA graph is being captured on another stream
                # (videonavqa_b200.graph) if the module has trained eagerly before.
                mod._leaf_params[n].grad = bucket[n] if n in bucket else g
                g = None
            out.append(g)
        if bucket:
            mod._bucket_filled = True
        return (None,) * 6 + tuple(out) + (None,) * (ctx.n_inputs - 6 - len(out))


def check_channels(num_input_channels, num_res_block_channels, hidden_size):
    """The tensor-core kernels stage operands in 32-channel TMA boxes and the recurrent kernels are instantiated for
    hidden sizes 32 / 64 / 128 (every configuration the reference ships: 512 -> 512/1024 channels, hidden 128)."""
    if num_input_channels % 32 or num_res_block_channels % 32:
        raise ValueError("videonavqa_b200 needs num_input_channels and num_res_block_channels to be multiples of 32 "
                         f"(got {num_input_channels}, {num_res_block_channels})")
    if hidden_size not in (32, 64, 128):
        raise ValueError(f"videonavqa_b200 supports hidden_size 32, 64 or 128 (got {hidden_size})")


class FilmModuleBase(nn.Module):
    """Behaviour common to FiLMAttnPretrainedStem / FiLMGlobalPoolingPretrainedStem (and the multi-hop variant)."""

    kind = "?"

    # ---- parameter plumbing -------------------------------------------------------------------------------
    def _collect_params(self):
        """(names, tensors): registered parameters in registration order, then the unregistered conv1x1_layers
        (quirk Q1: they are not in state_dict()/parameters() but take part in forward and receive .grad)."""
        names, tensors = [], []
        for n, p in self.named_parameters():
            names.append(n)
            tensors.append(p)
        for k, conv in enumerate(self.conv1x1_layers):
            names += [f"conv1x1_layers.{k}.weight", f"conv1x1_layers.{k}.bias"]
            tensors += [conv.weight, conv.bias]
        return names, tensors

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        # The reference moves conv1x1_layers / film_hidden to the GPU by hand at construction
        # (film_attn_pt_stem.py:102-103,137-138); follow the module so that .cuda() after construction also works.
        for conv in getattr(self, "conv1x1_layers", []):
            conv._apply(fn)
        fh = getattr(self, "film_hidden", None)
        if fh is not None:
            self.film_hidden = (fn(fh[0]), fn(fh[1]))
        if "_weight_cache" in self.__dict__:
            self.invalidate_weight_cache()
        return out

    def init_hidden(self):
        """Reset the carried question-LSTM state (models/film_attn_pt_stem.py:133-138)."""
        dev = self.embed.weight.device
        self.film_hidden = (torch.zeros(1, self.batch_size, self.hidden_size, device=dev),
                            torch.zeros(1, self.batch_size, self.hidden_size, device=dev))

    # ---- forward --------------------------------------------------------------------------------------------
    def _check_inputs(self, v_input, q_input, v_lens, q_lens):
        if not v_input.is_cuda:
            raise RuntimeError("videonavqa_b200 modules run on a CUDA sm_100 device only (no CPU fallback); "
                               "move the module and its inputs to the GPU")
        if v_input.shape[0] != self.batch_size:
            raise RuntimeError(f"batch size {v_input.shape[0]} != constructor batch_size {self.batch_size} "
                               "(the module is batch-size-locked like the reference, SURVEY Q7)")
        if tuple(v_input.shape[2:4]) != (10, 13):
            raise RuntimeError("stem feature maps must be 10 x 13 (hard-coded 130 in the reference)")

    def forward(self, v_input, q_input, v_lens, q_lens):
        self._check_inputs(v_input, q_input, v_lens, q_lens)
        ops.device_check()
        names, tensors = self._collect_params()
        self._param_names = names
        need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
        if need_grad and getattr(self, "_grad_bucket", None):
            # flat-bucket mode: backward assigns every .grad itself (see _FilmFunction.backward); the node hangs off a
            # fresh scalar leaf so that loss.backward() reaches it
            self._leaf_params = dict(zip(names, tensors))
            anchor = torch.zeros((), device=v_input.device, requires_grad=True)
            return _FilmFunction.apply(self, need_grad, v_input.float(), q_input, v_lens, q_lens,
                                       *[t.detach() for t in tensors], anchor)
        return _FilmFunction.apply(self, need_grad, v_input.float(), q_input, v_lens, q_lens, *tensors)

    def _host_lens(self, v_lens, q_lens, dev):
        """Clip / question lengths on the host.  The reference reads them from the device >= 70 times per forward
        (models/film_attn_pt_stem.py:159,203); here: no device read at all when the caller passes CPU tensors (they
        come from the DataLoader on the host), otherwise ONE read - skipped when the same, unmodified tensors were
        passed to the previous call (a training loop re-using its length tensors never waits for the GPU here)."""
        if not v_lens.is_cuda and not q_lens.is_cuda:
            return v_lens.tolist(), q_lens.tolist()
        key = (id(v_lens), v_lens._version, v_lens.data_ptr(), id(q_lens), q_lens._version, q_lens.data_ptr())
        hit = self.__dict__.get("_lens_cache")
        if hit is not None and hit[0] == key and hit[1]() is v_lens and hit[2]() is q_lens:
            return hit[3], hit[4]
        lens = torch.stack([v_lens.to(dev).long(), q_lens.to(dev).long()]).cpu()     # the one host sync
        v_host, q_host = lens[0].tolist(), lens[1].tolist()
        import weakref
        self.__dict__["_lens_cache"] = (key, weakref.ref(v_lens), weakref.ref(q_lens), v_host, q_host)
        return v_host, q_host

    def _prepare(self, v_input, q_input, v_lens, q_lens):
        dev = v_input.device
        v_host, q_host = self._host_lens(v_lens, q_lens, dev)
        L = q_input.shape[1]
        if min(q_host) < 1 or max(q_host) > L:
            raise ValueError(f"q_lens must lie in [1, {L}]")
        self._last_q_host = q_host
        q_lens_dev = q_lens.to(dev, non_blocking=True).long()
        q_lens32 = q_lens_dev.to(torch.int32)
        # film_hidden is kept in q_len-sorted order like the reference (film_attn_pt_stem.py:153-160); stable, so that
        # clips with equal question lengths keep their order across calls
        perm = q_lens_dev.sort(dim=0, descending=True, stable=True)[1]
        h = torch.empty(self.batch_size, self.hidden_size, device=dev)
        c = torch.empty_like(h)
        h[perm] = self.film_hidden[0][0].to(dev).detach()
        c[perm] = self.film_hidden[1][0].to(dev).detach()
        return v_host, q_lens32, perm, h, c, q_input.to(dev).long().contiguous()

    def _bn_state(self):
        return {"running_mean": self.bn_init.running_mean, "running_var": self.bn_init.running_var,
                "num_batches_tracked": self.bn_init.num_batches_tracked}

    def _finish(self, perm, h, c):
        self.film_hidden = (h[perm].unsqueeze(0), c[perm].unsqueeze(0))

    # ---- operand precision of the tensor-core kernels ---------------------------------------------------------
    @property
    def precision(self):
        """engine.Precision of this module: 'tf32' (fp32 storage, tf32 MMA operands) or 'fp16' (half trunk activations
        and operands, twice the MMA rate; fp32 accumulation, statistics, LSTMs and gradients).  Default from the
        environment variable VNQA_OPERANDS (DEFAULT_OPERANDS when unset); set with ``set_operands()``."""
        p = self.__dict__.get("_precision")
        if p is None:
            p = engine.Precision(os.environ.get("VNQA_OPERANDS", DEFAULT_OPERANDS),
                                 float(os.environ.get("VNQA_LOSS_SCALE", "4096")))
            self.__dict__["_precision"] = p
        return p

    def set_operands(self, operands, loss_scale=4096.0):
        self.__dict__["_precision"] = engine.Precision(operands, loss_scale)
        return self

    def invalidate_weight_cache(self, registered_only=False):
        """Drop the packed-weight caches (engine.WeightCache).  Needed after the parameters were written through
        ``.data`` or raw pointers (weight broadcasts, reference-style ``m.weight.data.fill_()``, the fused optimizer):
        those writes do not bump the autograd version counter the cache keys on."""
        if registered_only:
            self._cache.invalidate(list(self.parameters()))
        else:
            self._cache.invalidate()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_weight_cache()
        return out

    @property
    def _cache(self):
        c = self.__dict__.get("_weight_cache")
        if c is None:
            c = engine.WeightCache()
            self.__dict__["_weight_cache"] = c
        return c
