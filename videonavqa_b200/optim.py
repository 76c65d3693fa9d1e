"""Global-norm gradient clipping fused into Adam over flat buffers.

The reference's loop body is ``clip_grad_norm(model.parameters(), 1.0); optimizer.step()`` with
``torch.optim.Adam(model.parameters(), lr)`` (eval/q_and_v_eval.py:137-139, 333): ~40 multi-tensor launches over 283 MB
of gradients.  ``FusedClipAdam`` does the same arithmetic in two HBM-bound launches of libvnqa_b200.so
(csrc/optim.cu): the registered parameters are re-homed into ONE contiguous fp32 buffer (each ``nn.Parameter`` becomes
a view of it, so ``state_dict()`` / ``load_state_dict()`` / the kernels keep working on the same storage), the moments
live in two more flat buffers, and the gradients are written by the backward kernels straight into a flat bucket of the
same layout (``FlatBuffers.grad_views`` handed to the module as ``_grad_bucket``) - which is also the buffer the
data-parallel reducer all-reduces in place (videonavqa_b200.dp).

It subclasses ``torch.optim.Adam``: ``state_dict()`` has the standard ``step`` / ``exp_avg`` / ``exp_avg_sq`` entries
and loads into / from a plain ``torch.optim.Adam`` (checkpoint envelope of eval/q_and_v_eval.py:148-156, 336-346).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

ALIGN = 4       # floats: every tensor starts 16-byte aligned inside the flat buffers


class FlatBuffers:
    """Registered parameters of a module in one contiguous fp32 buffer + a gradient bucket of the same layout."""

    def __init__(self, params):
        self.params = list(params)
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("all parameters must be fp32 on one device")
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.param = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(off, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.param[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.grad_views = [self.grad[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def is_intact(self):
        """False when something re-allocated a parameter (module.to(), .cuda(), a checkpoint loaded by assignment)."""
        base = self.param.data_ptr()
        return all(p.data_ptr() == base + 4 * o for p, o in zip(self.params, self.offsets))


class FusedClipAdam(torch.optim.Adam):
    """clip_grad_norm_(params, max_norm) + Adam.step() in two launches.  ``attach(module)`` makes the module's
    backward write its gradients into the flat bucket (no per-tensor gradient allocations, nothing to pack for the
    all-reduce); without it ``step()`` gathers ``p.grad`` into the bucket first."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, capturable=False):
        """capturable: keep the step count on the device (csrc/optim.cu vnqa_clip_adam_dev) so that ``step()`` does not
        depend on host state and can be captured in a CUDA graph (videonavqa_b200.graph)."""
        params = list(params)
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        if len(self.param_groups) != 1:
            raise ValueError("FusedClipAdam supports a single parameter group (the reference uses one)")
        self.max_norm = float(max_norm)
        self.flat = FlatBuffers(self.param_groups[0]["params"])
        dev = self.flat.param.device
        self._m = torch.zeros_like(self.flat.param)
        self._v = torch.zeros_like(self.flat.param)
        self._steps = 0
        self._n_partials = _lib.lib().vnqa_opt_num_partials()
        self._partials = torch.zeros(self._n_partials, device=dev, dtype=torch.float64)
        self._norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self._module = None
        self.capturable = bool(capturable)
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int64) if self.capturable else None
        self._install_state()

    # ---- state plumbing ---------------------------------------------------------------------------------------
    def _install_state(self):
        """Standard torch.optim.Adam state entries, as views of the flat moment buffers."""
        for p, o in zip(self.flat.params, self.flat.offsets):
            self.state[p] = {"step": torch.tensor(float(self._steps)),
                             "exp_avg": self._m[o:o + p.numel()].view(p.shape),
                             "exp_avg_sq": self._v[o:o + p.numel()].view(p.shape)}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the loaded moments are fresh tensors: copy them back into the flat buffers and re-point the state at the views
        steps = 0
        with torch.no_grad():
            for p, o in zip(self.flat.params, self.flat.offsets):
                st = self.state.get(p, {})
                if "exp_avg" in st:
                    self._m[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
                    self._v[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                    steps = max(steps, int(float(st["step"])))
        self._steps = steps
        if self.capturable:
            self._step_dev.fill_(steps)
        self._install_state()

    def state_dict(self):
        if self.capturable:
            # the count lives on the device (graph replays do not run Python): refresh the standard entries
            self._steps = int(self._step_dev.item())
            for p in self.flat.params:
                self.state[p]["step"] = torch.tensor(float(self._steps))
        return super().state_dict()

    def attach(self, module):
        """Route the module's backward into the gradient bucket (engine.GradSink) and let step() invalidate the
        module's packed-weight caches (the kernels update the parameters through raw pointers)."""
        names = {id(p): n for n, p in module.named_parameters()}
        module._grad_bucket = {names[id(p)]: v for p, v in zip(self.flat.params, self.flat.grad_views) if id(p) in names}
        module._grad_flat = self.flat.grad
        self._module = module
        return self

    @property
    def grad_norm(self):
        """Total gradient norm seen by the last step() (what clip_grad_norm_ returns); a device tensor, no sync."""
        return self._norm

    def zero_grad(self, set_to_none=True):
        # bucket mode: the backward pass zeroes the bucket itself; p.grad stays a view of it
        if self._module is None:
            super().zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("FusedClipAdam does not take a closure")
        f = self.flat
        if not f.is_intact():
            raise RuntimeError("a parameter left the flat buffer (module.to()/cuda() after the optimizer was built?); "
                               "rebuild FusedClipAdam")
        if self._module is None or not getattr(self._module, "_bucket_filled", False):
            # plain autograd gradients: gather them (data movement only)
            views, grads = [], []
            for p, v in zip(f.params, f.grad_views):
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    views.append(v); grads.append(p.grad)
            if views:
                torch._foreach_copy_(views, grads)
        g = self.param_groups[0]
        beta1, beta2 = g["betas"]
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        L = _lib.lib()
        P = lambda x: ctypes.c_void_p(x.data_ptr())     # noqa: E731
        if self.max_norm > 0:
            _lib.check(L.vnqa_sumsq_partials(P(f.grad), f.total, P(self._partials), st), "vnqa_sumsq_partials")
        if self.capturable:
            _lib.check(L.vnqa_clip_adam_dev(P(f.param), P(f.grad), P(self._m), P(self._v), f.total, P(self._partials),
                                            self._n_partials, 1.0, self.max_norm, g["lr"], beta1, beta2, g["eps"],
                                            P(self._step_dev), P(self._norm), st), "vnqa_clip_adam_dev")
        else:
            self._steps += 1
            _lib.check(L.vnqa_clip_adam(P(f.param), P(f.grad), P(self._m), P(self._v), f.total, P(self._partials),
                                        self._n_partials, 1.0, self.max_norm, g["lr"], beta1, beta2, g["eps"],
                                        self._steps, P(self._norm), st), "vnqa_clip_adam")
            for p in f.params:
                self.state[p]["step"] += 1
        if self._module is not None:
            self._module.invalidate_weight_cache(registered_only=True)
            self._module._bucket_filled = False
        return None
