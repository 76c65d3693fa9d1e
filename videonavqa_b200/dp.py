"""Clip-sharded data parallelism for the FiLM hot path; new work, the reference is single-GPU.

One process per GPU (torchrun / torch.distributed, NCCL over NVLink 5 / NVSwitch).  Clips are independent, so the
forward needs no collective; the only exchange is ONE gradient all-reduce per step:

  * SUM, not mean: the reference loss is CrossEntropyLoss(reduction='sum') (eval.sh:15), so N ranks x b clips is
    the same optimisation step as one rank x N*b clips (per-replica BatchNorm statistics, as in ordinary DDP);
  * overlapped with backward: the engine reports every gradient the moment it is final (engine.GradSink) and the
    all-reduce of that tensor is enqueued asynchronously (NCCL runs it on its own stream, in-switch NVLS reduction
    when available) while the remaining backward kernels keep the SMs busy; the handles are waited for at the end
    of backward, before the optimizer;
  * the unregistered conv1x1_layers (quirk Q1) are broadcast from rank 0 at start-up - nothing else would ever
    synchronise them - and are never optimised, exactly as in the reference.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def broadcast_module(module, src=0, group=None):
    """Make every replica start from rank `src`'s weights: registered parameters, buffers (BatchNorm running
    statistics) and the unregistered conv1x1_layers."""
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    for conv in getattr(module, "conv1x1_layers", []):
        tensors += [conv.weight.data, conv.bias.data]
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


class GradAllReducer:
    """Installs itself as the module's gradient hook; all-reduces each gradient as soon as the backward pass
    produces it.  Call ``wait()`` after ``loss.backward()`` and before ``optimizer.step()``."""

    def __init__(self, module, group=None, registered_only=True):
        self.group = group
        self.handles = []
        self.bytes = 0
        self._registered = {n for n, _ in module.named_parameters()} if registered_only else None
        module._grad_hook = self

    def __call__(self, name, tensor):
        if self._registered is not None and name not in self._registered:
            return      # conv1x1_layers: never optimised in the reference, gradients stay local
        self.handles.append(dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.bytes += tensor.numel() * tensor.element_size()

    def wait(self):
        for h in self.handles:
            h.wait()
        n = self.bytes
        if self.handles:
            self.last_bytes = n
        self.handles, self.bytes = [], 0
        return n


def shard_batch(v_input, q_input, v_lens, q_lens, labels, rank, world):
    """Clip-wise shard of a global batch; every shard is re-sorted by descending video length (a precondition of
    forward, eval/q_and_v_eval.py:113-116)."""
    B = v_input.shape[0]
    assert B % world == 0, "global batch must divide by the number of ranks"
    idx = torch.arange(rank, B, world)          # strided: every rank gets a similar length mix
    vl = v_lens[idx]
    order = torch.argsort(vl, descending=True, stable=True)
    idx = idx[order]
    return v_input[idx], q_input[idx], v_lens[idx], q_lens[idx], labels[idx]
