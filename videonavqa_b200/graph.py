"""CUDA-graph capture of the fixed-shape train step.

One captured graph = ``model.init_hidden(); loss = loss_fn(model(v, q, v_lens, q_lens), y); loss.backward();
optimizer.step()`` - the loop body of eval/q_and_v_eval.py:119-139 - for ONE pattern of clip lengths (``v_lens`` decides
the launch geometry on the host, so it is part of the graph's identity; question lengths, tokens, frames and labels are
device data and may change between replays).  Replaying it issues the ~300 kernel launches of the step with a single
driver call: no Python, no ctypes, no tensor-map encoding, no allocator traffic.

The unregistered ``conv1x1_layers`` (quirk Q1: no optimizer ever updates them) are frozen into the graph through their
packed copies; re-capture after changing them by hand.

Requirements: ``FusedClipAdam(..., capturable=True).attach(model)`` (step count on the device, gradients in the flat
bucket), the attention or global-pooling model (the multi-hop model sizes a kernel by the longest question on the
host), a single GPU per graph.  Everything the step touches keeps its address: parameters and moments are flat buffers,
activations come from the graph's private pool.
"""
from __future__ import annotations

import torch

from .optim import FusedClipAdam


class GraphedTrainStep:
    def __init__(self, model, optimizer, v_input, q_input, v_lens, q_lens, labels, loss_fn=None, warmup=3):
        if not isinstance(optimizer, FusedClipAdam) or not optimizer.capturable or optimizer._module is not model:
            raise ValueError("GraphedTrainStep needs FusedClipAdam(..., capturable=True).attach(model)")
        if getattr(model, "kind", "") == "tmh":
            raise NotImplementedError("the multi-hop model reads the longest question length on the host")
        if v_lens.is_cuda:
            raise ValueError("pass v_lens as a CPU tensor: the clip lengths fix the launch geometry of the graph")
        if not model.training:
            raise ValueError("the train step needs train() mode")
        dev = v_input.device
        self.model, self.opt = model, optimizer
        self.loss_fn = loss_fn or torch.nn.CrossEntropyLoss(reduction="sum")       # eval.sh:15
        self.v_lens = v_lens.clone()
        # static inputs of the graph
        self.v = v_input.detach().float().contiguous().clone()
        self.q = q_input.to(dev).long().clone()
        self.q_lens = q_lens.to(dev).long().clone()
        self.labels = labels.to(dev).long().clone()
        # the eager warm-up steps (PyTorch asks for them on a side stream; they also fill every host-side cache: plans,
        # tensor maps, function attributes) must not count as training: snapshot, warm up, capture, restore
        f = optimizer.flat
        keep = [t.clone() for t in (f.param, optimizer._m, optimizer._v, optimizer._step_dev)]
        bufs = [b for _, b in model.named_buffers()]
        keep_bufs = [b.clone() for b in bufs]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        with torch.no_grad():
            for t, k in zip((f.param, optimizer._m, optimizer._v, optimizer._step_dev), keep):
                t.copy_(k)
            for b, k in zip(bufs, keep_bufs):
                b.copy_(k)
        # tensors that were created by the warm-up steps and whose ADDRESSES are baked into the graph (packed copies of
        # the never-optimised conv1x1_layers, quirk Q1) must outlive later cache invalidations
        self._keepalive = list(model._cache._c.values())
        model.invalidate_weight_cache(registered_only=True)
        model.init_hidden()

    def _step(self):
        self.model.init_hidden()                                                   # q_and_v_eval.py:119-120
        out = self.model(self.v, self.q, self.v_lens, self.q_lens)
        loss = self.loss_fn(out, self.labels)
        loss.backward()
        self.opt.step()
        self.opt.zero_grad()
        return loss.detach()

    def __call__(self, v_input=None, q_input=None, q_lens=None, labels=None):
        """Copy the new batch into the graph's inputs (any argument may be omitted to keep the previous one) and replay.
        Returns the loss tensor of the graph (overwritten by the next replay)."""
        for dst, src in ((self.v, v_input), (self.q, q_input), (self.q_lens, q_lens), (self.labels, labels)):
            if src is not None:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
