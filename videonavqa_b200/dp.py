"""Clip-sharded data parallelism for the FiLM hot path; new work, the reference is single-GPU.

One process per GPU.  Clips are independent, so the forward needs no collective; the only exchange is the gradient
all-reduce, done by the library's OWN NCCL communicator on its own communication stream (include/vnqa_b200.h
``vnqa_nccl_*``, csrc/nccl_comm.cu) - ``torch.distributed`` is only used as the host-side rendezvous that carries the
128-byte NCCL id (any backend; bench.py uses gloo):

  * SUM, not mean: the reference loss is CrossEntropyLoss(reduction='sum') (eval.sh:15), so N ranks x b clips is the
    same optimisation step as one rank x N*b clips (per-replica BatchNorm statistics, as in ordinary DDP);
  * no pack / unpack: the backward kernels write every registered parameter's gradient straight into ONE flat bucket
    (optim.FlatBuffers, parameters() order) that is all-reduced IN PLACE and then read by the fused optimizer;
  * overlapped: gradients become final in (roughly) reverse parameter order, so the reducer keeps a frontier - the
    lowest index above which every gradient is final - and each time it has moved down by >= ``min_bytes`` it enqueues
    one all-reduce of that contiguous slice behind an event on the compute stream.  The collective runs on the
    communication stream next to the remaining backward kernels (NVLS in-switch reduction at 8 GPUs; ring / LL at 2-4);
    ``wait()`` sends what is left and makes the compute stream wait for the communication stream.  The persistent
    tensor-core kernels tolerate losing SMs to the NCCL kernel: their CTA pairs allocate tensor memory only after a
    cluster barrier (round-1's 2/4-GPU dead-lock, profiles/r02_dp_hang_rootcause.md);
  * the unregistered conv1x1_layers (quirk Q1) are broadcast from rank 0 at start-up - nothing else would ever
    synchronise them - and are never reduced or optimised, exactly as in the reference.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

_READY = [False]


def _lib_nccl_path():
    env = os.environ.get("VNQA_NCCL_LIB")
    if env:
        return env
    try:        # the NCCL build PyTorch ships (same one torch.distributed would use)
        import nvidia.nccl
        p = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2")
        if os.path.exists(p):
            return p
    except Exception:  # noqa: BLE001
        pass
    return "libnccl.so.2"


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def init(rank=None, world=None):
    """Create the library's NCCL communicator for the current CUDA device.  The unique id travels through
    torch.distributed (must be initialised, any backend)."""
    import torch.distributed as dist
    if _READY[0]:
        return
    L = _lib.lib()
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    _lib.check(L.vnqa_nccl_load(_lib_nccl_path().encode()), "vnqa_nccl_load")
    box = [None]
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        _lib.check(L.vnqa_nccl_unique_id(buf), "vnqa_nccl_unique_id")
        box[0] = buf.raw
    dist.broadcast_object_list(box, src=0)
    idbuf = ctypes.create_string_buffer(bytes(box[0]), 128)
    _lib.check(L.vnqa_nccl_init(idbuf, rank, world), "vnqa_nccl_init")
    _READY[0] = True


def shutdown():
    if _READY[0]:
        _lib.lib().vnqa_nccl_destroy()
        _READY[0] = False


def world_size():
    return _lib.lib().vnqa_nccl_nranks() if _READY[0] else 1


class _LibraryNccl:
    """The product path: the library's own communicator / communication stream (csrc/nccl_comm.cu)."""

    def all_reduce(self, t, a=0, b=None, join=True):
        b = t.numel() if b is None else b
        assert _READY[0] and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        L = _lib.lib()
        _lib.check(L.vnqa_nccl_allreduce_sum_f32(ctypes.c_void_p(t.data_ptr() + 4 * a), b - a, _stream()),
                   "vnqa_nccl_allreduce")
        if join:
            self.join()

    def join(self):
        _lib.check(_lib.lib().vnqa_nccl_join(_stream()), "vnqa_nccl_join")

    def broadcast(self, t, src):
        assert _READY[0] and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        _lib.check(_lib.lib().vnqa_nccl_broadcast_f32(ctypes.c_void_p(t.data_ptr()), t.numel(), src, _stream()),
                   "vnqa_nccl_broadcast")
        self.join()

    def all_gather_f64(self, t):
        assert _READY[0] and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        out = torch.empty(world_size(), t.numel(), device=t.device, dtype=torch.float64)
        _lib.check(_lib.lib().vnqa_nccl_allgather_f64(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                      t.numel(), _stream()), "vnqa_nccl_allgather")
        self.join()
        return out


class _HostCollective:
    """torch.distributed (gloo) stand-in used ONLY by the CPU tests of the host logic (tests/test_dp_gloo.py): same
    call sequence, synchronous collectives on CPU tensors."""

    def all_reduce(self, t, a=0, b=None, join=True):
        import torch.distributed as dist
        dist.all_reduce(t.view(-1)[a:t.numel() if b is None else b])

    def join(self):
        pass

    def broadcast(self, t, src):
        import torch.distributed as dist
        dist.broadcast(t, src=src)

    def all_gather_f64(self, t):
        import torch.distributed as dist
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return torch.stack(out)


_COLL = [_LibraryNccl()]


def use_host_collective_for_tests():
    _COLL[0] = _HostCollective()


def all_reduce_(t):
    """In-place SUM all-reduce of a contiguous fp32 tensor, ordered after the current stream; the current stream is
    made to wait for the result (device-side, no host sync)."""
    _COLL[0].all_reduce(t)
    return t


def broadcast_(t, src=0):
    _COLL[0].broadcast(t, src)
    return t


def all_gather_f64(t):
    """(k,) float64 tensor -> (world, k)."""
    return _COLL[0].all_gather_f64(t)


def broadcast_module(module, flat=None, src=0):
    """Make every replica start from rank `src`'s weights: registered parameters (one broadcast of the flat buffer
    when there is one), buffers (BatchNorm running statistics) and the unregistered conv1x1_layers (Q1)."""
    with torch.no_grad():
        if flat is not None:
            broadcast_(flat.param, src)
        else:
            for p in module.parameters():
                broadcast_(p.data, src)
        for b in module.buffers():
            if b.dtype == torch.float32:
                broadcast_(b.data, src)
            else:       # num_batches_tracked (int64)
                f = b.data.to(torch.float32).reshape(-1).contiguous()
                b.data.copy_(broadcast_(f, src).view(b.shape).to(b.dtype))
        for conv in getattr(module, "conv1x1_layers", []):
            broadcast_(conv.weight.data, src)
            broadcast_(conv.bias.data, src)
    if hasattr(module, "invalidate_weight_cache"):
        module.invalidate_weight_cache()        # the writes above went through .data: no version bump


class FlatReducer:
    """Installs itself as the module's gradient hook (engine.GradSink).  ``wait()`` is called by the autograd node at
    the end of backward: after it returns (on the compute stream) the flat gradient bucket holds the SUM over ranks."""

    def __init__(self, module, flat, overlap=None, min_bytes=None):
        self.flat = flat
        self.overlap = (os.environ.get("VNQA_DP_OVERLAP", "1") == "1") if overlap is None else overlap
        if min_bytes is None:        # slice size from which the frontier is handed to NCCL (VNQA_DP_MIN_MB to tune)
            min_bytes = int(float(os.environ.get("VNQA_DP_MIN_MB", "16")) * (1 << 20))
        self.min_elems = max(1, min_bytes // 4)
        names = {id(p): n for n, p in module.named_parameters()}
        self.index = {names[id(p)]: i for i, p in enumerate(flat.params) if id(p) in names}
        self.n = len(flat.params)
        self.last_bytes = 0
        self.last_calls = 0
        self._reset()
        module._grad_hook = self

    def _reset(self):
        self.done = [False] * self.n
        self.lo = self.n          # every parameter with index >= lo has its final gradient
        self.sent = self.n        # ... and every parameter with index >= sent has been handed to NCCL
        self.calls = 0

    def _offset(self, i):
        return self.flat.offsets[i] if i < self.n else self.flat.total

    def _send(self, lo, hi):
        a, b = self._offset(lo), self._offset(hi)
        if b <= a:
            return
        _COLL[0].all_reduce(self.flat.grad, a, b, join=False)
        self.calls += 1

    def __call__(self, name, tensor):
        i = self.index.get(name)
        if i is None:
            return      # conv1x1_layers: never optimised in the reference, their gradients stay local
        self.done[i] = True
        while self.lo > 0 and self.done[self.lo - 1]:
            self.lo -= 1
        if self.overlap and self._offset(self.sent) - self._offset(self.lo) >= self.min_elems:
            self._send(self.lo, self.sent)
            self.sent = self.lo

    def wait(self):
        # parameters that received no gradient this step are zero in the bucket: reduce everything that is left
        self._send(0, self.sent)
        _COLL[0].join()
        self.last_bytes = self.flat.total * 4
        self.last_calls = self.calls
        self._reset()
        return self.last_bytes


def verify_allreduce(flat):
    """Bench / test check of the exchange itself: fill-independent - uses whatever is in the gradient bucket.  Returns
    a dict: every rank's reduced checksum is identical, and equals the sum of the per-rank local checksums."""
    g = flat.grad
    local = torch.stack([g.double().sum(), g.double().abs().sum()])
    locals_ = all_gather_f64(local)
    all_reduce_(g)
    red = torch.stack([g.double().sum(), g.double().abs().sum()])
    reds = all_gather_f64(red)
    want = float(locals_[:, 0].sum())
    scale = float(locals_[:, 1].sum()) + 1e-30
    got = reds[:, 0].tolist()
    return {"ranks_identical": bool((reds == reds[0]).all()), "reduced_checksum": got[0], "sum_of_local_checksums": want,
            "rel_err": abs(got[0] - want) / scale, "ok": bool((reds == reds[0]).all()) and abs(got[0] - want) / scale < 1e-5}


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
