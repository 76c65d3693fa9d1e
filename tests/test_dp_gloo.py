"""World-size-2 checks of the data-parallel HOST logic on CPU (gloo stands in for the library's NCCL communicator,
dp.use_host_collective_for_tests): the frontier-driven chunked all-reduce of the flat gradient bucket (dp.FlatReducer
driven through engine.GradSink exactly as the backward pass drives it), weight broadcast including the unregistered
conv1x1_layers, the exchange checksum check, and the clip-wise batch sharding."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from videonavqa_b200 import dp
        from videonavqa_b200.engine import GradSink
        from videonavqa_b200.models import FiLMGlobalPoolingPretrainedStem
        from videonavqa_b200.optim import FlatBuffers
        dp.use_host_collective_for_tests()
        torch.manual_seed(100 + rank)          # different initial weights per rank
        m = FiLMGlobalPoolingPretrainedStem(2, 16, 7, 32, 32, 8, 2, 32, vocab_size=20)
        flat = FlatBuffers(m.parameters())
        assert flat.is_intact() and all(p.data_ptr() == v.data_ptr() - (flat.grad.data_ptr() - flat.param.data_ptr())
                                        for p, v in zip(flat.params, flat.grad_views))
        dp.broadcast_module(m, flat)
        state = torch.cat([flat.param] + [c.weight.detach().flatten() for c in m.conv1x1_layers] +
                          [m.bn_init.running_var.flatten()])
        gathered = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(gathered, state)
        same = all(torch.equal(gathered[0], g) for g in gathered)

        # the backward pass as the engine drives it: buffers come from the bucket, gradients are published in
        # (roughly) reverse parameter order, the reducer sends contiguous slices as the frontier moves down
        red = dp.FlatReducer(m, flat, overlap=True, min_bytes=64)
        bucket = {n: v for (n, _), v in zip(m.named_parameters(), flat.grad_views)}
        names = [n for n, _ in m.named_parameters()]
        calls = []
        for step in range(2):                    # two steps: the reducer resets itself
            flat.grad.zero_()
            sink = GradSink(m._grad_hook, bucket)
            order = list(reversed(names))
            order[1], order[4] = order[4], order[1]          # not perfectly monotone
            for n in order[:-3] + ["conv1x1_layers.0.weight"]:
                if n.startswith("conv1x1_layers"):
                    sink[n] = torch.ones(2, 2)               # Q1: unregistered, stays local
                    continue
                t = sink.buffer(n, dict(m.named_parameters())[n])
                t.fill_(float(rank + 1) * (names.index(n) + 1))
                sink[n] = t
            nbytes = red.wait()                  # the last 3 parameters got no gradient: their slices stay zero
            calls.append(red.last_calls)
        ok_sum = True
        for n in names:
            want = 0.0 if n in order[-3:] else float(1 + 2) * (names.index(n) + 1)
            ok_sum &= bool((bucket[n] == want).all())
        ok_local = torch.equal(sink["conv1x1_layers.0.weight"], torch.ones(2, 2))

        flat.grad.copy_(torch.arange(flat.total, dtype=torch.float32) * (rank + 1) * 1e-3)
        chk = dp.verify_allreduce(flat)

        # sharding: every clip appears exactly once, each shard sorted by descending length
        B = 8
        v = torch.arange(B).float().view(B, 1, 1, 1, 1).expand(B, 1, 10, 13, 2).contiguous()
        v_lens = torch.tensor([35, 30, 22, 22, 17, 9, 4, 1])
        sv, sq, svl, sql, sl = dp.shard_batch(v, torch.zeros(B, 3, dtype=torch.long), v_lens, torch.ones(B).long(),
                                             torch.arange(B), rank, world)
        ids = sl.clone()
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        ok_shard = sorted(torch.cat(all_ids).tolist()) == list(range(B)) and bool((svl[:-1] >= svl[1:]).all()) \
            and torch.equal(sv[:, 0, 0, 0, 0].long(), sl)
        q.put((rank, same, ok_sum, ok_local, nbytes, flat.total * 4, calls, chk, ok_shard))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.is_available(), reason="host-logic test for the CPU container: with a GPU present the "
                    "modules place film_layer on it (reference CUDA branch) and the flat buffers must live on one device")
def test_dp_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, ok_sum, ok_local, nbytes, total_bytes, calls, chk, ok_shard in res:
        assert same, "broadcast_module left replicas different"
        assert ok_sum, "all-reduced gradients are not the SUM over ranks"
        assert ok_local, "conv1x1_layers gradients must stay local (never optimised in the reference)"
        assert nbytes == total_bytes
        assert calls[0] == calls[1] and calls[0] > 2, f"expected several chunked all-reduces per step, got {calls}"
        assert chk["ok"] and chk["ranks_identical"], chk
        assert ok_shard
    assert res[0][6] == res[1][6], "ranks issued different numbers of collectives"
