"""World-size-2 checks of the data-parallel host logic on CPU (gloo): the overlapped gradient all-reduce
(GradAllReducer driven through engine.GradSink), weight broadcast including the unregistered conv1x1_layers, and
the clip-wise batch sharding."""
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
        torch.manual_seed(100 + rank)          # different initial weights per rank
        m = FiLMGlobalPoolingPretrainedStem(2, 16, 7, 32, 32, 8, 2, 32, vocab_size=20)
        dp.broadcast_module(m)
        flat = torch.cat([p.detach().flatten() for p in m.parameters()] +
                         [c.weight.detach().flatten() for c in m.conv1x1_layers] +
                         [m.bn_init.running_var.flatten()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered)

        red = dp.GradAllReducer(m)
        sink = GradSink(m._grad_hook)
        names = [n for n, _ in m.named_parameters()]
        g_local = {}
        for i, n in enumerate(names[:5] + ["conv1x1_layers.0.weight"]):
            t = torch.full((3, 2), float(rank + 1) * (i + 1))
            g_local[n] = t.clone()
            sink[n] = t                          # engine reports a finished gradient -> async all-reduce starts
        nbytes = red.wait()
        ok_sum = all(torch.equal(sink[n], g_local[n] * 0 + (1 + 2) * (i + 1)) for i, n in enumerate(names[:5]))
        ok_local = torch.equal(sink["conv1x1_layers.0.weight"], g_local["conv1x1_layers.0.weight"])   # Q1: not reduced

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
        q.put((rank, same, ok_sum, ok_local, nbytes, ok_shard))
    finally:
        dist.destroy_process_group()


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
    for rank, same, ok_sum, ok_local, nbytes, ok_shard in res:
        assert same, "broadcast_module left replicas different"
        assert ok_sum, "all-reduced gradients are not the SUM over ranks"
        assert ok_local, "conv1x1_layers gradients must stay local (never optimised in the reference)"
        assert nbytes == 5 * 6 * 4
        assert ok_shard
