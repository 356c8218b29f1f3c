"""World-size-2 gloo run of the multi-GPU plumbing (alg_b200/distributed.py) on CPU: weight broadcast from rank 0,
per-rank sample seeds, max-over-ranks timing and the whole-job rate bench.py reports."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from alg_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert D.env_rank() == (rank, world, rank)
        g = torch.Generator().manual_seed(7)
        sd = {"b.weight": torch.randn(5, 3, generator=g), "a.bias": torch.randn(4, generator=g).bfloat16()}
        if rank != 0:  # non-root ranks start from garbage of the right shape, like bench.py's torch.empty
            sd = {k: torch.full_like(v, float("nan")) for k, v in sd.items()}
        D.broadcast_state_dict(sd)
        # flat arena: one collective for the whole model, views stay aligned and typed
        views, arena = D.arena_state_dict({"w": ((5, 3), torch.float32), "b": ((7,), torch.bfloat16), "s": ((1, 2, 4), torch.float32)}, "cpu")
        if rank == 0:
            for i, k in enumerate(sorted(views)):
                views[k].copy_(torch.arange(views[k].numel(), dtype=torch.float32).view(views[k].shape) + 10 * i)
        else:
            arena.fill_(255)
        D.broadcast_arena(arena, chunk_bytes=64)
        assert all(v.data_ptr() % 16 == 0 for v in views.values())
        assert [float(views[k].float().flatten()[1]) for k in sorted(views)] == [1.0, 11.0, 21.0]
        assert views["w"].dtype == torch.float32 and views["b"].dtype == torch.bfloat16 and views["s"].shape == (1, 2, 4)
        ms = D.max_over_ranks([10.0 + rank, 5.0 - rank], "cpu")
        q.put((rank, D.sample_seed(rank), {k: v.float().sum().item() for k, v in sd.items()}, ms,
               D.aggregate_rate(81 / 50, world, ms[0])))
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, sums0, ms0, rate0), (r1, s1, sums1, ms1, rate1) = res
    assert (s0, s1) == (42, 43)
    assert sums0 == sums1 and all(v == v for v in sums0.values())  # rank 1 received rank 0's weights (no NaN left)
    assert ms0 == ms1 == [11.0, 5.0]
    assert abs(rate0 - 2 * (81 / 50) / 0.011) < 1e-9 and rate0 == rate1
