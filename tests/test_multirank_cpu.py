"""world_size-2 gloo tests (CPU) of the multi-rank host logic bench.py uses: clip sharding, max-over-ranks timing,
result gather.  The data path itself has no collective (clips are independent)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    clips = bench.shard_clips(7, rank, world)
    mine = bench.shard_longest_first([60 * 5] * 4, world)[rank]
    labels = torch.full((len(mine), 3, 4, 5), rank + 1, dtype=torch.uint8)
    maps = bench.gather_label_maps(labels, rank, world)
    assert (maps is None) == (rank != 0)
    if rank == 0:
        assert [int(m.unique()) for m in maps] == [1, 2] and all(m.shape == (2, 3, 4, 5) and m.dtype == torch.uint8 for m in maps)
    dev_ms, e2e_s, sums = bench.reduce_over_ranks(10.0 + rank, 2.0 - rank, 100.0 * (rank + 1), torch.device("cpu"), rank, world)
    q.put((rank, clips, dev_ms, e2e_s, sums))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_max_and_gather():
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, c0, ms0, e0, s0), (r1, c1, ms1, e1, s1) = out
    assert c0 == [0, 2, 4, 6] and c1 == [1, 3, 5]            # every clip exactly once
    assert ms0 == ms1 == 11.0 and e0 == e1 == 2.0            # MAX over ranks on both ranks
    assert s0 == [100.0, 200.0] and s1 is None               # gather lands on rank 0 only


def test_single_rank_is_a_no_op():
    import bench
    assert bench.shard_clips(3, 0, 1) == [0, 1, 2]
    assert bench.reduce_over_ranks(1.0, 2.0, 3.0, torch.device("cpu"), 0, 1) == (1.0, 2.0, [3.0])


def test_longest_first_sharding_balances_and_covers_every_clip():
    import bench
    costs = [60 * n for n in (1, 2, 3, 4, 5)] * 3 + [104 * 5]
    for world in (1, 2, 4, 8):
        shards = bench.shard_longest_first(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(len(costs)))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)                 # LPT bound
    assert bench.shard_longest_first(costs, 2)[0][0] == len(costs) - 1   # the longest clip goes first, to rank 0
    eq = bench.shard_longest_first([300] * 16, 8)
    assert all(len(s) == 2 for s in eq)                              # equal clips -> equal shards (the gather needs that)
