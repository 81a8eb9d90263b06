"""world_size-2 gloo tests (CPU) of the multi-GPU sharding logic in comfyui_propainter_nodes_b200.parallel."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from comfyui_propainter_nodes_b200 import parallel as P
from oracle import propainter_oracle as O


def test_shard_range_partitions():
    for n in (0, 1, 7, 16, 79, 80):
        for world in (1, 2, 3, 8):
            parts = [P.shard_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = P.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            assert sorted(sum((P.round_robin(n, world, r) for r in range(world)), [])) == list(range(n))


def test_composite_order_matches_reference_loop():
    sched = O.window_schedule(80, 10, 10, 80)
    ids, first = P.composite_order(sched)
    assert len(ids) == 170 and sum(first) == 80
    # frames at multiples of 5 (except the ends) are visited by 3 windows
    assert ids.count(5) == 3 and ids.count(0) == 2 and ids.count(3) == 2


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # variable-length all-gather == concatenation in rank order
        n = 7
        sizes = P.shard_sizes(n, world)
        lo, hi = P.shard_range(n, world, rank)
        full = torch.arange(n * 6, dtype=torch.float32).view(n, 2, 3)
        got = P.all_gather_variable(full[lo:hi].clone(), sizes)
        assert torch.equal(got, full)
        # an empty shard on one rank
        sizes = [3, 0]
        loc = full[:3].clone() if rank == 0 else full[:0].clone()
        assert torch.equal(P.all_gather_variable(loc, sizes), full[:3])
        # window sharding + gathered predictions reproduce the single-process composite
        T, H, W = 12, 4, 5
        sched = O.window_schedule(T, 4, 3, 80)
        g = torch.Generator().manual_seed(0)
        pred_all = torch.rand(sum(len(nb) for nb, _ in sched), H, W, 3, generator=g)   # stand-in for window outputs
        wlo, whi = P.shard_range(len(sched), world, rank)
        off = [0]
        for nb, _ in sched:
            off.append(off[-1] + len(nb))
        mine = pred_all[off[wlo]:off[whi]].clone()
        wsizes = [off[b] - off[a] for a, b in P.window_shards(len(sched), world)]
        gathered = P.all_gather_variable(mine, wsizes)
        assert torch.equal(gathered, pred_all)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_all_gather_variable_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret.get(0) and ret.get(1)


def test_flow_chunks_and_teams():
    """Chunking of complete_flow equals the reference's loop (propainter_inference.py:115-139); teams cover the ranks."""
    assert P.flow_chunks(79, 80) == [(0, 79, 0, 79)]
    ch = P.flow_chunks(239, 80)
    assert ch == [(0, 80, 0, 85), (80, 160, 75, 165), (160, 239, 155, 239)]
    for L, sub in ((25, 12), (29, 10), (239, 80), (100, 100), (101, 100)):
        got = P.flow_chunks(L, sub)
        ref = []
        if L <= sub:
            ref = [(0, L, 0, L)]
        else:
            for f in range(0, L, sub):          # the reference's index arithmetic
                s_f, e_f = max(0, f - 5), min(L, f + sub + 5)
                pad_s, pad_e = max(0, f) - s_f, e_f - min(L, f + sub)
                ref.append((s_f + pad_s, e_f - pad_e, s_f, e_f))
        assert got == ref, (L, sub)
        assert [c[0] for c in got][1:] == [c[1] for c in got][:-1] and got[-1][1] == L
    for n_chunks in (1, 2, 3, 7):
        for world in (1, 2, 3, 4, 8):
            n_teams, size = P.flow_teams(n_chunks, world)
            assert 1 <= n_teams <= n_chunks and size >= 1 and n_teams * size <= world
    assert P.flow_teams(1, 8) == (1, 8) and P.flow_teams(3, 8) == (3, 2) and P.flow_teams(3, 2) == (2, 1)


def test_balanced_ranges_minimise_the_heaviest_share():
    import itertools
    for costs, world in (([1, 1, 1, 1], 2), ([5, 1, 1, 1, 1, 1], 2), ([3, 3, 3, 9, 1, 1, 1], 3), ([2] * 16, 8), ([7], 4), ([], 3)):
        parts = P.balanced_ranges(costs, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(costs)
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        got = max(sum(costs[a:b]) for a, b in parts) if costs else 0
        # brute force over all cut positions
        best = min((max(sum(costs[a:b]) for a, b in zip((0,) + c, c + (len(costs),)))
                    for c in itertools.combinations_with_replacement(range(len(costs) + 1), world - 1)), default=0) if costs else 0
        assert abs(got - best) < 1e-9, (costs, world, parts)
    sched = O.window_schedule(80, 10, 10, 80)
    parts = P.balanced_ranges([P.window_cost(nb, r) for nb, r in sched], 8)
    loads = [sum(P.window_cost(*w) for w in sched[a:b]) for a, b in parts]
    assert max(loads) <= 1.25 * (sum(loads) / 8)
