"""CPU, world_size 2 over gloo: the multi-GPU host logic (work assignment, max-over-ranks timing).  The data path has
no collective (SCAs are independent), so this is all there is to distribute."""

import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from romanimpreprocess_b200 import sharding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = [(e, s) for e in range(4) for s in range(1, 19)]  # 4 exposures x 18 SCAs
    mine = sharding.assign_items(items, rank, world)
    rate, units, tmax = sharding.job_throughput(len(mine), 1.0 + rank, None)  # rank 1 is the slow one
    few = sharding.assign_items([(e, 7) for e in range(5)], rank, world)  # one SCA only: exposures are spread
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, rate, units, tmax, few))


def test_two_ranks_partition_and_timing():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, rate0, units0, tmax0, few0), (r1, m1, rate1, units1, tmax1, few1) = res
    assert (r0, r1) == (0, 1)
    all_items = [(e, s) for e in range(4) for s in range(1, 19)]
    assert sorted(m0 + m1) == sorted(all_items) and not set(m0) & set(m1)  # disjoint cover
    assert {s for _, s in m0}.isdisjoint({s for _, s in m1})  # an SCA's CALDIR lives on one rank only
    assert len(m0) == len(m1) == 36
    assert units0 == units1 == 72 and tmax0 == tmax1 == 2.0  # max over ranks
    assert rate0 == rate1 == pytest.approx(36.0)
    assert sorted(few0 + few1) == [(e, 7) for e in range(5)] and abs(len(few0) - len(few1)) <= 1


def test_single_process_needs_no_group():
    sys.path.insert(0, ROOT)
    from romanimpreprocess_b200 import sharding

    assert sharding.job_throughput(10, 2.0) == (5.0, 10, 2.0)
    items = [(0, s) for s in range(1, 19)]
    parts = [sharding.assign_items(items, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == items
    assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert sharding.resident_scas(parts[0]) == [1, 9, 17]


def test_balanced_assignment_18_scas_on_8_ranks():
    sys.path.insert(0, ROOT)
    from romanimpreprocess_b200 import sharding

    items = [(e, s) for e in range(4) for s in range(1, 19)]  # 4 exposures x 18 SCAs = 72 items
    parts = [sharding.assign_items_balanced(items, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == sorted(items)  # disjoint cover
    assert {len(p) for p in parts} == {9}  # 72 / 8, where SCA-major dealing gives 12/12/8/8/...
    assert sharding.imbalance(items, 8) == 1.0
    assert sharding.imbalance(items, 8, sharding.assign_items) == pytest.approx(12 / 9.0)
    assert max(len(sharding.resident_scas(p)) for p in parts) <= 4  # 2 own SCAs + at most the 2 leftover ones
    for w in (1, 2, 4):
        ps = [sharding.assign_items_balanced(items, r, w) for r in range(w)]
        assert sorted(sum(ps, [])) == sorted(items) and max(map(len, ps)) - min(map(len, ps)) <= 1
