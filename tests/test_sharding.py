"""CPU suite: the N>1 host logic (task sharding + metric reduction) on gloo, world_size 2.
The data path has no collective; each rank runs its slice through a local engine
(here the oracle stands in for the GPU, which is what lets this run without one)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle
from lamsa_b200 import sharding, workload


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tasks, keep = workload.gen_microbench(3000, seed=77, qmax=200)
    mine = sharding.shard_slices(len(tasks), world, chunk=256)[rank]
    res, cig, secs = _oracle.oracle_run(tasks[mine], 1)
    (tmax,), (cells, ntask) = sharding.reduce_metrics(dist, "cpu", [0.5 + rank], [int(res["cells"].sum()), len(mine)])
    ret[rank] = (tmax, cells, ntask, mine, res["score"].copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world, port = 2, 29000 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        out = dict(ret)
    tasks, keep = workload.gen_microbench(3000, seed=77, qmax=200)
    full, _, _ = _oracle.oracle_run(tasks, 2)
    slices = [out[r][3] for r in range(world)]
    # disjoint cover of the stream
    allidx = np.sort(np.concatenate(slices))
    assert (allidx == np.arange(len(tasks))).all()
    # both ranks agree on the reduced numbers: max of times, sum of cells / tasks
    for r in range(world):
        assert out[r][0] == 1.5 and out[r][1] == int(full["cells"].sum()) and out[r][2] == len(tasks)
    merged = sharding.merge_results([out[r][4] for r in range(world)], slices, len(tasks))
    assert (merged == full["score"]).all()


def test_shard_slices_edge_cases():
    for n, world, chunk in [(0, 2, 16), (1, 8, 16), (17, 3, 4), (4096, 8, 4096), (10000, 4, 333)]:
        s = sharding.shard_slices(n, world, chunk)
        assert len(s) == world
        cat = np.sort(np.concatenate(s)) if n else np.zeros(0, int)
        assert (cat == np.arange(n)).all()


# ---- the chaining half of the path shards the same way: reads are independent ------------------
def _sdp_worker(rank, world, port, ret):
    import _sdp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = _sdp.gen_reads(90, seed=21, mode="pacbio", repeat_frac=0.3, sv_rate=0.5, miss_frac=0.2, read_len=(500, 5000))
    mine = sharding.shard_slices(len(rs), world, chunk=8)[rank]
    (s1, o1), (s2, o2), pairs = _sdp.oracle_run(rs.subset(mine))
    (tmax,), (npairs, nreads) = sharding.reduce_metrics(dist, "cpu", [1.0 + rank], [int(pairs.sum()), len(mine)])
    ret[rank] = (tmax, npairs, nreads, mine, (s1.copy(), o1.copy()), (s2.copy(), o2.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_read_sharding_of_the_chaining():
    import _sdp
    world, port = 2, 31000 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_sdp_worker, args=(world, port, ret), nprocs=world, join=True)
        out = dict(ret)
    rs = _sdp.gen_reads(90, seed=21, mode="pacbio", repeat_frac=0.3, sv_rate=0.5, miss_frac=0.2, read_len=(500, 5000))
    (f1, fo1), (f2, fo2), fpairs = _sdp.oracle_run(rs)
    for r in range(world):
        assert out[r][0] == 2.0 and out[r][1] == int(fpairs.sum()) and out[r][2] == len(rs)
    # every read's skeleton streams are the same whether it was chained alone in its shard or in the full set
    for r in range(world):
        mine, (s1, o1), (s2, o2) = out[r][3], out[r][4], out[r][5]
        for k, g in enumerate(mine):
            assert np.array_equal(s1[o1[k]:o1[k + 1]], f1[fo1[g]:fo1[g + 1]])
            assert np.array_equal(s2[o2[k]:o2[k + 1]], f2[fo2[g]:fo2[g + 1]])
