"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: replica aggregation and the
landmark partition used for one large window."""
import os

import numpy as np
import torch.multiprocessing as mp

from scavislam_b200 import dist as sdist
from scavislam_b200 import synth


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # replicas: each rank "measured" rank-dependent numbers
    t, n = sdist.reduce_job_totals([10.0 + rank, 3.0 - rank], [100 * (rank + 1), 7], dist)
    # landmark shards: partial sums over shards must add up to the full-window quantity
    pb = synth.make_window(6, 120, seed=21)
    sh, idx = sdist.shard_landmarks(pb, rank, world)
    import torch
    part = torch.tensor([float(sh.E), float(sh.L), float(sh.C), float(sh.e_obs.sum())], dtype=torch.float64)
    dist.all_reduce(part)
    q.put((rank, t, n, part.tolist(), sh.E, sh.C))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pb = synth.make_window(6, 120, seed=21)
    for rank, t, n, part, E, C in res:
        assert t == [11.0, 3.0]                 # max over ranks
        assert n == [300.0, 14.0]               # sum over ranks
        assert part[0] == pb.E and part[1] == pb.L and part[2] == pb.C
        assert abs(part[3] - pb.e_obs.sum()) < 1e-6
    assert res[1][5] == 0 and res[0][5] == pb.C  # constraints live on rank 0 only


def test_shards_partition_the_window():
    pb = synth.make_config("C1")
    seen = np.zeros(pb.L, int)
    edges = 0
    for r in range(4):
        sh, idx = sdist.shard_landmarks(pb, r, 4)
        seen[idx] += 1
        edges += sh.E
        assert sh.P == pb.P and np.array_equal(sh.pose_qt, pb.pose_qt)
        # every shard edge refers to a landmark of the shard, with the original psi
        assert sh.e_point.max(initial=-1) < sh.L
        np.testing.assert_array_equal(sh.psi, pb.psi[idx])
    assert (seen == 1).all() and edges == pb.E


def test_global_pose_pairs_cover_every_shard():
    pb = synth.make_config("C1")
    pairs = {tuple(p) for p in sdist.global_pose_pairs(pb)}
    for r in range(3):
        sh, _ = sdist.shard_landmarks(pb, r, 3)
        assert {tuple(p) for p in sdist.global_pose_pairs(sh)} <= pairs
    # co-visible neighbours in a chain window
    assert (0, 1) in pairs and all(i < j for i, j in pairs)
