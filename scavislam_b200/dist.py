"""Multi-GPU plumbing of the BA path (one process per GPU, torch.distributed).

Two ways the path shards:
  * replicas: every rank optimises an independent window (BASELINE config C4); no data-path
    collective, only the timing reduction at the end -- `window_for_rank`, `reduce_job_totals`;
  * one large window split by landmarks (config C5): `shard_landmarks` gives each rank the
    landmarks l with l % world == rank (their edges follow), poses are replicated, and the reduced
    camera system is summed across ranks once per Levenberg trial -- `global_pose_pairs` is the
    block pattern every rank must agree on.
"""
from __future__ import annotations

import numpy as np


def window_for_rank(rank: int):
    """Rank 0 measures the C2 window itself; the other ranks get the same 200-keyframe / 20k-landmark
    window (identical structure, so that the per-rank work of the weak-scaling run is equal) with an
    independent realisation of the measurement and initialisation noise."""
    from . import synth
    if rank == 0:
        return synth.make_config("C2")
    P, L, idx = synth.CONFIGS["C2"]
    return synth.make_window(P, L, 1234 + idx, name=f"C4[{rank}]", noise_seed=9000 + rank)


def reduce_job_totals(times_s, counts, dist=None, device="cpu"):
    """Whole-job aggregation: max over ranks of every time, sum over ranks of every count."""
    import torch
    t = torch.tensor(list(times_s), dtype=torch.float64, device=device)
    n = torch.tensor(list(counts), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return [float(x) for x in t], [float(x) for x in n]


def shard_landmarks(pb, rank: int, world: int):
    """Landmark partition of one window: rank r keeps landmarks l % world == r with all their
    observation edges; poses are replicated; pose-pose constraints stay on rank 0."""
    keep_l = (np.arange(pb.L) % world) == rank
    new_index = np.cumsum(keep_l) - 1
    keep_e = keep_l[pb.e_point]
    sh = pb.copy()
    sh.L = int(keep_l.sum())
    sh.psi = np.ascontiguousarray(pb.psi[keep_l])
    sh.E = int(keep_e.sum())
    sh.e_point = np.ascontiguousarray(new_index[pb.e_point[keep_e]].astype(np.int32))
    for k in ("e_pose", "e_anchor", "e_obs", "e_info"):
        setattr(sh, k, np.ascontiguousarray(getattr(pb, k)[keep_e]))
    if rank != 0:
        sh.C = 0
        sh.c_i, sh.c_j = pb.c_i[:0].copy(), pb.c_j[:0].copy()
        sh.c_T, sh.c_Lambda = pb.c_T[:0].copy(), pb.c_Lambda[:0].copy()
    if pb.truth_psi is not None:
        sh.truth_psi = pb.truth_psi[keep_l]
    sh.name = f"{pb.name}[{rank}/{world}]"
    return sh, np.nonzero(keep_l)[0]


def global_pose_pairs(pb):
    """Unordered pose pairs coupled by a landmark track or a constraint: the block pattern of the
    reduced camera system of the WHOLE window (every rank needs the same one).  A handle that is given
    this pattern (svs_ba_set_structure) does not complete its tracks with zero-weight edges -- a shard's
    padding would add pairs the other shards' layouts lack; the in-library sharded window
    (svs_ba_set_problem_sharded) derives the padded pattern itself and keeps the padding."""
    order = np.argsort(pb.e_point, kind="stable")
    ep, ef = pb.e_point[order], pb.e_pose[order]
    anchors = pb.e_anchor[order]
    pairs = set()
    start = 0
    bounds = np.flatnonzero(np.diff(ep)) + 1
    for b in list(bounds) + [len(ep)]:
        ps = set(ef[start:b].tolist())
        if b > start:
            ps.add(int(anchors[start]))
        ps = sorted(ps)
        for i in range(len(ps)):
            for j in range(i + 1, len(ps)):
                pairs.add((ps[i], ps[j]))
        start = b
    for i, j in zip(pb.c_i.tolist(), pb.c_j.tolist()):
        pairs.add((min(i, j), max(i, j)))
    if not pairs:
        return np.zeros((0, 2), np.int32)
    return np.array(sorted(pairs), np.int32)


class _DevArray:
    """Raw device pointer -> object torch.as_tensor understands (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class ShardedWindow:
    """One large window (BASELINE config C5) split by landmarks over `group` = a list of
    BundleAdjuster handles.  With torch.distributed every rank passes ONE handle and the sums go
    through dist.all_reduce (NCCL over NVLink); without it, several handles on one GPU are summed
    in place (used to test the sharding arithmetic on a single device).

    Per Levenberg trial: every shard builds its partial reduced camera system, the partial systems
    (S, bp, bc) are summed, every shard solves the identical system and back-substitutes its own
    landmarks, the three scalars of the gain ratio are summed, every shard takes the same decision."""

    def __init__(self, adjusters, pb, rank=0, world=None, dist=None, device=None):
        import torch
        self.torch = torch
        self.dist = dist
        self.ba = list(adjusters)
        world = world if world is not None else (dist.get_world_size() if dist is not None else len(self.ba))
        pairs = global_pose_pairs(pb)
        self.index = []
        for k, b in enumerate(self.ba):
            r = rank if dist is not None else k
            sh, idx = shard_landmarks(pb, r, world)
            b.set_structure(pairs)
            b.set_problem(sh)
            self.index.append(idx)
        self.L = pb.L
        self.views = []
        for b in self.ba:
            pS, nS, pbp, pbc, nb, ptot = b.system_buffers()
            dev = device if device is not None else torch.cuda.current_device()
            mk = lambda p, n: torch.as_tensor(_DevArray(p, n), device=f"cuda:{dev}")
            self.views.append((mk(pS, nS), mk(pbp, nb), mk(pbc, nb), mk(ptot, 3)))

    def _sum(self, which):
        ts = [v[which] for v in self.views]
        if self.dist is not None:
            for t in ts:
                self.dist.all_reduce(t)
        elif len(ts) > 1:
            total = ts[0].clone()
            for t in ts[1:]:
                total += t
            for t in ts:
                t.copy_(total)
        self.torch.cuda.synchronize()

    def optimize(self, num_iters, robust=True, huber_delta=1.0, lambda_init=50.0, max_trials=5):
        for b in self.ba:
            b.lm_begin(lambda_init, max_trials)
        it, ok = 0, True
        while it < num_iters and ok:
            for b in self.ba:
                b.trial_build(robust, huber_delta)
            for w in (0, 1, 2):
                self._sum(w)
            for b in self.ba:
                b.trial_solve(robust, huber_delta)
            self._sum(3)
            res = [b.trial_decide() for b in self.ba]
            again, stop, it = res[0]
            assert all(r == res[0] for r in res), "shards disagree on the Levenberg decision"
            if not again and stop:
                ok = False
        return it, self.ba[0].lm_stats()

    def poses(self):
        return self.ba[0].poses()

    def points_local(self):
        """[(global landmark indices, psi)] for the shards this process holds."""
        return [(idx, b.points()) for idx, b in zip(self.index, self.ba)]
