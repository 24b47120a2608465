"""Worker of tests/test_dist_gpu.py::test_in_library_nccl_window (launched with torch.distributed.run,
one process per GPU): the landmark-sharded window driven INSIDE the library -- svs_ba_comm_init,
svs_ba_set_problem_sharded, svs_ba_optimize with its per-trial ncclAllReduce -- against the CPU oracle
on the whole window.  Prints 'NCCL_WORKER_OK' on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scavislam_b200 import capi, synth
    from oracle import pyoracle as po
    ids = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    pb = synth.make_window(40, 3000, seed=77)
    ba = capi.BundleAdjuster(device=local)
    ba.comm_init(world, rank, ids[0])
    ba.set_problem_sharded(pb)
    it, st = ba.optimize(5)
    poses, psi = ba.poses(), ba.points_all()
    p_o, s_o, st_o = po.optimize(pb, 5)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert it == st_o["iterations"], (it, st_o["iterations"])
    assert st["trials_iter"] == st_o["trials_iter"]
    np.testing.assert_allclose(st["chi2_iter"], st_o["chi2_iter"], rtol=1e-7)
    assert rel(poses, p_o) < 1e-6, rel(poses, p_o)
    assert rel(psi, s_o) < 1e-6, rel(psi, s_o)
    # a second call on the same communicator: the one-call path of a back-end tick
    ba.set_problem_sharded(pb)
    it2, st2 = ba.optimize(2)
    assert it2 == 2 and st2["trials_iter"] == st_o["trials_iter"][:2]
    # tracks with visibility drop-outs: every rank completes ITS tracks with zero-weight edges, so the block pattern all
    # ranks agree on must contain the pose pairs of every rank's padding (svs_ba_set_problem_sharded derives it from the
    # whole window with the same rule) -- a mismatch would sum different blocks in the all-reduce
    pbd = synth.with_dropouts(pb, 0.2, seed=5)
    ba.set_problem_sharded(pbd)
    itd, std = ba.optimize(4)
    p_d, s_d, st_d = po.optimize(pbd, 4)
    assert itd == st_d["iterations"] and std["trials_iter"] == st_d["trials_iter"]
    np.testing.assert_allclose(std["chi2_iter"], st_d["chi2_iter"], rtol=1e-7)
    assert rel(ba.poses(), p_d) < 1e-6 and rel(ba.points_all(), s_d) < 1e-6
    ba.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(f"NCCL_WORKER_OK world={world} pose_rel={rel(poses, p_o):.2e} psi_rel={rel(psi, s_o):.2e}")


if __name__ == "__main__":
    main()
