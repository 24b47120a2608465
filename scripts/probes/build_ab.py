"""Developer probe (GPU box): per-trial kernel times of the BA step for C2, its two structure variants and C5 under
the environment's build-kernel switches (SVS_BUILD_STATIC, SVS_BUILD_NO_LPT, SVS_BUILD_CHUNK), plus parity of the
result against a run of the same process's reference settings is left to tests/test_ba_gpu.py.
usage: python scripts/probes/build_ab.py [C2 C2d C2l C5]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scavislam_b200 import capi, synth

names = sys.argv[1:] or ["C2", "C2d", "C2l", "C5"]
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SVS_BUILD")) or "default"
ba = capi.BundleAdjuster()
for name in names:
    base = synth.make_config("C5" if name == "C5" else "C2")
    pb = {"C2d": lambda: synth.with_dropouts(base, 0.2, seed=1), "C2l": lambda: synth.with_loop_closures(base, 10, seed=1)}.get(name, lambda: base)()
    ba.set_problem(pb)
    rows = []
    for rep in range(6):
        ba.reset_state()
        it, st = ba.optimize(10)
        rows.append([st[k] / max(st["trials_total"], 1) for k in ("ms_build", "ms_solve", "ms_update", "ms_total")])
    med = np.median(np.array(rows[1:]), axis=0)
    print(f"[{tag}] {name}: E={pb.E} per trial ms build {med[0]:.4f} solve {med[1]:.4f} update {med[2]:.4f} total {med[3]:.4f}"
          f"  chi2_final {st['chi2_iter'][-1]:.9g} trials {st['trials_total']}", flush=True)
