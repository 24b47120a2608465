"""Developer probe (GPU box): k_solve phase cycles (SVS_SOLVE_TIMING=1, last launch) for C2, C2 + 10 loop closures and C5."""
import os
import sys
os.environ.setdefault("SVS_SOLVE_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from scavislam_b200 import capi, synth

ba = capi.BundleAdjuster()
for name in sys.argv[1:] or ["C2", "C2l", "C5"]:
    base = synth.make_config("C5" if name == "C5" else "C2")
    pb = synth.with_loop_closures(base, 10, seed=1) if name == "C2l" else base
    ba.set_problem(pb)
    ba.optimize(10)
    ba.reset_state()
    sys.stderr.write(f"== {name}: P {pb.P} C {pb.C}\n")
    sys.stderr.flush()
    it, st = ba.optimize(10)
    sys.stderr.write(f"   nnzb_S {st['nnzb_S']} nnzb_L {st['nnzb_L']} ms_solve per trial {st['ms_solve'] / st['trials_total']:.4f}\n")
