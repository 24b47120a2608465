"""Developer probe (GPU box): k_build_wave time vs task chunk (SVS_BUILD_CHUNK) on C2."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from scavislam_b200 import synth, capi
    ba = capi.BundleAdjuster(); pb = synth.make_config("C2"); ba.set_problem(pb)
    best = 1e9
    for r in range(4):
        ba.reset_state(); it, st = ba.optimize(10)
        best = min(best, st["ms_build"])
    print("chunk", os.environ.get("SVS_BUILD_CHUNK", "default"), "ms_build/10it %.4f" % best, "ms_total %.4f" % st["ms_total"])
else:
    for c in ("", "4", "6", "8", "10", "12", "16", "24"):
        env = dict(os.environ)
        if c:
            env["SVS_BUILD_CHUNK"] = c
        subprocess.run([sys.executable, __file__, "x"], env=env)
