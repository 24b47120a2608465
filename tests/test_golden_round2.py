"""The oracles of the rows added in round 2 (residual image, window selection, keyframe growth) against their
committed outputs (tests/golden/round2_golden.npz, scripts/make_golden_round2.py).  Oracle outputs, not reference
outputs -- the reference has no golden vectors (PARITY UNPINNED) -- so the file guards against drift only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import make_golden_round2 as gen   # noqa: E402  (the generator's compute() is the single definition of the cases)

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "round2_golden.npz"))


def test_round2_oracles_reproduce_their_golden_outputs():
    out = gen.compute()
    assert set(out) == set(G.files)
    for k, v in out.items():
        g = G[k]
        v = np.asarray(v)
        if v.dtype.kind in "iub":
            np.testing.assert_array_equal(v, g, err_msg=k)
        elif k.startswith("res_"):      # FP32 per-pixel arithmetic without contraction: bit-exact
            np.testing.assert_array_equal(v, g, err_msg=k)
        else:
            np.testing.assert_allclose(v, g, rtol=1e-12, atol=1e-14, err_msg=k)
