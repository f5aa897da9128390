"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/baselines.npz by running the UNMODIFIED reference ``rrwm_ged`` and
``ga_ged`` (utils/ged_env.py:424-477) on its own dense K (``construct_k``), single-threaded, on small synthetic pairs.

The functions return only the projected assignment; the iterate ``v`` handed to the final ``hungarian_lap(-v)`` is captured
by wrapping the reference module's ``hungarian_lap`` for the duration of the call (the function itself is untouched).

    python oracle/gen_golden_baselines.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import OUT, pack_ragged, pairs_blob, small_pairs, to_data  # noqa: E402


def captured(ref, fn, *a, **kw):
    """Run ``fn`` and return (result, list of the matrices passed to hungarian_lap)."""
    seen, orig = [], ref.hungarian_lap

    def spy(mat, n1, n2):
        seen.append(mat.detach().clone().numpy())
        return orig(mat, n1, n2)
    ref.hungarian_lap = spy
    try:
        out = fn(*a, **kw)
    finally:
        ref.hungarian_lap = orig
    return out, seen


def main():
    torch.set_num_threads(1)
    ref = ref_shim.load_reference()
    env = ref_shim.make_reference_env("ipfp", "AIDS-20-30")
    pairs = small_pairs(4, 5, 9, 11) + small_pairs(4, 10, 16, 12)
    blob = pairs_blob(pairs)
    out = {k: [] for k in ("rrwm_v1", "rrwm_v5", "rrwm_v", "rrwm_x", "rrwm_ged", "ga_soft", "ga_x", "ga_ged", "hun_x")}
    for g1, g2 in pairs:
        d1, d2 = to_data(g1), to_data(g2)
        k = env.construct_k(d1, d2).squeeze(0)
        n1, n2 = g1.n, g2.n
        out["hun_x"].append(ref.hungarian_ged(k, n1, n2).numpy())
        for key, it in (("rrwm_v1", 1), ("rrwm_v5", 5), ("rrwm_v", 100)):
            x, seen = captured(ref, ref.rrwm_ged, k, n1, n2, max_iter=it)
            out[key].append(-seen[-1])                       # the last hungarian_lap call is the projection of -v (:476)
        out["rrwm_x"].append(x.numpy())
        out["rrwm_ged"].append(float(env.comp_ged(x, k)))
        x, seen = captured(ref, ref.ga_ged, k, n1, n2)
        # hungarian_lap call t of the hard stage receives -K v_t; v_0 (the soft result) is not passed anywhere, so keep K v_0
        out["ga_soft"].append(-seen[0])
        out["ga_x"].append(x.numpy())
        out["ga_ged"].append(float(env.comp_ged(x, k)))
        print(n1, n2, "rrwm", out["rrwm_ged"][-1], "ga", out["ga_ged"][-1], "hard iterations", len(seen))
    for key in ("rrwm_v1", "rrwm_v5", "rrwm_v", "rrwm_x", "ga_soft", "ga_x", "hun_x"):
        blob[key], blob["mat_off"] = pack_ragged(out[key], np.float32)
    blob["rrwm_ged"] = np.asarray(out["rrwm_ged"], np.float32)
    blob["ga_ged"] = np.asarray(out["ga_ged"], np.float32)
    np.savez_compressed(os.path.join(OUT, "baselines.npz"), **blob)


if __name__ == "__main__":
    main()
