"""Attribution of device-vs-oracle differences (BASELINE north_star: replayed runs are exact except where a slot
uniform lies within 1e-12 of a CDF boundary, or — for MH — an accept decision is a numerical tie).

`attribute(...)` compares final columns at 1e-9 relative and requires every differing particle to be FRAGILE in
the oracle's bookkeeping (oracle/ref.py: OracleState.fragile_tie / fragile_anc, propagated through every
resampling step).  The counts go to gpurun_out/parity_attribution.jsonl (copied to profiles/ after a GPU run)."""
import json
import os

import numpy as np

REL = 1e-9
_OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_attribution.jsonl")


def attribute(label, state, ost, cols, rel=REL):
    n = ost.n
    diff = np.zeros(n, dtype=bool)
    for c in cols:
        a, b = state[c], ost.cols[c]
        assert a.shape == b.shape, c
        diff |= (np.abs(a - b) > rel * (1.0 + np.abs(b))).reshape(n, -1).any(axis=1)
    tie = diff & ost.fragile_tie
    anc = diff & ost.fragile_anc & ~ost.fragile_tie
    unexplained = diff & ~ost.fragile_tie & ~ost.fragile_anc
    rec = {"test": label, "n": n, "rel_tol": rel, "differing_particles": int(diff.sum()),
           "attributed_near_tie_accept": int(tie.sum()), "attributed_near_boundary_ancestor": int(anc.sum()),
           "unexplained": int(unexplained.sum()),
           "oracle_near_tie_accepts_seen": int(ost.n_near_tie), "oracle_near_boundary_uniforms_seen": int(ost.n_near_boundary)}
    try:
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        with open(_OUT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    print(f"[attribution] {rec}")
    assert rec["unexplained"] == 0, rec
    return rec
