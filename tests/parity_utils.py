"""Helpers shared by the CPU and GPU parity tests."""
import numpy as np

# north_star tolerance: scores within 1e-3 RELATIVE of the reference's fp32 result
SCORE_RTOL = 1e-3


def check_topk(got_pids, got_scores, ref_pids, ref_scores, tol, all_ref_pids=None, all_ref_scores=None):
    """Position by position the scores agree within ``tol`` (relative, floor 1.0 on the scale); pids are
    identical except inside groups of scores tied within the tolerance.  When the list is truncated
    (top-k of n) pass the reference's full ranking so that a tie straddling the cut is recognised."""
    got_pids, ref_pids = np.asarray(got_pids), np.asarray(ref_pids)
    got_scores = np.asarray(got_scores, dtype=np.float64)
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    assert got_pids.shape == ref_pids.shape, (got_pids.shape, ref_pids.shape)
    scale = np.maximum(np.abs(ref_scores), 1.0)
    err = np.abs(got_scores - ref_scores)
    assert np.all(err <= tol * scale), f"max score error {err.max()} (rel {np.max(err / scale)})"
    pool_p = ref_pids if all_ref_pids is None else np.asarray(all_ref_pids)
    pool_s = ref_scores if all_ref_scores is None else np.asarray(all_ref_scores, dtype=np.float64)
    for i in np.nonzero(got_pids != ref_pids)[0]:
        j = np.nonzero(pool_p == got_pids[i])[0]
        assert j.size >= 1, f"pid {got_pids[i]} at rank {i} is not in the reference ranking"
        assert np.abs(pool_s[j] - ref_scores[i]).min() <= 2 * tol * scale[i], \
            f"rank {i}: got pid {got_pids[i]}, reference pid {ref_pids[i]} and the scores are not tied"
    return float(np.max(err / scale)) if err.size else 0.0
