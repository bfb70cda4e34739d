"""Comparison rules shared by the parity tests (SURVEY.md §8d "Parity protocol").

  * integer / index results: exact;
  * scores: |a - b| <= rtol * max(|a|, |b|) + atol, rtol = 1e-3 for the 16-bit-in / fp32-accumulate
    paths (the tolerance BASELINE.json's north_star states), atol a few fp32 ulps of the score scale;
  * top-k ids: identical, except that two entries may swap / differ when their ORACLE scores are
    within the same tolerance of each other (a tie the accumulation order may break either way).
"""
from __future__ import annotations

import numpy as np

RTOL_16BIT = 1e-3   # north_star: "scores within 1e-3 relative for the bf16-in/fp32-accumulate path"
RTOL_FP32 = 2e-5    # exact-fp32 path: only the summation order differs


def assert_scores_close(got, want, rtol=RTOL_16BIT, atol=1e-6, what="scores"):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    both_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    err = np.abs(got - want)
    tol = rtol * np.maximum(np.abs(got), np.abs(want)) + atol
    bad = ~both_inf & ~(err <= tol)
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {bad.size} outside tolerance; worst abs err "
                           f"{float(np.nanmax(np.where(both_inf, 0, err))):.3e} at {np.argwhere(bad)[:5].tolist()}")


def assert_topk_matches(got_scores, got_ids, all_scores, passing, k, rtol=RTOL_16BIT, atol=1e-6, id_base=0):
    """Check one query's top-k against the oracle's FULL score vector.

    all_scores: oracle scores of every row (float32 [n]); passing: bool [n].
    Valid result: ids distinct, all passing; each returned score matches the oracle score of that id;
    the returned score multiset is the true top-k up to tolerance (no non-returned passing row beats
    the k-th returned score by more than tolerance); order is score-descending; exact ties ordered
    by ascending id; padding (-1, -inf) only when fewer than k rows pass.
    """
    got_scores = np.asarray(got_scores, dtype=np.float64)
    got_ids = np.asarray(got_ids, dtype=np.int64)
    n_pass = int(np.count_nonzero(passing))
    n_valid = min(k, n_pass)
    assert (got_ids[n_valid:] == -1).all() and np.isneginf(got_scores[n_valid:]).all(), "padding must be (-1, -inf)"
    ids = got_ids[:n_valid] - id_base
    assert len(set(ids.tolist())) == n_valid, "duplicate ids in the top-k"
    assert ((ids >= 0) & (ids < len(all_scores))).all(), "id out of range"
    assert np.asarray(passing)[ids].all(), "a filtered-out row was returned"
    ref = np.asarray(all_scores, dtype=np.float64)
    assert_scores_close(got_scores[:n_valid], ref[ids], rtol, atol, "returned scores vs oracle score of the same id")
    # descending, ties by ascending id
    d = np.diff(got_scores[:n_valid])
    assert (d <= 0).all(), "scores not descending"
    tie = d == 0
    assert (np.diff(ids)[tie] > 0).all(), "exact ties must be ordered by ascending id"
    # nothing outside beats the k-th by more than tolerance
    if n_valid and n_pass > n_valid:
        kth = ref[ids].min()
        rest = np.asarray(passing).copy()
        rest[ids] = False
        best_rest = ref[rest].max()
        assert best_rest <= kth + rtol * max(abs(kth), abs(best_rest)) + atol, (
            f"a non-returned row scores {best_rest} > k-th returned {kth}")
    # ids identical to the oracle's except at ties inside tolerance
    want = np.argsort(-np.where(passing, ref, -np.inf), kind="stable")[:n_valid]
    diff = set(ids.tolist()) ^ set(want.tolist())
    if diff:
        kth = ref[want].min()
        for i in diff:
            assert abs(ref[i] - kth) <= rtol * max(abs(ref[i]), abs(kth)) + atol, (
                f"id {i} differs from the oracle's top-k outside a score tie (score {ref[i]}, k-th {kth})")
