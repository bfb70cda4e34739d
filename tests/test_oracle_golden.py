"""The oracle against the golden vectors produced by the reference's own code (CPU, no GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dense as odense
from oracle import maxsim as omaxsim
from tests._cases import MAXSIM_CASES, RERANK_CASES, make_dense_case, make_maxsim_case, make_rerank_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def maxsim_gold():
    return np.load(os.path.join(GOLD, "maxsim_golden.npz"))


@pytest.fixture(scope="module")
def rerank_gold():
    with open(os.path.join(GOLD, "rerank_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(MAXSIM_CASES))
def test_maxsim_oracle_matches_reference_output(name, maxsim_gold):
    """oracle.maxsim == ColBERTReranker._compute_maxsim_scores (rerankers.py:215-265) on the same inputs."""
    q, docs = make_maxsim_case(MAXSIM_CASES[name])
    got = omaxsim.maxsim_scores(q, docs)
    want = maxsim_gold[name]
    assert got.shape == want.shape
    # both are fp32 sums of the same fp32 matmul; only the summation order may differ
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-4)


def test_packed_oracle_equals_list_oracle():
    q, docs = make_maxsim_case(MAXSIM_CASES["ragged"])
    toks = torch.cat(docs)
    off = np.cumsum([0] + [d.shape[0] for d in docs])
    got = omaxsim.maxsim_scores_packed(q, None, toks, off)
    np.testing.assert_allclose(got[0], omaxsim.maxsim_scores(q, docs), rtol=2e-6, atol=1e-4)
    cand = np.array([[8, 0, 3]])
    got_c = omaxsim.maxsim_scores_packed(q, None, toks, off, cand)
    np.testing.assert_allclose(got_c[0], omaxsim.maxsim_scores(q, [docs[8], docs[0], docs[3]]), rtol=2e-6, atol=1e-4)


def test_all_ones_weights_reproduce_the_stale_known_answer():
    """Reference's stale test: row-max mocked to ones over 32 query tokens -> 32.0
    (tests/ test_colbert.py:105-112) == all-ones weights; HEAD's rule drops two tokens -> 30.0."""
    q = torch.ones(32, 4)
    d = torch.full((5, 4), 0.25)  # every similarity is exactly 1.0
    assert omaxsim.maxsim_scores(q, [d], weights=np.ones(32, np.float32))[0] == 32.0
    assert omaxsim.maxsim_scores(q, [d])[0] == 30.0


@pytest.mark.parametrize("name", sorted(RERANK_CASES))
def test_rerank_tail_oracle_matches_reference_output(name, rerank_gold):
    """oracle.hybrid_rerank == ColBERTReranker.rerank / batch_rerank_queries (rerankers.py:267-349, :563-662)."""
    spec = RERANK_CASES[name]
    case = make_rerank_case(spec)
    gold = rerank_gold[name]

    def run(qi, top_k, restrict=None):
        scores = omaxsim.maxsim_scores(case["queries"][qi], case["docs"])
        other = case["bge"] if spec["use_bge"] else None
        return omaxsim.hybrid_rerank(scores, other, 0.8, 0.2, top_k)

    got = run(0, spec["top_k"])
    assert [i for i, _ in got] == [i for i, _ in gold["rerank"]]
    np.testing.assert_allclose([s for _, s in got], [s for _, s in gold["rerank"]], rtol=1e-5, atol=1e-6)
    if spec.get("batch"):
        for qi in range(spec["n_queries"]):
            want = gold["batch"][f"q-{qi}"]
            scores = omaxsim.maxsim_scores(case["queries"][qi], case["docs"])
            if spec["use_bge"]:
                first = omaxsim.hybrid_rerank(scores, None, top_k=spec["top_k"] * 2)   # :604
                keep = [i for i, _ in first]
                second = omaxsim.hybrid_rerank([scores[i] for i in keep], [case["bge"][i] for i in keep],
                                               0.8, 0.2, spec["top_k"])
                got_b = [(keep[i], s) for i, s in second]
            else:
                got_b = omaxsim.hybrid_rerank(scores, None, top_k=spec["top_k"])
            assert [i for i, _ in got_b] == [i for i, _ in want]
            np.testing.assert_allclose([s for _, s in got_b], [s for _, s in want], rtol=1e-5, atol=1e-6)


def test_stable_tie_order_is_input_order(rerank_gold):
    ids = [i for i, _ in rerank_gold["ties"]["rerank"]]
    assert ids.index(1) < ids.index(4) and ids.index(2) < ids.index(7)


# ------------------------------------------------------------------------------------ dense
def test_dense_oracle_against_bruteforce_loops():
    c, q = make_dense_case(3, 257, 64)
    mask = np.random.default_rng(0).random(257) < 0.5
    s, i = odense.topk(c.numpy(), q.numpy(), 10, mask)
    cf, qf = c.float().numpy(), q.float().numpy()
    qn = qf / np.linalg.norm(qf)
    brute = sorted(((float(np.dot(cf[r], qn)), r) for r in range(257) if mask[r]), key=lambda t: (-t[0], t[1]))[:10]
    assert i.tolist() == [r for _, r in brute]
    np.testing.assert_allclose(s, [v for v, _ in brute], rtol=1e-5, atol=1e-6)


def test_dense_oracle_agrees_with_independent_cosine_knn_libraries():
    """The dense oracle stays PARITY UNPINNED against the reference (qdrant-client is absent, oracle/dense.py), but it
    is not alone: two independent exact-cosine implementations installed here — scikit-learn's brute-force
    NearestNeighbors(metric="cosine") and scipy's cdist — return the same ids and scores on unnormalised fp32 rows
    (inv_norm path) and on unit rows, with and without a filter."""
    from scipy.spatial.distance import cdist
    from sklearn.neighbors import NearestNeighbors

    rng = np.random.default_rng(11)
    for n, d, k, p in ((4000, 96, 10, 1.0), (3001, 1024, 25, 0.4), (513, 8, 100, 0.7)):
        c = (rng.standard_normal((n, d)) * rng.uniform(0.2, 3.0, size=(n, 1))).astype(np.float32)
        q = rng.standard_normal(d).astype(np.float32)
        bits = rng.random(n) < p
        inv = (1.0 / np.linalg.norm(c.astype(np.float64), axis=1)).astype(np.float32)
        s, i = odense.topk(c, q, k, bits, inv_norm=inv)
        rows = np.flatnonzero(bits)
        nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(c[rows].astype(np.float64))
        dist, idx = nn.kneighbors(q[None].astype(np.float64))
        assert i.tolist() == rows[idx[0]].tolist()
        np.testing.assert_allclose(s, 1.0 - dist[0], rtol=2e-5, atol=2e-6)
        full = 1.0 - cdist(q[None].astype(np.float64), c.astype(np.float64), metric="cosine")[0]
        np.testing.assert_allclose(odense.scores_f32(c, q, odense.COSINE, inv), full, rtol=2e-5, atol=2e-6)


def test_dense_oracle_padding_and_ties():
    c = torch.zeros(6, 8, dtype=torch.float16)
    c[:, 0] = 1.0  # six identical rows -> six exactly tied scores
    q = torch.zeros(8, dtype=torch.float16)
    q[0] = 1.0
    s, i = odense.topk(c.numpy(), q.numpy(), 4, np.array([1, 0, 1, 1, 1, 1], bool), id_base=100)
    assert i.tolist() == [100, 102, 103, 104]  # ascending id inside the tie
    s, i = odense.topk(c.numpy(), q.numpy(), 4, np.array([0, 0, 1, 0, 0, 0], bool))
    assert i.tolist() == [2, -1, -1, -1] and np.isneginf(s[1:]).all()


def test_mask_pack_roundtrip():
    for n in (1, 31, 32, 33, 1000):
        b = np.random.default_rng(n).random(n) < 0.3
        w = odense.pack_mask(b)
        assert w.dtype == np.uint32 and w.shape == ((n + 31) // 32,)
        assert (odense.unpack_mask(w, n) == b).all()
        for r in (0, n - 1):
            assert bool((w[r >> 5] >> (r & 31)) & 1) == bool(b[r])  # LSB-first, as rs_dense_topk reads it


def test_merge_oracle():
    s = np.array([[[0.9, 0.5, 0.1]], [[0.9, 0.6, -np.inf]]], dtype=np.float32)
    i = np.array([[[7, 3, 1]], [[2, 9, -1]]], dtype=np.int64)
    ms, mi = odense.merge_topk(s, i, 4)
    assert mi.tolist() == [[2, 7, 9, 3]]
