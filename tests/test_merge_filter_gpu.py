"""top-k merge, filter->bitmask and rerank-tail kernels vs the oracle (integer / index work: exact)."""
import json
import os
import random

import numpy as np
import pytest
import torch

import automative_rag_b200 as rag
from automative_rag_b200 import distributed as D
from automative_rag_b200.filters import INT_MISSING
from oracle import dense as odense
from oracle import filters as ofilters
from oracle import maxsim as omaxsim
from tests._cases import RERANK_CASES, make_rerank_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("nl,nq,k_in,k_out", [(1, 1, 10, 10), (2, 3, 10, 10), (8, 5, 100, 100), (8, 2, 1000, 1000),
                                              (8, 1, 2048, 2048), (4, 7, 33, 50), (3, 2, 5, 2)])
def test_topk_merge_matches_oracle(engine, nl, nq, k_in, k_out):
    rng = np.random.default_rng(nl * 1000 + k_in)
    scores = np.sort(rng.standard_normal((nl, nq, k_in)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    scores[:, :, ::3] = np.round(scores[:, :, ::3], 1)  # force exact ties across lists
    scores = np.sort(scores, axis=2)[:, :, ::-1].copy()
    ids = rng.permutation(nl * nq * k_in).astype(np.int64).reshape(nl, nq, k_in) * 7
    if k_in > 5:  # padding at the tail of the last list
        scores[-1, :, -2:] = -np.inf
        ids[-1, :, -2:] = -1
    ws, wi = odense.merge_topk(scores, ids, k_out)
    s, i = engine.topk_merge(torch.from_numpy(scores).to(engine.device), torch.from_numpy(ids).to(engine.device), k_out)
    np.testing.assert_array_equal(i.cpu().numpy(), wi)
    np.testing.assert_array_equal(s.cpu().numpy(), ws)


@pytest.mark.parametrize("nl,nq,k_in,k_out,mode", [
    (148, 3, 100, 100, "sorted"),      # the batched dense search: 148 ranges x top-100 (pruned: ~300 of 14800 keys sorted)
    (37, 4, 100, 100, "sorted"),
    (148, 2, 16, 10, "sorted"),        # k_out < k_in
    (64, 2, 100, 10, "sorted"),
    (100, 2, 100, 100, "unsorted"),    # the bound holds for any order: same result, less pruning
    (120, 2, 100, 100, "sparse"),      # most lists empty or nearly empty -> the bound switches off
    (16, 3, 1000, 1000, "ties"),       # heavy exact ties at the bound
])
def test_topk_merge_pruned_paths(engine, nl, nq, k_in, k_out, mode):
    rng = np.random.default_rng(nl + k_in)
    scores = rng.standard_normal((nl, nq, k_in)).astype(np.float32)
    if mode == "ties":
        scores = np.round(scores, 1)
    if mode != "unsorted":
        scores = np.sort(scores, axis=2)[:, :, ::-1].copy()
    ids = rng.permutation(nl * nq * k_in).astype(np.int64).reshape(nl, nq, k_in) * 3
    if mode == "sparse":
        keep = rng.integers(0, 4, size=(nl, nq))            # 0..3 valid entries per list
        for l in range(nl):
            for q in range(nq):
                scores[l, q, keep[l, q]:] = -np.inf
                ids[l, q, keep[l, q]:] = -1
        scores[0, :, :] = np.sort(rng.standard_normal((nq, k_in)).astype(np.float32), axis=1)[:, ::-1]
        ids[0, :, :] = np.arange(10**6, 10**6 + nq * k_in).reshape(nq, k_in)
    ws, wi = odense.merge_topk(scores, ids, k_out)
    s, i = engine.topk_merge(torch.from_numpy(scores).to(engine.device), torch.from_numpy(ids).to(engine.device), k_out)
    np.testing.assert_array_equal(i.cpu().numpy(), wi)
    np.testing.assert_array_equal(s.cpu().numpy(), ws)


def test_topk_merge_consumes_the_gathered_wire_buffer_in_place(engine):
    nq, k, world = 3, 10, 4
    rng = np.random.default_rng(0)
    bufs, all_s, all_i = [], [], []
    for r in range(world):
        buf = torch.zeros(D.wire_words(nq, k), dtype=torch.int32)
        s, i = D.wire_views(buf, nq, k)
        sv = np.sort(rng.standard_normal((nq, k)).astype(np.float32), axis=1)[:, ::-1].copy()
        iv = (rng.permutation(nq * k).reshape(nq, k) + r * 1000).astype(np.int64)
        s.copy_(torch.from_numpy(sv))
        i.copy_(torch.from_numpy(iv))
        bufs.append(buf)
        all_s.append(sv)
        all_i.append(iv)
    gathered = torch.cat(bufs).to(engine.device)
    gs, gi = D.gathered_views(gathered, world, nq, k)
    s, i = engine.topk_merge(gs, gi, k)
    ws, wi = odense.merge_topk(np.stack(all_s), np.stack(all_i), k)
    np.testing.assert_array_equal(i.cpu().numpy(), wi)
    np.testing.assert_array_equal(s.cpu().numpy(), ws)


def _random_payload(rng):
    md = {}
    if rng.random() < 0.9:
        md["manufacturer"] = rng.choice(["Toyota", "Honda", "BMW", "Tesla"])
    if rng.random() < 0.8:
        md["year"] = rng.choice([2019, 2020, 2021, 2022, 2023])
    if rng.random() < 0.7:
        md["category"] = rng.choice(["sedan", "suv", "truck"])
    return {"page_content": "x", "metadata": md}


@pytest.mark.parametrize("flt", [{"manufacturer": "Toyota"}, {"manufacturer": ["Toyota", "BMW"], "year": 2021},
                                 {"year": [2019, 2023], "category": "suv"}, {"manufacturer": "Nobody"}, {"year": 2021.5},
                                 {"manufacturer": "Tesla", "year": 2022, "category": ["suv", "sedan"]}])
def test_filter_mask_kernel_matches_oracle(engine, flt):
    from automative_rag_b200.vectorstore import Collection

    rng = random.Random(1)
    n = 10_007
    payloads = [_random_payload(rng) for _ in range(n)]
    col = Collection("t", 8, engine)
    col.upsert([f"id{i}" for i in range(n)], torch.randn(n, 8), payloads)
    deleted = np.zeros(n, bool)
    dele = rng.sample(range(n), 300) + [n - 1, 31, 32]
    deleted[dele] = True
    col.delete([f"id{i}" for i in dele])
    want = ofilters.filter_mask(payloads, flt, deleted)
    got = col.device_mask(rag.build_filter(flt))
    got_bits = odense.unpack_mask(got.cpu().numpy().view(np.uint32), n)
    assert (got_bits == want).all()
    # the device column encoding itself
    assert col.columns["year"][:n].cpu().tolist() == [p["metadata"].get("year", INT_MISSING) for p in payloads]


def test_filter_on_unindexed_key_uses_host_predicate(engine):
    from automative_rag_b200.vectorstore import Collection

    n = 500
    payloads = [{"page_content": "", "metadata": {"custom": i % 3, "manufacturer": "A"}} for i in range(n)]
    col = Collection("t2", 8, engine)
    col.upsert([str(i) for i in range(n)], torch.randn(n, 8), payloads)
    got = col.device_mask(rag.build_filter({"custom": 1}))
    bits = odense.unpack_mask(got.cpu().numpy().view(np.uint32), n)
    assert (bits == (np.arange(n) % 3 == 1)).all()


@pytest.mark.parametrize("name", sorted(RERANK_CASES))
def test_rerank_postprocess_kernel_matches_reference_golden(engine, name):
    """rs_rerank_postprocess vs the reference's own rerank() outputs (tests/golden/rerank_golden.json)."""
    spec = RERANK_CASES[name]
    case = make_rerank_case(spec)
    gold = json.load(open(os.path.join(GOLD, "rerank_golden.json")))[name]["rerank"]
    scores = omaxsim.maxsim_scores(case["queries"][0], case["docs"])  # fp32 oracle scores as kernel input
    s = torch.from_numpy(scores).to(engine.device).reshape(1, -1)
    o = torch.from_numpy(case["bge"]).to(engine.device).reshape(1, -1) if spec["use_bge"] else None
    idx, out = engine.rerank_postprocess(s, o, spec["top_k"] or s.shape[1])
    assert idx[0].tolist() == [i for i, _ in gold]
    np.testing.assert_allclose(out[0].cpu().numpy(), [v for _, v in gold], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,hybrid", [(1, False), (5000, False), (16384, True), (777, True)])
def test_rerank_postprocess_large_and_tied_inputs(engine, n, hybrid):
    """Stable order at sizes up to the device limit, with many exactly tied scores."""
    rng = np.random.default_rng(n)
    a = np.round(rng.standard_normal((2, n)).astype(np.float32), 1)      # heavy ties
    b = np.round(rng.standard_normal((2, n)).astype(np.float32), 1) if hybrid else None
    top_k = min(n, 50)
    idx, out = engine.rerank_postprocess(torch.from_numpy(a).to(engine.device),
                                         None if b is None else torch.from_numpy(b).to(engine.device), top_k)
    for q in range(2):
        want = omaxsim.hybrid_rerank(a[q].tolist(), None if b is None else b[q].tolist(), 0.8, 0.2, top_k)
        got_idx = idx[q].cpu().tolist()
        if not hybrid:
            assert got_idx == [i for i, _ in want]        # exact: pure integer / ordering work
        else:
            # blended scores are fp32 on the device and float64 in Python: order may differ only inside fp32 ties
            ws = np.array([s for _, s in want])
            np.testing.assert_allclose(out[q].cpu().numpy(), ws, rtol=1e-5, atol=1e-6)
            assert sorted(got_idx) == sorted(i for i, _ in want) or np.isclose(ws[-1], out[q, -1].item(), atol=1e-6)


def test_rerank_postprocess_rejects_oversized_candidate_sets(engine):
    with pytest.raises(ValueError):
        engine.rerank_postprocess(torch.zeros(1, 20000, device=engine.device), None, 10)


def test_signed_zero_scores_tie(engine):
    """-0.0 and +0.0 are one score: ordered by id in the merge (as numpy's lexsort does) and by input position in the
    rerank tail (as Python's stable sort does, rerankers.py:377-380)."""
    dev = engine.device
    scores = np.array([[[0.5, -0.0, 0.0, -0.0, 0.0, -1.0]]], dtype=np.float32)
    ids = np.array([[[7, 50, 40, 30, 20, 1]]], dtype=np.int64)
    s, i = engine.topk_merge(torch.from_numpy(scores).to(dev), torch.from_numpy(ids).to(dev), 6)
    assert i.cpu().tolist() == [[7, 20, 30, 40, 50, 1]]
    idx, out = engine.rerank_postprocess(torch.tensor([[-0.0, 0.0, -0.0, 1.0]], device=dev), None, 4)
    assert idx.cpu().tolist() == [[3, 0, 1, 2]]


def test_owned_candidates_kernel_matches_the_host_rule(engine):
    """rs_owned_candidates (sharded MaxSim, round-robin document ownership) against distributed.owned_candidates:
    int32 and int64 ids, padding, every rank of worlds 1 / 2 / 8, with and without the modulo-pool mapping."""
    from automative_rag_b200.distributed import owned_candidates

    g = torch.Generator().manual_seed(5)
    for dtype in (torch.int32, torch.int64):
        cand = torch.randint(0, 1_000_000, (37, 211), generator=g, dtype=dtype)
        cand[::5, ::7] = -1
        for world in (1, 2, 8):
            for rank in range(world):
                for pool in (0, 12_345):
                    got = engine.owned_candidates(cand.to(engine.device), world, rank, pool).cpu()
                    c = cand if pool == 0 else torch.where(cand >= 0, cand % pool, cand)
                    want = owned_candidates(c, world, rank)
                    assert got.dtype == torch.int32 and torch.equal(got, want)
