"""Batched dense top-k (tcgen05 GEMM + fused per-query top-k) vs the CPU oracle and vs the scan kernel."""
import numpy as np
import pytest
import torch

from automative_rag_b200 import _ffi
from oracle import dense as odense
from tests._cases import bernoulli_mask
from tests._parity import assert_topk_matches

pytestmark = pytest.mark.gpu


def _case(seed, n, d, nq, dtype, normalise=True):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(n, d, generator=g)
    q = torch.randn(nq, d, generator=g)
    if normalise:
        c = c / c.norm(dim=1, keepdim=True)
        q = q / q.norm(dim=1, keepdim=True)
    return c.to(dtype), q.to(dtype)


def _run(engine, c, q, k, impl, **kw):
    engine.set_dense_impl(impl)
    try:
        s, i = engine.dense_topk(c.to(engine.device), q.to(engine.device), k, **kw)
        torch.cuda.synchronize()
        assert engine.last_dense_impl == impl
    finally:
        engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    return s.cpu().numpy(), i.cpu().numpy()


def _check_all(got_s, got_i, c, q, k, bits=None, metric=_ffi.RS_METRIC_COSINE, inv_norm=None, id_base=0):
    cf, qf = c.float().numpy(), q.float().numpy()
    passing = np.ones(c.shape[0], bool) if bits is None else bits
    for j in range(q.shape[0]):
        all_scores = odense.scores_f32(cf, qf[j], metric, inv_norm)
        assert_topk_matches(got_s[j], got_i[j], all_scores, passing, k, id_base=id_base)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,d,nq,k", [(256, 64, 32, 1), (1000, 128, 100, 10), (5000, 1024, 256, 100),
                                      (70001, 1024, 300, 100), (20000, 512, 1024, 128), (300, 64, 33, 128),
                                      (3000, 1024, 4, 10), (40_000, 512, 7, 100), (9000, 256, 31, 128),
                                      (6000, 128, 129, 10)])
def test_batched_matches_oracle(engine, dtype, n, d, nq, k):
    c, q = _case(n + nq, n, d, nq, dtype)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05)
    _check_all(s, i, c, q, k)


def test_batched_with_mask_ip_inv_norm_and_id_base(engine):
    n, d, nq, k = 30_011, 1024, 64, 50
    c, q = _case(7, n, d, nq, torch.bfloat16, normalise=False)
    bits = bernoulli_mask(8, n, 0.3)
    mask = torch.from_numpy(odense.pack_mask(bits).view(np.int32).copy()).to(engine.device)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=mask, metric=_ffi.RS_METRIC_IP, id_base=5_000_000_000)
    _check_all(s, i, c, q, k, bits, metric=_ffi.RS_METRIC_IP, id_base=5_000_000_000)
    inv = (1.0 / np.linalg.norm(c.float().numpy(), axis=1)).astype(np.float32)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=mask, inv_norm=torch.from_numpy(inv).to(engine.device))
    _check_all(s, i, c, q, k, bits, inv_norm=inv)


def test_batched_few_passing_rows_pads(engine):
    n, d, nq, k = 4096, 128, 40, 20
    c, q = _case(9, n, d, nq, torch.float16)
    bits = np.zeros(n, bool)
    bits[[5, 77, 4095]] = True
    mask = torch.from_numpy(odense.pack_mask(bits).view(np.int32).copy()).to(engine.device)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=mask)
    _check_all(s, i, c, q, k, bits)
    assert (i[:, 3:] == -1).all() and np.isneginf(s[:, 3:]).all()


def test_batched_exact_ties_ordered_by_id(engine):
    n, d, nq, k = 9000, 128, 48, 16
    c, q = _case(10, n, d, nq, torch.bfloat16)
    dup = list(range(100, 8000, 311))
    c[dup] = c[100].clone()
    q[:] = c[100].clone()
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05)
    for j in range(nq):
        assert i[j].tolist() == dup[:k]


def test_batched_all_rows_equal_keeps_lowest_ids(engine):
    """Every row scores the same: the cross-range bound equals that score, rows EQUAL to it must survive, and the
    (score desc, id asc) order leaves exactly ids 0..k-1."""
    n, d, nq, k = 40_000, 128, 64, 100
    c, q = _case(12, n, d, nq, torch.bfloat16)
    c[:] = c[0].clone()
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05)
    for j in range(nq):
        assert i[j].tolist() == list(range(k))
        assert (s[j] == s[j][0]).all()


@pytest.mark.parametrize("order", ["ascending", "descending", "one_range_holds_all"])
def test_batched_adversarial_row_orders(engine, order):
    """Row orders that stress the running thresholds: scores rising with the row id (every row beats the local
    bound), falling (bounds tight from the first tile), and all winners inside one range (the other ranges publish
    low bounds, which must not cut anything)."""
    n, d, nq, k = 60_000, 128, 40, 100
    g = torch.Generator().manual_seed(13)
    base = torch.randn(d, generator=g)
    base = base / base.norm()
    noise = torch.randn(n, d, generator=g) * 0.05
    if order == "one_range_holds_all":
        w = torch.full((n,), 0.1)
        w[41_000:41_300] = torch.linspace(0.5, 1.0, 300)
    else:
        w = torch.linspace(0.1, 1.0, n) if order == "ascending" else torch.linspace(1.0, 0.1, n)
    c = (w[:, None] * base[None, :] + noise).to(torch.bfloat16)
    q = (base[None, :] + 0.01 * torch.randn(nq, d, generator=g)).to(torch.bfloat16)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, metric=_ffi.RS_METRIC_IP)
    _check_all(s, i, c, q, k, metric=_ffi.RS_METRIC_IP)


def test_batched_masked_out_ranges(engine):
    """Whole ranges filtered out: they never publish a bound, so no cross-range cut may happen — and the result is
    still the exact top-k of the passing rows."""
    n, d, nq, k = 50_000, 128, 48, 64
    c, q = _case(14, n, d, nq, torch.float16)
    bits = np.zeros(n, bool)
    bits[30_000:30_500] = True
    bits[49_990:] = True
    mask = torch.from_numpy(odense.pack_mask(bits).view(np.int32).copy()).to(engine.device)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=mask)
    _check_all(s, i, c, q, k, bits)


def test_batched_agrees_with_scan_and_auto_dispatch(engine):
    n, d, nq, k = 50_000, 1024, 128, 100
    c, q = _case(11, n, d, nq, torch.bfloat16)
    bs, bi = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05)
    ss, si = _run(engine, c, q, k, _ffi.RS_DENSE_SCAN)
    np.testing.assert_allclose(bs, ss, rtol=1e-3, atol=1e-6)
    same = (bi == si).mean()
    assert same > 0.999, f"only {same:.4f} of ids identical between the GEMM and the scan kernel"
    engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    engine.dense_topk(c.to(engine.device), q.to(engine.device), k)
    assert engine.last_dense_impl == _ffi.RS_DENSE_TCGEN05    # a batch goes to the tensor cores
    engine.dense_topk(c.to(engine.device), q[:4].to(engine.device), k)
    assert engine.last_dense_impl == _ffi.RS_DENSE_TCGEN05    # ... from four queries on (one pass over the corpus)
    engine.dense_topk(c.to(engine.device), q[:3].to(engine.device), k)
    assert engine.last_dense_impl == _ffi.RS_DENSE_SCAN       # fewer over a small corpus: a loop of scans
    engine.dense_topk(c.to(engine.device), q[:1].to(engine.device), k)
    assert engine.last_dense_impl == _ffi.RS_DENSE_SCAN       # a single query to the HBM-bound scan
    big = torch.zeros(200_000, 64, dtype=torch.bfloat16, device=engine.device)
    big[:, 0] = 1
    q2 = torch.ones(2, 64, dtype=torch.bfloat16, device=engine.device)
    s2, i2 = engine.dense_topk(big, q2, 3)
    assert engine.last_dense_impl == _ffi.RS_DENSE_TCGEN05    # two queries over a large corpus: one pass again
    assert i2.cpu().tolist() == [[0, 1, 2], [0, 1, 2]]        # all rows tie: the lowest ids


def test_config3_reduced_rows_properties(engine):
    """BASELINE config 3 shape (1024 queries, d=1024, bf16, top-100) over 2M rows (the full 10M x 1024 corpus
    is 20 GB; 2M keeps the test in seconds): returned scores recomputed from the rows, threshold check on a
    sample, and exact agreement with the scan kernel for a few queries."""
    n, d, nq, k = 2_000_000, 1024, 1024, 100
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(4)
    c = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
    for lo in range(0, n, 250_000):
        blk = torch.randn(250_000, d, generator=g, device=dev)
        c[lo: lo + 250_000] = (blk / blk.norm(dim=1, keepdim=True)).bfloat16()
    q = torch.randn(nq, d, generator=torch.Generator(device=dev).manual_seed(5), device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).bfloat16()
    engine.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
    s, i = engine.dense_topk(c, q, k)
    engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    torch.cuda.synchronize()
    assert (i >= 0).all() and (i < n).all()
    assert (torch.diff(s, dim=1) <= 0).all()
    qn = q.float() / q.float().norm(dim=1, keepdim=True)
    # (i) scores of the returned ids recomputed with plain torch
    for j in (0, 1, 511, 1023):
        ref = (c[i[j]].float() @ qn[j])
        torch.testing.assert_close(s[j], ref, rtol=1e-3, atol=1e-6)
        assert len(set(i[j].tolist())) == k
    # (ii) threshold check: no row of a 200k-row sample beats the k-th returned score by more than tolerance
    sample = torch.randint(0, n, (200_000,), generator=torch.Generator(device=dev).manual_seed(6), device=dev)
    for j in (3, 700):
        sc = c[sample].float() @ qn[j]
        kth = s[j, -1]
        returned = torch.isin(sample, i[j])
        assert (sc[~returned] <= kth + 1e-3 * kth.abs() + 1e-6).all()
    # (iii) the scan kernel agrees on full queries
    engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
    ss, si = engine.dense_topk(c, q[:4], k)
    engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    torch.testing.assert_close(s[:4], ss, rtol=1e-3, atol=1e-6)
    assert (i[:4] == si).float().mean() > 0.99


@pytest.mark.parametrize("nq", [5, 130, 300])
def test_batched_with_a_filter_per_query(engine, nq):
    """the reference sends a metadata filter per request (retrieval_tasks.py:74-79), so a batch
    carries one mask per query ([nq, words]).  The batched tcgen05 kernel tests each query's own mask word in the
    epilogue; every query is checked against the oracle with ITS filter, and AUTO now picks the batched kernel."""
    n, d, k = 50_021, 256, 24
    c, q = _case(300 + nq, n, d, nq, torch.bfloat16)
    bits = np.stack([bernoulli_mask(1000 + j, n, [0.9, 0.3, 0.02][j % 3]) for j in range(nq)])
    bits[1] = False
    bits[1, [3, 999, 50_020]] = True                                    # a query with three passing rows: padded result
    masks = torch.from_numpy(np.stack([odense.pack_mask(b).view(np.int32) for b in bits]).copy()).to(engine.device)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=masks)
    cf, qf = c.float().numpy(), q.float().numpy()
    for j in range(nq):
        assert_topk_matches(s[j], i[j], odense.scores_f32(cf, qf[j]), bits[j], k)
    assert (i[1, 3:] == -1).all()
    engine.dense_topk(c.to(engine.device), q.to(engine.device), k, mask=masks)
    assert engine.last_dense_impl == _ffi.RS_DENSE_TCGEN05


@pytest.mark.parametrize("scale", [1e17, 1e-17])
def test_batched_first_tile_threshold_extreme_magnitudes(engine, scale):
    """The first tile's threshold is bisected between its lowest and highest score: inner products near the ends
    of the fp32 range must neither overflow the midpoint nor lose rows."""
    n, d, nq, k = 4096, 128, 16, 10
    c, q = _case(91, n, d, nq, torch.bfloat16, normalise=False)
    c = (c.float() * scale).to(torch.bfloat16)
    q = (q.float() * scale).to(torch.bfloat16)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, metric=_ffi.RS_METRIC_IP)
    _check_all(s, i, c, q, k, metric=_ffi.RS_METRIC_IP)


@pytest.mark.parametrize("passing_in_first_tile", [0, 9, 10, 11])
def test_batched_first_tile_threshold_with_sparse_filter(engine, passing_in_first_tile):
    """k = 10 with 0 / k-1 / k / k+1 admissible rows in each range's first tile: the bisected threshold may only be
    used when the tile holds at least k admissible rows."""
    n, d, nq, k = 148 * 256 * 3, 64, 8, 10
    c, q = _case(92, n, d, nq, torch.bfloat16)
    bits = np.zeros(n, bool)
    g = np.random.default_rng(5)
    for t in range(n // 256):
        m = passing_in_first_tile if t < 148 else 40
        bits[t * 256 + g.choice(256, m, replace=False)] = True
    mask = torch.from_numpy(odense.pack_mask(bits).view(np.int32).copy()).to(engine.device)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=mask)
    _check_all(s, i, c, q, k, bits)


# ---------------------------------------------------------------------------------- 128 < k <= 1024: two-level lists
@pytest.mark.parametrize("n,d,nq,k,p", [(200_000, 128, 16, 1000, 1.0), (60_000, 1024, 300, 500, 1.0),
                                        (300_000, 64, 5, 1024, 0.5), (150_000, 256, 130, 129, 1.0),
                                        (90_000, 128, 2, 300, 0.2)])
def test_batched_long_lists_match_oracle(engine, n, d, nq, k, p):
    """k up to 1024 in ONE batched pass (config 5's stage 1 with k1 = 1000, `tests/test_retrieval.py:206-258` with a large
    retrieval_k): every corpus range keeps its best <= 128 rows, the merge takes the k best of all lists; on random data
    no range holds more of the answer than it can keep, so nothing is re-run."""
    c, q = _case(n + k, n, d, nq, torch.bfloat16)
    bits = None if p == 1.0 else bernoulli_mask(k, n, p)
    kw = {}
    if bits is not None:
        kw["mask"] = torch.from_numpy(odense.pack_mask(bits).view(np.int32).copy()).to(engine.device)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, **kw)
    assert engine.last_dense_redo == 0
    _check_all(s, i, c, q, k, bits)


def test_batched_long_lists_redo_when_a_range_overflows(engine):
    """The exactness check of the two-level lists: (a) the whole answer sits in ONE 256-row tile (one range, which keeps
    at most 128 rows), (b) thousands of exactly tied top scores.  Both must be detected and re-run through the
    single-query scan, and the result must equal the oracle's (ties in ascending id)."""
    n, d, nq, k = 100_000, 128, 6, 200
    c, q = _case(91, n, d, nq, torch.float16)
    c = c * 0.1                                            # background rows score low
    qn = (q[0].float() / q[0].float().norm())
    for r in range(256):                                   # rows 0..255 (tile 0): distinct high scores for query 0
        c[r] = (qn * (1.0 - r * 1e-3)).half()
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05)
    assert engine.last_dense_redo >= 1
    _check_all(s, i, c, q, k)
    assert i[0].tolist() == list(range(k))
    # (b) 3000 copies of one row: an exactly tied top block larger than k
    c2, q2 = _case(92, 50_000, 128, 4, torch.float16)
    c2[1000:4000] = (q2[1].float() / q2[1].float().norm()).half()
    s, i = _run(engine, c2, q2, 300, _ffi.RS_DENSE_TCGEN05)
    assert engine.last_dense_redo >= 1
    _check_all(s, i, c2, q2, 300)
    assert i[1].tolist() == list(range(1000, 1300))


def test_auto_takes_long_lists_to_the_batched_kernel(engine):
    n, d, nq, k = 250_000, 128, 8, 1000
    c, q = _case(93, n, d, nq, torch.float16)
    engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    s, i = engine.dense_topk(c.to(engine.device), q.to(engine.device), k)
    torch.cuda.synchronize()
    assert engine.last_dense_impl == _ffi.RS_DENSE_TCGEN05
    _check_all(s.cpu().numpy(), i.cpu().numpy(), c, q, k)
    s2, i2 = _run(engine, c, q, k, _ffi.RS_DENSE_SCAN)
    assert (i2 == i.cpu().numpy()).mean() > 0.99            # ids equal the scan's except near-ties


def test_batched_long_lists_per_query_masks_ip_inv_norm_id_base(engine):
    """The long-list path with everything else the entry point takes: a filter per query, inner product with and
    without inv_norm, an id offset (the shard offset of `ShardedDenseIndex`)."""
    n, d, nq, k = 120_000, 128, 9, 400
    c, q = _case(95, n, d, nq, torch.bfloat16, normalise=False)
    bits = np.stack([bernoulli_mask(100 + j, n, 0.25 + 0.05 * j) for j in range(nq)])
    mask = torch.stack([torch.from_numpy(odense.pack_mask(b).view(np.int32).copy()) for b in bits]).to(engine.device)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=mask, metric=_ffi.RS_METRIC_IP, id_base=7_000_000_000)
    cf, qf = c.float().numpy(), q.float().numpy()
    for j in range(nq):
        assert_topk_matches(s[j], i[j], odense.scores_f32(cf, qf[j], _ffi.RS_METRIC_IP), bits[j], k, id_base=7_000_000_000)
    inv = (1.0 / np.linalg.norm(cf, axis=1)).astype(np.float32)
    s, i = _run(engine, c, q, k, _ffi.RS_DENSE_TCGEN05, mask=mask, inv_norm=torch.from_numpy(inv).to(engine.device))
    for j in range(nq):
        assert_topk_matches(s[j], i[j], odense.scores_f32(cf, qf[j], _ffi.RS_METRIC_COSINE, inv), bits[j], k)
