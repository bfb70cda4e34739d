"""Dense top-k parity: CUDA scan (through the C ABI) vs the CPU oracle on identical rounded inputs."""
import numpy as np
import pytest
import torch

from automative_rag_b200 import _ffi
from oracle import dense as odense
from tests._cases import bernoulli_mask, make_dense_case
from tests._parity import assert_topk_matches

pytestmark = pytest.mark.gpu


def _to_np32(t: torch.Tensor) -> np.ndarray:
    return t.float().cpu().numpy()


def _mask_dev(bits, device):
    return torch.from_numpy(odense.pack_mask(bits).view(np.int32).copy()).to(device)


def _check(engine, corpus, query, k, mask_bits=None, metric=_ffi.RS_METRIC_COSINE, inv_norm=None, id_base=0):
    dev = engine.device
    c, q = corpus.to(dev), query.to(dev)
    mask = None if mask_bits is None else _mask_dev(mask_bits, dev)
    inv = None if inv_norm is None else torch.from_numpy(inv_norm).to(dev)
    engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
    s, i = engine.dense_topk(c, q, k, mask=mask, inv_norm=inv, metric=metric, id_base=id_base)
    torch.cuda.synchronize()
    assert engine.last_dense_impl == _ffi.RS_DENSE_SCAN
    all_scores = odense.scores_f32(_to_np32(corpus), _to_np32(query), metric, inv_norm)
    passing = np.ones(corpus.shape[0], bool) if mask_bits is None else mask_bits
    assert_topk_matches(s[0].cpu().numpy(), i[0].cpu().numpy(), all_scores, passing, k, id_base=id_base)
    return s, i


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,d,k", [(1, 64, 1), (7, 8, 10), (1000, 1024, 10), (4097, 1024, 100), (20011, 128, 1000),
                                   (3000, 264, 17), (50000, 1024, 2048), (777, 2048, 5), (300, 4096, 3)])
def test_scan_matches_oracle(engine, dtype, n, d, k):
    corpus, query = make_dense_case(1000 + n, n, d, dtype)
    _check(engine, corpus, query, k)


@pytest.mark.parametrize("p", [1.0, 0.5, 0.1, 0.01, 0.0])
def test_scan_with_filter_mask(engine, p):
    n, d, k = 30001, 1024, 10
    corpus, query = make_dense_case(2, n, d)
    bits = bernoulli_mask(3, n, p)
    _check(engine, corpus, query, k, bits)


def test_mask_edge_patterns(engine):
    n, d = 5000, 1024
    corpus, query = make_dense_case(4, n, d)
    for bits in (np.arange(n) % 8 == 3, np.arange(n) < 9, np.arange(n) >= n - 3, (np.arange(n) // 8) % 2 == 0):
        _check(engine, corpus, query, 10, bits)
    one = np.zeros(n, bool)
    one[n - 1] = True
    _check(engine, corpus, query, 10, one)  # fewer passing rows than k -> (-inf, -1) padding


@pytest.mark.parametrize("n,d,k", [(200_003, 64, 10), (150_000, 128, 33), (400_001, 256, 32), (90_000, 1024, 32),
                                   (90_000, 1024, 33), (90_000, 1024, 48), (90_000, 1024, 49), (33, 1024, 10), (4737, 1024, 10)])
def test_guided_work_distribution_shapes(engine, n, d, k):
    """Shapes that exercise the scan's work counter and both cross-CTA merges: small d (a grab is several mask words,
    32-row tiles), many grabs per CTA, k on both sides of the tournament limit (48; 32 / 33 were its round-1 sides), fewer words than SMs."""
    corpus, query = make_dense_case(77 + n, n, d)
    bits = bernoulli_mask(n, n, 0.6)
    _check(engine, corpus, query, k, bits)
    _check(engine, corpus, query, k)


@pytest.mark.parametrize("n,d,p", [(70_001, 16, 0.2), (50_000, 48, 0.1), (50_000, 80, 0.3), (40_003, 136, 0.1),
                                   (60_000, 512, 0.1), (120_000, 1024, 0.05), (20_000, 2048, 0.1)])
def test_sparse_filter_gather_tiles(engine, n, d, p):
    """Sparse filters: passing rows of sparse mask words travel as gather tiles — four rows per TMA tile::gather4
    through the row tensor map when d % 16 == 0 and d <= 1024, one bulk copy per run of rows otherwise (d = 136,
    2048 here) and for the last partial tile.  Reference: the payload filter of vectorstore.py:188-197."""
    corpus, query = make_dense_case(300 + d, n, d)
    _check(engine, corpus, query, 10, bernoulli_mask(d, n, p))
    runs = (np.arange(n) // 3) % 7 == 0  # runs of three passing rows: gather tiles straddle runs
    _check(engine, corpus, query, 17, runs)


def test_back_to_back_launches_share_nothing(engine):
    """40 queries in one call (chained launches under programmatic dependent launch, rotating work counters and
    alternating workspaces), three times over: every query's result equals its own single-launch result."""
    n, d, k, nq = 60_000, 256, 10, 40
    corpus, _ = make_dense_case(5, n, d)
    g = torch.Generator().manual_seed(6)
    queries = torch.randn(nq, d, generator=g).half()
    dev = engine.device
    c, q = corpus.to(dev), queries.to(dev)
    engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
    try:
        single = [engine.dense_topk(c, q[j], k) for j in range(nq)]
        for _ in range(3):
            s, i = engine.dense_topk(c, q, k)
            for j in range(nq):
                assert torch.equal(i[j], single[j][1][0]) and torch.equal(s[j], single[j][0][0])
    finally:
        engine.set_dense_impl(_ffi.RS_DENSE_AUTO)


def test_scan_trace_records_every_phase(engine):
    """rs_set_scan_trace: every CTA stamps its phases in order; exactly one CTA (the last to publish) runs the merge."""
    n, d, k = 300_000, 1024, 10
    corpus, query = make_dense_case(8, n, d)
    dev = engine.device
    c, q = corpus.to(dev), query.to(dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    trace = torch.zeros(8, sms, 8, dtype=torch.int64, device=dev)
    engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
    engine.set_scan_trace(trace)
    try:
        s, i = engine.dense_topk(c, q, k)
        torch.cuda.synchronize()
    finally:
        engine.set_scan_trace(None)
        engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    t = trace.cpu()
    used = [b for b in range(8) if int(t[b, :, 0].max()) > 0]
    assert len(used) == 1
    tt = t[used[0]]
    assert (tt[:, 0] > 0).all()                       # one CTA per SM
    assert (tt[:, 1:7] >= tt[:, 0:6]).all()           # phases in order
    assert int((tt[:, 7] > 0).sum()) == 1             # a single merging CTA
    s2, i2 = engine.dense_topk(c, q, k)               # tracing does not change the result
    assert torch.equal(i, i2) and torch.equal(s, s2)


def test_inner_product_and_inv_norm(engine):
    n, d = 9000, 1024
    corpus, query = make_dense_case(5, n, d, normalise=False)
    _check(engine, corpus, query, 20, metric=_ffi.RS_METRIC_IP)
    inv = (1.0 / np.linalg.norm(_to_np32(corpus), axis=1)).astype(np.float32)
    _check(engine, corpus, query, 20, metric=_ffi.RS_METRIC_COSINE, inv_norm=inv, id_base=10**10)


def test_exact_ties_are_ordered_by_id(engine):
    n, d, k = 4000, 128, 16
    corpus, query = make_dense_case(6, n, d)
    corpus[100:2000:37] = corpus[100]  # many identical rows -> exactly equal scores
    query = corpus[100].clone()        # ... which are also the best match
    s, i = _check(engine, corpus, query, k)
    dup = list(range(100, 2000, 37))
    assert i[0, : min(k, len(dup))].tolist() == dup[:k]


def test_multi_query_and_per_query_masks(engine):
    n, d, k, nq = 6000, 1024, 10, 3
    dev = engine.device
    corpus, _ = make_dense_case(8, n, d)
    g = torch.Generator().manual_seed(9)
    queries = torch.randn(nq, d, generator=g).to(torch.float16)
    bits = np.stack([bernoulli_mask(20 + j, n, 0.3) for j in range(nq)])
    mask = torch.stack([_mask_dev(b, dev) for b in bits])
    engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
    s, i = engine.dense_topk(corpus.to(dev), queries.to(dev), k, mask=mask)
    for j in range(nq):
        all_scores = odense.scores_f32(_to_np32(corpus), _to_np32(queries[j]))
        assert_topk_matches(s[j].cpu().numpy(), i[j].cpu().numpy(), all_scores, bits[j], k)


def test_host_entry_point(engine):
    n, d, k = 20000, 1024, 10
    corpus, query = make_dense_case(10, n, d)
    bits = bernoulli_mask(11, n, 0.5)
    words = torch.from_numpy(odense.pack_mask(bits).view(np.int32).copy())
    all_scores = odense.scores_f32(_to_np32(corpus), _to_np32(query))
    c = corpus.to(engine.device)
    engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    s, i = engine.dense_topk_host(c, query, k, mask_host=words)
    assert s.device.type == "cpu" and i.device.type == "cpu"
    assert_topk_matches(s[0].numpy(), i[0].numpy(), all_scores, bits, k)
    s, i = engine.dense_topk_host(c, query, k, mask_dev=words.to(engine.device))
    assert_topk_matches(s[0].numpy(), i[0].numpy(), all_scores, bits, k)
    s, i = engine.dense_topk_host(c, query, k)
    assert_topk_matches(s[0].numpy(), i[0].numpy(), all_scores, np.ones(n, bool), k)


def test_repeated_calls_are_deterministic(engine):
    corpus, query = make_dense_case(12, 40000, 1024)
    c, q = corpus.to(engine.device), query.to(engine.device)
    engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
    ref_s, ref_i = engine.dense_topk(c, q, 50)
    for _ in range(5):
        s, i = engine.dense_topk(c, q, 50)
        assert torch.equal(s, ref_s) and torch.equal(i, ref_i)


def test_argument_errors(engine):
    dev = engine.device
    c = torch.zeros(16, 64, dtype=torch.float16, device=dev)
    q = torch.zeros(64, dtype=torch.float16, device=dev)
    with pytest.raises(ValueError):
        engine.dense_topk(c, q, 0)
    with pytest.raises(ValueError):
        engine.dense_topk(c, q, 4096)
    with pytest.raises(ValueError):
        engine.dense_topk(torch.zeros(16, 60, dtype=torch.float16, device=dev), torch.zeros(60, dtype=torch.float16, device=dev), 4)
    with pytest.raises(ValueError):
        engine.dense_topk(c.float(), q.float(), 4)
    with pytest.raises(ValueError):
        engine.dense_topk(c, q.cpu(), 4)


def test_config2_full_size_1m_x_1024(engine):
    """BASELINE config 2 at full size: 1M x 1024 fp16, top-10, masks p in {1.0, 0.5, 0.1}; full CPU oracle."""
    n, d, k = 1_000_000, 1024, 10
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(1)
    corpus = torch.empty(n, d, dtype=torch.float16, device=dev)
    for lo in range(0, n, 100_000):  # generated on the device in chunks
        blk = torch.randn(100_000, d, generator=g, device=dev)
        corpus[lo: lo + 100_000] = (blk / blk.norm(dim=1, keepdim=True)).half()
    q = torch.randn(d, generator=torch.Generator(device=dev).manual_seed(2), device=dev)
    q = (q / q.norm()).half()
    host = corpus.cpu().numpy()
    qf = q.float().cpu().numpy()
    all_scores = np.empty(n, np.float32)
    for lo in range(0, n, 100_000):
        all_scores[lo: lo + 100_000] = host[lo: lo + 100_000].astype(np.float32) @ qf
    all_scores *= np.float32(1.0 / np.sqrt(np.dot(qf, qf)))
    engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
    for p in (1.0, 0.5, 0.1):
        bits = np.ones(n, bool) if p == 1.0 else bernoulli_mask(3, n, p)
        mask = None if p == 1.0 else _mask_dev(bits, dev)
        s, i = engine.dense_topk(corpus, q, k, mask=mask)
        assert_topk_matches(s[0].cpu().numpy(), i[0].cpu().numpy(), all_scores, bits, k)
