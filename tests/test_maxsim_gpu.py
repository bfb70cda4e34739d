"""MaxSim parity: the three CUDA kernel families (through the C ABI) vs the reference's golden outputs
and the CPU oracle on identical rounded inputs."""
import os

import numpy as np
import pytest
import torch

from automative_rag_b200 import _ffi
from automative_rag_b200.rerankers import pack_documents
from oracle import maxsim as omaxsim
from tests._cases import MAXSIM_CASES, make_maxsim_case
from tests._parity import RTOL_16BIT, RTOL_FP32, assert_scores_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(MAXSIM_CASES))
def test_fp32_path_matches_the_reference_golden(engine, name):
    """Exact-fp32 kernel vs ColBERTReranker._compute_maxsim_scores outputs stored by make_golden.py."""
    q, docs = make_maxsim_case(MAXSIM_CASES[name])
    toks, off = pack_documents(docs, engine.device, torch.float32)
    got = engine.maxsim(q.to(engine.device), toks, off)
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_SIMT
    want = np.load(os.path.join(GOLD, "maxsim_golden.npz"))[name]
    assert_scores_close(got[0].cpu().numpy(), want, rtol=RTOL_FP32, atol=1e-3, what=name)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name", sorted(MAXSIM_CASES))
def test_mma_path_matches_oracle_on_rounded_inputs(engine, name, dtype):
    q, docs = make_maxsim_case(MAXSIM_CASES[name])
    scale = 0.125 if dtype == torch.float16 else 1.0  # keep fp16 products well inside range
    q16 = (q * scale).to(dtype)
    d16 = [(d * scale).to(dtype) for d in docs]
    toks, off = pack_documents(d16, engine.device, dtype)
    engine.set_maxsim_impl(_ffi.RS_MAXSIM_MMA)
    try:
        got, arg = engine.maxsim(q16.to(engine.device), toks, off, want_argmax=True)
    finally:
        engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    want, want_arg = omaxsim.maxsim_scores(q16, d16, return_argmax=True)
    assert_scores_close(got[0].cpu().numpy(), want, rtol=RTOL_16BIT, atol=1e-4, what=name)
    # argmax: identical except where two doc tokens tie within tolerance
    ga = arg[0].cpu().numpy()
    qf = q16[0].float()
    for j, d in enumerate(d16):
        sim = (qf @ d.float().T).numpy()
        rows = np.arange(sim.shape[0])
        assert ((ga[j] >= 0) & (ga[j] < d.shape[0])).all()
        np.testing.assert_allclose(sim[rows, ga[j]], sim[rows, want_arg[j]], rtol=RTOL_16BIT, atol=1e-4)


def _batch_case(seed, nq, lq, d, lens, dtype):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(nq, lq, d, generator=g).to(dtype)
    docs = [torch.randn(n, d, generator=g).to(dtype) for n in lens]
    return q, docs


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("nq,lq,d,lens", [
    (4, 32, 128, [300] * 10),                                  # one 128-row tile, docs straddle 256-token tiles
    (8, 32, 128, [256] * 6),                                   # boundaries exactly on tile edges
    (9, 32, 128, [1, 2, 31, 32, 33, 255, 256, 257, 511, 513, 700]),  # ragged incl. 1-token docs
    (64, 32, 64, [180] * 40),                                  # d = 64, many query tiles -> several CTAs per range
    (5, 20, 128, [64, 100, 300, 17]),                          # lq < 32 -> zero-padded query rows (TMA OOB fill)
    (6, 45, 128, [90] * 12),                                   # lq in (32, 64] -> two warps per query, fixed-order sum
    (3, 100, 64, [50, 60, 70]),                                # lq in (64, 128]
    (16, 32, 128, [300] * 300),                                # more docs than SMs / groups
])
def test_tcgen05_path_matches_oracle(engine, nq, lq, d, lens, dtype):
    q, docs = _batch_case(nq * 100 + lq, nq, lq, d, lens, dtype)
    toks, off = pack_documents(docs, engine.device, dtype)
    g = torch.Generator().manual_seed(5)
    w = torch.rand(nq, lq, generator=g)
    w[:, 0] = 0
    for weight in (None, w):
        engine.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05)
        try:
            got = engine.maxsim(q.to(engine.device), toks, off,
                                q_weight=None if weight is None else weight.to(engine.device))
        finally:
            engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
        assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05
        want = omaxsim.maxsim_scores_packed(q, None if weight is None else weight, torch.cat(docs),
                                            off.cpu().numpy())
        assert_scores_close(got.cpu().numpy(), want, rtol=RTOL_16BIT, atol=1e-3, what="tcgen05 maxsim")


def test_auto_dispatch_picks_tcgen05_for_shared_candidates(engine):
    q, docs = _batch_case(1, 8, 32, 128, [100] * 20, torch.bfloat16)
    toks, off = pack_documents(docs, engine.device, torch.bfloat16)
    engine.maxsim(q.to(engine.device), toks, off)
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05
    engine.maxsim(q[:1].to(engine.device), toks, off)  # a single query: the document-streaming tcgen05 kernel
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05_CAND
    engine.maxsim(q[:1].to(engine.device), toks, off, want_argmax=True)  # argmax: epilogue of the same tcgen05 kernel
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05_CAND
    q7, docs7 = _batch_case(2, 2, 32, 768, [256] * 5, torch.float16)  # the deployed hidden size (rerankers.py:118-120)
    toks7, off7 = pack_documents(docs7, engine.device, torch.float16)
    engine.maxsim(q7.to(engine.device), toks7, off7)
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05_CAND
    q9, docs9 = _batch_case(3, 2, 32, 80, [40] * 3, torch.float16)   # d % 64 != 0: the general mma.sync kernel
    toks9, off9 = pack_documents(docs9, engine.device, torch.float16)
    engine.maxsim(q9.to(engine.device), toks9, off9)
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_MMA


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name", sorted(MAXSIM_CASES))
def test_tcgen05_candidate_kernel_argmax_and_token_maxima(engine, name, dtype):
    """the explanations path (arg-max document token and the maximum itself per query token,
    rerankers.py:489-501) and the deployed d = 768 run on the tcgen05 candidate kernel, checked against the oracle
    on the reference's golden cases (incl. `deployed768`: 32 x 768 query vs 256-token documents)."""
    q, docs = make_maxsim_case(MAXSIM_CASES[name])
    if q.shape[-1] % 64:
        pytest.skip("d not a multiple of 64: mma.sync path")
    scale = 0.125 if dtype == torch.float16 else 1.0
    q16 = (q * scale).to(dtype)
    d16 = [(d * scale).to(dtype) for d in docs]
    toks, off = pack_documents(d16, engine.device, dtype)
    got, arg, tmax = engine.maxsim(q16.to(engine.device), toks, off, want_argmax=True, want_tokmax=True)
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05_CAND
    want, want_arg = omaxsim.maxsim_scores(q16, d16, return_argmax=True)
    assert_scores_close(got[0].cpu().numpy(), want, rtol=RTOL_16BIT, atol=1e-4, what=name)
    ga, gm = arg[0].cpu().numpy(), tmax[0].cpu().numpy()
    qf = q16[0].float() if q16.dim() == 3 else q16.float()
    for j, d in enumerate(d16):
        sim = (qf @ d.float().T).numpy()
        rows = np.arange(sim.shape[0])
        assert ((ga[j] >= 0) & (ga[j] < d.shape[0])).all()
        # the arg-max token scores the row maximum (ties within tolerance may pick another token) ...
        np.testing.assert_allclose(sim[rows, ga[j]], sim[rows, want_arg[j]], rtol=RTOL_16BIT, atol=1e-4)
        # ... and out_tokmax is that maximum
        np.testing.assert_allclose(gm[j], sim.max(axis=1), rtol=RTOL_16BIT, atol=1e-4)
    # scores without the extra outputs are bit-identical (same kernel, same order of operations)
    plain = engine.maxsim(q16.to(engine.device), toks, off)
    if engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05_CAND:
        assert torch.equal(plain, got)


@pytest.mark.parametrize("nq,lq,d,lens,nc", [
    (3, 32, 768, [256] * 12, 5),       # deployed shape with candidate lists
    (2, 32, 384, [1, 33, 129, 300], 0),  # six K blocks, one per ring stage
    (5, 64, 256, [90, 200, 31], 2),     # two warps per query, four K blocks
    (2, 32, 704, [140, 260], 0),        # eleven K blocks, one per ring stage (lq_pad * d <= 24576 is the limit)
])
def test_candidate_tcgen05_k_pipeline_matches_oracle(engine, nq, lq, d, lens, nc):
    """Rows wider than one shared-memory stage: the chunk is streamed K block by K block (d / 64 blocks)."""
    dtype = torch.bfloat16
    q, docs = _batch_case(nq * 7 + d, nq, lq, d, lens, dtype)
    q, docs = q * 0.25, [x * 0.25 for x in docs]
    toks, off = pack_documents(docs, engine.device, dtype)
    rng = np.random.default_rng(d)
    cand = np.stack([rng.permutation(len(docs))[:nc] for _ in range(nq)]).astype(np.int32) if nc else None
    engine.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05_CAND)
    try:
        got = engine.maxsim(q.to(engine.device), toks, off, cand=None if cand is None else torch.from_numpy(cand).to(engine.device))
    finally:
        engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    want = omaxsim.maxsim_scores_packed(q, None, torch.cat(docs), off.cpu().numpy(), cand)
    assert_scores_close(got.cpu().numpy(), want, rtol=RTOL_16BIT, atol=1e-3, what=f"d = {d}")


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_candidate_lists_mma_path(engine, dtype):
    nq, lq, d = 6, 32, 128
    q, docs = _batch_case(77, nq, lq, d, [50, 120, 300, 7, 64, 200, 33, 90], dtype)
    toks, off = pack_documents(docs, engine.device, dtype)
    rng = np.random.default_rng(0)
    cand = np.stack([rng.permutation(len(docs))[:5] for _ in range(nq)]).astype(np.int32)
    engine.set_maxsim_impl(_ffi.RS_MAXSIM_MMA)
    try:
        got = engine.maxsim(q.to(engine.device), toks, off, cand=torch.from_numpy(cand).to(engine.device))
    finally:
        engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_MMA
    want = omaxsim.maxsim_scores_packed(q, None, torch.cat(docs), off.cpu().numpy(), cand)
    assert_scores_close(got.cpu().numpy(), want, rtol=RTOL_16BIT, atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("nq,lq,d,lens,nc", [
    (6, 32, 128, [50, 120, 300, 7, 64, 200, 33, 90], 5),          # ragged documents, 1..3 chunks each
    (1, 32, 128, [180] * 100, 0),                                  # BASELINE config 1 shape: one query, shared docs
    (3, 32, 128, [1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 511, 513, 700], 9),  # chunk-edge lengths
    (40, 32, 64, [180] * 30, 30),                                  # d = 64; more pairs than SMs, many query switches
    (5, 20, 128, [64, 100, 300, 17], 3),                           # lq < 32: zero-padded query rows
    (4, 45, 128, [90] * 12, 7),                                    # lq in (32, 64]: two epilogue warps
    (3, 70, 64, [50, 60, 70, 300], 4),                             # lq in (64, 96]: three
    (2, 128, 128, [140, 260, 20], 3),                              # lq = 128: four
    (300, 32, 128, [40] * 8, 2),                                   # a query switch every other pair
])
def test_candidate_tcgen05_path_matches_oracle(engine, nq, lq, d, lens, nc, dtype):
    """Per-query candidate lists (and shared lists with few queries, nc == 0) through the document-streaming
    tcgen05 kernel vs the CPU oracle on the same rounded inputs; default and explicit token weights."""
    q, docs = _batch_case(nq * 10 + lq, nq, lq, d, lens, dtype)
    toks, off = pack_documents(docs, engine.device, dtype)
    rng = np.random.default_rng(nq)
    cand = None
    if nc:
        cand = np.stack([rng.permutation(len(docs))[:nc] if nc <= len(docs) else rng.integers(0, len(docs), nc)
                         for _ in range(nq)]).astype(np.int32)
    g = torch.Generator().manual_seed(5)
    w = torch.rand(nq, lq, generator=g)
    w[:, 0] = 0
    for weight in (None, w):
        engine.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05_CAND)
        try:
            got = engine.maxsim(q.to(engine.device), toks, off,
                                q_weight=None if weight is None else weight.to(engine.device),
                                cand=None if cand is None else torch.from_numpy(cand).to(engine.device))
        finally:
            engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
        assert engine.last_maxsim_impl == _ffi.RS_MAXSIM_TCGEN05_CAND
        want = omaxsim.maxsim_scores_packed(q, None if weight is None else weight, torch.cat(docs),
                                            off.cpu().numpy(), cand)
        assert_scores_close(got.cpu().numpy(), want, rtol=RTOL_16BIT, atol=1e-3, what="tcgen05 candidate maxsim")


def test_candidate_tcgen05_empty_documents_and_bad_indices(engine):
    """C-ABI edge: a zero-token document (the Python layer refuses those, as torch.max does in the reference) and a
    candidate index outside the collection both score -inf; their neighbours are unaffected."""
    dtype, d, lq = torch.bfloat16, 128, 32
    q, docs = _batch_case(3, 2, lq, d, [300, 45, 128], dtype)
    dev = engine.device
    toks = torch.cat(docs).to(dev)
    off = torch.tensor([0, 300, 300, 345, 345, 473], dtype=torch.int32, device=dev)   # docs 1 and 3 are empty
    cand = torch.tensor([[0, 1, 2, 3, 4, 7], [4, -1, 3, 2, 0, 1]], dtype=torch.int32, device=dev)
    engine.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05_CAND)
    try:
        got = engine.maxsim(q.to(dev), toks, off, cand=cand).cpu().numpy()
    finally:
        engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    real = {0: 0, 2: 1, 4: 2}
    want = omaxsim.maxsim_scores_packed(q, None, torch.cat(docs), np.array([0, 300, 345, 473]))
    for qi in range(2):
        for c, j in enumerate(cand[qi].tolist()):
            if j in real:
                assert abs(got[qi, c] - want[qi, real[j]]) <= RTOL_16BIT * abs(want[qi, real[j]]) + 1e-3
            else:
                assert got[qi, c] == -np.inf


def test_candidate_tcgen05_equals_mma_at_config4b_scale(engine):
    """BASELINE config 4b shape scaled to 64 queries x 1000 own candidates out of a 4096-document pool (300 tokens,
    bf16): the tcgen05 candidate kernel and the mma.sync kernel agree on every (query, candidate) pair."""
    nq, lq, d, pool, ld, nc = 64, 32, 128, 4096, 300, 1000
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(8)
    q = torch.randn(nq, lq, d, generator=g, device=dev).to(torch.bfloat16)
    toks = torch.randn(pool * ld, d, generator=g, device=dev).to(torch.bfloat16)
    off = (torch.arange(pool + 1, device=dev) * ld).to(torch.int32)
    cand = torch.stack([torch.randperm(pool, generator=g, device=dev)[:nc] for _ in range(nq)]).to(torch.int32)
    out = {}
    for impl in (_ffi.RS_MAXSIM_TCGEN05_CAND, _ffi.RS_MAXSIM_MMA):
        engine.set_maxsim_impl(impl)
        try:
            out[impl] = engine.maxsim(q, toks, off, cand=cand).cpu().numpy()
        finally:
            engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    assert_scores_close(out[_ffi.RS_MAXSIM_TCGEN05_CAND], out[_ffi.RS_MAXSIM_MMA], rtol=RTOL_16BIT, atol=1e-3)


def test_weights_reproduce_all_three_conventions(engine):
    """w = [0,1..1,0] == reference HEAD (a4); all-ones == canonical ColBERT / stale test's 32.0; mask-based == a7."""
    q = torch.ones(1, 32, 64)
    d = torch.full((5, 64), 1.0 / 64)  # every similarity is exactly 1.0
    toks, off = pack_documents([d], engine.device, torch.float32)
    dev = engine.device
    assert engine.maxsim(q.to(dev), toks, off)[0, 0].item() == 30.0
    assert engine.maxsim(q.to(dev), toks, off, q_weight=torch.ones(1, 32, device=dev))[0, 0].item() == 32.0
    m = torch.zeros(1, 32, device=dev)
    m[0, 1:9] = 1
    assert engine.maxsim(q.to(dev), toks, off, q_weight=m)[0, 0].item() == 8.0


def test_empty_inputs(engine):
    dev = engine.device
    q = torch.zeros(1, 4, 64, dtype=torch.float16, device=dev)
    out = engine.maxsim(q, torch.zeros(1, 64, dtype=torch.float16, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
    assert out.shape == (1, 0)


def test_config4_full_size_properties(engine):
    """BASELINE config 4a at full size (256x32 queries vs 1000x300-token shared candidates, bf16):
    tcgen05 == mma.sync on every (query, doc) pair, and == the CPU oracle on a sample of queries."""
    nq, lq, d, nd, ld = 256, 32, 128, 1000, 300
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(6)
    q = torch.randn(nq, lq, d, generator=g, device=dev).to(torch.bfloat16)
    toks = torch.randn(nd * ld, d, generator=torch.Generator(device=dev).manual_seed(7), device=dev).to(torch.bfloat16)
    off = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
    engine.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05)
    a = engine.maxsim(q, toks, off)
    engine.set_maxsim_impl(_ffi.RS_MAXSIM_MMA)
    b = engine.maxsim(q, toks, off)
    engine.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    assert_scores_close(a.cpu().numpy(), b.cpu().numpy(), rtol=RTOL_16BIT, atol=1e-3, what="tcgen05 vs mma.sync")
    sample = [0, 1, 127, 128, 255]
    want = omaxsim.maxsim_scores_packed(q[sample].cpu(), None, toks.cpu(), off.cpu().numpy())
    assert_scores_close(a[sample].cpu().numpy(), want, rtol=RTOL_16BIT, atol=1e-3, what="tcgen05 vs oracle")
