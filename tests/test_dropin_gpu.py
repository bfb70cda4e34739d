"""The drop-in classes behind the reference's call signatures, end to end on the GPU."""
import json
import os
import zlib

import numpy as np
import pytest
import torch

import automative_rag_b200 as rag
from oracle import dense as odense
from oracle import filters as ofilters
from tests._cases import RERANK_CASES, make_rerank_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class FakeEmbeddings:
    """Deterministic text -> 1024-d vector (the bge-m3 forward pass is out of scope and injected)."""

    def __init__(self, dim=1024):
        self.dim = dim

    def _vec(self, text):
        g = torch.Generator().manual_seed(zlib.crc32(text.encode()))
        return torch.randn(self.dim, generator=g).tolist()

    def embed_query(self, text):
        return self._vec(text)

    def embed_documents(self, texts):
        return [self._vec(t) for t in texts]


def _docs(n):
    makers = ["Toyota", "Honda", "BMW"]
    return [rag.Document(page_content=f"chunk {i} about {makers[i % 3]}",
                         metadata={"manufacturer": makers[i % 3], "year": 2020 + i % 4, "category": ["sedan", "suv"][i % 2]})
            for i in range(n)]


@pytest.fixture()
def store():
    client = rag.B200Client(0)
    return rag.B200VectorStore(client, f"col-{np.random.randint(1 << 30)}", FakeEmbeddings())


def test_embedding_function_is_mandatory():
    with pytest.raises(ValueError, match="Embedding function is required"):
        rag.B200VectorStore(rag.B200Client(0), "x", None)  # vectorstore.py:39-40


def _oracle_search(emb, docs, query, k, flt, deleted=None):
    c = np.asarray(emb.embed_documents([d.page_content for d in docs]), dtype=np.float32)
    c = c / np.linalg.norm(c, axis=1, keepdims=True)
    c16 = c.astype(np.float16)
    inv = (1.0 / np.linalg.norm(c16.astype(np.float32), axis=1)).astype(np.float32)
    q16 = np.asarray(emb.embed_query(query), dtype=np.float32).astype(np.float16)
    payloads = [{"page_content": d.page_content, "metadata": d.metadata} for d in docs]
    mask = ofilters.filter_mask(payloads, flt or {}, deleted)
    return odense.topk(c16, q16, k, mask, odense.COSINE, inv)


@pytest.mark.parametrize("flt", [None, {"manufacturer": "Toyota"}, {"manufacturer": ["Honda", "BMW"], "year": 2021},
                                 {"category": "suv", "year": [2020, 2023]}, {"manufacturer": "Nobody"}])
def test_similarity_search_with_score_matches_oracle(store, flt):
    docs = _docs(500)
    ids = store.add_documents(docs)
    assert len(ids) == 500 and len(set(ids)) == 500
    res = store.similarity_search_with_score("What is the horsepower?", k=8, metadata_filter=flt)
    ws, wi = _oracle_search(store.embedding_function, docs, "What is the horsepower?", 8, flt)
    n_valid = int((wi >= 0).sum())
    assert len(res) == n_valid
    assert [d.page_content for d, _ in res] == [docs[i].page_content for i in wi[:n_valid]]
    np.testing.assert_allclose([s for _, s in res], ws[:n_valid], rtol=1e-3, atol=1e-6)
    assert all(isinstance(s, float) for _, s in res)


def test_add_delete_and_stats(store):
    docs = _docs(100)
    ids = store.add_documents(docs)
    assert all("ingestion_time" in d.metadata and d.metadata["id"] for d in docs)  # vectorstore.py:141-152
    assert store.get_stats()["vectors_count"] == 100
    top = store.similarity_search_with_score("query text", k=3)
    best = top[0][0].page_content
    victim = ids[[d.page_content for d in docs].index(best)]
    store.delete_by_ids([victim, "no-such-id"])
    store.delete_by_ids([])
    assert store.get_stats()["vectors_count"] == 99
    again = store.similarity_search_with_score("query text", k=3)
    assert best not in [d.page_content for d, _ in again]
    assert again[0][0].page_content == top[1][0].page_content
    assert store.add_documents([]) == []
    found = store.search_by_metadata({"manufacturer": "BMW", "year": 2021}, limit=5)
    assert 0 < len(found) <= 5 and all(d.metadata["manufacturer"] == "BMW" and d.metadata["year"] == 2021 for d in found)


def test_filter_failure_falls_back_to_unfiltered(store, monkeypatch):
    store.add_documents(_docs(50))
    col = store.collection

    def boom(flt):
        if flt is not None:
            raise RuntimeError("filter exploded")
        return None

    monkeypatch.setattr(col, "device_mask", boom)
    res = store.similarity_search_with_score("q", k=5, metadata_filter={"manufacturer": "Toyota"})
    assert len(res) == 5  # vectorstore.py:199-207: log, retry unfiltered


# ------------------------------------------------------------------------------------ reranker
def _reranker(case, use_bge, nq):
    docs_by_text = {f"doc-{i}": t for i, t in enumerate(case["docs"])}

    class Bge:
        def predict(self, pairs):
            return np.asarray([case["bge"][int(p[1].split("-")[1])] for p in pairs], dtype=np.float32)

    return rag.B200ColBERTReranker(
        device="cuda:0", use_fp16=False, use_bge_reranker=use_bge, cross_encoder=Bge() if use_bge else None,
        query_encoder=lambda text: case["queries"][int(text.split("-")[1])],
        doc_encoder=lambda texts: [docs_by_text[t] for t in texts])


@pytest.mark.parametrize("name", sorted(RERANK_CASES))
def test_reranker_dropin_matches_reference_golden(name):
    """B200ColBERTReranker.rerank / batch_rerank_queries vs the reference's own outputs on the same
    (stubbed-encoder) inputs: same documents, same order, same scores."""
    spec = RERANK_CASES[name]
    case = make_rerank_case(spec)
    gold = json.load(open(os.path.join(GOLD, "rerank_golden.json")))[name]
    rr = _reranker(case, spec["use_bge"], spec["n_queries"])
    docs = [rag.Document(page_content=f"doc-{i}", metadata={"i": i}) for i in range(spec["n_docs"])]
    res = rr.rerank("q-0", docs, spec["top_k"])
    assert [d.metadata["i"] for d, _ in res] == [i for i, _ in gold["rerank"]]
    np.testing.assert_allclose([s for _, s in res], [s for _, s in gold["rerank"]], rtol=2e-5, atol=1e-4)
    if spec.get("batch"):
        b = rr.batch_rerank_queries([f"q-{i}" for i in range(spec["n_queries"])], docs, spec["top_k"])
        assert list(b) == list(gold["batch"])
        for key, want in gold["batch"].items():
            assert [d.metadata["i"] for d, _ in b[key]] == [i for i, _ in want]
            np.testing.assert_allclose([s for _, s in b[key]], [s for _, s in want], rtol=2e-5, atol=1e-4)


def test_reranker_edge_behaviour():
    case = make_rerank_case(RERANK_CASES["colbert_only"])
    rr = _reranker(case, False, 1)
    assert rr.rerank("q-0", []) == []                      # rerankers.py:281-282
    assert rr._compute_maxsim_scores(case["queries"][0], []) == []
    assert rr.batch_rerank_queries([], []) == {}           # :577-578
    s = rr._compute_maxsim_scores(case["queries"][0], case["docs"])
    assert isinstance(s, list) and len(s) == len(case["docs"]) and all(isinstance(v, float) for v in s)
    s2 = rr._compute_maxsim_scores(case["queries"][0].squeeze(0), case["docs"])  # [Lq, D] accepted (:234-238)
    assert s == s2


def test_hybrid_retriever_composition(store):
    docs = _docs(200)
    store.add_documents(docs)
    g = torch.Generator().manual_seed(0)
    table = {d.page_content: torch.randn(40, 64, generator=g) for d in docs}
    rr = rag.B200ColBERTReranker(device="cuda:0", use_fp16=True, use_bge_reranker=False,
                                 query_encoder=lambda t: torch.randn(1, 32, 64, generator=torch.Generator().manual_seed(1)),
                                 doc_encoder=lambda texts: [table[t] for t in texts])
    hr = rag.HybridRetriever(store, rr, top_k=20, rerank_top_k=5)
    out, seconds = hr.retrieve("q", metadata_filter={"manufacturer": "Honda"})
    assert seconds > 0
    assert len(out) == 5 and all(d.metadata["manufacturer"] == "Honda" for d, _ in out)
    assert [s for _, s in out] == sorted([s for _, s in out], reverse=True)


# ------------------------------------------------------------------------------------------ stream ordering
def _big_collection(n, dim=256, seed=11):
    """A collection filled through Collection.upsert with device-generated vectors (the embedding model is out of
    scope) and a payload whose fields cycle with the row index, so filters have a closed-form answer."""
    client = rag.B200Client(0)
    col = client.create_collection(f"big-{np.random.randint(1 << 30)}", size=dim)
    g = torch.Generator(device=col.device).manual_seed(seed)
    vec = torch.randn(n, dim, generator=g, device=col.device)
    makers = ["Toyota", "Honda", "BMW", "Ford", "Kia"]
    payloads = [{"page_content": f"c{i}", "metadata": {"manufacturer": makers[i % 5], "year": 2000 + i % 7,
                                                         "category": ["sedan", "suv", "truck"][i % 3]}}
                for i in range(n)]
    return client, col, vec, payloads


def test_search_is_ordered_after_upsert_and_mask_build_on_the_callers_stream():
    """the host entry point used a private non-blocking stream, so a search could start before
    the upsert / rs_filter_mask enqueued just before it on torch's stream had finished.  400k rows are appended and
    IMMEDIATELY searched with a 16-clause filter (a long mask kernel), with no synchronise in between; repeated so a
    race would show.  The expected answer is computed with torch on the same device after a full synchronise."""
    from automative_rag_b200.filters import FieldCondition, Filter, MatchValue, Range

    n, dim, k = 400_000, 256, 16
    client, col, vec, payloads = _big_collection(n, dim)
    conds = [FieldCondition(key="metadata.manufacturer", match=MatchValue(value="Honda")),
             FieldCondition(key="metadata.year", range=Range(gte=2003, lte=2003))]
    flt = Filter(must=(conds * 8))  # 16 clauses: the mask kernel reads 16 columns x 400k rows
    q = torch.randn(dim, generator=torch.Generator().manual_seed(5))
    for rep in range(3):
        ids = [f"p{rep}-{i}" for i in range(n)]
        if rep:
            col.delete([f"p{rep - 1}-{i}" for i in range(n)])
        col.upsert(ids, vec + rep, payloads)                      # device writes on torch's current stream
        mask = col.device_mask(flt)                                # rs_filter_mask on torch's current stream
        scores, rows = col.engine.dense_topk_host(
            col.vectors[: col.n], q.to(col.dtype).contiguous(), k, mask_dev=mask, inv_norm=col.inv_norm[: col.n],
            metric=rag._ffi.RS_METRIC_COSINE)
        torch.cuda.synchronize()
        base = rep * n
        i = torch.arange(n, device=col.device)
        passing = ((i % 5) == 1) & ((i % 7) == 3)
        stored = col.vectors[base: base + n].float()
        ref = (stored @ q.to(col.device).half().float()) * col.inv_norm[base: base + n] / q.half().float().norm().to(col.device)
        ref = torch.where(passing, ref, torch.full_like(ref, float("-inf")))
        ws, wi = torch.topk(ref, k)
        assert rows[0].tolist() == (wi + base).cpu().tolist(), f"rep {rep}: stale mask / corpus read by the scan"
        np.testing.assert_allclose(scores[0].numpy(), ws.cpu().numpy(), rtol=1e-3, atol=1e-6)


def test_calls_on_different_streams_do_not_overlap_on_the_shared_workspace(engine):
    """scans on two streams share one workspace; the handle orders a call after the previous one
    when the stream changes.  Alternating streams with no host synchronise must give the single-stream answers."""
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(3)
    c = torch.randn(300_000, 256, generator=g, device=dev)
    c = (c / c.norm(dim=1, keepdim=True)).half()
    qs = torch.randn(12, 256, generator=g, device=dev).half()
    want = [engine.dense_topk(c, qs[j], 50) for j in range(12)]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    got = []
    for j in range(12):
        with torch.cuda.stream(s1 if j % 2 == 0 else s2):
            got.append(engine.dense_topk(c, qs[j], 50))
    torch.cuda.synchronize()
    for (ws, wi), (gs, gi) in zip(want, got):
        assert torch.equal(wi, gi) and torch.equal(ws, gs)
