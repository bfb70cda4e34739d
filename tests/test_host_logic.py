"""Host-side logic that needs no GPU: packing, shard bounds, the all-gather wire format."""
import numpy as np
import pytest
import torch

from automative_rag_b200 import distributed as D
from automative_rag_b200.rerankers import pack_documents


def test_shard_bounds_partition_rows():
    for n in (0, 1, 7, 8, 1000, 12_500_001):
        for w in (1, 2, 4, 8):
            spans = [D.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("nq,k", [(1, 1), (1, 10), (3, 7), (4, 1000)])
def test_wire_views_roundtrip(nq, k):
    words = D.wire_words(nq, k)
    assert words % 2 == 0 and words >= 3 * nq * k
    world = 3
    bufs = []
    for r in range(world):
        buf = torch.zeros(words, dtype=torch.int32)
        s, i = D.wire_views(buf, nq, k)
        s.copy_(torch.arange(nq * k, dtype=torch.float32).view(nq, k) + 0.5 + r)
        i.copy_(torch.arange(nq * k, dtype=torch.int64).view(nq, k) + (r << 40))
        bufs.append(buf)
    gs, gi = D.gathered_views(torch.cat(bufs), world, nq, k)
    assert gs.shape == gi.shape == (world, nq, k)
    for r in range(world):
        assert gs[r, nq - 1, k - 1].item() == nq * k - 1 + 0.5 + r
        assert gi[r, 0, 0].item() == (r << 40)
    # every [nq, k] list is dense, lists are strided: the layout rs_topk_merge consumes without a copy
    assert gs.stride(2) == 1 and gs.stride(1) == k and gi.stride(2) == 1 and gi.stride(1) == k
    assert gs.stride(0) == words and gi.stride(0) == words // 2


def test_pack_documents():
    docs = [torch.randn(3, 8), torch.randn(1, 5, 8), torch.randn(2, 8)]
    toks, off = pack_documents(docs, torch.device("cpu"), torch.float16)
    assert toks.shape == (10, 8) and toks.dtype == torch.float16
    assert off.tolist() == [0, 3, 8, 10] and off.dtype == torch.int32
    with pytest.raises(ValueError):
        pack_documents([torch.randn(0, 8)], torch.device("cpu"), torch.float16)


def test_hybrid_retriever_matches_the_reference_tests():
    """The reference's own expectations for HybridRetriever (tests/test_retrieval.py:191-276), with stand-ins for the
    store and the reranker exactly as the reference mocks them: argument forwarding, (results, execution_time), the
    rerank_top_k cut without reranking — plus the per-mode sizes of mode_config.py."""
    from unittest.mock import MagicMock

    from automative_rag_b200.documents import Document
    from automative_rag_b200.retriever import HybridRetriever, retrieval_params

    store, reranker = MagicMock(), MagicMock()
    r = HybridRetriever(vector_store=store, reranker=reranker, top_k=20, rerank_top_k=5)
    assert r.vector_store is store and r.reranker is reranker and r.top_k == 20 and r.rerank_top_k == 5

    r = HybridRetriever(vector_store=store, reranker=reranker, top_k=10, rerank_top_k=3)
    documents = [(Document(page_content=f"Doc {i}", metadata={"id": str(i)}), 1.0 - 0.1 * i) for i in (1, 2, 3, 4)]
    reranked = [documents[2], documents[0], documents[1]]
    store.similarity_search_with_score.return_value = documents
    reranker.rerank.return_value = reranked
    results, seconds = r.retrieve(query="What is the horsepower?", metadata_filter={"manufacturer": "Toyota"}, rerank=True)
    assert results == reranked and seconds > 0
    store.similarity_search_with_score.assert_called_once_with(query="What is the horsepower?", k=10,
                                                               metadata_filter={"manufacturer": "Toyota"})
    reranker.rerank.assert_called_once_with(query="What is the horsepower?", documents=[d for d, _ in documents], top_k=3)

    store.reset_mock(), reranker.reset_mock()
    store.similarity_search_with_score.return_value = documents
    results, seconds = r.retrieve(query="What is the horsepower?", metadata_filter={"manufacturer": "Toyota"}, rerank=False)
    assert results == documents[:3] and seconds > 0
    store.similarity_search_with_score.assert_called_once()
    reranker.rerank.assert_not_called()

    # per-mode sizes (mode_config.py): retrieval_k for the search, final_k after reranking; unknown -> facts
    assert retrieval_params("debate") == (40, 18) and retrieval_params("FACTS") == (20, 8) and retrieval_params("nope") == (20, 8)
    store.reset_mock(), reranker.reset_mock()
    store.similarity_search_with_score.return_value = documents
    r.retrieve(query="q", mode="tradeoffs")
    assert store.similarity_search_with_score.call_args.kwargs["k"] == 35
    assert reranker.rerank.call_args.kwargs["top_k"] == 15

    store.similarity_search_with_score.return_value = []
    results, seconds = r.retrieve(query="q")
    assert results == [] and seconds > 0


def test_collection_storage_logic_on_cpu_tensors():
    """The store's host-side bookkeeping (row storage, dictionary-encoded payload columns, tombstones, upsert-replaces,
    rebuild of the columns) does not depend on the GPU: run it over CPU tensors with a stand-in for the engine (whose
    only job here would be rs_filter_mask / rs_dense_topk, not exercised)."""
    from types import SimpleNamespace

    from automative_rag_b200.filters import INT_MISSING
    from automative_rag_b200.vectorstore import INTEGER_FIELDS, KEYWORD_FIELDS, Collection

    col = Collection("c", 16, SimpleNamespace(device=torch.device("cpu")), capacity=4)
    g = torch.Generator().manual_seed(0)
    vec = torch.randn(6, 16, generator=g)
    pay = [{"page_content": f"t{i}", "metadata": {"manufacturer": ["Toyota", "Honda"][i % 2], "year": 2020 + i,
                                                  "model": None, "ingestion_time": True}} for i in range(6)]
    col.upsert([f"id{i}" for i in range(6)], vec, pay)
    assert col.n == 6 and col.capacity >= 6 and col.capacity % 32 == 0
    stored = col.vectors[:6].float()
    np.testing.assert_allclose(stored.norm(dim=1).numpy(), 1.0, atol=2e-3)                 # unit rows (Distance.COSINE)
    np.testing.assert_allclose((col.inv_norm[:6] * stored.norm(dim=1)).numpy(), 1.0, atol=1e-6)
    assert col.columns["manufacturer"][:6].tolist() == [0, 1, 0, 1, 0, 1]                  # dictionary codes
    assert col.columns["year"][:6].tolist() == [2020, 2021, 2022, 2023, 2024, 2025]
    assert col.columns["model"][:6].tolist() == [-1] * 6                                   # absent keyword
    assert col.columns["ingestion_time"][:6].tolist() == [INT_MISSING] * 6                 # a bool is not an integer
    assert set(col.columns) == set(KEYWORD_FIELDS + INTEGER_FIELDS)

    assert col.delete(["id1", "nope", "id4"]) == 2 and col.deleted == 2
    assert col.tombstone[0].item() == (1 << 1) | (1 << 4)
    assert "id1" not in col.id_to_row and col.id_to_row["id5"] == 5

    col.upsert(["id0"], torch.randn(1, 16, generator=g), [{"page_content": "new", "metadata": {"manufacturer": "Kia"}}])
    assert col.n == 7 and col.id_to_row["id0"] == 6 and col.deleted == 3                   # upsert replaces the point
    assert col.columns["manufacturer"][6].item() == 2 and col.keyword_dicts["manufacturer"]["Kia"] == 2

    before = {f: col.columns[f][: col.n].clone() for f in col.columns}
    col.columns["year"][: col.n] = 0                                                       # damage an "index"
    assert sorted(col.rebuild_columns()) == sorted(KEYWORD_FIELDS + INTEGER_FIELDS)
    for f in col.columns:
        assert torch.equal(col.columns[f][: col.n], before[f]), f                          # ... and it is restored
        assert (col.columns[f][col.n:] == INT_MISSING).all()


def test_explanations_logic_matches_the_reference_golden():
    """B200ColBERTReranker._explain_colbert_matches (token strings, contexts, per-token similarities, mask-based score
    rule, ordering) against the output of the reference's own function (tests/golden/explain_golden.json, made by
    make_golden.py).  The device call is replaced by a stand-in backed by the CPU oracle: what is under test is the
    host logic around rs_maxsim(out_argmax, q_weight); the kernel's argmax / weights are covered by test_maxsim_gpu."""
    import json
    import os
    from types import SimpleNamespace

    from automative_rag_b200.documents import Document
    from automative_rag_b200.rerankers import B200ColBERTReranker
    from oracle import maxsim as omaxsim
    from tests._cases import EXPLAIN_CASE, make_explain_case

    tok, q_emb, doc_embs = make_explain_case()

    def fake_maxsim(q, tokens, offs, q_weight=None, cand=None, want_argmax=False, want_tokmax=False):
        off = offs.tolist()
        docs = [tokens[off[i]: off[i + 1]] for i in range(len(off) - 1)]
        w = None if q_weight is None else q_weight[0]
        sc, arg = omaxsim.maxsim_scores(q[0], docs, weights=None if w is None else w.numpy(), return_argmax=True)
        out = torch.from_numpy(np.asarray(sc, dtype=np.float32)[None].copy())
        tmax = torch.stack([(q[0].float() @ d.float().T).max(dim=1).values for d in docs])[None]  # rs_maxsim's out_tokmax
        res = (out,) + ((torch.tensor(np.stack(arg)[None], dtype=torch.int32),) if want_argmax else ()) \
            + ((tmax,) if want_tokmax else ())
        return res if len(res) > 1 else out

    rr = object.__new__(B200ColBERTReranker)          # the constructor needs a B200; only the host logic runs here
    rr.engine = SimpleNamespace(device=torch.device("cpu"), maxsim=fake_maxsim)
    rr.tokenizer, rr.compute_dtype, rr.batch_size = tok, torch.float32, 2
    rr.max_query_length, rr.max_doc_length = EXPLAIN_CASE["max_query_length"], EXPLAIN_CASE["max_doc_length"]
    rr.query_encoder = lambda query: q_emb
    rr.doc_encoder = lambda texts: [doc_embs[t] for t in texts]
    docs = [Document(page_content=t, metadata={"i": i}) for i, t in enumerate(EXPLAIN_CASE["docs"])]
    got = rr._explain_colbert_matches(EXPLAIN_CASE["query"], docs, EXPLAIN_CASE["num_explanations"])
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "explain_golden.json")))
    assert [r["document"].metadata["i"] for r in got] == [w["doc"] for w in want]
    for r, w in zip(got, want):
        assert abs(r["score"] - w["score"]) <= 1e-4 * abs(w["score"])
        assert [(e["query_token"], e["doc_token"], e["context"]) for e in r["explanations"]] == \
               [(e["query_token"], e["doc_token"], e["context"]) for e in w["explanations"]]
        np.testing.assert_allclose([e["similarity"] for e in r["explanations"]],
                                   [e["similarity"] for e in w["explanations"]], rtol=1e-5)


def _oracle_backed_reranker(case, use_bge):
    """B200ColBERTReranker with the two device calls replaced by the CPU oracle (the constructor needs a B200)."""
    from types import SimpleNamespace

    from automative_rag_b200.rerankers import B200ColBERTReranker
    from oracle import maxsim as omaxsim

    def fake_maxsim(q, tokens, offs, q_weight=None, cand=None, want_argmax=False, want_tokmax=False):
        out = omaxsim.maxsim_scores_packed(q, q_weight, tokens, offs.numpy(), cand)
        return torch.from_numpy(out)

    def fake_postprocess(scores, other, top_k, w_a=0.8, w_b=0.2):
        idx, out = [], []
        for r in range(scores.shape[0]):
            ranked = omaxsim.hybrid_rerank(scores[r].tolist(), None if other is None else other[r].tolist(), w_a, w_b, top_k)
            idx.append([i for i, _ in ranked])
            out.append([v for _, v in ranked])
        return torch.tensor(idx, dtype=torch.int32), torch.tensor(out, dtype=torch.float32)

    class Bge:
        def predict(self, pairs):
            return np.asarray([case["bge"][int(p[1].split("-")[1])] for p in pairs], dtype=np.float32)

    docs_by_text = {f"doc-{i}": t for i, t in enumerate(case["docs"])}
    rr = object.__new__(B200ColBERTReranker)
    rr.engine = SimpleNamespace(device=torch.device("cpu"), maxsim=fake_maxsim, rerank_postprocess=fake_postprocess)
    rr.compute_dtype, rr.batch_size, rr.colbert_weight, rr.bge_weight = torch.float32, 16, 0.8, 0.2
    rr.use_bge_reranker, rr.bge_reranker = use_bge, (Bge() if use_bge else None)
    rr.query_encoder = lambda text: case["queries"][int(text.split("-")[1])]
    rr.doc_encoder = lambda texts: [docs_by_text[t] for t in texts]
    return rr


def test_rerank_host_logic_matches_the_reference_golden_on_cpu():
    """rerank / batch_rerank_queries host logic (ColBERT-only, hybrid blend on the top 2*top_k, ties, batches) against
    the reference's own outputs (tests/golden/rerank_golden.json) with the device calls served by the CPU oracle —
    the CPU twin of test_dropin_gpu.py::test_reranker_dropin_matches_reference_golden."""
    import json
    import os

    from automative_rag_b200.documents import Document
    from tests._cases import RERANK_CASES, make_rerank_case

    gold_all = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rerank_golden.json")))
    for name, spec in RERANK_CASES.items():
        case, gold = make_rerank_case(spec), gold_all[name]
        rr = _oracle_backed_reranker(case, spec["use_bge"])
        docs = [Document(page_content=f"doc-{i}", metadata={"i": i}) for i in range(spec["n_docs"])]
        res = rr.rerank("q-0", docs, spec["top_k"])
        assert [d.metadata["i"] for d, _ in res] == [i for i, _ in gold["rerank"]], name
        np.testing.assert_allclose([s for _, s in res], [s for _, s in gold["rerank"]], rtol=2e-5, atol=1e-4)
        if spec.get("batch"):
            b = rr.batch_rerank_queries([f"q-{i}" for i in range(spec["n_queries"])], docs, spec["top_k"])
            assert list(b) == list(gold["batch"])
            for key, want in gold["batch"].items():
                assert [d.metadata["i"] for d, _ in b[key]] == [i for i, _ in want], (name, key)
                np.testing.assert_allclose([s for _, s in b[key]], [s for _, s in want], rtol=2e-5, atol=1e-4)
        assert rr.rerank("q-0", []) == [] and rr.batch_rerank_queries([], docs) == {}


def _oracle_backed_store(dim=64):
    """B200VectorStore over CPU tensors: the engine's two calls on this path (rs_filter_mask, rs_dense_topk_host) are
    served by the CPU oracle, so the class's own logic — normalisation on insert, payload columns, tombstones, filter
    compilation, the retry rule, Document assembly — runs without a GPU."""
    import zlib
    from types import SimpleNamespace

    from automative_rag_b200.filters import pack_bits
    from automative_rag_b200.vectorstore import B200Client, B200VectorStore
    from oracle import dense as odense

    class FakeEmbeddings:
        def _vec(self, text):
            return torch.randn(dim, generator=torch.Generator().manual_seed(zlib.crc32(text.encode()))).tolist()

        def embed_query(self, text):
            return self._vec(text)

        def embed_documents(self, texts):
            return [self._vec(t) for t in texts]

    def filter_mask(columns, value_sets, n, tombstone=None, out=None):
        ok = np.ones(n, dtype=bool)
        for col, vals in zip(columns, value_sets):
            ok &= np.isin(col.numpy(), np.asarray(list(vals), dtype=np.int32))
        words = torch.from_numpy(pack_bits(ok))
        return words if tombstone is None else words & ~tombstone

    def dense_topk_host(corpus, query, k, mask_dev=None, inv_norm=None, metric=1, **_):
        n = corpus.shape[0]
        bits = None
        if mask_dev is not None:
            w = mask_dev.numpy().view(np.uint32)
            bits = ((w[np.arange(n) // 32] >> (np.arange(n) % 32)) & 1).astype(bool)
        s, i = odense.topk(corpus.numpy(), query.numpy().reshape(-1), k, bits, odense.COSINE,
                           None if inv_norm is None else inv_norm.numpy())
        return torch.from_numpy(s)[None], torch.from_numpy(i)[None]

    client = object.__new__(B200Client)
    client.engine = SimpleNamespace(device=torch.device("cpu"), filter_mask=filter_mask, dense_topk_host=dense_topk_host)
    client.collections = {}
    return B200VectorStore(client, "cpu-col", FakeEmbeddings())


def test_vector_store_host_logic_on_cpu():
    from automative_rag_b200.documents import Document
    from oracle import dense as odense
    from oracle import filters as ofilters

    store = _oracle_backed_store()
    makers = ["Toyota", "Honda", "BMW"]
    docs = [Document(page_content=f"chunk {i} about {makers[i % 3]}",
                     metadata={"manufacturer": makers[i % 3], "year": 2020 + i % 4, "category": ["sedan", "suv"][i % 2],
                               "custom": f"c{i % 5}"}) for i in range(300)]
    ids = store.add_documents(docs)
    assert len(set(ids)) == 300 and store.get_stats()["vectors_count"] == 300
    assert all("ingestion_time" in d.metadata and d.metadata["id"] for d in docs)          # vectorstore.py:141-152

    emb = store.embedding_function
    c = np.asarray(emb.embed_documents([d.page_content for d in docs]), dtype=np.float32)
    c16 = (c / np.linalg.norm(c, axis=1, keepdims=True)).astype(np.float16)
    inv = (1.0 / np.linalg.norm(c16.astype(np.float32), axis=1)).astype(np.float32)
    payloads = [{"page_content": d.page_content, "metadata": d.metadata} for d in docs]
    for flt in (None, {"manufacturer": "Toyota"}, {"manufacturer": ["Honda", "BMW"], "year": 2021},
                {"category": "suv", "year": [2020, 2023]}, {"manufacturer": "Nobody"}, {"custom": "c3"}):
        res = store.similarity_search_with_score("What is the horsepower?", k=7, metadata_filter=flt)
        q16 = np.asarray(emb.embed_query("What is the horsepower?"), dtype=np.float32).astype(np.float16)
        ws, wi = odense.topk(c16, q16, 7, ofilters.filter_mask(payloads, flt or {}, None), odense.COSINE, inv)
        nv = int((wi >= 0).sum())
        assert [d.page_content for d, _ in res] == [docs[i].page_content for i in wi[:nv]], flt
        np.testing.assert_allclose([s for _, s in res], ws[:nv], rtol=1e-6)
        assert all(isinstance(s, float) for _, s in res)

    # delete: tombstoned rows never come back; unknown ids are ignored; stats follow
    top = store.similarity_search_with_score("query text", k=3)
    victim = ids[[d.page_content for d in docs].index(top[0][0].page_content)]
    store.delete_by_ids([victim, "no-such-id"])
    store.delete_by_ids([])
    assert store.get_stats()["vectors_count"] == 299
    again = store.similarity_search_with_score("query text", k=3)
    assert again[0][0].page_content == top[1][0].page_content
    assert store.get_embedding(victim) is None and len(store.get_embedding(ids[5])) == 64

    # vectorstore.py:199-207: a filtered search that raises is logged and retried without the filter
    col = store.collection
    real = col.device_mask
    col.device_mask = lambda flt: (_ for _ in ()).throw(RuntimeError("boom")) if flt is not None else real(None)
    assert len(store.similarity_search_with_score("q", k=5, metadata_filter={"manufacturer": "Toyota"})) == 5
    col.device_mask = real

    found = store.search_by_metadata({"manufacturer": "BMW", "year": 2021}, limit=5)
    assert 0 < len(found) <= 5 and all(d.metadata["manufacturer"] == "BMW" and d.metadata["year"] == 2021 for d in found)
    rep = store.repair_indices()
    assert rep["success"] and not rep["errors"] and "metadata.manufacturer" in rep["recreated_indices"]
    assert [d.page_content for d, _ in store.similarity_search_with_score("q", k=4, metadata_filter={"year": 2022})] == \
           [d.page_content for d, _ in store.similarity_search_with_score("q", k=4, metadata_filter={"year": [2022]})]
    assert store.add_documents([]) == []


def test_collection_persistence_metadata_search_and_lazy_columns_on_cpu(tmp_path):
    """VERDICT r1 items 5 / 12 and ADVICE r1 (low): save / load round trip (vectors, 1/|row|, payload columns, keyword
    dictionaries, tombstones, ids, payloads); search_by_metadata through the device mask; a column created on first use
    for an un-indexed payload key (`custom_filters`, query_models.py:22-28); values a column cannot express (a
    list-valued keyword, which Qdrant matches on any element; a numeric string) send the filter to the host evaluator
    instead of silently never matching; an integral float year matches like the integer."""
    from automative_rag_b200.documents import Document
    from automative_rag_b200.vectorstore import Collection

    store = _oracle_backed_store()
    makers = ["Toyota", "Honda", "BMW"]
    docs = [Document(page_content=f"chunk {i}", metadata={"manufacturer": makers[i % 3], "year": 2020 + i % 4,
                                                           "custom": f"c{i % 5}", "doors": 2 + i % 3}) for i in range(120)]
    docs[7].metadata["year"] = 2021.0                       # integral float: encoded as 2021
    ids = store.add_documents(docs)
    store.delete_by_ids([ids[3], ids[40]])
    col = store.collection
    assert "custom" not in col.columns
    # un-indexed keys: first use builds the column (keyword / integer by the stored values), later upserts fill it
    found = store.search_by_metadata({"custom": "c3", "doors": 4}, limit=100)
    want = [d for i, d in enumerate(docs) if i % 5 == 3 and 2 + i % 3 == 4 and i not in (3, 40)]
    assert [d.page_content for d in found] == [d.page_content for d in want]
    assert "custom" in col.keyword_dicts and "doors" in col.int_fields
    store.add_documents([Document(page_content="late", metadata={"manufacturer": "BMW", "year": 2021, "custom": "c3", "doors": 4})])
    assert store.search_by_metadata({"custom": "c3", "doors": 4}, limit=100)[-1].page_content == "late"
    assert [d.page_content for d in store.search_by_metadata({"year": 2021}, limit=3)] == ["chunk 1", "chunk 5", "chunk 7"]
    assert len(store.search_by_metadata({"manufacturer": "Toyota"}, limit=7)) == 7          # limit honoured, scroll order
    assert store.search_by_metadata({"manufacturer": "Nobody"}) == []

    # round trip
    before = store.similarity_search_with_score("What is the horsepower?", k=6, metadata_filter={"manufacturer": ["BMW", "Honda"]})
    store.save(str(tmp_path / "col"))
    fresh = _oracle_backed_store()
    fresh.load(str(tmp_path / "col"))
    c2 = fresh.collection
    assert c2.n == col.n and c2.deleted == 2 and c2.ids == col.ids and c2.keyword_dicts == col.keyword_dicts
    assert torch.equal(c2.vectors[: c2.n], col.vectors[: col.n]) and torch.equal(c2.inv_norm[: c2.n], col.inv_norm[: col.n])
    assert all(torch.equal(c2.columns[f][: c2.n], col.columns[f][: col.n]) for f in col.columns)
    assert fresh.get_embedding(ids[3]) is None and fresh.get_embedding(ids[5]) == store.get_embedding(ids[5])
    after = fresh.similarity_search_with_score("What is the horsepower?", k=6, metadata_filter={"manufacturer": ["BMW", "Honda"]})
    assert [(d.page_content, s) for d, s in after] == [(d.page_content, s) for d, s in before]
    fresh.delete_by_ids([ids[5]])                                                            # the loaded store stays usable
    assert fresh.get_stats()["vectors_count"] == store.get_stats()["vectors_count"] - 1

    # values the column cannot express: the filter must agree with the host evaluator, not silently drop the row
    odd = _oracle_backed_store()
    odd.add_documents([Document(page_content="multi", metadata={"manufacturer": ["Toyota", "Lexus"], "year": 2020}),
                       Document(page_content="plain", metadata={"manufacturer": "Toyota", "year": "2020"}),
                       Document(page_content="other", metadata={"manufacturer": "Honda", "year": 2020})])
    assert [d.page_content for d in odd.search_by_metadata({"manufacturer": "Toyota"})] == ["multi", "plain"]
    assert odd.collection.unencodable.get("manufacturer") and odd.collection.unencodable.get("year")
    res = odd.similarity_search_with_score("q", k=3, metadata_filter={"manufacturer": "Lexus"})
    assert [d.page_content for d, _ in res] == ["multi"]
    assert isinstance(Collection.load(str(tmp_path / "col"), fresh.client.engine), Collection)
