"""CPU restatement of exact filtered cosine / dot top-k — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/__init__.py): the arithmetic behind
QdrantStore.similarity_search_with_score (/root/reference/src/core/query/retrieval/
vectorstore.py:166-214) lives in third-party code absent from /root/reference — qdrant-client
1.13.3 local mode (poetry.lock:5310-5311; the backend of `QdrantClient(location=":memory:")`,
tests/conftest.py:80) and the Qdrant server.  This file restates the published algorithm of
qdrant-client local mode:
  * `qdrant_client/local/distances.py::cosine_similarity` — normalise the query and the vectors
    in float32 (zero norms guarded), `np.dot(vectors, query)`;  DOT — plain `np.dot`;
  * `qdrant_client/local/local_collection.py::search` — payload-filter mask AND not-deleted,
    `np.argsort(scores)[::-1]` (larger is better), walk the order skipping masked rows until
    `limit` hits.
anchored on the reference's call sites: collection created with Distance.COSINE
(vectorstore.py:52-57,75-81), search with `k` and an optional filter (:192-196,:209-212).

Tie order: qdrant's is unspecified; the engine's contract is (score desc, id asc), which this
oracle produces with a stable sort.  Never imported by the product package.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

IP, COSINE = 0, 1


def pack_mask(bits: np.ndarray) -> np.ndarray:
    """bool [n] -> uint32 [ceil(n/32)], LSB-first (bit i&31 of word i>>5 == row i passes)."""
    bits = np.asarray(bits, dtype=bool)
    n = bits.shape[0]
    padded = np.zeros(((n + 31) // 32) * 32, dtype=np.uint8)
    padded[:n] = bits
    return np.packbits(padded.reshape(-1, 32), axis=1, bitorder="little").view("<u4").reshape(-1).copy()


def unpack_mask(words: np.ndarray, n: int) -> np.ndarray:
    b = np.unpackbits(np.asarray(words, dtype="<u4").view(np.uint8), bitorder="little")
    return b[:n].astype(bool)


def scores_f32(corpus: np.ndarray, query: np.ndarray, metric: int = COSINE,
               inv_norm: Optional[np.ndarray] = None) -> np.ndarray:
    """All n scores in float32.  `corpus` may be fp16 (or bf16 pre-upcast): it is upcast first.

    COSINE with inv_norm=None assumes rows are stored unit-length (Qdrant normalises on insert for
    Distance.COSINE) and normalises only the query — the same convention as rs_dense_topk.
    """
    c = np.asarray(corpus, dtype=np.float32)
    q = np.asarray(query, dtype=np.float32)
    s = c @ q
    if metric == COSINE:
        qn = float(np.sqrt(np.dot(q, q)))
        s = s * np.float32(1.0 / qn if qn > 0 else 0.0)
        if inv_norm is not None:
            s = s * np.asarray(inv_norm, dtype=np.float32)
    return s.astype(np.float32)


def topk(corpus: np.ndarray, query: np.ndarray, k: int, mask: Optional[np.ndarray] = None,
         metric: int = COSINE, inv_norm: Optional[np.ndarray] = None, id_base: int = 0
         ) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k (scores float32 [k], ids int64 [k]); padded with (-inf, -1) when < k rows pass.

    `mask` is a bool [n] array (True = passes) or None.
    """
    s = scores_f32(corpus, query, metric, inv_norm)
    n = s.shape[0]
    keep = np.ones(n, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
    idx = np.nonzero(keep)[0]
    # stable sort on -score keeps ascending row order inside ties: (score desc, id asc)
    order = idx[np.argsort(-s[idx], kind="stable")][:k]
    out_s = np.full(k, -np.inf, dtype=np.float32)
    out_i = np.full(k, -1, dtype=np.int64)
    out_s[: len(order)] = s[order]
    out_i[: len(order)] = order + id_base
    return out_s, out_i


def merge_topk(scores: np.ndarray, ids: np.ndarray, k_out: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard lists [nlists, nq, k_in] -> [nq, k_out] in (score desc, id asc) order."""
    nl, nq, k_in = scores.shape
    out_s = np.full((nq, k_out), -np.inf, dtype=np.float32)
    out_i = np.full((nq, k_out), -1, dtype=np.int64)
    for q in range(nq):
        s = scores[:, q, :].reshape(-1)
        i = ids[:, q, :].reshape(-1)
        ok = i >= 0
        s, i = s[ok], i[ok]
        order = np.lexsort((i, -s.astype(np.float64)))[:k_out]
        out_s[q, : len(order)] = s[order]
        out_i[q, : len(order)] = i[order]
    return out_s, out_i
