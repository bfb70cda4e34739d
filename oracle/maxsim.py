"""CPU restatement of ColBERT MaxSim scoring and the rerank tail — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/core/query/llm/rerankers.py:
  * `_compute_maxsim_scores` :215-265 — per document: `S = Q @ D.T` (:247), `max over doc
    tokens` (:250), `sum(max_sim[1:-1])` when Lq > 2 else `sum(max_sim)` (:255-261).
    No L2 normalisation, no padding mask on the document side.
  * `_explain_colbert_matches` :489-501 — same S, plus argmax per query token, and a score that
    sums only query tokens with attention_mask == 1 that are not [CLS]/[SEP] (a 0/1 weight).
  * `_colbert_rerank` :377-380 — stable `sorted(zip(docs, scores), reverse=True)`.
  * `rerank` :298-343 — min-max normalise ColBERT scores over the candidate set (all equal ->
    1.0), same for the second (BGE) score vector, `0.8*c + 0.2*b`, stable sort, `[:top_k]`.

Pinned against the reference's own function by tests/golden/make_golden.py (see
oracle/__init__.py).  Never imported by the product package.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def reference_weights(lq: int) -> np.ndarray:
    """Per-query-token weights equivalent to rerankers.py:255-261."""
    w = np.ones(lq, dtype=np.float32)
    if lq > 2:
        w[0] = 0.0
        w[-1] = 0.0
    return w


def _as_f32(t) -> torch.Tensor:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    return t.detach().to("cpu").to(torch.float32)


def maxsim_scores(
    query_embeddings,
    doc_embeddings_list: Sequence,
    weights: Optional[np.ndarray] = None,
    return_argmax: bool = False,
):
    """Scores of one query against a list of documents, fp32 on CPU.

    `query_embeddings` is [Lq, D] or [1, Lq, D] (squeezed as rerankers.py:234-238); inputs in
    fp16/bf16 are upcast to fp32 first (the parity protocol: the oracle consumes the identical
    rounded values).  Returns float32 [n_docs] (and int32 [n_docs, Lq] argmax if requested).
    """
    q = _as_f32(query_embeddings)
    if q.dim() == 3 and q.size(0) == 1:
        q = q.squeeze(0)
    lq = q.size(0)
    w = torch.from_numpy(reference_weights(lq) if weights is None else np.asarray(weights, dtype=np.float32))
    scores = np.zeros(len(doc_embeddings_list), dtype=np.float32)
    argmax = np.zeros((len(doc_embeddings_list), lq), dtype=np.int32)
    for i, d in enumerate(doc_embeddings_list):
        sim = torch.matmul(q, _as_f32(d).T)          # :247
        max_sim, idx = sim.max(dim=1)                # :250 (:492 for the indices)
        nz = w != 0
        scores[i] = float((max_sim[nz] * w[nz]).sum())   # :255-261 / :495-501 as a weighted sum
        argmax[i] = idx.numpy().astype(np.int32)
    return (scores, argmax) if return_argmax else scores


def maxsim_scores_packed(q, q_weight, doc_tokens, doc_offsets, cand=None):
    """Batched form over a packed token buffer, vectorised per document.

    q [nq, Lq, D]; q_weight [nq, Lq] or None; doc_tokens [T, D]; doc_offsets [nd+1];
    cand None (every query x every doc) or [nq, nc] doc indices.  Returns float32 [nq, nd|nc].
    """
    q = _as_f32(q)
    toks = _as_f32(doc_tokens)
    off = np.asarray(doc_offsets, dtype=np.int64)
    nq, lq, _ = q.shape
    nd = len(off) - 1
    if q_weight is None:
        w = torch.from_numpy(np.tile(reference_weights(lq), (nq, 1)))
    else:
        w = _as_f32(q_weight)
    if cand is None:
        out = torch.empty(nq, nd, dtype=torch.float32)
        qf = q.reshape(nq * lq, -1)
        for j in range(nd):
            sim = qf @ toks[off[j]:off[j + 1]].T            # [nq*lq, Ld]
            m = sim.max(dim=1).values.reshape(nq, lq)
            out[:, j] = (m * w).sum(dim=1)
        return out.numpy()
    cand = np.asarray(cand, dtype=np.int64)
    out = np.empty(cand.shape, dtype=np.float32)
    for qi in range(nq):
        for c, j in enumerate(cand[qi]):
            sim = q[qi] @ toks[off[j]:off[j + 1]].T
            out[qi, c] = float((sim.max(dim=1).values * w[qi]).sum())
    return out


def stable_rank(scores: Sequence[float]) -> List[int]:
    """Indices in the order of `sorted(zip(docs, scores), key=score, reverse=True)` (:377-380)."""
    return [i for i, _ in sorted(enumerate(scores), key=lambda x: x[1], reverse=True)]


def minmax(values: Sequence[float]) -> List[float]:
    """rerankers.py:302-310 / :319-327 — all-equal input maps to 1.0."""
    lo, hi = min(values), max(values)
    rng = hi - lo
    if rng > 0:
        return [(v - lo) / rng for v in values]
    return [1.0 for _ in values]


def hybrid_rerank(colbert_scores: Sequence[float], other_scores: Optional[Sequence[float]],
                  w_a: float = 0.8, w_b: float = 0.2, top_k: Optional[int] = None) -> List[Tuple[int, float]]:
    """The tail of `rerank` (:298-343) as (input index, final score) pairs.

    `other_scores[i]` is the second model's score for input document i (the reference computes it
    for documents in ColBERT order, :314-317; indexing by document is equivalent).
    """
    order = stable_rank(colbert_scores)
    if other_scores is None:
        out = [(i, float(colbert_scores[i])) for i in order]
        return out[:top_k] if top_k is not None else out
    a = minmax([float(colbert_scores[i]) for i in order])
    b = minmax([float(other_scores[i]) for i in order])
    combined = [w_a * a[r] + w_b * b[r] for r in range(len(order))]
    ranked = sorted(zip(order, combined), key=lambda x: x[1], reverse=True)
    return ranked[:top_k] if top_k is not None else ranked
