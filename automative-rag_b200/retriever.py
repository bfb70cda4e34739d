"""Retrieve-then-rerank composition.

The reference's tests describe `HybridRetriever(vector_store, reranker, top_k, rerank_top_k)` doing
`search(k=top_k)` -> `reranker.rerank(query=, documents=, top_k=rerank_top_k)` and returning
`(results, execution_time)` (tests/test_retrieval.py:206-276); the class itself is gone from HEAD.  This restores it over the
B200 vector store and reranker.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Tuple, Union

from .documents import Document


# Per-mode sizes of the two stages: (retrieval_k for the dense search, final_k kept after reranking) — the values of
# the reference's mode table (src/core/query/llm/mode_config.py:29-135, read through get_retrieval_params :155-164;
# retrieve_documents_task passes mode.retrieval_k to the vector store, retrieval_tasks.py:74-79).  An unknown mode
# falls back to "facts", as the reference's `.get(mode, FACTS)` does.
MODE_RETRIEVAL_PARAMS: Dict[str, Tuple[int, int]] = {
    "facts": (20, 8),
    "features": (30, 12),
    "tradeoffs": (35, 15),
    "scenarios": (30, 12),
    "debate": (40, 18),
    "quotes": (25, 10),
}


def retrieval_params(mode) -> Tuple[int, int]:
    """(retrieval_k, final_k) of a query mode (a string or the reference's QueryMode str-enum)."""
    key = getattr(mode, "value", mode)
    return MODE_RETRIEVAL_PARAMS.get(str(key).lower(), MODE_RETRIEVAL_PARAMS["facts"])


class HybridRetriever:
    def __init__(self, vector_store, reranker=None, top_k: int = 20, rerank_top_k: int = 5):
        self.vector_store = vector_store
        self.reranker = reranker
        self.top_k = top_k
        self.rerank_top_k = rerank_top_k

    def retrieve(self, query: str, metadata_filter: Optional[Dict[str, Union[str, List[str], int, List[int]]]] = None,
                 rerank: bool = True, mode=None) -> Tuple[List[Tuple[Document, float]], float]:
        """search(k) -> rerank(top_k).  Returns (results, execution_time in seconds), the shape the reference's tests
        unpack (tests/test_retrieval.py:236-240, :265-269).  `mode` (optional) takes both sizes from the reference's
        per-mode table instead of the constructor's top_k / rerank_top_k."""
        start = time.perf_counter()
        top_k, rerank_top_k = (self.top_k, self.rerank_top_k) if mode is None else retrieval_params(mode)
        results = self.vector_store.similarity_search_with_score(query=query, k=top_k, metadata_filter=metadata_filter)
        if results:
            if rerank and self.reranker is not None:
                docs = [doc for doc, _ in results]
                results = self.reranker.rerank(query=query, documents=docs, top_k=rerank_top_k)
            else:
                results = results[:rerank_top_k]
        return results, max(time.perf_counter() - start, 1e-9)
