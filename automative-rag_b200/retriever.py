"""Retrieve-then-rerank composition.

The reference's tests describe `HybridRetriever(vector_store, reranker, top_k, rerank_top_k)` doing
`search(k=top_k)` -> `reranker.rerank(query=, documents=, top_k=rerank_top_k)`
(tests/test_retrieval.py:206-258); the class itself is gone from HEAD.  This restores it over the
B200 vector store and reranker.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

from .documents import Document


class HybridRetriever:
    def __init__(self, vector_store, reranker=None, top_k: int = 20, rerank_top_k: int = 5):
        self.vector_store = vector_store
        self.reranker = reranker
        self.top_k = top_k
        self.rerank_top_k = rerank_top_k

    def retrieve(self, query: str, metadata_filter: Optional[Dict[str, Union[str, List[str], int, List[int]]]] = None,
                 rerank: bool = True) -> List[Tuple[Document, float]]:
        initial = self.vector_store.similarity_search_with_score(query=query, k=self.top_k,
                                                                 metadata_filter=metadata_filter)
        if not initial:
            return []
        if rerank and self.reranker is not None:
            docs = [doc for doc, _ in initial]
            return self.reranker.rerank(query=query, documents=docs, top_k=self.rerank_top_k)
        return initial[: self.rerank_top_k]
