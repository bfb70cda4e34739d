"""B200-native retrieval-scoring engine behind the call signatures of jliang87/Automative-RAG's
vector-store search and ColBERT reranker.

  B200VectorStore       <- QdrantStore           (src/core/query/retrieval/vectorstore.py)
  B200ColBERTReranker   <- ColBERTReranker       (src/core/query/llm/rerankers.py)
  HybridRetriever       <- the retrieve-then-rerank composition of tests/test_retrieval.py:206-258
  Engine                   ctypes handle on librag_b200.so (include/rag_b200.h)

Importing the package does not load the CUDA library; constructing an Engine / store / reranker
does, and raises when the library or a B200 is missing — there is no CPU fallback.
"""
from . import _ffi
from ._ffi import Engine, EngineError, get_engine, load_library
from .distributed import (ShardedCandidateMaxSim, ShardedDenseIndex, ShardedMaxSim, owned_candidates,
                          partition_candidates, shard_bounds)
from .documents import Document
from .filters import FieldCondition, Filter, MatchValue, Range, build_filter
from .rerankers import B200ColBERTReranker, pack_documents
from .retriever import HybridRetriever
from .vectorstore import B200Client, B200VectorStore, Collection

__all__ = [
    "Engine", "EngineError", "get_engine", "load_library", "Document", "Filter", "FieldCondition", "MatchValue",
    "Range", "build_filter", "B200Client", "B200VectorStore", "Collection", "B200ColBERTReranker", "pack_documents",
    "HybridRetriever", "ShardedDenseIndex", "ShardedMaxSim", "ShardedCandidateMaxSim", "partition_candidates",
    "owned_candidates",
    "shard_bounds",
]
