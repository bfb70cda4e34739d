"""Seeded synthetic inputs shared by the golden generator and the tests."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

# name -> spec.  "lens" overrides a uniform doc length with ragged lengths.
MAXSIM_CASES: Dict[str, dict] = {
    # BASELINE config 1: 1 query x 32 tokens vs 100 docs x 180 tokens, dim 128, fp32
    "config1": dict(seed=0, lq=32, d=128, n_docs=100, ld=180),
    # deployed shape of the reference: Lq=32, Ld=256 (docs padded to max_doc_length), 768-d hidden state
    "deployed768": dict(seed=11, lq=32, d=768, n_docs=8, ld=256),
    # ragged documents incl. a 1-token doc and one longer than any tile
    "ragged": dict(seed=12, lq=32, d=128, n_docs=9, lens=[1, 2, 7, 31, 32, 33, 180, 300, 517]),
    # Lq <= 2: the reference sums ALL query tokens (rerankers.py:259-261)
    "lq2": dict(seed=13, lq=2, d=64, n_docs=5, ld=20),
    "lq1": dict(seed=14, lq=1, d=64, n_docs=5, ld=20),
    # Lq = 3: exactly one content token survives [1:-1]
    "lq3": dict(seed=15, lq=3, d=64, n_docs=6, ld=17),
    # non-multiple-of-16 query length, larger than 32
    "lq45": dict(seed=16, lq=45, d=128, n_docs=7, ld=64),
    # single document
    "one_doc": dict(seed=17, lq=32, d=128, n_docs=1, ld=300),
}

RERANK_CASES: Dict[str, dict] = {
    "colbert_only": dict(seed=21, lq=32, d=64, n_docs=12, ld=40, use_bge=False, top_k=5, n_queries=1),
    "hybrid": dict(seed=22, lq=32, d=64, n_docs=12, ld=40, use_bge=True, top_k=5, n_queries=1),
    "hybrid_all": dict(seed=23, lq=32, d=64, n_docs=9, ld=33, use_bge=True, top_k=None, n_queries=1),
    # duplicated documents -> exactly tied ColBERT scores: pins the stable-sort tie rule
    "ties": dict(seed=24, lq=32, d=64, n_docs=10, ld=24, use_bge=False, top_k=None, n_queries=1, dup=[(1, 4), (2, 7)]),
    "batch": dict(seed=25, lq=32, d=64, n_docs=14, ld=30, use_bge=False, top_k=4, n_queries=3, batch=True),
    "batch_hybrid": dict(seed=26, lq=32, d=64, n_docs=14, ld=30, use_bge=True, top_k=4, n_queries=3, batch=True),
}


def make_maxsim_case(spec: dict) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """fp32 query [1, lq, d] and a list of fp32 docs [ld_i, d], iid N(0,1), torch.Generator(seed)."""
    g = torch.Generator().manual_seed(spec["seed"])
    q = torch.randn(1, spec["lq"], spec["d"], generator=g)
    lens = spec.get("lens") or [spec["ld"]] * spec["n_docs"]
    docs = [torch.randn(n, spec["d"], generator=g) for n in lens]
    return q, docs


def make_rerank_case(spec: dict) -> dict:
    g = torch.Generator().manual_seed(spec["seed"])
    queries = [torch.randn(1, spec["lq"], spec["d"], generator=g) for _ in range(spec["n_queries"])]
    docs = [torch.randn(spec["ld"], spec["d"], generator=g) for _ in range(spec["n_docs"])]
    for a, b in spec.get("dup", []):
        docs[b] = docs[a].clone()
    bge = torch.randn(spec["n_docs"], generator=g).numpy().astype(np.float32)
    return {"queries": queries, "docs": docs, "bge": bge}


# ---------------------------------------------------------------------------------- dense
def make_dense_case(seed: int, n: int, d: int, dtype: torch.dtype = torch.float16, normalise: bool = True,
                    device: str = "cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Corpus [n, d] iid N(0,1), L2-normalised in fp32, cast to `dtype`; one query by the same law."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    c = torch.randn(n, d, generator=g)
    if normalise:
        c = c / c.norm(dim=1, keepdim=True)
    q = torch.randn(d, generator=g)
    if normalise:
        q = q / q.norm()
    return c.to(dtype).to(device), q.to(dtype).to(device)


def bernoulli_mask(seed: int, n: int, p: float) -> np.ndarray:
    """bool [n], True = row passes, Bernoulli(p)."""
    return np.random.default_rng(seed).random(n) < p
