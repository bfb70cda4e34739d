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


# ---------------------------------------------------------------------------------- explanations
class StubTokenizer:
    """Whitespace word-piece look-alike with BERT's interface subset the explanations path uses
    (rerankers.py:432-443, :460-473, :503): [CLS] w1 w2 ... [SEP] [PAD]...; words longer than 6 characters are split
    into a head and a '##' continuation so the wordpiece branches run."""

    def __init__(self):
        self.vocab = ["[PAD]", "[CLS]", "[SEP]", "[UNK]"]
        self.index = {t: i for i, t in enumerate(self.vocab)}

    def _pieces(self, text):
        out = []
        for w in text.lower().split():
            out.extend([w[:6], "##" + w[6:]] if len(w) > 6 else [w])
        return out

    def _id(self, tok):
        if tok not in self.index:
            self.index[tok] = len(self.vocab)
            self.vocab.append(tok)
        return self.index[tok]

    def __call__(self, texts, add_special_tokens=True, max_length=32, padding="max_length", truncation=True,
                 return_tensors="pt"):
        from types import SimpleNamespace

        ids, masks = [], []
        for t in texts:
            toks = ["[CLS]"] + self._pieces(t)[: max_length - 2] + ["[SEP]"]
            row = [self._id(x) for x in toks]
            masks.append([1] * len(row) + [0] * (max_length - len(row)))
            ids.append(row + [0] * (max_length - len(row)))
        return SimpleNamespace(input_ids=torch.tensor(ids), attention_mask=torch.tensor(masks))

    def convert_ids_to_tokens(self, ids):
        return [self.vocab[i] for i in ids]


EXPLAIN_CASE = dict(
    seed=31, d=64, max_query_length=16, max_doc_length=24, num_explanations=4,
    query="which engine has the strongest horsepower rating",
    docs=["the turbocharged engine delivers impressive horsepower figures on track",
          "comfortable seats and a quiet cabin make long trips pleasant",
          "horsepower rating of the strongest variant exceeds expectations easily"],
)


def make_explain_case(spec: dict = EXPLAIN_CASE):
    """Token embeddings with unambiguous matches: every distinct word piece gets a random unit vector, a token's
    embedding is its piece's vector plus small position noise, specials / padding get their own vectors — so a query
    token's best document token is the same piece when the document has it (margin ~1 vs ~0.1), also in fp16."""
    tok = StubTokenizer()
    g = torch.Generator().manual_seed(spec["seed"])
    table = {}

    def vec(piece):
        if piece not in table:
            v = torch.randn(spec["d"], generator=g)
            table[piece] = v / v.norm()
        return table[piece]

    def embed(text, max_length):
        enc = tok([text], max_length=max_length)
        pieces = tok.convert_ids_to_tokens(enc.input_ids[0].tolist())
        rows = [vec(p) + 0.02 * torch.randn(spec["d"], generator=g) for p in pieces]
        return torch.stack(rows)

    q = embed(spec["query"], spec["max_query_length"]).unsqueeze(0)              # [1, Lq, d]
    docs = {t: embed(t, spec["max_doc_length"]) for t in spec["docs"]}           # text -> [Ld, d]
    return tok, q, docs
