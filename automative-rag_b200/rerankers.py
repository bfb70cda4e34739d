"""B200ColBERTReranker — drop-in for the reference's ColBERTReranker on the MaxSim path.

Mirrors src/core/query/llm/rerankers.py of the reference: `_compute_maxsim_scores(query_embeddings,
doc_embeddings_list) -> List[float]` (:215-265) is the exact seam; `rerank` (:267-349),
`_colbert_rerank` (:351-385), `batch_rerank_queries` (:563-662) and
`rerank_with_explanations` / `_explain_colbert_matches` (:387-561) keep their signatures and
result shapes.  The encoders (BERT forward passes, the BGE cross-encoder) are out of scope and are
injected: `query_encoder(text) -> Tensor [1, Lq, D]`, `doc_encoder(texts) -> List[Tensor [Ld_i, D]]`,
`cross_encoder.predict(pairs) -> scores`.

Where the reference loops over documents launching matmul / max / sum and syncing with `.item()`
per document (:244-263), this class packs the candidate token embeddings once and makes ONE
`rs_maxsim` call; the sort / min-max / blend tail (:302-343) is one `rs_rerank_postprocess` call.
"""
from __future__ import annotations

import logging
import time
from operator import attrgetter
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _ffi
from .documents import Document

logger = logging.getLogger(__name__)
_shape_of = attrgetter("shape")
_dtype_of = attrgetter("dtype")


# Pinned staging for host-resident embeddings (the reference's CPU call shape): the rows are concatenated straight
# into page-locked memory and go up in one DMA; a pageable source would be copied once more by the driver.
_PINNED: Dict[Tuple[int, torch.dtype], Tuple[torch.Tensor, "torch.cuda.Event"]] = {}


def _stage_host_rows(mats: Sequence[torch.Tensor], rows: int, device: torch.device) -> torch.Tensor:
    cols, dt = mats[0].shape[1], mats[0].dtype
    key = (device.index if device.index is not None else torch.cuda.current_device(), dt)
    buf, ev = _PINNED.get(key, (None, None))
    if ev is not None:
        ev.synchronize()  # the previous upload out of this buffer has finished
    if buf is None or buf.numel() < rows * cols:
        buf = torch.empty(max(rows * cols, 1 << 20), dtype=dt).pin_memory()
    view = buf[: rows * cols].view(rows, cols)
    torch.cat(list(mats), dim=0, out=view)
    with torch.cuda.device(device):
        on_dev = view.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
    _PINNED[key] = (buf, ev)
    return on_dev


def pack_documents(doc_embeddings_list: Sequence[torch.Tensor], device: torch.device, dtype: torch.dtype
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """List of [Ld_i, D] tensors -> (tokens [sum Ld, D] on `device`, offsets int32 [n+1] on `device`)."""
    lens = []
    mats = []
    for d in doc_embeddings_list:
        if d.dim() == 3 and d.size(0) == 1:
            d = d.squeeze(0)
        if d.dim() != 2:
            raise ValueError(f"document embedding must be [Ld, D], got {tuple(d.shape)}")
        if d.size(0) == 0:
            raise ValueError("document with zero tokens")  # torch.max over an empty dim raises in the reference too
        lens.append(d.size(0))
        mats.append(d)
    device = torch.device(device)
    if len({(m.device, m.dtype) for m in mats}) == 1:
        # one concatenation where the tensors live, then ONE transfer / cast (100 small H2D copies cost more
        # than the scoring kernel itself)
        if mats[0].device.type == "cpu" and device.type == "cuda":
            tokens = _stage_host_rows(mats, sum(lens), device).to(dtype).contiguous()
        else:
            tokens = torch.cat(mats, dim=0).to(device=device, dtype=dtype).contiguous()
    else:
        tokens = torch.cat([m.to(device=device, dtype=dtype) for m in mats], dim=0).contiguous()
    offs = torch.zeros(len(lens) + 1, dtype=torch.int64)
    offs[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0)
    if int(offs[-1]) >= 2**31:
        raise ValueError("too many document tokens for int32 offsets")
    return tokens, offs.to(torch.int32).to(device)


class B200ColBERTReranker:
    def __init__(
            self,
            model_name: str = "colbertv2.0",
            device: Optional[str] = None,
            max_query_length: int = 32,
            max_doc_length: int = 256,
            batch_size: int = 16,
            use_fp16: bool = True,
            similarity_metric: str = "maxsim",
            checkpoint_path: Optional[str] = None,
            use_bge_reranker: bool = True,
            colbert_weight: float = 0.8,
            bge_weight: float = 0.2,
            bge_model_name: str = "BAAI/bge-reranker-base",
            *,
            query_encoder: Optional[Callable[[str], torch.Tensor]] = None,
            doc_encoder: Optional[Callable[[List[str]], List[torch.Tensor]]] = None,
            cross_encoder: Any = None,
            tokenizer: Any = None,
            compute_dtype: Optional[torch.dtype] = None,
    ):
        self.device = device or "cuda:0"
        self.engine = _ffi.get_engine(self.device)  # raises without a B200: no CPU fallback
        self.model_name = model_name
        self.max_query_length = max_query_length
        self.max_doc_length = max_doc_length
        self.batch_size = batch_size
        self.use_fp16 = use_fp16
        self.amp_enabled = use_fp16
        self.similarity_metric = similarity_metric
        self.checkpoint_path = checkpoint_path
        self.colbert_weight = colbert_weight
        self.bge_weight = bge_weight
        self.bge_model_name = bge_model_name
        self.query_encoder = query_encoder
        self.doc_encoder = doc_encoder
        self.tokenizer = tokenizer
        self.bge_reranker = cross_encoder
        # rerankers.py:102-104 — a BGE model that fails to load silently disables the hybrid branch
        self.use_bge_reranker = bool(use_bge_reranker and cross_encoder is not None)
        # the reference scores under fp16 autocast on CUDA (:245) and in fp32 on CPU
        self.compute_dtype = compute_dtype or (torch.float16 if use_fp16 else torch.float32)

    # -- encoders (out of scope: injected) ----------------------------------------------------
    def _encode_query(self, query: str) -> torch.Tensor:  # rerankers.py:133-165
        if self.query_encoder is None:
            raise ValueError("no query_encoder configured: the ColBERT encoder forward pass is injected")
        return self.query_encoder(query)

    def _encode_documents_batched(self, documents: List[Document]) -> List[torch.Tensor]:  # :167-213
        if self.doc_encoder is None:
            raise ValueError("no doc_encoder configured: the ColBERT encoder forward pass is injected")
        out: List[torch.Tensor] = []
        for i in range(0, len(documents), self.batch_size):
            out.extend(self.doc_encoder([d.page_content for d in documents[i:i + self.batch_size]]))
        return out

    # -- the hot path -------------------------------------------------------------------------
    def _maxsim_device(self, query_embeddings: torch.Tensor, doc_embeddings_list: Sequence[torch.Tensor],
                       q_weight: Optional[torch.Tensor] = None, want_argmax: bool = False, want_tokmax: bool = False):
        q = query_embeddings
        if q.dim() == 2:
            q = q.unsqueeze(0)
        dev = self.engine.device
        q = q.to(device=dev, dtype=self.compute_dtype).contiguous()
        tokens, offs = pack_documents(doc_embeddings_list, dev, self.compute_dtype)
        if q_weight is not None:
            q_weight = q_weight.to(device=dev, dtype=torch.float32).reshape(q.shape[0], q.shape[1]).contiguous()
        return self.engine.maxsim(q, tokens, offs, q_weight=q_weight, want_argmax=want_argmax, want_tokmax=want_tokmax)

    def _compute_maxsim_scores(self, query_embeddings: torch.Tensor,
                               doc_embeddings_list: List[torch.Tensor]) -> List[float]:
        """rerankers.py:215-265, one kernel launch instead of 3 launches + 1 sync per document."""
        if len(doc_embeddings_list) == 0:
            return []
        q = query_embeddings
        if q.dim() == 3 and q.size(0) != 1:
            raise ValueError("_compute_maxsim_scores takes one query ([Lq, D] or [1, Lq, D])")
        fast = self._list_call(q, doc_embeddings_list)
        if fast is not None:
            return fast
        scores = self._maxsim_device(q, doc_embeddings_list)
        return scores[0].tolist()

    def _list_call(self, q: torch.Tensor, docs: Sequence[torch.Tensor]) -> Optional[List[float]]:
        """The reference's call shape in ONE C call (rs_maxsim_list) when the list is uniform: every document a
        contiguous [Ld, D] tensor of the query's dtype, all on the host or all on the engine's GPU.  Anything else
        (mixed devices / dtypes, [1, Ld, D] entries, views) takes the general pack-and-call path."""
        q2 = q[0] if q.dim() == 3 else q
        dev = self.engine.device
        if q2.dim() != 2 or not q2.is_contiguous() or q2.dtype not in (torch.float16, torch.bfloat16, torch.float32):
            return None
        if q2.is_cuda and q2.device != dev:
            return None
        d = q2.shape[1]
        if d % 8:
            return None
        # one C-level pass per property (map / set / all) instead of a Python loop over the documents: this check is
        # on the latency path of every rerank call
        try:
            lens = [n for n, dd in map(_shape_of, docs) if dd == d]
        except ValueError:  # an entry that is not 2-D
            return None
        if len(lens) != len(docs):
            return None
        if set(map(_dtype_of, docs)) != {q2.dtype} or not all(map(torch.Tensor.is_contiguous, docs)):
            return None
        where = set(map(torch.Tensor.get_device, docs))  # -1 = host
        if len(where) != 1 or (where != {-1} and where != {dev.index}):
            return None
        if 0 in lens:
            raise ValueError("document with zero tokens")  # torch.max over an empty dim raises in the reference too
        return self.engine.maxsim_list(q2, docs, self.compute_dtype, doc_lens=lens)

    # -- rerank tail --------------------------------------------------------------------------
    def _order(self, scores: torch.Tensor, other: Optional[torch.Tensor], top_k: Optional[int]
               ) -> List[Tuple[int, float]]:
        """(input index, final score) in reference order for one query's score vector(s)."""
        n = scores.numel()
        k = n if top_k is None else min(top_k, n)
        if n > 16384:
            raise ValueError(f"rerank tail handles at most 16384 candidates per query on the device (got {n})")
        idx, out = self.engine.rerank_postprocess(
            scores.reshape(1, n), None if other is None else other.reshape(1, n), k,
            self.colbert_weight, self.bge_weight)
        return list(zip(idx[0].tolist(), out[0].tolist()))

    def _colbert_rerank(self, query: str, documents: List[Document]) -> List[Tuple[Document, float]]:
        """rerankers.py:351-385."""
        if not documents:
            return []
        start_time = time.time()
        query_embeddings = self._encode_query(query)
        doc_embeddings_list = self._encode_documents_batched(documents)
        scores = self._maxsim_device(query_embeddings, doc_embeddings_list)[0]
        ranked = [(documents[i], s) for i, s in self._order(scores, None, None)]
        logger.info(f"ColBERT scoring completed in {time.time() - start_time:.2f}s for {len(documents)} documents")
        return ranked

    def rerank(self, query: str, documents: List[Document], top_k: Optional[int] = None
               ) -> List[Tuple[Document, float]]:
        """rerankers.py:267-349."""
        if not documents:
            return []
        if not self.use_bge_reranker:
            results = self._colbert_rerank(query, documents)
            return results[:top_k] if top_k is not None else results
        query_embeddings = self._encode_query(query)
        doc_embeddings_list = self._encode_documents_batched(documents)
        scores = self._maxsim_device(query_embeddings, doc_embeddings_list)[0]
        pairs = [[query, doc.page_content] for doc in documents]
        bge = torch.as_tensor(self.bge_reranker.predict(pairs), dtype=torch.float32).to(self.engine.device)
        return [(documents[i], s) for i, s in self._order(scores, bge, top_k)]

    def batch_rerank_queries(self, queries: List[str], documents: List[Document], top_k: Optional[int] = None
                             ) -> Dict[str, List[Tuple[Document, float]]]:
        """rerankers.py:563-662: documents encoded once, ALL queries scored in one rs_maxsim call
        (the shared-candidate shape the tcgen05 kernel is built for)."""
        if not documents or not queries:
            return {}
        doc_embeddings_list = self._encode_documents_batched(documents)
        q = torch.cat([self._encode_query(qs).reshape(1, -1, doc_embeddings_list[0].shape[-1]) for qs in queries], dim=0)
        scores = self._maxsim_device(q, doc_embeddings_list)  # [nq, nd]
        results: Dict[str, List[Tuple[Document, float]]] = {}
        for qi, query in enumerate(queries):
            if self.use_bge_reranker:
                # the reference blends only the top 2*top_k ColBERT candidates (:604)
                first = self._order(scores[qi], None, top_k * 2 if top_k else None)
                keep = [i for i, _ in first]
                sub = scores[qi][torch.tensor(keep, device=scores.device)]
                pairs = [[query, documents[i].page_content] for i in keep]
                bge = torch.as_tensor(self.bge_reranker.predict(pairs), dtype=torch.float32).to(scores.device)
                ranked = self._order(sub.contiguous(), bge, top_k if top_k else None)
                results[query] = [(documents[keep[i]], s) for i, s in ranked]
            else:
                results[query] = [(documents[i], s) for i, s in self._order(scores[qi].contiguous(), None, top_k or None)]
        return results

    # -- explanations (rerankers.py:387-561) ---------------------------------------------------
    def rerank_with_explanations(self, query: str, documents: List[Document], top_k: Optional[int] = None,
                                 num_explanations: int = 5) -> List[Dict]:
        reranked_docs = self.rerank(query, documents, top_k)
        docs = [doc for doc, _ in reranked_docs]
        return self._explain_colbert_matches(query, docs, num_explanations)

    def _explain_colbert_matches(self, query: str, documents: List[Document], num_explanations: int = 5
                                 ) -> List[Dict]:
        if not documents:
            return []
        if self.tokenizer is None:
            raise ValueError("explanations need the tokenizer the encoders use")
        tok = self.tokenizer
        q_enc = tok([query], add_special_tokens=True, max_length=self.max_query_length, padding="max_length",
                    truncation=True, return_tensors="pt")
        q_tokens = tok.convert_ids_to_tokens(q_enc.input_ids[0].tolist())
        q_mask = q_enc.attention_mask[0].tolist()
        query_embeddings = self._encode_query(query)
        lq = query_embeddings.shape[-2]
        # score rule of the explanations path (:495-501) as a 0/1 weight per query token
        weight = torch.tensor([[1.0 if (m == 1 and t not in ("[CLS]", "[SEP]")) else 0.0
                                for m, t in zip(q_mask[:lq], q_tokens[:lq])]], dtype=torch.float32)
        doc_embeddings_list = self._encode_documents_batched(documents)
        # one rs_maxsim call: scores, the arg-max document token per query token (:492) and the maximum itself — the
        # "similarity" of an explanation (:493-501) — all out of the kernel's TMEM epilogue
        scores, argmax, tokmax = self._maxsim_device(query_embeddings, doc_embeddings_list, q_weight=weight,
                                                     want_argmax=True, want_tokmax=True)
        argmax_h, tokmax_h = argmax[0].tolist(), tokmax[0].tolist()
        results = []
        for j, doc in enumerate(documents):
            d_enc = tok([doc.page_content], add_special_tokens=True, max_length=self.max_doc_length,
                        padding="max_length", truncation=True, return_tensors="pt")
            d_tokens = tok.convert_ids_to_tokens(d_enc.input_ids[0].tolist())
            d_mask = d_enc.attention_mask[0].tolist()
            idx, sims = argmax_h[j], tokmax_h[j]
            explanations = []
            for qidx in range(min(lq, len(q_tokens))):
                if (q_mask[qidx] == 0 or q_tokens[qidx] in ("[PAD]", "[CLS]", "[SEP]", "[UNK]")
                        or q_tokens[qidx].startswith("##")):
                    continue
                didx = idx[qidx]
                if didx < len(d_tokens) and d_mask[didx] == 1:
                    matched = d_tokens[didx]
                    if matched not in ("[PAD]", "[CLS]", "[SEP]", "[UNK]"):
                        ctx_tokens = [t for t in d_tokens[max(0, didx - 2): min(len(d_tokens), didx + 3)]
                                      if t not in ("[PAD]", "[CLS]", "[SEP]", "[UNK]")]
                        context = ""
                        for t in ctx_tokens:
                            context = context[:-1] + t[2:] if t.startswith("##") else context + t + " "
                        explanations.append({"query_token": q_tokens[qidx], "doc_token": matched,
                                             "context": context.strip(), "similarity": sims[qidx]})
            explanations.sort(key=lambda x: x["similarity"], reverse=True)
            results.append({"document": doc, "score": float(scores[0, j]), "explanations": explanations[:num_explanations]})
        results.sort(key=lambda x: x["score"], reverse=True)
        return results
