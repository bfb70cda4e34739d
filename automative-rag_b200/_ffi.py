"""ctypes binding of the C ABI in include/rag_b200.h (librag_b200.so).

PyTorch tensors appear only at this boundary: the wrappers take `tensor.data_ptr()` and sizes and
never copy.  There is NO CPU fallback: if the library is missing or no sm_100 device is visible
the constructors raise.  ctypes drops the GIL for the duration of every foreign call.
"""
from __future__ import annotations

import ctypes as C
import os
from array import array
from typing import List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librag_b200.so")

ABI_VERSION = 2
RS_OK, RS_ERR_INVALID_ARG, RS_ERR_UNSUPPORTED, RS_ERR_CUDA, RS_ERR_NO_DEVICE, RS_ERR_NOMEM, RS_ERR_COMM = 0, -1, -2, -3, -4, -5, -6
RS_COMM_HANDLE_BYTES, RS_COMM_MAX_WORLD = 128, 8
RS_F16, RS_BF16, RS_F32 = 0, 1, 2
RS_METRIC_IP, RS_METRIC_COSINE = 0, 1
RS_MAXSIM_AUTO, RS_MAXSIM_MMA, RS_MAXSIM_TCGEN05, RS_MAXSIM_SIMT, RS_MAXSIM_TCGEN05_CAND = 0, 1, 2, 3, 4
RS_DENSE_AUTO, RS_DENSE_SCAN, RS_DENSE_TCGEN05 = 0, 1, 2

# every symbol include/rag_b200.h declares: (restype, argtypes)
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "rs_abi_version": (C.c_int, []),
    "rs_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "rs_destroy": (C.c_int, [_P]),
    "rs_last_error": (C.c_char_p, [_P]),
    "rs_launch_count": (_I64, [_P]),
    "rs_set_dense_impl": (C.c_int, [_P, C.c_int]),
    "rs_set_maxsim_impl": (C.c_int, [_P, C.c_int]),
    "rs_set_scan_trace": (C.c_int, [_P, _P]),
    "rs_scan_plan": (C.c_int, [_I32, _I32, C.POINTER(_I64)]),
    "rs_scan_plan_chained": (C.c_int, [_I32, _I32, C.POINTER(_I64)]),
    "rs_set_profiling": (C.c_int, [_P, C.c_int]),
    "rs_last_call_stats": (C.c_int, [_P, _P]),
    "rs_last_dense_impl": (C.c_int, [_P]),
    "rs_last_dense_redo": (C.c_int, [_P]),
    "rs_last_maxsim_impl": (C.c_int, [_P]),
    "rs_dense_topk": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _I32, _P, _I32, _P, _I64, _I32, _I64, _P, _P, _P]),
    "rs_dense_topk_host": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _I32, _P, _I32, _P, _P, _I64, _I32, _I64, _P, _P, _P]),
    "rs_topk_merge": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I64, _I64, _P, _P, _P]),
    "rs_maxsim": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _I64, _P, _I32, _P, _I32, _P, _P, _P, _P]),
    "rs_maxsim_list": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _I32, _I32, _P, _P]),
    "rs_rerank_postprocess": (C.c_int, [_P, _P, _P, _I32, _I32, _F, _F, _I32, _P, _P, _P]),
    "rs_filter_mask": (C.c_int, [_P, C.POINTER(_P), _I32, C.POINTER(_I32), C.POINTER(_I32), _P, _I64, _P, _P]),
    "rs_comm_export": (C.c_int, [_P, _I32, _I32, _I64, _P]),
    "rs_comm_open": (C.c_int, [_P, _P]),
    "rs_comm_close": (C.c_int, [_P]),
    "rs_comm_info": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I64)]),
    "rs_allgather_topk": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "rs_allgather": (C.c_int, [_P, _P, _I64, _P, _P]),
    "rs_allreduce_max_f32": (C.c_int, [_P, _P, _I64, _P, _P]),
    "rs_owned_candidates": (C.c_int, [_P, _P, _I32, _I64, _I32, _I32, _I64, _P, _P]),
    "rs_dense_topk_sharded_host": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _I32, _P, _I32, _P, _I64, _I32, _I64, _P, _P, _P]),
}

RS_CALL_NONE, RS_CALL_DENSE_TOPK, RS_CALL_MAXSIM, RS_CALL_TOPK_MERGE, RS_CALL_ALLGATHER_TOPK = 0, 1, 2, 3, 4


class CallStats(C.Structure):
    """rs_call_stats of include/rag_b200.h."""
    _fields_ = [("entry", _I32), ("kernel_family", _I32), ("launches", _I32), ("queries", _I32),
                ("bytes_scanned", _I64), ("flops", C.c_double), ("device_ms", _F), ("merge_ms", _F)]


_lib = None


class EngineError(RuntimeError):
    """A C-ABI call returned a CUDA / device error status."""


def load_library() -> C.CDLL:
    """dlopen librag_b200.so and bind every declared symbol.  Raises if the library is absent —
    the product path never falls back to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.rs_abi_version() != ABI_VERSION:
        raise EngineError(f"ABI version mismatch: library reports {lib.rs_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float16:
        return RS_F16
    if dt == torch.bfloat16:
        return RS_BF16
    if dt == torch.float32:
        return RS_F32
    raise ValueError(f"unsupported embedding dtype {dt}; use float16, bfloat16 or float32")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Engine:
    """One `rs_handle`: bound to one CUDA device, calls serialised by the caller."""

    def __init__(self, device: int | str | torch.device = 0):
        self._lib = load_library()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise EngineError("the retrieval-scoring engine runs on CUDA devices only (no CPU fallback)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
        h = _P()
        rc = self._lib.rs_create(self.device.index, C.byref(h))
        if rc != RS_OK:
            msg = self._lib.rs_last_error(None).decode()
            raise EngineError(f"rs_create failed ({rc}): {msg}")
        self._h = h

    # -- plumbing -----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rs_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        if rc == RS_OK:
            return
        msg = self._lib.rs_last_error(self._h).decode()
        if rc in (RS_ERR_INVALID_ARG, RS_ERR_UNSUPPORTED):  # RS_ERR_COMM and CUDA errors raise EngineError
            raise ValueError(f"{what}: {msg}")
        raise EngineError(f"{what} failed ({rc}): {msg}")

    def _dev(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if t.device != self.device:
            raise ValueError(f"{name} must live on {self.device}, got {t.device}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        return t

    @property
    def launch_count(self) -> int:
        return int(self._lib.rs_launch_count(self._h))

    def set_dense_impl(self, impl: int) -> None:
        self._check(self._lib.rs_set_dense_impl(self._h, impl), "rs_set_dense_impl")

    def set_maxsim_impl(self, impl: int) -> None:
        self._check(self._lib.rs_set_maxsim_impl(self._h, impl), "rs_set_maxsim_impl")

    def set_scan_trace(self, trace: Optional[torch.Tensor]) -> None:
        """Diagnostics: int64 device tensor [8, num_sms, 8] receiving per-CTA phase time stamps (None detaches)."""
        self._check(self._lib.rs_set_scan_trace(self._h, trace.data_ptr() if trace is not None else None), "rs_set_scan_trace")

    def set_profiling(self, on: bool) -> None:
        """Per-call statistics on/off (two CUDA event records per scoring call, no synchronisation)."""
        self._check(self._lib.rs_set_profiling(self._h, 1 if on else 0), "rs_set_profiling")

    def last_call_stats(self) -> dict:
        """Statistics of the most recent profiled call (waits for it): the engine-side counterpart of the reference's
        `search_time_ms` / `docs_per_second` debug fields (src/services/system_service.py:336-372)."""
        st = CallStats()
        self._check(self._lib.rs_last_call_stats(self._h, C.byref(st)), "rs_last_call_stats")
        d = {name: getattr(st, name) for name, _ in CallStats._fields_}
        if st.device_ms > 0:
            d["gb_per_s"] = st.bytes_scanned / st.device_ms / 1e6
            d["tflop_per_s"] = st.flops / st.device_ms / 1e9
        return d

    @property
    def last_dense_impl(self) -> int:
        return int(self._lib.rs_last_dense_impl(self._h))

    @property
    def last_dense_redo(self) -> int:
        """Queries the last batched call with k > 128 re-ran through the exact single-query scan."""
        return int(self._lib.rs_last_dense_redo(self._h))

    @property
    def last_maxsim_impl(self) -> int:
        return int(self._lib.rs_last_maxsim_impl(self._h))

    # -- dense --------------------------------------------------------------------------------
    def dense_topk(self, corpus: torch.Tensor, queries: torch.Tensor, k: int, *, mask: Optional[torch.Tensor] = None,
                   inv_norm: Optional[torch.Tensor] = None, metric: int = RS_METRIC_COSINE, id_base: int = 0,
                   out_scores: Optional[torch.Tensor] = None, out_ids: Optional[torch.Tensor] = None):
        """corpus [n, d] fp16/bf16, queries [nq, d] or [d]; mask int32/uint32 words [ceil(n/32)] or
        [nq, ceil(n/32)].  Returns (scores [nq, k] fp32, ids [nq, k] int64) on the device."""
        self._dev(corpus, "corpus")
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        self._dev(queries, "queries")
        if queries.dtype != corpus.dtype:
            raise ValueError("queries must have the corpus dtype")
        n, d = corpus.shape
        nq = queries.shape[0]
        if queries.shape[1] != d:
            raise ValueError(f"query dim {queries.shape[1]} != corpus dim {d}")
        stride = 0
        if mask is not None:
            self._dev(mask, "mask")
            words = (n + 31) // 32
            if mask.dtype not in (torch.int32, torch.uint32) or mask.shape[-1] != words:
                raise ValueError(f"mask must be int32 words with last dim {words}")
            if mask.dim() == 2:
                if mask.shape[0] != nq:
                    raise ValueError("per-query mask must have nq rows")
                stride = words
        if inv_norm is not None:
            self._dev(inv_norm, "inv_norm")
            if inv_norm.dtype != torch.float32 or inv_norm.numel() != n:
                raise ValueError("inv_norm must be float32 [n]")
        if out_scores is None:
            out_scores = torch.empty(nq, k, dtype=torch.float32, device=self.device)
        if out_ids is None:
            out_ids = torch.empty(nq, k, dtype=torch.int64, device=self.device)
        rc = self._lib.rs_dense_topk(self._h, _ptr(corpus), n, d, dtype_code(corpus.dtype), _ptr(inv_norm), metric,
                                     _ptr(queries), nq, _ptr(mask), stride, k, id_base, _ptr(out_scores),
                                     _ptr(out_ids), _stream_ptr(self.device))
        self._check(rc, "rs_dense_topk")
        return out_scores, out_ids

    def dense_topk_host(self, corpus: torch.Tensor, queries_host: torch.Tensor, k: int, *,
                        mask_host: Optional[torch.Tensor] = None, mask_dev: Optional[torch.Tensor] = None,
                        inv_norm: Optional[torch.Tensor] = None, metric: int = RS_METRIC_COSINE, id_base: int = 0,
                        out_scores: Optional[torch.Tensor] = None, out_ids: Optional[torch.Tensor] = None):
        """Per-request call: HOST queries / mask in, HOST (scores, ids) out; corpus stays on the device."""
        self._dev(corpus, "corpus")
        if queries_host.dim() == 1:
            queries_host = queries_host.unsqueeze(0)
        if queries_host.device.type != "cpu" or not queries_host.is_contiguous():
            raise ValueError("queries_host must be a contiguous CPU tensor")
        if queries_host.dtype != corpus.dtype:
            raise ValueError("queries must have the corpus dtype")
        n, d = corpus.shape
        nq = queries_host.shape[0]
        stride = 0
        words = (n + 31) // 32
        for m, nm in ((mask_host, "mask_host"), (mask_dev, "mask_dev")):
            if m is not None:
                if m.dtype not in (torch.int32, torch.uint32) or m.shape[-1] != words or not m.is_contiguous():
                    raise ValueError(f"{nm} must be contiguous int32 words with last dim {words}")
                if m.dim() == 2:
                    stride = words
        if mask_host is not None and mask_host.device.type != "cpu":
            raise ValueError("mask_host must be a CPU tensor")
        if mask_dev is not None:
            self._dev(mask_dev, "mask_dev")
        if out_scores is None:
            out_scores = torch.empty(nq, k, dtype=torch.float32)
        if out_ids is None:
            out_ids = torch.empty(nq, k, dtype=torch.int64)
        rc = self._lib.rs_dense_topk_host(self._h, _ptr(corpus), n, d, dtype_code(corpus.dtype), _ptr(inv_norm), metric,
                                          _ptr(queries_host), nq, _ptr(mask_host), _ptr(mask_dev), stride, k, id_base,
                                          _ptr(out_scores), _ptr(out_ids), _stream_ptr(self.device))
        self._check(rc, "rs_dense_topk_host")
        return out_scores, out_ids

    def topk_merge(self, scores: torch.Tensor, ids: torch.Tensor, k_out: int):
        """scores/ids [nlists, nq, k_in] -> ([nq, k_out], [nq, k_out])."""
        if scores.device != self.device or ids.device != self.device:
            raise ValueError(f"scores and ids must live on {self.device}")
        if scores.dtype != torch.float32 or ids.dtype != torch.int64 or scores.shape != ids.shape or scores.dim() != 3:
            raise ValueError("scores float32 / ids int64 of identical shape [nlists, nq, k_in] expected")
        nl, nq, k_in = scores.shape
        for t in (scores, ids):  # each list dense, lists may be strided (views into the gathered wire buffer)
            if t.stride(2) != 1 or t.stride(1) != k_in:
                raise ValueError("each [nq, k_in] list must be contiguous")
        out_s = torch.empty(nq, k_out, dtype=torch.float32, device=self.device)
        out_i = torch.empty(nq, k_out, dtype=torch.int64, device=self.device)
        rc = self._lib.rs_topk_merge(self._h, _ptr(scores), _ptr(ids), nl, nq, k_in, k_out, scores.stride(0),
                                     ids.stride(0), _ptr(out_s), _ptr(out_i), _stream_ptr(self.device))
        self._check(rc, "rs_topk_merge")
        return out_s, out_i

    # -- MaxSim -------------------------------------------------------------------------------
    def maxsim(self, q: torch.Tensor, doc_tokens: torch.Tensor, doc_offsets: torch.Tensor, *,
               q_weight: Optional[torch.Tensor] = None, cand: Optional[torch.Tensor] = None,
               want_argmax: bool = False, want_tokmax: bool = False):
        """q [nq, lq, d]; doc_tokens [T, d]; doc_offsets int32 [nd+1]; cand None or int32 [nq, nc].
        Returns scores [nq, nd|nc] fp32; with want_argmax also argmax int32 [nq, nd|nc, lq]; with want_tokmax also
        the per-query-token maxima fp32 [nq, nd|nc, lq] (in that order)."""
        self._dev(q, "q")
        self._dev(doc_tokens, "doc_tokens")
        self._dev(doc_offsets, "doc_offsets")
        if q.dim() != 3 or doc_tokens.dim() != 2 or q.shape[2] != doc_tokens.shape[1]:
            raise ValueError("q must be [nq, lq, d] and doc_tokens [T, d] with the same d")
        if q.dtype != doc_tokens.dtype:
            raise ValueError("q and doc_tokens must share a dtype")
        if doc_offsets.dtype != torch.int32 or doc_offsets.dim() != 1:
            raise ValueError("doc_offsets must be int32 [nd + 1]")
        nq, lq, d = q.shape
        nd = doc_offsets.numel() - 1
        nc = 0
        if cand is not None:
            self._dev(cand, "cand")
            if cand.dtype != torch.int32 or cand.dim() != 2 or cand.shape[0] != nq:
                raise ValueError("cand must be int32 [nq, nc]")
            nc = cand.shape[1]
        if q_weight is not None:
            self._dev(q_weight, "q_weight")
            if q_weight.dtype != torch.float32 or tuple(q_weight.shape) != (nq, lq):
                raise ValueError("q_weight must be float32 [nq, lq]")
        ndo = nc if cand is not None else nd
        out = torch.empty(nq, ndo, dtype=torch.float32, device=self.device)
        arg = torch.empty(nq, ndo, lq, dtype=torch.int32, device=self.device) if want_argmax else None
        tmax = torch.empty(nq, ndo, lq, dtype=torch.float32, device=self.device) if want_tokmax else None
        rc = self._lib.rs_maxsim(self._h, _ptr(q), nq, lq, d, dtype_code(q.dtype), _ptr(q_weight), _ptr(doc_tokens),
                                 doc_tokens.shape[0], _ptr(doc_offsets), nd, _ptr(cand), nc, _ptr(out), _ptr(arg),
                                 _ptr(tmax), _stream_ptr(self.device))
        self._check(rc, "rs_maxsim")
        res = (out,) + ((arg,) if want_argmax else ()) + ((tmax,) if want_tokmax else ())
        return res if len(res) > 1 else out

    def maxsim_list(self, q: torch.Tensor, docs: Sequence[torch.Tensor], compute_dtype: torch.dtype,
                    q_weight: Optional[torch.Tensor] = None, doc_lens: Optional[Sequence[int]] = None) -> List[float]:
        """One query [lq, d] against a list of [Ld_i, d] tensors — all on the host or all on this device, one dtype —
        scored in `compute_dtype`; returns the scores as Python floats.  One C call: staged upload, gather / convert
        launch, rs_maxsim, results through mapped pinned memory (no torch op in between)."""
        nd = len(docs)
        lq, d = q.shape
        on_host = not docs[0].is_cuda
        # the per-document work stays inside C loops (map / array.array): at 100 documents a Python-level loop with
        # ctypes element conversion was most of a 0.13 ms call whose kernel takes ~15 us
        ptrs = array("Q", map(torch.Tensor.data_ptr, docs))
        lens = array("i", doc_lens if doc_lens is not None else [t.shape[0] for t in docs])
        out = array("f", bytes(4 * nd))
        w = None
        if q_weight is not None:
            w = q_weight.detach().to("cpu", torch.float32).contiguous().view(-1)
        rc = self._lib.rs_maxsim_list(self._h, q.data_ptr(), 0 if q.is_cuda else 1, lq, d, dtype_code(q.dtype),
                                      dtype_code(compute_dtype), None if w is None else w.data_ptr(),
                                      ptrs.buffer_info()[0], lens.buffer_info()[0], nd, 1 if on_host else 0,
                                      out.buffer_info()[0], _stream_ptr(self.device))
        self._check(rc, "rs_maxsim_list")
        return out.tolist()

    def rerank_postprocess(self, scores: torch.Tensor, other: Optional[torch.Tensor], top_k: int,
                           w_a: float = 0.8, w_b: float = 0.2):
        """scores [nq, n] (+ other [nq, n]) -> (idx int32 [nq, top_k], final scores fp32 [nq, top_k])."""
        self._dev(scores, "scores")
        if scores.dtype != torch.float32 or scores.dim() != 2:
            raise ValueError("scores must be float32 [nq, n]")
        if other is not None:
            self._dev(other, "other")
            if other.dtype != torch.float32 or other.shape != scores.shape:
                raise ValueError("other must match scores")
        nq, n = scores.shape
        top_k = min(top_k, n)
        idx = torch.empty(nq, top_k, dtype=torch.int32, device=self.device)
        out = torch.empty(nq, top_k, dtype=torch.float32, device=self.device)
        rc = self._lib.rs_rerank_postprocess(self._h, _ptr(scores), _ptr(other), nq, n, w_a, w_b, top_k, _ptr(idx),
                                             _ptr(out), _stream_ptr(self.device))
        self._check(rc, "rs_rerank_postprocess")
        return idx, out

    # -- multi-GPU exchange over peer memory -----------------------------------------------------
    def comm_init(self, group=None, slot_bytes: int = 4 << 20) -> None:
        """Open the peer-memory exchange with the ranks of `group` (one process per GPU): export this rank's wire
        block, exchange the 128-byte handles through torch.distributed, map the peers.  `slot_bytes` bounds one
        rank's contribution per collective (nq * k * 12 bytes for top-k lists, 4 bytes per score)."""
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        blob = C.create_string_buffer(RS_COMM_HANDLE_BYTES)
        self._check(self._lib.rs_comm_export(self._h, world, rank, slot_bytes, blob), "rs_comm_export")
        blobs: list = [None] * world
        dist.all_gather_object(blobs, bytes(blob.raw), group=group)
        allb = C.create_string_buffer(b"".join(blobs), RS_COMM_HANDLE_BYTES * world)
        self._check(self._lib.rs_comm_open(self._h, allb), "rs_comm_open")
        dist.barrier(group)  # every rank has mapped every block before the first collective

    def comm_close(self) -> None:
        self._check(self._lib.rs_comm_close(self._h), "rs_comm_close")

    @property
    def comm_world(self) -> int:
        w = _I32(0)
        self._lib.rs_comm_info(self._h, C.byref(w), None, None)
        return int(w.value)

    def allgather_topk(self, local_scores: torch.Tensor, local_ids: torch.Tensor, k_out: int,
                       out_scores: Optional[torch.Tensor] = None, out_ids: Optional[torch.Tensor] = None):
        """local [nq, k_in] (this rank's top-k, global ids) -> merged global ([nq, k_out], [nq, k_out]) on every rank."""
        self._dev(local_scores, "local_scores")
        self._dev(local_ids, "local_ids")
        if local_scores.dtype != torch.float32 or local_ids.dtype != torch.int64 or local_scores.shape != local_ids.shape \
                or local_scores.dim() != 2:
            raise ValueError("local_scores float32 / local_ids int64 of identical shape [nq, k_in] expected")
        nq, k_in = local_scores.shape
        if out_scores is None:
            out_scores = torch.empty(nq, k_out, dtype=torch.float32, device=self.device)
        if out_ids is None:
            out_ids = torch.empty(nq, k_out, dtype=torch.int64, device=self.device)
        rc = self._lib.rs_allgather_topk(self._h, _ptr(local_scores), _ptr(local_ids), nq, k_in, k_out, _ptr(out_scores),
                                         _ptr(out_ids), _stream_ptr(self.device))
        self._check(rc, "rs_allgather_topk")
        return out_scores, out_ids

    def allgather(self, local: torch.Tensor) -> torch.Tensor:
        """[...] on every rank -> [world, ...] (same shape and dtype on every rank; bytes a multiple of 16)."""
        self._dev(local, "local")
        out = torch.empty((self.comm_world,) + tuple(local.shape), dtype=local.dtype, device=self.device)
        rc = self._lib.rs_allgather(self._h, _ptr(local), local.numel() * local.element_size(), _ptr(out),
                                    _stream_ptr(self.device))
        self._check(rc, "rs_allgather")
        return out

    def owned_candidates(self, cand: torch.Tensor, world: int, rank: int, pool: int = 0) -> torch.Tensor:
        """Global candidate ids (int32 / int64, any shape; negative = padding; modulo `pool` first when pool > 0) ->
        int32 local document indices of THIS rank under round-robin ownership, -1 elsewhere.  One launch."""
        self._dev(cand, "cand")
        if cand.dtype not in (torch.int32, torch.int64) or not cand.is_contiguous():
            raise ValueError("cand must be a contiguous int32 / int64 tensor")
        out = torch.empty(cand.shape, dtype=torch.int32, device=self.device)
        rc = self._lib.rs_owned_candidates(self._h, _ptr(cand), 1 if cand.dtype == torch.int64 else 0, cand.numel(),
                                           world, rank, pool, _ptr(out), _stream_ptr(self.device))
        self._check(rc, "rs_owned_candidates")
        return out

    def allreduce_max(self, local: torch.Tensor) -> torch.Tensor:
        """Element-wise max over ranks of a float32 tensor."""
        self._dev(local, "local")
        if local.dtype != torch.float32:
            raise ValueError("allreduce_max takes float32")
        out = torch.empty_like(local)
        rc = self._lib.rs_allreduce_max_f32(self._h, _ptr(local), local.numel(), _ptr(out), _stream_ptr(self.device))
        self._check(rc, "rs_allreduce_max_f32")
        return out

    def dense_topk_sharded_host(self, corpus: torch.Tensor, queries_host: torch.Tensor, k: int, *,
                                mask_dev: Optional[torch.Tensor] = None, inv_norm: Optional[torch.Tensor] = None,
                                metric: int = RS_METRIC_COSINE, id_base: int = 0,
                                out_scores: Optional[torch.Tensor] = None, out_ids: Optional[torch.Tensor] = None):
        """Per-request call against this rank's row shard: HOST queries in, merged global HOST (scores, ids) out."""
        self._dev(corpus, "corpus")
        if queries_host.dim() == 1:
            queries_host = queries_host.unsqueeze(0)
        if queries_host.device.type != "cpu" or not queries_host.is_contiguous() or queries_host.dtype != corpus.dtype:
            raise ValueError("queries_host must be a contiguous CPU tensor of the corpus dtype")
        n, d = corpus.shape
        nq = queries_host.shape[0]
        stride = 0
        if mask_dev is not None:
            self._dev(mask_dev, "mask_dev")
            stride = (n + 31) // 32 if mask_dev.dim() == 2 else 0
        if out_scores is None:
            out_scores = torch.empty(nq, k, dtype=torch.float32)
        if out_ids is None:
            out_ids = torch.empty(nq, k, dtype=torch.int64)
        rc = self._lib.rs_dense_topk_sharded_host(self._h, _ptr(corpus), n, d, dtype_code(corpus.dtype), _ptr(inv_norm),
                                                  metric, _ptr(queries_host), nq, _ptr(mask_dev), stride, k, id_base,
                                                  _ptr(out_scores), _ptr(out_ids), _stream_ptr(self.device))
        self._check(rc, "rs_dense_topk_sharded_host")
        return out_scores, out_ids

    def filter_mask(self, columns: Sequence[torch.Tensor], value_sets: Sequence[Sequence[int]], n: int,
                    tombstone: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """AND over clauses; clause c: columns[c] (int32 [n], device) in value_sets[c].  Returns int32 words."""
        if len(columns) != len(value_sets):
            raise ValueError("one value set per column")
        words = (n + 31) // 32
        if out is None:
            out = torch.empty(words, dtype=torch.int32, device=self.device)
        ncl = len(columns)
        cols = (_P * max(ncl, 1))()
        for i, c in enumerate(columns):
            self._dev(c, "column")
            if c.dtype != torch.int32 or c.numel() != n:
                raise ValueError("columns must be int32 [n]")
            cols[i] = c.data_ptr()
        flat: List[int] = []
        offs = [0]
        for vs in value_sets:
            flat.extend(int(v) for v in vs)
            offs.append(len(flat))
        vals = (_I32 * max(len(flat), 1))(*flat)
        offa = (_I32 * len(offs))(*offs)
        if tombstone is not None:
            self._dev(tombstone, "tombstone")
        rc = self._lib.rs_filter_mask(self._h, cols, ncl, vals, offa, _ptr(tombstone), n, _ptr(out),
                                      _stream_ptr(self.device))
        self._check(rc, "rs_filter_mask")
        return out


_engines: dict = {}


def get_engine(device: int | str | torch.device = 0) -> Engine:
    """Process-wide engine per device (the reference keeps module-level singletons,
    src/core/background/models.py:16-19)."""
    dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
    idx = dev.index if dev.index is not None else 0
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]
