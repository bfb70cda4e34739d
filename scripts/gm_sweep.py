"""Batched dense, <= 128 queries x 1M x 1024 bf16: ms per batch for k = 10 / 40 / 100 / 128 (cross-range bound groups)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
n, d = 1_000_000, 1024
g = torch.Generator(device=dev).manual_seed(4)
c = torch.randn(n, d, generator=g, device=dev).bfloat16()
for k in (10, 40, 100, 128):
    row = f"k={k:<4}"
    for nq in (4, 32, 128, 512):
        q = torch.randn(nq, d, generator=g, device=dev).bfloat16()
        for _ in range(3): eng.dense_topk(c, q, k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.dense_topk(c, q, k)
        b.record(); torch.cuda.synchronize()
        row += f"  nq={nq}: {a.elapsed_time(b) / 10:6.3f} ms"
    print(row, flush=True)
