"""Batches of <= 128 queries, one tcgen05 pass over 1M x 1024 bf16 rows: ms per batch by batch size and k.
Run as it is and with RS_DENSE_NO_SMALL=1 (3-stage ring with two query-tile slots per stage) for the A/B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
n, d = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000, 1024
g = torch.Generator(device=dev).manual_seed(4)
c = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 1_000_000):
    c[lo:lo + 1_000_000] = torch.randn(min(1_000_000, n - lo), d, generator=g, device=dev).bfloat16()
for k in (10, 100, 1000):
    for nq in (2, 16, 64, 128):
        q = torch.randn(nq, d, generator=g, device=dev).bfloat16()
        for _ in range(3): eng.dense_topk(c, q, k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.dense_topk(c, q, k)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"n={n} k={k:<5} nq={nq:<4} {ms:7.3f} ms/batch = {n*d*2/ms/1e6/6545.9:.3f} of HBM", flush=True)
