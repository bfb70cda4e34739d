"""Unfiltered single-query scans chained inside one call (64 queries, a launch each): us per query by corpus size and k.
Run once as it is and once with RS_SCAN_NO_CORESIDENT=1 for the A/B of the two-CTAs-per-SM plan."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi

eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
g = torch.Generator(device=dev).manual_seed(1)
for d in (1024, 256):
    c = torch.randn(4_000_000 * 1024 // d // 4, d, generator=g, device=dev, dtype=torch.float16)
    q = torch.randn(64, d, generator=g, device=dev, dtype=torch.float16)
    for k in (10, 100, 200):
        for n in (31_250, 125_000, 250_000, 500_000, 1_000_000):
            n = n * 1024 // d
            if n > c.shape[0]:
                continue
            os_, oi_ = torch.empty(64, k, device=dev), torch.empty(64, k, dtype=torch.int64, device=dev)
            for _ in range(2):
                eng.dense_topk(c[:n], q, k, out_scores=os_, out_ids=oi_)
            torch.cuda.synchronize()
            reps = 3 if n >= 500_000 else 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                eng.dense_topk(c[:n], q, k, out_scores=os_, out_ids=oi_)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * 64)
            print(f"d={d:<5} k={k:<4} n={n:<8} {us:7.1f} us/q  {n*d*2/us/1e3:7.0f} GB/s = {n*d*2/us/1e3/6545.9:.3f} of HBM", flush=True)
