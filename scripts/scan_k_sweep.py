"""Single-query scan over 1M x 1024 fp16 by k (the cross-CTA merge is a tournament up to k = 32 and a streamed merge
above): us per lone query, and parity of every result against torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
g = torch.Generator(device=dev).manual_seed(1)
for n in (1_000_000, 125_000):
    c = torch.randn(n, 1024, generator=g, device=dev, dtype=torch.float16)
    q = torch.randn(1024, generator=g, device=dev, dtype=torch.float16)
    ref = (c.float() @ q.float()) / q.float().norm()
    for k in (10, 32, 33, 100, 500, 1000, 2048):
        s, i = eng.dense_topk(c, q, k)
        rs, ri = torch.topk(ref, k)
        # neighbours closer than the accumulation-order noise may swap: compare the score lists, and every returned
        # id against its own reference score
        ok = bool(torch.allclose(rs, s[0], rtol=1e-3, atol=1e-6) and torch.allclose(ref[i[0]], s[0], rtol=1e-3, atol=1e-6)
                  and i[0].unique().numel() == k)
        for _ in range(5): eng.dense_topk(c, q, k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30): eng.dense_topk(c, q, k)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 30 * 1e3
        print(f"n={n:<8} k={k:<5} {us:7.1f} us/query = {n*2048/us/1e3/6545.9:.3f} of HBM  parity {ok}", flush=True)
