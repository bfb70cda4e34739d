"""Same-box A/B of the scan's optional inputs at 1M x 1024 fp16, k=10: metric (ip / cosine + inv_norm) x mask (none / all ones)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi

eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
n, d, NQ = 1_000_000, 1024, 64
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(n, d, generator=g, device=dev, dtype=torch.float16)
q = torch.randn(NQ, d, generator=g, device=dev, dtype=torch.float16)
inv = (1.0 / c.float().norm(dim=1)).contiguous()
ones = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device=dev)

def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / NQ * 1e3

for rep in range(2):
    for name, kw in (("ip", {"metric": _ffi.RS_METRIC_IP}), ("ip+mask", {"metric": _ffi.RS_METRIC_IP, "mask": ones}),
                     ("cos+inv_norm", {"metric": _ffi.RS_METRIC_COSINE, "inv_norm": inv}),
                     ("cos+inv_norm+mask", {"metric": _ffi.RS_METRIC_COSINE, "inv_norm": inv, "mask": ones}),
                     ("cos(no inv_norm)", {"metric": _ffi.RS_METRIC_COSINE})):
        us = timed(lambda: eng.dense_topk(c, q, 10, **kw))
        print(f"{name:20s} {us:7.1f} us/query  {n * 2048 / us / 1e3:7.0f} GB/s", flush=True)
