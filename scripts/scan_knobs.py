"""Time the 1M and 4M scans for one setting of the RS_SCAN_* knobs (read from the environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
d = 1024; NQ = 16
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(4_000_000, d, generator=g, device=dev, dtype=torch.float16)
q = torch.randn(NQ, d, generator=g, device=dev, dtype=torch.float16)
def timed(fn, iters=6, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / NQ * 1e3
knobs = {k: v for k, v in os.environ.items() if k.startswith("RS_SCAN")}
out = []
for n in (250_000, 1_000_000, 4_000_000):
    us = timed(lambda: eng.dense_topk(c[:n], q, 10))
    out.append(f"{n}: {us:.1f} us {n*2048/us/1e3:.0f} GB/s")
print(knobs, " | ".join(out))
