"""Masked-scan timings (1M x 1024) for one setting of the RS_SCAN_* knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.filters import pack_bits
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
d = 1024; NQ = 16; n = 1_000_000
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(n, d, generator=g, device=dev, dtype=torch.float16)
q = torch.randn(NQ, d, generator=g, device=dev, dtype=torch.float16)
def timed(fn, iters=6, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / NQ * 1e3
out = []
for p in (1.0, 0.9, 0.5, 0.25, 0.1, 0.01):
    bits = np.random.default_rng(3).random(n) < p
    m = torch.from_numpy(pack_bits(bits)).to(dev)
    us = timed(lambda: eng.dense_topk(c, q, 10, mask=m))
    out.append(f"p={p}: {us:.0f}us {bits.sum()*2048/us/1e3:.0f}GB/s")
print({k: v for k, v in os.environ.items() if k.startswith("RS_SCAN")}, " | ".join(out))
