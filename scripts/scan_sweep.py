"""Scan-kernel sweep: time per query vs corpus rows / k / mask density (CUDA events, 30 launches each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.filters import pack_bits

eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
d = 1024
nmax = 4_000_000
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(nmax, d, generator=g, device=dev, dtype=torch.float16)
NQ = 16  # queries per C-ABI call: one Python call, NQ back-to-back nq=1 scan launches (PDL between them)
q = torch.randn(NQ, d, generator=g, device=dev, dtype=torch.float16)

def timed(fn, iters=8, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / NQ * 1e3  # us per query

print("rows      us/query   GB/s")
res = {}
for n in (62_500, 125_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000):
    us = timed(lambda: eng.dense_topk(c[:n], q, 10))
    res[n] = us
    print(f"{n:9d} {us:9.1f} {n * 2048 / us / 1e3:8.1f}")
slope = (res[4_000_000] - res[1_000_000]) / 3e6
print(f"asymptotic {2048 / slope / 1e3:.1f} GB/s, fixed cost at 1M ~ {res[1_000_000] - slope * 1e6:.1f} us")
n = 1_000_000
for p in (0.9, 0.5, 0.25, 0.1, 0.01):
    bits = np.random.default_rng(3).random(n) < p
    m = torch.from_numpy(pack_bits(bits)).to(dev)
    us = timed(lambda: eng.dense_topk(c[:n], q, 10, mask=m))
    print(f"mask p={p}: {us:.1f} us, {bits.sum() * 2048 / us / 1e3:.1f} GB/s of passing bytes")
blk = (np.arange(n) // 4096) % 2 == 0   # clustered filter: alternate 4096-row blocks
m = torch.from_numpy(pack_bits(blk)).to(dev)
us = timed(lambda: eng.dense_topk(c[:n], q, 10, mask=m))
print(f"mask blocks-of-4096 (50%): {us:.1f} us, {blk.sum() * 2048 / us / 1e3:.1f} GB/s")
for k in (1, 10, 100, 1000, 2048):
    us = timed(lambda: eng.dense_topk(c[:n], q, k))
    print(f"k={k}: {us:.1f} us")
