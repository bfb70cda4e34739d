"""A/B of the scan's grab size (RS_SCAN_GRAB_MAX / RS_SCAN_UNIT_WORDS are read once per process): prints one line."""
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
NQ = 16
q = torch.randn(NQ, d, generator=g, device=dev, dtype=torch.float16)

def timed(fn, iters=12, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / NQ * 1e3

out = [f"gmax={os.environ.get('RS_SCAN_GRAB_MAX', '-')} umin={os.environ.get('RS_SCAN_UNIT_WORDS', '-')}"]
for n in (125_000, 1_000_000, 4_000_000):
    us = timed(lambda: eng.dense_topk(c[:n], q, 10))
    out.append(f"n={n}: {us:.1f}us {n * 2048 / us / 1e3:.0f}GB/s")
n = 1_000_000
for p in (0.5, 0.1, 0.01):
    bits = np.random.default_rng(3).random(n) < p
    m = torch.from_numpy(pack_bits(bits)).to(dev)
    us = timed(lambda: eng.dense_topk(c[:n], q, 10, mask=m))
    out.append(f"p={p}: {us:.1f}us")
print(" | ".join(out), flush=True)
