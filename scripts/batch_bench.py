"""BASELINE config 3: 1024 queries x N x 1024 bf16, top-100, tcgen05 GEMM + fused top-k. Prints ms per batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
d = 1024
g = torch.Generator(device=dev).manual_seed(4)
c = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
step = 500_000
for lo in range(0, n, step):
    m = min(step, n - lo)
    blk = torch.randn(m, d, generator=g, device=dev)
    c[lo: lo + m] = (blk / blk.norm(dim=1, keepdim=True)).bfloat16()
del blk
q = torch.randn(nq, d, generator=torch.Generator(device=dev).manual_seed(5), device=dev)
q = (q / q.norm(dim=1, keepdim=True)).bfloat16()
eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
for _ in range(2): eng.dense_topk(c, q, k)
torch.cuda.synchronize()
iters = 5
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters): eng.dense_topk(c, q, k)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
fl = 2.0 * nq * n * d
print(f"n={n} nq={nq} k={k}: {ms:.2f} ms/batch, {nq/ms*1e3:.0f} q/s, {fl/ms/1e9:.0f} TFLOP/s, corpus {n*d*2/ms/1e6:.0f} GB/s")
