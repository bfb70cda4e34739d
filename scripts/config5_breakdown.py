"""One-GPU breakdown of a config-5 query: 12.5M x 1024 fp16 shard, top-k1 scan (k1 in 10/100/1000), then MaxSim over the
1000 winners and the rerank tail.  CUDA events per stage, averaged over 20 queries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
n, D, DT, LQ, LD, P = 12_500_000, 1024, 128, 32, 300, 125_000
corpus = torch.empty(n, D, dtype=torch.float16, device=dev)
g = torch.Generator(device=dev).manual_seed(100)
for a in range(0, n, 500_000):
    blk = torch.randn(500_000, D, generator=g, device=dev)
    corpus[a:a + 500_000] = (blk / blk.norm(dim=1, keepdim=True)).half()
del blk
pool = torch.randn(P * LD, DT, generator=g, device=dev).bfloat16()
pool_off = (torch.arange(P + 1, dtype=torch.int64) * LD).to(torch.int32).to(dev)
nq = 20
q = torch.randn(nq, D, generator=g, device=dev); q = (q / q.norm(dim=1, keepdim=True)).half()
qtok = torch.randn(nq, LQ, DT, generator=g, device=dev).bfloat16()

def timed(fn, reps=nq):
    fn(0); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for j in range(reps): fn(j)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

floor = n * D * 2 / 6545.9e6
print(f"scan floor at the measured HBM peak: {floor:.3f} ms")
for k1 in (10, 100, 1000):
    ms = timed(lambda j: eng.dense_topk(corpus, q[j:j + 1], k1))
    print(f"stage 1 scan k1={k1}: {ms:.3f} ms ({n * D * 2 / ms / 1e6:.0f} GB/s)")
s1, ids = eng.dense_topk(corpus, q[0:1], 1000)
cand = (ids % P).to(torch.int32)
ms = timed(lambda j: eng.maxsim(qtok[j:j + 1], pool, pool_off, cand=cand))
print(f"stage 2 MaxSim over 1000 candidates: {ms * 1e3:.1f} us")
sc = eng.maxsim(qtok[0:1], pool, pool_off, cand=cand)
ms = timed(lambda j: eng.rerank_postprocess(sc, None, 10))
print(f"rerank tail (stable order of 1000, top-10): {ms * 1e3:.1f} us")
