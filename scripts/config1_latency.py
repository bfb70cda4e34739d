"""BASELINE config 1 through the drop-in `_compute_maxsim_scores`: 1 query x 32 tokens vs 100 docs x 180 tokens, d = 128.
Host fp32 tensors (exact-fp32 and fp16 compute) and embeddings already on the GPU; ms per call, best of 3 runs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
dev = rag.get_engine(0).device
g1 = torch.Generator().manual_seed(0)
q1 = torch.randn(1, 32, 128, generator=g1)
d1 = [torch.randn(180, 128, generator=g1) for _ in range(100)]
def timed(rr, q, d, n=200):
    for _ in range(10): rr._compute_maxsim_scores(q, d)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(n): rr._compute_maxsim_scores(q, d)
        best = min(best, (time.perf_counter() - t0) / n)
    return best * 1e3
for name, fp16 in (("host fp32 -> exact fp32", False), ("host fp32 -> fp16 compute", True)):
    rr = rag.B200ColBERTReranker(device=str(dev), use_fp16=fp16, use_bge_reranker=False)
    print(f"{name}: {timed(rr, q1, d1):.3f} ms per call", flush=True)
rr = rag.B200ColBERTReranker(device=str(dev), use_fp16=True, use_bge_reranker=False)
q1d, d1d = q1.to(dev).half(), [t.to(dev).half() for t in d1]
print(f"device fp16 inputs: {timed(rr, q1d, d1d, 500):.3f} ms per call", flush=True)
d1h = [t.half() for t in d1]
print(f"host fp16 inputs: {timed(rr, q1.half(), d1h):.3f} ms per call", flush=True)
from oracle import maxsim as omaxsim
ref = torch.tensor(omaxsim.maxsim_scores(q1, d1), dtype=torch.float64)
rr32 = rag.B200ColBERTReranker(device=str(dev), use_fp16=False, use_bge_reranker=False)
got = torch.tensor(rr32._compute_maxsim_scores(q1, d1), dtype=torch.float64)
got_dev = torch.tensor(rr32._compute_maxsim_scores(q1.to(dev), [t.to(dev) for t in d1]), dtype=torch.float64)
print("max relative diff vs the CPU port: host list", float(((ref - got).abs() / ref.abs()).max()), "device list",
      float(((ref - got_dev).abs() / ref.abs()).max()))
