"""Where does one tcgen05 batch pass over the corpus beat a loop of single-query scans?  N x 1024 bf16, k = 10 / 100.
usage: batch_crossover.py [rows ...]   (default 1e6)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
d = 1024
ns = [int(float(x)) for x in sys.argv[1:]] or [1_000_000]
g = torch.Generator(device=dev).manual_seed(4)
def timed(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for n, k in [(n, k) for n in ns for k in (10, 40, 100)]:
    c = torch.randn(n, d, generator=g, device=dev); c = (c / c.norm(dim=1, keepdim=True)).bfloat16()
    for nq in (2, 3, 4, 8, 16, 32, 64, 128, 256):
        q = torch.randn(nq, d, generator=g, device=dev); q = (q / q.norm(dim=1, keepdim=True)).bfloat16()
        out = {}
        for impl, nm in ((_ffi.RS_DENSE_SCAN, "scan loop"), (_ffi.RS_DENSE_TCGEN05, "tcgen05 batch")):
            eng.set_dense_impl(impl)
            try:
                out[nm] = timed(lambda: eng.dense_topk(c, q, k))
                res = eng.dense_topk(c, q, k)
                out[nm + " ids"] = res[1]
            except Exception as e:
                out[nm] = float("nan")
            finally:
                eng.set_dense_impl(_ffi.RS_DENSE_AUTO)
        same = (out["scan loop ids"] == out["tcgen05 batch ids"]).float().mean().item() if "tcgen05 batch ids" in out else -1
        print(f"n={n} k={k} nq={nq}: scan loop {out['scan loop']:.3f} ms, tcgen05 batch {out['tcgen05 batch']:.3f} ms, ids equal {same:.4f}", flush=True)
