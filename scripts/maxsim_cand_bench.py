"""BASELINE config 4b (per-query candidates) and config 1 (one query vs 100 docs): candidate tcgen05 kernel vs mma.sync."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi

eng = rag.get_engine(0); dev = eng.device
d, ld = 128, 300

def timed(fn, iters, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def run(name, nq, lq, pool_docs, nc, ld, iters, shared=False):
    g = torch.Generator(device=dev).manual_seed(8)
    q = torch.randn(nq, lq, d, generator=g, device=dev).bfloat16()
    toks = torch.randn(pool_docs * ld, d, generator=g, device=dev).bfloat16()
    off = (torch.arange(pool_docs + 1, dtype=torch.int32) * ld).to(dev)
    cand = None if shared else torch.randint(0, pool_docs, (nq, nc), generator=g, device=dev, dtype=torch.int32)
    n_out = pool_docs if shared else nc
    res = {}
    for impl, nm in ((_ffi.RS_MAXSIM_TCGEN05_CAND, "tcgen05_cand"), (_ffi.RS_MAXSIM_MMA, "mma.sync")):
        eng.set_maxsim_impl(impl)
        try:
            ms = timed(lambda: eng.maxsim(q, toks, off, cand=cand), iters)
            res[nm] = eng.maxsim(q, toks, off, cand=cand)
        finally:
            eng.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
        nbytes = float(nq) * n_out * ld * d * 2
        print(f"{name} {nm}: {ms*1e3:.1f} us, {nbytes/ms/1e6:.0f} GB/s of token bytes, {2*nq*lq*n_out*ld*d/ms/1e9:.1f} TFLOP/s", flush=True)
    a, b = res["tcgen05_cand"], res["mma.sync"]
    print(f"{name} max rel diff {((a-b).abs()/(b.abs()+1e-3)).max().item():.2e}", flush=True)

run("config4b 256q x 1000 of 20000 docs x 300 tok", 256, 32, 20_000, 1000, 300, 5)
run("config5-stage2 1q x 1000 of 20000 docs x 300 tok", 1, 32, 20_000, 1000, 300, 20)
run("config1 1q x 100 shared docs x 180 tok", 1, 32, 100, 0, 180, 50, shared=True)
run("deployed 1q x 40 shared docs x 256 tok", 1, 32, 40, 0, 256, 50, shared=True)
