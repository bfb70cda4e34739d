"""BASELINE config 4a (256 queries x 32 tokens vs 1000 shared candidates x 300 tokens, d = 128, bf16) on the
shared-candidate tcgen05 kernel: parity against the general mma.sync kernel (all 256k scores) and a CPU oracle sample,
run-to-run determinism, timing (short burst and a sustained loop), plus lq = 48 / 100 shapes that sum across warps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from oracle import maxsim as omaxsim

eng = rag.get_engine(0); dev = eng.device
nq, lq, d, nd, ld = 256, 32, 128, 1000, 300
g = torch.Generator(device=dev).manual_seed(6)
q = torch.randn(nq, lq, d, generator=g, device=dev).bfloat16()
toks = torch.randn(nd * ld, d, generator=torch.Generator(device=dev).manual_seed(7), device=dev).bfloat16()
off = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
flops = 2.0 * nq * lq * nd * ld * d


def timed(fn, iters, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


eng.set_maxsim_impl(_ffi.RS_MAXSIM_MMA)
ref = eng.maxsim(q, toks, off)
eng.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05)
got = eng.maxsim(q, toks, off)
again = eng.maxsim(q, toks, off)
torch.cuda.synchronize()
err = ((got - ref).abs() / ref.abs().clamp_min(1e-6)).max().item()
print(f"4a: max rel diff vs mma.sync {err:.2e}; bitwise repeatable {torch.equal(got, again)}")
sel = [0, 17, 255]
want = omaxsim.maxsim_scores_packed(q[sel].cpu(), None, toks.cpu(), off.cpu().numpy())
print(f"4a: max rel diff vs CPU oracle (3 queries x 1000 docs) {np.abs(got[sel].cpu().numpy() - want).max() / np.abs(want).max():.2e}")
ms = timed(lambda: eng.maxsim(q, toks, off), 20)
print(f"4a burst (20 launches): {ms*1e3:.1f} us/batch  {flops/ms/1e9:.0f} TFLOP/s")
ms = timed(lambda: eng.maxsim(q, toks, off), 2000)
print(f"4a sustained (2000 launches): {ms*1e3:.1f} us/batch  {flops/ms/1e9:.0f} TFLOP/s")
# ragged documents and query lengths that span 2 / 4 warps, odd query counts (half-empty pair tiles)
for (nq2, lq2, nd2, d2) in ((37, 48, 211, 128), (9, 100, 57, 64), (300, 32, 3, 128), (5, 32, 500, 64), (129, 17, 40, 128)):
    gg = torch.Generator().manual_seed(nq2 * 1000 + lq2)
    lens = torch.randint(1, 400, (nd2,), generator=gg)
    o2 = torch.zeros(nd2 + 1, dtype=torch.int32); o2[1:] = lens.cumsum(0).to(torch.int32)
    t2 = torch.randn(int(o2[-1]), d2, generator=gg).bfloat16()
    q2 = torch.randn(nq2, lq2, d2, generator=gg).bfloat16()
    w2 = torch.rand(nq2, lq2, generator=gg)
    eng.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05)
    a = eng.maxsim(q2.to(dev), t2.to(dev), o2.to(dev), q_weight=w2.to(dev))
    b = eng.maxsim(q2.to(dev), t2.to(dev), o2.to(dev), q_weight=w2.to(dev))
    want = omaxsim.maxsim_scores_packed(q2, w2, t2, o2.numpy())
    e = np.abs(a.cpu().numpy() - want).max() / np.abs(want).max()
    print(f"nq {nq2} lq {lq2} nd {nd2} d {d2}: max rel err vs oracle {e:.2e}, repeatable {torch.equal(a, b)}, impl {eng.last_maxsim_impl}")
    assert e < 1e-3 and torch.equal(a, b)
assert err < 1e-3
