"""Soak test: thousands of scan / MaxSim / batched launches over random shapes, every result checked against plain
torch on the GPU.  Looks for intermittent protocol bugs (a hang surfaces as a trap after ~1 s, never as a stuck GPU).
    python scripts/soak.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.filters import pack_bits

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
eng = rag.get_engine(0); dev = eng.device
rng = np.random.default_rng(0)
g = torch.Generator(device=dev).manual_seed(0)
big = {dt: torch.randn(400_000, 1024, generator=g, device=dev).to(dt) for dt in (torch.float16, torch.bfloat16)}
t0 = time.time(); launches = 0; cases = 0
while time.time() - t0 < budget:
    dt = (torch.float16, torch.bfloat16)[rng.integers(2)]
    d = int(rng.choice([64, 128, 256, 1024]))
    n = int(rng.choice([1, 7, 33, 1000, 4097, 65_537, 150_000, 390_000]))
    k = int(rng.choice([1, 10, 37, 100, 1000]))
    nq = int(rng.choice([1, 1, 3, 5, 40, 64, 200]))
    c = big[dt][:n, :d].contiguous()
    q = torch.randn(nq, d, generator=g, device=dev).to(dt)
    if rng.random() < 0.15:   # coarse values: thousands of exactly tied scores (first-tile bisection / list cut fall-backs)
        c = (c.float() * 0.7).round().to(dt)
        q = q.float().round().to(dt)
    p = float(rng.choice([1.0, 1.0, 0.9, 0.5, 0.05, 0.0]))
    bits = None if p == 1.0 else (rng.random(n) < p)
    mask = None if bits is None else torch.from_numpy(pack_bits(bits)).to(dev)
    # the batched kernel: k <= 128 directly, k = 1000 through the per-range lists (needs enough corpus ranges)
    batched_ok = nq >= 2 and n >= 256 and (k <= 128 or n >= 65_537)
    impl = _ffi.RS_DENSE_TCGEN05 if (batched_ok and rng.random() < 0.7) else _ffi.RS_DENSE_SCAN
    eng.set_dense_impl(impl)
    s, i = eng.dense_topk(c, q, k, mask=mask, metric=_ffi.RS_METRIC_IP)
    launches += nq if impl == _ffi.RS_DENSE_SCAN else 2
    ref = q.float() @ c.float().T
    if bits is not None:
        ref = torch.where(torch.from_numpy(bits).to(dev)[None, :], ref, torch.full_like(ref, float("-inf")))
    kk = min(k, n)
    rs, ri = torch.topk(ref, kk, dim=1)
    npass = n if bits is None else int(bits.sum())
    nv = min(k, npass)
    assert (i[:, nv:] == -1).all() and torch.isinf(s[:, nv:]).all(), (n, d, k, nq, p, "padding")
    if nv:
        torch.testing.assert_close(s[:, :nv], rs[:, :nv], rtol=1e-3, atol=1e-3)
        got = torch.gather(ref, 1, i[:, :nv].clamp_min(0))
        torch.testing.assert_close(got, s[:, :nv], rtol=1e-3, atol=1e-3)   # every returned id really has that score
    cases += 1
    if cases % 10 == 0:   # MaxSim: shared candidates (both tcgen05 kernels by shape), per-query candidates, mma.sync
        nqm, lq = int(rng.choice([1, 3, 8, 40])), int(rng.choice([32, 20, 70]))
        lens = rng.integers(1, 400, size=int(rng.integers(1, 60))).tolist()
        qe = torch.randn(nqm, lq, 128, generator=g, device=dev).bfloat16()
        docs = [torch.randn(L, 128, generator=g, device=dev).bfloat16() for L in lens]
        toks, off = rag.pack_documents(docs, dev, torch.bfloat16)
        out = eng.maxsim(qe, toks, off)
        w = torch.ones(lq, device=dev); w[0] = 0; w[-1] = 0
        want = torch.stack([((qe.float() @ dd.float().T).max(dim=2).values * w).sum(dim=1) for dd in docs], dim=1)
        torch.testing.assert_close(out, want, rtol=2e-3, atol=2e-2)
        nc = int(rng.integers(1, len(lens) + 1))
        cand = torch.from_numpy(rng.integers(0, len(lens), size=(nqm, nc)).astype(np.int32)).to(dev)
        for impl in (_ffi.RS_MAXSIM_TCGEN05_CAND, _ffi.RS_MAXSIM_MMA):
            eng.set_maxsim_impl(impl)
            outc = eng.maxsim(qe, toks, off, cand=cand)
            eng.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
            torch.testing.assert_close(outc, torch.gather(want, 1, cand.long()), rtol=2e-3, atol=2e-2)
        launches += 3
torch.cuda.synchronize()
print(f"soak ok: {cases} dense cases, ~{launches} launches in {time.time() - t0:.0f} s, engine launch count {eng.launch_count}")
