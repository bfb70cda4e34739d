"""Small driver for ncu: a few launches of each hot kernel at BASELINE sizes.

    python scripts/profile_kernels.py scan      # dense_scan_kernel, 1M x 1024 fp16, top-10
    python scripts/profile_kernels.py scan_p01  # same with a Bernoulli(0.1) filter (gather tiles)
    python scripts/profile_kernels.py maxsim    # maxsim_tc5_kernel, config 4a
    python scripts/profile_kernels.py maxsim_mma
    python scripts/profile_kernels.py maxsim_cand # maxsim_cand_tc5_kernel, config 4b
    python scripts/profile_kernels.py dense_batch
    python scripts/profile_kernels.py dense_batch_c5   # 16 queries x 12.5M rows, k = 1000 (long lists)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import automative_rag_b200 as rag  # noqa: E402
from automative_rag_b200 import _ffi  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "scan"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = rag.get_engine(0)
dev = eng.device
if what == "scan":
    n, d = 1_000_000, 1024
    g = torch.Generator(device=dev).manual_seed(1)
    c = torch.randn(n, d, generator=g, device=dev, dtype=torch.float16)
    q = torch.randn(d, generator=g, device=dev, dtype=torch.float16)
    mask = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device=dev)
    eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
    for _ in range(iters):
        eng.dense_topk(c, q, 10, mask=mask)
elif what == "scan_p01":  # sparse filter: gather tiles, four rows per TMA tile::gather4
    import numpy as np
    from automative_rag_b200.filters import pack_bits
    n, d = 1_000_000, 1024
    g = torch.Generator(device=dev).manual_seed(1)
    c = torch.randn(n, d, generator=g, device=dev, dtype=torch.float16)
    q = torch.randn(d, generator=g, device=dev, dtype=torch.float16)
    mask = torch.from_numpy(pack_bits(np.random.default_rng(3).random(n) < 0.1)).to(dev)
    eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
    for _ in range(iters):
        eng.dense_topk(c, q, 10, mask=mask)
elif what in ("maxsim", "maxsim_mma"):
    nq, lq, d, nd, ld = 256, 32, 128, 1000, 300
    g = torch.Generator(device=dev).manual_seed(6)
    q = torch.randn(nq, lq, d, generator=g, device=dev).bfloat16()
    toks = torch.randn(nd * ld, d, generator=g, device=dev).bfloat16()
    off = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
    eng.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05 if what == "maxsim" else _ffi.RS_MAXSIM_MMA)
    for _ in range(iters):
        eng.maxsim(q, toks, off)
elif what == "maxsim_cand":
    nq, lq, d, pool, ld, nc = 256, 32, 128, 20_000, 300, 1000
    g = torch.Generator(device=dev).manual_seed(8)
    q = torch.randn(nq, lq, d, generator=g, device=dev).bfloat16()
    toks = torch.randn(pool * ld, d, generator=g, device=dev).bfloat16()
    off = (torch.arange(pool + 1, dtype=torch.int32) * ld).to(dev)
    cand = torch.randint(0, pool, (nq, nc), generator=g, device=dev, dtype=torch.int32)
    eng.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05_CAND)
    for _ in range(iters):
        eng.maxsim(q, toks, off, cand=cand)
elif what == "dense_batch_c5":  # config 5's stage 1 with the step's 16 queries batched: 12.5M x 1024 fp16, k1 = 1000
    n, d, nq, k = 12_500_000, 1024, 16, 1000
    g = torch.Generator(device=dev).manual_seed(100)
    c = torch.empty(n, d, device=dev, dtype=torch.float16)
    for lo in range(0, n, 500_000):
        blk = torch.randn(min(500_000, n - lo), d, generator=g, device=dev)
        c[lo:lo + 500_000] = (blk / blk.norm(dim=1, keepdim=True)).half()
    q = torch.randn(nq, d, generator=g, device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).half()
    eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
    for _ in range(iters):
        eng.dense_topk(c, q, k)
    assert eng.last_dense_redo == 0
elif what in ("dense_batch", "dense_batch_10m"):  # dense_batch_10m = BASELINE config 3 at its full size
    n, d, nq, k = (2_000_000 if what == "dense_batch" else 10_000_000), 1024, 1024, 100
    g = torch.Generator(device=dev).manual_seed(4)
    c = torch.empty(n, d, device=dev, dtype=torch.bfloat16)
    for lo in range(0, n, 1_000_000):
        c[lo:lo + 1_000_000] = torch.randn(min(1_000_000, n - lo), d, generator=g, device=dev).bfloat16()
    q = torch.randn(nq, d, generator=g, device=dev, dtype=torch.bfloat16)
    eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
    for _ in range(iters):
        eng.dense_topk(c, q, k)
torch.cuda.synchronize()
print("ok", what, "launches", eng.launch_count)
