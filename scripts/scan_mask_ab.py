"""Single-query scan over 1M x 1024 fp16 with Bernoulli(p) filters: us per query and fraction of the HBM peak on the
passing bytes, plus parity of every result against torch on the same device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.filters import pack_bits
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
n, d, k = 1_000_000, 1024, 10
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(n, d, generator=g, device=dev); c = (c / c.norm(dim=1, keepdim=True)).half()
q = torch.randn(4, d, generator=g, device=dev).half()
for p in (1.0, 0.9, 0.5, 0.25, 0.1, 0.03, 0.01):
    bits = np.random.default_rng(3).random(n) < p
    m = torch.from_numpy(pack_bits(bits)).to(dev)
    ok = True
    for j in range(4):
        s, i = eng.dense_topk(c, q[j], k, mask=m)
        ref = (c.float() @ q[j].float()) / q[j].float().norm()
        ref = torch.where(torch.from_numpy(bits).to(dev), ref, torch.full_like(ref, float("-inf")))
        rs, ri = torch.topk(ref, k)
        ok &= bool(torch.equal(ri, i[0]) and torch.allclose(rs, s[0], rtol=1e-3, atol=1e-6))
    for _ in range(5): eng.dense_topk(c, q[0], k, mask=m)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): eng.dense_topk(c, q[0], k, mask=m)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 50 * 1e3
    byt = int(bits.sum()) * d * 2 + n // 8
    print(f"p={p:<5} {us:7.1f} us/query  {byt/us/1e3:7.0f} GB/s on passing bytes = {byt/us/1e3/6545.9:.3f} of HBM peak  parity {ok}")
