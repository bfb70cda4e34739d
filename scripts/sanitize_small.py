"""One small launch of every kernel (for compute-sanitizer: memcheck / racecheck / synccheck, one tool per run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.filters import pack_bits
eng = rag.get_engine(0); dev = eng.device
g = torch.Generator().manual_seed(0)
# dense scan: contiguous + gather tiles, ragged tail, k small and large
c = torch.randn(5003, 1024, generator=g).half().to(dev); q = torch.randn(2, 1024, generator=g).half().to(dev)
bits = np.random.default_rng(0).random(5003) < 0.4
m = torch.from_numpy(pack_bits(bits)).to(dev)
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
for k in (10, 300):
    eng.dense_topk(c, q, k); eng.dense_topk(c, q, k, mask=m)
s, i = eng.dense_topk(c, q[:1], 10, mask=m)
# batched dense (tcgen05)
c2 = torch.randn(2048, 128, generator=g).bfloat16().to(dev); q2 = torch.randn(64, 128, generator=g).bfloat16().to(dev)
eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05); eng.dense_topk(c2, q2, 20); eng.set_dense_impl(_ffi.RS_DENSE_AUTO)
# merge / postprocess / filter mask
sc = torch.randn(3, 2, 16, generator=g).sort(dim=2, descending=True).values.to(dev); ids = torch.arange(96).view(3, 2, 16).to(dev)
eng.topk_merge(sc, ids, 16)
eng.rerank_postprocess(torch.randn(2, 50, generator=g).to(dev), torch.randn(2, 50, generator=g).to(dev), 8)
col = torch.randint(0, 4, (5003,), generator=g, dtype=torch.int32).to(dev)
eng.filter_mask([col], [[1, 3]], 5003)
# MaxSim: tcgen05, mma.sync (+argmax, candidate lists), fp32 SIMT
qe = torch.randn(8, 32, 128, generator=g).bfloat16(); docs = [torch.randn(n, 128, generator=g).bfloat16() for n in (300, 1, 33, 257, 64)]
toks, off = rag.pack_documents(docs, dev, torch.bfloat16)
eng.maxsim(qe.to(dev), toks, off)
eng.set_maxsim_impl(_ffi.RS_MAXSIM_MMA); eng.maxsim(qe.to(dev), toks, off, want_argmax=True)
eng.maxsim(qe.to(dev), toks, off, cand=torch.tensor([[0, 4, 2]] * 8, dtype=torch.int32).to(dev)); eng.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
toks32, off32 = rag.pack_documents([d.float() for d in docs], dev, torch.float32)
eng.maxsim(qe[:1].float().to(dev), toks32, off32)
torch.cuda.synchronize()
print("sanitize_small ok; launches", eng.launch_count, "top ids", i[0, :3].tolist())
