"""MaxSim config 4a timing for one setting of the RS_MAXSIM_* knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
nq, lq, d, nd, ld = 256, 32, 128, 1000, 300
g = torch.Generator(device=dev).manual_seed(6)
q = torch.randn(nq, lq, d, generator=g, device=dev).bfloat16()
toks = torch.randn(nd * ld, d, generator=g, device=dev).bfloat16()
off = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
eng.set_maxsim_impl(_ffi.RS_MAXSIM_TCGEN05)
for _ in range(5): eng.maxsim(q, toks, off)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): eng.maxsim(q, toks, off)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
print({k: v for k, v in os.environ.items() if k.startswith("RS_MAXSIM")}, f"{ms*1e3:.1f} us/batch, {2.0*nq*lq*nd*ld*d/ms/1e9:.0f} TFLOP/s")
