"""A few scan launches per corpus size (for `ncu --metrics gpu__time_duration.sum`: in-kernel time without launch gaps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
d = 1024
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(1_000_000, d, generator=g, device=dev, dtype=torch.float16)
q = torch.randn(4, d, generator=g, device=dev, dtype=torch.float16)
for n in (4736, 62_500, 250_000, 1_000_000):
    eng.dense_topk(c[:n], q, 10)   # 4 launches each
torch.cuda.synchronize()
print("ok")
