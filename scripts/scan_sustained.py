"""Sustained single-query scan: ~2.5 s of back-to-back 64-query calls (1M x 1024 fp16, k=10) with nvidia-smi clock /
power sampling.  The board reaches its power cap after ~0.5 s and lowers the SM clock; this shows what the kernel does
then.  Knobs come from the environment (RS_SCAN_TILE_BYTES, RS_SCAN_STAGES, ...)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
n, d, nq = 1_000_000, 1024, 64
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(n, d, generator=g, device=dev, dtype=torch.float16)
q = torch.randn(nq, d, generator=g, device=dev, dtype=torch.float16)
for _ in range(2): eng.dense_topk(c, q, 10)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.PIPE, text=True)
res = []
for blk in range(5):   # 5 blocks of 25 calls = 8000 queries ~ 2.4 s
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(25): eng.dense_topk(c, q, 10)
    b.record(); torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / (25 * nq) * 1e3)
p.terminate(); out = p.communicate()[0]
rows = [l.split(",") for l in out.strip().splitlines() if l.count(",") == 1]
knobs = {k: v for k, v in os.environ.items() if k.startswith("RS_SCAN")}
print(f"{knobs}: us/query per 0.5 s block {[round(x, 1) for x in res]}; clocks {[int(float(r[0])) for r in rows][::3]}; "
      f"power {[int(float(r[1])) for r in rows][::3]}", flush=True)
