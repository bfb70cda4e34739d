"""Is the batched dense kernel power/clock bound?  ~3 s of back-to-back config-3 batches (2M rows) with nvidia-smi
clock / power sampling, for the CTA-pair and (RS_DENSE_NO_PAIR=1) single-CTA variants."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
eng = rag.get_engine(0); dev = eng.device
n, d, nq, k = 2_000_000, 1024, 1024, 100
g = torch.Generator(device=dev).manual_seed(4)
c = torch.randn(n, d, generator=g, device=dev).bfloat16()
q = torch.randn(nq, d, generator=g, device=dev).bfloat16()
eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
for _ in range(3): eng.dense_topk(c, q, k)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
iters = 600
for _ in range(iters): eng.dense_topk(c, q, k)
b.record(); torch.cuda.synchronize()
p.terminate(); out = p.communicate()[0]
ms = a.elapsed_time(b) / iters
print(f"pair={'off' if os.environ.get('RS_DENSE_NO_PAIR') else 'on'}: back-to-back {iters} batches: {ms:.2f} ms each, {2.0*nq*n*d/ms/1e9:.0f} TFLOP/s")
rows = [l.split(",") for l in out.strip().splitlines() if l.count(",") == 2]
clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
print("sm clock MHz samples:", [int(x) for x in clk][:60:3])
print("power W samples:", [int(x) for x in pw][:60:3])
print("power cap active:", sum("Active" in r[2] and "Not" not in r[2] for r in rows), "of", len(rows))
