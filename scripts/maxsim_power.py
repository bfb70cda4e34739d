"""Is MaxSim config 4a power/clock bound?  (a) isolated launches with idle gaps, (b) a 2 s back-to-back loop with
nvidia-smi clock / power sampling."""
import os, sys, time, subprocess
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
for _ in range(3): eng.maxsim(q, toks, off)
torch.cuda.synchronize()
iso = []
for _ in range(10):
    time.sleep(0.05)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.maxsim(q, toks, off); b.record(); torch.cuda.synchronize()
    iso.append(a.elapsed_time(b) * 1e3)
print("isolated launches (us):", [round(x) for x in iso])
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
n = 4000
for _ in range(n): eng.maxsim(q, toks, off)
b.record(); torch.cuda.synchronize()
p.terminate(); out = p.communicate()[0]
print(f"back-to-back {n} launches: {a.elapsed_time(b)/n*1e3:.1f} us each")
rows = [l.split(",") for l in out.strip().splitlines() if l.count(",") == 2]
clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
print("sm clock MHz samples:", [int(c) for c in clk][:40])
print("power W samples:", [int(x) for x in pw][:40])
print("power cap active:", sum("Active" in r[2] and "Not" not in r[2] for r in rows), "of", len(rows))
