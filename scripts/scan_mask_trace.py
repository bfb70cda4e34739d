"""Filtered single-query scan (1M x 1024 fp16, Bernoulli(p) mask): us per query for 1 / 16 / 64 queries per call (one
launch per query, chained with PDL inside a call) and the phase breakdown of the last launch (rs_set_scan_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.filters import pack_bits

eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
sms = torch.cuda.get_device_properties(0).multi_processor_count
n, d, k = 1_000_000, 1024, 10
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(n, d, generator=g, device=dev, dtype=torch.float16)
trace = torch.zeros(8, sms, 8, dtype=torch.int64, device=dev)
NAMES = ["entry", "bar_init", "prologue", "first_tile", "stream", "compact", "publish", "merge"]
for p in (1.0, 0.5, 0.25, 0.1, 0.03, 0.01):
    bits = np.random.default_rng(3).random(n) < p
    m = torch.from_numpy(pack_bits(bits)).to(dev)
    byt = int(bits.sum()) * d * 2 + n // 8
    for nq in (1, 16, 64):
        q = torch.randn(nq, d, generator=g, device=dev, dtype=torch.float16)
        os_, oi_ = torch.empty(nq, k, device=dev), torch.empty(nq, k, dtype=torch.int64, device=dev)
        eng.set_scan_trace(None)
        for _ in range(3):
            eng.dense_topk(c, q, k, mask=m, out_scores=os_, out_ids=oi_)
        torch.cuda.synchronize()
        reps = max(1, 64 // nq)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.dense_topk(c, q, k, mask=m, out_scores=os_, out_ids=oi_)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nq)
        trace.zero_(); eng.set_scan_trace(trace)
        eng.dense_topk(c, q, k, mask=m, out_scores=os_, out_ids=oi_); torch.cuda.synchronize()
        t = trace.cpu()
        used = [i for i in range(8) if t[i, :, 0].max() > 0]
        last = max(used, key=lambda i: int(t[i, :, 0].max()))
        tt = t[last].double()
        grid = int((tt[:, 0] > 0).sum()); tt = tt[:grid]; t0 = tt[:, 0].min()
        line = f"p={p:<5} nq={nq:<3} {us:7.1f} us/q = {byt/us/1e3/6545.9:.3f} of HBM |"
        for i in range(1, 7):
            dlt = (tt[:, i] - tt[:, i - 1]) / 1e3
            line += f" {NAMES[i]} {dlt.median():.1f}/{dlt.max():.1f}"
        mm = tt[:, 7].max(); last_pub = tt[:, 6].max()
        line += f" | publish@{(last_pub-t0)/1e3:.1f} merge {(mm-last_pub)/1e3:.1f} total {(mm-t0)/1e3:.1f}"
        print(line, flush=True)
eng.set_scan_trace(None)
