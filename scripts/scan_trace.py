"""Where does a single-query scan launch spend its time?  Attaches the rs_set_scan_trace buffer and prints, per corpus
size, the phase durations (median / max over CTAs) of the last launch of a back-to-back series."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi

eng = rag.get_engine(0); dev = eng.device
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
sms = torch.cuda.get_device_properties(0).multi_processor_count
d = 1024
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(1_000_000, d, generator=g, device=dev, dtype=torch.float16)
trace = torch.zeros(8, sms, 8, dtype=torch.int64, device=dev)
NAMES = ["entry", "bar_init", "prologue", "first_tile", "stream", "compact", "publish", "merge"]
KS = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (10, 100)
NS = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (4736, 62_500, 125_000, 250_000, 1_000_000)
for k in KS:
    for n in NS:
        for nq in (1, 8):
            q = torch.randn(nq, d, generator=g, device=dev, dtype=torch.float16)
            eng.set_scan_trace(None)
            for _ in range(3):
                eng.dense_topk(c[:n], q, k)
            trace.zero_(); eng.set_scan_trace(trace)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.dense_topk(c[:n], q, k); e1.record(); torch.cuda.synchronize()
            t = trace.cpu()
            used = [i for i in range(8) if t[i, :, 0].max() > 0]
            # last launch = the block with the largest entry stamp
            last = max(used, key=lambda i: int(t[i, :, 0].max()))
            tt = t[last].double()
            grid = int((tt[:, 0] > 0).sum())
            tt = tt[:grid]
            t0 = tt[:, 0].min()
            line = f"k={k} n={n} nq={nq} grid={grid} call={e0.elapsed_time(e1)*1e3/nq:.1f}us/q |"
            for i in range(1, 7):
                dlt = (tt[:, i] - tt[:, i - 1]) / 1e3
                line += f" {NAMES[i]} {dlt.median():.1f}/{dlt.max():.1f}"
            m = tt[:, 7].max()
            last_pub = tt[:, 6].max()
            line += f" | entry spread {(tt[:,0].max()-t0)/1e3:.1f} publish@{(last_pub-t0)/1e3:.1f} merge {(m-last_pub)/1e3:.1f} total {(m-t0)/1e3:.1f}"
            if len(used) > 1:
                prev = t[(last - 1) % 8].double()[:grid]
                line += f" | prev.merge_end-this.entry0 {(prev[:,7].max()-t0)/1e3:.1f}"
            print(line, flush=True)
eng.set_scan_trace(None)
