"""Per-request latency of the host entry point (rs_dense_topk_host): H2D query -> scan -> results in mapped host memory
-> stream synchronise, next to the device entry point + torch.cuda.synchronize().
(An experiment that replaced the synchronise by polling a host-mapped completion flag set by the merging CTA measured
the same latency to within 1 us at every size — profiles/r01_host_call_latency.txt — and was dropped.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import automative_rag_b200 as rag
eng = rag.get_engine(0); dev = eng.device
d, k = 1024, 10
g = torch.Generator(device=dev).manual_seed(1)
c = torch.randn(1_000_000, d, generator=g, device=dev, dtype=torch.float16)
qh = torch.randn(64, d, dtype=torch.float16).pin_memory()
out_s = torch.empty(1, k, dtype=torch.float32).pin_memory()
out_i = torch.empty(1, k, dtype=torch.int64).pin_memory()
mode = "stream-sync"
for n in (4_000, 125_000, 1_000_000):
    cn = c[:n]
    for j in range(64): eng.dense_topk_host(cn, qh[j], k, out_scores=out_s, out_ids=out_i)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        for j in range(64): eng.dense_topk_host(cn, qh[j], k, out_scores=out_s, out_ids=out_i)
    dt = (time.perf_counter() - t0) / (64 * reps)
    # same queries through the device entry, one at a time with a synchronise (kernel + launch only)
    qd = qh.to(dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for j in range(64):
            eng.dense_topk(cn, qd[j], k); torch.cuda.synchronize()
    dt_dev = (time.perf_counter() - t0) / (64 * reps)
    s, i = eng.dense_topk(cn, qd[63], k)
    ok = torch.equal(i.cpu(), out_i) and torch.allclose(s.cpu(), out_s)
    print(f"{mode}: n={n}: host call {dt*1e6:.1f} us/query, device call + synchronize {dt_dev*1e6:.1f} us/query, results equal {ok}", flush=True)
