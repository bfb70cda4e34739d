#!/usr/bin/env python
"""bench.py — queries/sec of the dense top-k hot path (BASELINE.json config 2) on N B200s.

Workload (config.workload = "config2"): exact cosine top-10 over a 1M x 1024 fp16 corpus (rows
L2-normalised in fp32 then rounded, seed 1), SINGLE-QUERY searches, metadata-filter bitmask passed
(p = 1.0: every bit set, so the mask words are read and tested but no row is skipped).

A step = one batch of `--queries-per-step` (default 64) independent single-query searches.  Every
query is its own nq=1 scan launch that re-reads its whole shard from HBM (the corpus, 2 GB, is 16x
the 126 MB L2, so nothing is cached between queries).  With N > 1 (torchrun, one process per GPU)
the corpus is row-sharded (strong scaling: 1M rows total), every rank scans its shard for all
queries of the step, ONE NCCL all-gather moves the step's k (id, score) pairs and one merge launch
produces the global top-k on every rank.

  value     queries/s, inputs resident in HBM, CUDA events, max over ranks
  e2e       the same metric through the public host entry point (rs_dense_topk_host at N = 1, the sharded index
            at N > 1): the step's queries start in pinned host memory, its (score, id) pairs end in host memory,
            H2D + D2H + synchronise inside the timed region once per step; e2e.per_request is the same with one
            call, copy pair and synchronise per query (one request in flight)
  roofline  dense_scan_kernel: algorithmic bytes per launch / average launch duration (events over
            the timed region / launches) vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference
            the reference's CPU scoring for this path — the numpy restatement of qdrant-client local
            mode (oracle/dense.py; the arithmetic is third-party and absent from /root/reference) —
            on the same corpus and queries, all host threads numpy's BLAS uses, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ROWS, DIM, TOPK = 1_000_000, 1024, 10
METRIC, UNIT = "queries/sec dense top-k (1M x 1024 fp16, top-10, single query, filter bitmask)", "queries/s"


def log(*a):
    """Progress on stderr (stdout carries exactly one JSON line)."""
    if os.environ.get("BENCH_VERBOSE", "1") != "0":
        print("[bench]", *a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries-per-step", type=int, default=64)
    ap.add_argument("--rows", type=int, default=N_ROWS, help="total corpus rows (default: config 2)")
    ap.add_argument("--k", type=int, default=TOPK)
    ap.add_argument("--mask-p", type=float, default=1.0, help="Bernoulli pass probability of the filter mask")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary (MaxSim / masked) measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-queries", type=int, default=0, help="queries in the CPU sample (0 = auto, ~10-30 s)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ data
def make_corpus(rows_lo: int, rows_hi: int, device):
    """Rows [lo, hi) of the seed-1 corpus, generated on `device` in 100k-row chunks whose content
    depends only on the chunk index (so every shard layout sees the same global corpus)."""
    import torch

    n = rows_hi - rows_lo
    out = torch.empty(n, DIM, dtype=torch.float16, device=device)
    chunk = 100_000
    for c in range(rows_lo // chunk, (rows_hi + chunk - 1) // chunk):
        g = torch.Generator(device=device).manual_seed(1_000_003 * 1 + c)
        blk = torch.randn(chunk, DIM, generator=g, device=device)
        blk = (blk / blk.norm(dim=1, keepdim=True)).half()
        lo, hi = max(rows_lo, c * chunk), min(rows_hi, (c + 1) * chunk)
        out[lo - rows_lo: hi - rows_lo] = blk[lo - c * chunk: hi - c * chunk]
    return out


def make_queries(nq: int):
    import torch

    g = torch.Generator().manual_seed(2)
    q = torch.randn(nq, DIM, generator=g)
    return (q / q.norm(dim=1, keepdim=True)).half()


def make_mask_words(rows_lo: int, rows_hi: int, p: float):
    """Bit-packed Bernoulli(p) mask (seed 3) for rows [lo, hi) as int32 words (numpy)."""
    import numpy as np

    from automative_rag_b200.filters import pack_bits

    n = rows_hi - rows_lo
    if p >= 1.0:
        bits = np.ones(n, dtype=bool)
    else:
        rng = np.random.default_rng(3)
        full = rng.random(N_ROWS if rows_hi <= N_ROWS else rows_hi) < p
        bits = full[rows_lo:rows_hi]
    return pack_bits(bits), bits


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_qps(c32, queries_host, mask_bits, k: int, n_queries: int):
    """Time the CPU restatement of the reference's scoring on `n_queries` queries of the workload.
    `c32` is the corpus already upcast to float32 (qdrant-local keeps float32 vectors; untimed)."""
    from oracle import dense as odense

    odense.topk(c32, queries_host[0], k, mask_bits)  # warm-up (page-in, BLAS threads)
    t0 = time.perf_counter()
    for j in range(n_queries):
        odense.topk(c32, queries_host[j % len(queries_host)], k, mask_bits)
    dt = time.perf_counter() - t0
    return n_queries / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU path for this metric (see module docstring)."""
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = torch.get_num_threads()
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    corpus = make_corpus(0, args.rows, dev).cpu().numpy()
    queries = make_queries(args.queries_per_step).numpy().astype(np.float32)
    _, bits = make_mask_words(0, args.rows, args.mask_p)
    from oracle import dense as odense

    c32 = corpus.astype(np.float32)
    del corpus
    per_step = 2  # bounded sample: 2 queries of the workload per step (~0.1-0.3 s each)
    for _ in range(max(1, min(args.warmup, 2))):
        odense.topk(c32, queries[0], args.k, bits)
    t0 = time.perf_counter()
    for s in range(args.steps):
        for j in range(per_step):
            odense.topk(c32, queries[(s * per_step + j) % len(queries)], args.k, bits)
    dt = time.perf_counter() - t0
    qps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2", "rows": args.rows, "dim": DIM, "k": args.k, "mask_p": args.mask_p,
                   "queries_per_step": args.queries_per_step, "sampled_queries_per_step": per_step},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} full-corpus queries per step x {args.steps} steps (numpy restatement "
                                   "of qdrant-client local mode; fp16 corpus upcast to fp32 once, untimed)"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import automative_rag_b200 as rag
    from automative_rag_b200 import _ffi
    from automative_rag_b200.distributed import ShardedDenseIndex, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = rag.get_engine(local_rank)
    eng.set_dense_impl(_ffi.RS_DENSE_SCAN)

    nq, k = args.queries_per_step, args.k
    lo, hi = shard_bounds(args.rows, world, rank)
    corpus = make_corpus(lo, hi, dev)
    words, bits = make_mask_words(lo, hi, args.mask_p)
    mask = torch.from_numpy(words).to(dev)
    queries_host = make_queries(nq).pin_memory()
    queries = queries_host.to(dev)
    index = ShardedDenseIndex(corpus, lo, engine=eng, metric=_ffi.RS_METRIC_COSINE)
    torch.cuda.synchronize()

    def step_device():
        return index.search(queries, k, mask)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log(f"rank {rank}/{world}: corpus rows [{lo}, {hi}) resident, warm-up")
    # clocks are sampled every 20 ms from the warm-up on (same load as the timed steps): a sharded timed region
    # can be shorter than one nvidia-smi sampling period.  Exactly W warm-up steps are run.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)  # let nvidia-smi come up
    for _ in range(args.warmup):
        step_device()
    barrier()
    log("timed region")
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    qps = nq * args.steps / (elapsed_ms * 1e-3)
    log(f"value {qps:.1f} q/s; end-to-end leg")

    # ---- end to end: the step's queries start in pinned host memory and its results end in host memory, with
    # the copies and the synchronise inside the timed region.
    #   e2e.value        one public call per STEP: the 64 queries go up in one H2D copy, 64 single-query scans
    #                    (N > 1: + all-gather + merge), the 64 x k pairs come back, one synchronise
    #   e2e.per_request  one public call per QUERY, each with its own H2D, D2H and synchronise (a latency-bound
    #                    serving loop with a single request in flight)
    out_s = torch.empty(nq, k, dtype=torch.float32).pin_memory()
    out_i = torch.empty(nq, k, dtype=torch.int64).pin_memory()
    out_s1 = torch.empty(1, k, dtype=torch.float32).pin_memory()
    out_i1 = torch.empty(1, k, dtype=torch.int64).pin_memory()
    q_dev = torch.empty(nq, DIM, dtype=torch.float16, device=dev)

    def step_e2e():
        if world == 1:
            # the C-ABI host entry point: H2D(queries) -> scans -> (score, id) pairs in host memory -> synchronise
            eng.dense_topk_host(corpus, queries_host, k, mask_dev=mask, metric=_ffi.RS_METRIC_COSINE,
                                id_base=lo, out_scores=out_s, out_ids=out_i)
        else:
            q_dev.copy_(queries_host, non_blocking=True)
            s, i = index.search(q_dev, k, mask)
            out_s.copy_(s, non_blocking=True)
            out_i.copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def step_per_request():
        for j in range(nq):
            if world == 1:
                eng.dense_topk_host(corpus, queries_host[j], k, mask_dev=mask, metric=_ffi.RS_METRIC_COSINE,
                                    id_base=lo, out_scores=out_s1, out_ids=out_i1)
            else:
                q_dev[:1].copy_(queries_host[j: j + 1], non_blocking=True)
                s, i = index.search(q_dev[:1], k, mask)
                out_s1.copy_(s, non_blocking=True)
                out_i1.copy_(i, non_blocking=True)
                torch.cuda.current_stream().synchronize()

    def timed_host_loop(fn, steps):
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return nq * steps / dt

    e2e_steps = args.steps
    e2e_qps = timed_host_loop(step_e2e, e2e_steps)
    per_request_steps = max(2, args.steps // 4)
    per_request_qps = timed_host_loop(step_per_request, per_request_steps)
    log(f"e2e {e2e_qps:.1f} q/s per step call, {per_request_qps:.1f} q/s per request; sanity check, extras, cpu baseline")

    # ---- sanity check of the timed configuration against an independent torch computation on the
    # GPU (not the oracle, not our kernel): fp32 matmul + mask + torch.topk over this rank's shard
    check = None
    if rank == 0:
        s, i = eng.dense_topk(corpus, queries[:1], k, mask=mask, id_base=lo)
        ref = (corpus.float() @ queries[0].float()) / queries[0].float().norm()
        ref = torch.where(torch.from_numpy(bits).to(dev), ref, torch.full_like(ref, float("-inf")))
        rs, ri = torch.topk(ref, k)
        check = bool(torch.equal(ri + lo, i[0]) and torch.allclose(rs, s[0], rtol=1e-3, atol=1e-6))
    # ... and the end-to-end leg returned, in host memory, what the device-resident leg computes (every rank takes
    # part: with N > 1 both legs contain the all-gather)
    step_e2e()
    s_all, i_all = step_device()
    same = bool(torch.equal(i_all.cpu(), out_i) and torch.allclose(s_all.cpu(), out_s, rtol=1e-6, atol=0))
    if rank == 0:
        check = check and same

    # ---- roofline of the scan kernel
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    rows_local = hi - lo
    passing = int(bits.sum())
    alg_bytes = passing * DIM * 2 + (rows_local + 7) // 8 + DIM * 2 + k * 12
    scan_launches = nq * args.steps
    avg_launch_ms = elapsed_ms / scan_launches  # launches are back to back on one stream
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this exact workload from the committed
    # `ncu --set full` capture (profiles/r01_dense_scan_v4_ncu.txt): 2.048201 GB + 3.71 MB; null for other shapes
    traffic = 2_051_912_488 if (world == 1 and args.rows == N_ROWS and k == TOPK and args.mask_p >= 1.0) else None
    roofline = {"bound": "hbm", "kernel": "dense_scan_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": avg_launch_ms * 1e3}

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = extra_measurements(eng, corpus, queries, dev, hbm_peak, peaks)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        host = corpus.cpu().numpy().astype(np.float32)
        qh = queries_host.numpy().astype(np.float32)
        probe_q = args.cpu_queries or 0
        if probe_q == 0:
            v, dt = cpu_reference_qps(host, qh, bits, k, 2)
            probe_q = int(min(64, max(4, 15.0 / (dt / 2))))
        v, dt = cpu_reference_qps(host, qh, bits, k, probe_q)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{probe_q} full-corpus queries of the same workload in {dt:.1f} s (numpy restatement "
                                  "of qdrant-client local mode, oracle/dense.py)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 in / f32 accumulate", "data": "synthetic",
            "config": {"workload": "config2", "rows": args.rows, "rows_per_gpu": rows_local, "dim": DIM, "k": k,
                       "mask_p": args.mask_p, "queries_per_step": nq, "parallelism": f"row-shard x{world}",

                       "l2": (f"inputs larger than L2: every query re-reads its {rows_local * DIM * 2 / 1e6:.0f} MB shard "
                              "(126 MB L2, loads carry an evict_first hint)")},
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": nq * DIM * 2, "d2h_bytes_per_step": nq * k * 12,
                    "steps": e2e_steps, "calls_per_step": 1,
                    "per_request": {"value": per_request_qps, "unit": UNIT, "calls_per_step": nq,
                                    "steps": per_request_steps}},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "parity_spot_check": check, "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def extra_measurements(eng, corpus, queries, dev, hbm_peak, peaks):
    """Secondary numbers (not the headline): masked scans, MaxSim config 4a on the tensor roof."""
    import numpy as np
    import torch

    from automative_rag_b200 import _ffi

    out = {}

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    n = corpus.shape[0]
    for p in (0.5, 0.1):
        words, bits = make_mask_words(0, n, p)
        m = torch.from_numpy(words).to(dev)
        ms = timed(lambda: eng.dense_topk(corpus, queries[:1], TOPK, mask=m), 50)
        alg = int(bits.sum()) * DIM * 2 + n // 8 + DIM * 2 + TOPK * 12
        out[f"dense_mask_p{p}"] = {"ms_per_query": ms, "queries_per_s": 1e3 / ms, "achieved_gbs": alg / ms / 1e6,
                                   "frac_of_hbm": alg / ms / 1e6 / hbm_peak}
    for kk in (100, 1000):
        ms = timed(lambda: eng.dense_topk(corpus, queries[:1], kk), 30)
        out[f"dense_k{kk}"] = {"ms_per_query": ms, "queries_per_s": 1e3 / ms}

    # BASELINE config 3: 1024 queries x 10M x 1024 bf16, top-100 — tcgen05 GEMM + fused per-query top-k
    try:
        del corpus
        torch.cuda.empty_cache()
        n3, nq3, k3 = 10_000_000, 1024, 100
        c3 = torch.empty(n3, DIM, dtype=torch.bfloat16, device=dev)
        g3 = torch.Generator(device=dev).manual_seed(4)
        for lo in range(0, n3, 500_000):
            blk = torch.randn(500_000, DIM, generator=g3, device=dev)
            c3[lo: lo + 500_000] = (blk / blk.norm(dim=1, keepdim=True)).bfloat16()
        del blk
        q3 = torch.randn(nq3, DIM, generator=torch.Generator(device=dev).manual_seed(5), device=dev)
        q3 = (q3 / q3.norm(dim=1, keepdim=True)).bfloat16()
        eng.set_dense_impl(_ffi.RS_DENSE_TCGEN05)
        ms = timed(lambda: eng.dense_topk(c3, q3, k3), 5, warm=2)
        fl = 2.0 * nq3 * n3 * DIM
        tf_sus = peaks.get("bf16_tflops_sustained", 1400.0)
        out["dense_batch_config3"] = {
            "ms_per_batch": ms, "queries_per_s": nq3 / ms * 1e3, "rows": n3, "queries": nq3, "k": k3,
            "roofline": {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": tf_sus, "unit": "TFLOP/s",
                         "frac": fl / ms / 1e9 / tf_sus, "traffic": None,
                         "peak_source": "measured sustained" if "bf16_tflops_sustained" in peaks else "fallback"}}
        del c3
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["dense_batch_config3"] = {"error": str(e)}
    finally:
        eng.set_dense_impl(_ffi.RS_DENSE_SCAN)

    # BASELINE config 1: the reference's own call shape — 1 query x 32 tokens vs 100 docs x 180 tokens, d=128,
    # fp32 tensors on the HOST — through the drop-in `_compute_maxsim_scores` (H2D, one kernel, D2H inside the
    # timed region), next to the CPU port of the reference loop (rerankers.py:244-263) on this box's cores.
    try:
        import time as _time

        import automative_rag_b200 as rag
        from oracle import maxsim as omaxsim

        g1 = torch.Generator().manual_seed(0)
        q1 = torch.randn(1, 32, 128, generator=g1)
        d1 = [torch.randn(180, 128, generator=g1) for _ in range(100)]
        for name, fp16 in (("fp32_exact", False), ("fp16", True)):
            rr = rag.B200ColBERTReranker(device=str(dev), use_fp16=fp16, use_bge_reranker=False)
            for _ in range(5):
                rr._compute_maxsim_scores(q1, d1)
            t0 = _time.perf_counter()
            for _ in range(50):
                rr._compute_maxsim_scores(q1, d1)
            dt = (_time.perf_counter() - t0) / 50
            out[f"maxsim_config1_{name}"] = {"ms_per_query_e2e": dt * 1e3, "queries_per_s": 1.0 / dt}
        # the deployed situation: the encoder left the embeddings on the GPU (reference: use_fp16 on CUDA)
        rr = rag.B200ColBERTReranker(device=str(dev), use_fp16=True, use_bge_reranker=False)
        q1d, d1d = q1.to(dev).half(), [t.to(dev).half() for t in d1]
        for _ in range(5):
            rr._compute_maxsim_scores(q1d, d1d)
        t0 = _time.perf_counter()
        for _ in range(50):
            rr._compute_maxsim_scores(q1d, d1d)
        dt = (_time.perf_counter() - t0) / 50
        out["maxsim_config1_fp16_device_inputs"] = {"ms_per_query_e2e": dt * 1e3, "queries_per_s": 1.0 / dt}
        omaxsim.maxsim_scores(q1, d1)
        t0 = _time.perf_counter()
        for _ in range(20):
            omaxsim.maxsim_scores(q1, d1)
        dt = (_time.perf_counter() - t0) / 20
        out["maxsim_config1_cpu_port"] = {"ms_per_query": dt * 1e3, "queries_per_s": 1.0 / dt,
                                          "cores": torch.get_num_threads(), "kind": "port"}
    except Exception as e:  # noqa: BLE001
        out["maxsim_config1"] = {"error": str(e)}

    # MaxSim config 4a: 256 queries x 32 tokens vs 1000 shared candidates x 300 tokens, d=128, bf16
    nq, lq, d, nd, ld = 256, 32, 128, 1000, 300
    g = torch.Generator(device=dev).manual_seed(6)
    q = torch.randn(nq, lq, d, generator=g, device=dev).bfloat16()
    toks = torch.randn(nd * ld, d, generator=torch.Generator(device=dev).manual_seed(7), device=dev).bfloat16()
    off = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
    flops = 2.0 * nq * lq * nd * ld * d
    tf_peak = peaks.get("bf16_tflops", 1590.0)
    for name, impl in (("tcgen05", _ffi.RS_MAXSIM_TCGEN05), ("mma_sync", _ffi.RS_MAXSIM_MMA)):
        try:
            eng.set_maxsim_impl(impl)
            ms = timed(lambda: eng.maxsim(q, toks, off), 20)
            out[f"maxsim_4a_{name}"] = {
                "ms_per_batch": ms, "queries_per_s": nq / ms * 1e3,
                "roofline": {"bound": "tensor", "achieved": flops / ms / 1e9, "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": flops / ms / 1e9 / tf_peak, "traffic": None,
                             "peak_source": "measured burst" if "bf16_tflops" in peaks else "fallback"}}
        except Exception as e:  # noqa: BLE001
            out[f"maxsim_4a_{name}"] = {"error": str(e)}
        finally:
            eng.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    # BASELINE config 4b: per-query candidate lists (the retrieve-then-rerank shape): 256 queries, each with its
    # own 1000 documents drawn from a 20000-document pool (1.5 GB of bf16 tokens, 12x the L2) -> HBM-bound:
    # 256 * 1000 * 300 * 128 * 2 B = 19.66 GB of token reads per batch.  AUTO picks the document-streaming tcgen05
    # kernel (maxsim_cand_tc5.cu); traffic = dram bytes of one launch from profiles/r01_maxsim_cand_v1_ncu.txt.
    try:
        pool_docs, nc = 20_000, 1000
        ptoks = torch.randn(pool_docs * ld, d, generator=torch.Generator(device=dev).manual_seed(8), device=dev).bfloat16()
        poff = (torch.arange(pool_docs + 1, dtype=torch.int32) * ld).to(dev)
        cand = torch.randint(0, pool_docs, (nq, nc), generator=torch.Generator(device=dev).manual_seed(9), device=dev,
                             dtype=torch.int32)
        ms = timed(lambda: eng.maxsim(q, ptoks, poff, cand=cand), 5, warm=2)
        nbytes = float(nq) * nc * ld * d * 2
        out["maxsim_4b_per_query_candidates"] = {
            "ms_per_batch": ms, "queries_per_s": nq / ms * 1e3,
            "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                         "frac": nbytes / ms / 1e6 / hbm_peak, "traffic": 19_725_812_248, "peak_source": "measured"},
            "impl": {_ffi.RS_MAXSIM_MMA: "mma.sync", _ffi.RS_MAXSIM_TCGEN05_CAND: "tcgen05_cand"}.get(eng.last_maxsim_impl, "?")}
    except Exception as e:  # noqa: BLE001
        out["maxsim_4b_per_query_candidates"] = {"error": str(e)}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
