#!/usr/bin/env python
"""bench.py — queries/sec of the dense top-k hot path (BASELINE.json config 2) on N B200s, plus the north-star
workloads (sharded MaxSim config 4a / 4b, the end-to-end config 5, a strong-scaled 12.5M-row scan) at every N.

Headline workload (config.workload = "config2"): exact cosine top-10 over a 1M x 1024 fp16 corpus (rows
L2-normalised in fp32 then rounded, seed 1), SINGLE-QUERY searches, metadata-filter bitmask passed
(p = 1.0: every bit set, so the mask words are read and tested but no row is skipped).

A step = one batch of `--queries-per-step` (default 64) independent single-query searches.  Every
query is its own nq=1 scan launch that re-reads its whole shard from HBM (the corpus, 2 GB, is 16x
the 126 MB L2, so nothing is cached between queries).  With N > 1 (torchrun, one process per GPU)
the corpus is row-sharded (strong scaling: 1M rows total), every rank scans its shard for all
queries of the step, and ONE exchange kernel over NVLink peer memory (rs_allgather_topk: push to every peer +
flag + k-way merge, comm.cu) leaves the global top-k on every rank.

  value     queries/s, inputs resident in HBM, CUDA events, max over ranks
  e2e       the same metric through the public host entry point (rs_dense_topk_host at N = 1,
            rs_dense_topk_sharded_host at N > 1): the step's queries start in pinned host memory, its (score, id)
            pairs end in host memory, H2D + D2H + synchronise inside the timed region once per step;
            e2e.per_request is the same with one call, copy pair and synchronise per query (one request in flight)
  roofline  dense_scan_kernel: algorithmic bytes per launch / average launch duration (events over
            the timed region / launches) vs MEASURED_PEAKS.json hbm_gbs
  extra     N = 1 only: masked scans, k = 100 / 1000, config 3 (batched tcgen05), config 1 (the reference's call shape);
            every N: config 4a / 4b with the candidates sharded by rank, config 5 weak-scaled (12.5M rows per GPU,
            top-1000 -> exchange -> MaxSim over the winners -> top-10) with the SURVEY §8d parity block run in the
            same process (returned scores recomputed on the CPU, threshold check on >= 1M sampled rows per shard,
            full oracle on a 1M-row slice per shard, oracle MaxSim + stable rerank of the winners), and a
            strong-scaled 12.5M-row scan
  cpu_baseline / --impl reference
            the reference's CPU scoring for this path — the numpy restatement of qdrant-client local
            mode (oracle/dense.py; the arithmetic is third-party and absent from /root/reference; PARITY
            UNPINNED) — on the same corpus and queries, all host threads, bounded sample.
"""
from __future__ import annotations

import argparse
import datetime
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
    ap.add_argument("--no-extra", action="store_true", help="skip every secondary measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-queries", type=int, default=0, help="queries in the CPU sample (0 = auto, ~10-30 s)")
    ap.add_argument("--c5-rows-per-gpu", type=int, default=12_500_000, help="config 5: corpus rows per GPU (weak scaling)")
    ap.add_argument("--c5-pool-docs-per-gpu", type=int, default=125_000, help="config 5: candidate-token pool documents per GPU")
    ap.add_argument("--c5-queries", type=int, default=16)
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity block of config 5")
    return ap.parse_args()


def config_dict(args, world: int) -> dict:
    """config of the JSON line — identical for the GPU arm and the reference arm at the same N."""
    rows_local = (args.rows + world - 1) // world
    return {"workload": "config2", "rows": args.rows, "rows_per_gpu": rows_local, "dim": DIM, "k": args.k,
            "mask_p": args.mask_p, "queries_per_step": args.queries_per_step, "parallelism": f"row-shard x{world}",
            "l2": (f"inputs larger than L2: every query re-reads its {rows_local * DIM * 2 / 1e6:.0f} MB shard "
                   "(126 MB L2, loads carry an evict_first hint)")}


def set_host_threads() -> int:
    """Use every host core for the CPU arm whatever the launcher put into OMP_NUM_THREADS (torchrun sets it to 1,
    which starved the reference arm at N > 1 in round 1).  Returns the thread count actually in use."""
    import torch

    n = int(os.environ.get("BENCH_CPU_THREADS", "0")) or (os.cpu_count() or 1)
    torch.set_num_threads(n)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=n)
        used = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if used:
            n = max(used)
    except Exception:  # noqa: BLE001
        pass
    return n


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
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    import numpy as np
    import torch

    threads = set_host_threads()
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
        "config": config_dict(args, max(world, args.gpus, 1)),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} full-corpus queries of the workload per step x {args.steps} steps (numpy "
                                   "restatement of qdrant-client local mode, oracle/dense.py — parity unpinned; fp16 corpus "
                                   "upcast to fp32 once, untimed)"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """What every measurement needs: ranks, device, engine, peaks, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import automative_rag_b200 as rag

        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=300))
        self.eng = rag.get_engine(self.local_rank)
        self.exchange = "single GPU"
        if self.world > 1:
            try:
                self.eng.comm_init(slot_bytes=8 << 20)
                self.exchange = "peer memory (rs_allgather_topk / rs_allreduce_max_f32 / rs_allgather, comm.cu)"
            except Exception as e:  # noqa: BLE001 — e.g. CUDA IPC unavailable in this container: NCCL plumbing instead
                log(f"rank {self.rank}: peer exchange unavailable ({e}); falling back to NCCL all-gather + rs_topk_merge")
                self.exchange = f"NCCL all_gather + rs_topk_merge (peer exchange unavailable: {e})"
            flag = torch.tensor([self.eng.comm_world], device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) != self.world and self.eng.comm_world:
                self.eng.comm_close()  # all ranks or none
        self.peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                self.peaks = json.load(f)
        except Exception:  # noqa: BLE001
            pass
        self.hbm_peak, self.peak_src = (self.peaks["hbm_gbs"], "measured") if "hbm_gbs" in self.peaks else (6650.0, "fallback")
        self.tf_burst = self.peaks.get("bf16_tflops", 1590.0)
        self.tf_sustained = self.peaks.get("bf16_tflops_sustained", 1400.0)

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, fn, iters, warm=3):
        """ms per call: CUDA events on the current stream, barrier + synchronise on both sides, max over ranks."""
        import torch

        for _ in range(warm):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        self.barrier()
        return self.max_over_ranks(a.elapsed_time(b) / iters)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from automative_rag_b200 import _ffi
    from automative_rag_b200.distributed import ShardedDenseIndex, shard_bounds

    cx = Ctx(args)
    world, rank, dev, eng = cx.world, cx.rank, cx.dev, cx.eng
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

    log(f"rank {rank}/{world}: corpus rows [{lo}, {hi}) resident, exchange: {cx.exchange}; warm-up")
    # clocks are sampled every 20 ms from the warm-up on (same load as the timed steps): a sharded timed region
    # can be shorter than one nvidia-smi sampling period.  Exactly W warm-up steps are run.
    sampler = ClockSampler(cx.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)  # let nvidia-smi come up
    for _ in range(args.warmup):
        step_device()
    cx.barrier()
    log("timed region")
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    cx.barrier()
    elapsed_ms = cx.max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    qps = nq * args.steps / (elapsed_ms * 1e-3)
    log(f"value {qps:.1f} q/s; end-to-end leg")

    # ---- end to end: the step's queries start in pinned host memory and its results end in host memory, with
    # the copies and the synchronise inside the timed region, through the C-ABI host entry points.
    #   e2e.value        one public call per STEP: the 64 queries go up in one H2D copy, 64 single-query scans
    #                    (N > 1: + the fused exchange), the 64 x k pairs come back, one synchronise
    #   e2e.per_request  one public call per QUERY, each with its own H2D, D2H and synchronise (a latency-bound
    #                    serving loop with a single request in flight)
    out_s = torch.empty(nq, k, dtype=torch.float32).pin_memory()
    out_i = torch.empty(nq, k, dtype=torch.int64).pin_memory()
    out_s1 = torch.empty(1, k, dtype=torch.float32).pin_memory()
    out_i1 = torch.empty(1, k, dtype=torch.int64).pin_memory()
    q_dev = torch.empty(nq, DIM, dtype=torch.float16, device=dev)
    peer = world > 1 and eng.comm_world == world

    def host_call(qh, os_, oi_):
        if world == 1:
            eng.dense_topk_host(corpus, qh, k, mask_dev=mask, metric=_ffi.RS_METRIC_COSINE, id_base=lo,
                                out_scores=os_, out_ids=oi_)
        elif peer:
            eng.dense_topk_sharded_host(corpus, qh, k, mask_dev=mask, metric=_ffi.RS_METRIC_COSINE, id_base=lo,
                                        out_scores=os_, out_ids=oi_)
        else:  # NCCL plumbing (only when the peer exchange could not be opened)
            n_ = qh.shape[0] if qh.dim() == 2 else 1
            q_dev[:n_].copy_(qh.reshape(n_, DIM), non_blocking=True)
            s, i = index.search(q_dev[:n_], k, mask)
            os_.copy_(s, non_blocking=True)
            oi_.copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def step_e2e():
        host_call(queries_host, out_s, out_i)

    def step_per_request():
        for j in range(nq):
            host_call(queries_host[j], out_s1, out_i1)

    def timed_host_loop(fn, steps):
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        cx.barrier()
        return nq * steps / cx.max_over_ranks(time.perf_counter() - t0)

    e2e_steps = args.steps
    e2e_qps = timed_host_loop(step_e2e, e2e_steps)
    per_request_steps = max(2, args.steps // 4)
    per_request_qps = timed_host_loop(step_per_request, per_request_steps)
    log(f"e2e {e2e_qps:.1f} q/s per step call, {per_request_qps:.1f} q/s per request; sanity check, extras, cpu baseline")

    # ---- sanity check of the timed configuration against an independent torch computation on the
    # GPU (not the oracle, not our kernel): fp32 matmul + mask + torch.topk over this rank's shard
    s, i = eng.dense_topk(corpus, queries[:1], k, mask=mask, id_base=lo)
    ref = (corpus.float() @ queries[0].float()) / queries[0].float().norm()
    ref = torch.where(torch.from_numpy(bits).to(dev), ref, torch.full_like(ref, float("-inf")))
    rs, ri = torch.topk(ref, k)
    check = bool(torch.equal(ri + lo, i[0]) and torch.allclose(rs, s[0], rtol=1e-3, atol=1e-6))
    # ... and the end-to-end leg returned, in host memory, what the device-resident leg computes (every rank takes
    # part: with N > 1 both legs contain the exchange)
    step_e2e()
    s_all, i_all = step_device()
    same = bool(torch.equal(i_all.cpu(), out_i) and torch.allclose(s_all.cpu(), out_s, rtol=1e-6, atol=0))
    check = cx.all_ok(check and same)

    # ---- roofline of the scan kernel
    rows_local = hi - lo
    passing = int(bits.sum())
    alg_bytes = passing * DIM * 2 + (rows_local + 7) // 8 + DIM * 2 + k * 12
    scan_launches = nq * args.steps
    avg_launch_ms = elapsed_ms / scan_launches  # launches are back to back on one stream
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this exact workload from the committed
    # `ncu --set full` capture of the SAME kernel source (profiles/dense_scan_traffic.json names the source hash it
    # was taken from); null when the kernel changed since, or for any other shape
    traffic = None
    if world == 1 and args.rows == N_ROWS and k == TOPK and args.mask_p >= 1.0:
        traffic = committed_traffic("dense_scan.cu", "config2")
    roofline = {"bound": "hbm", "kernel": "dense_scan_kernel", "achieved": achieved, "peak": cx.hbm_peak, "unit": "GB/s",
                "frac": achieved / cx.hbm_peak, "traffic": traffic, "peak_source": cx.peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": avg_launch_ms * 1e3}

    extra = {}
    if not args.no_extra:
        if world == 1:
            try:
                extra.update(extra_single_gpu(cx, corpus, queries))
            except Exception as e:  # noqa: BLE001
                extra["single_gpu_error"] = repr(e)
        del corpus, index, mask
        torch.cuda.empty_cache()
        eng.set_dense_impl(_ffi.RS_DENSE_AUTO)
        eng.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
        for name, fn in (("config4_sharded", extra_config4_sharded), ("config5", extra_config5)):
            log(f"extra: {name}")
            try:
                extra[name] = fn(cx)
            except Exception as e:  # noqa: BLE001 — keep the headline line even if an extra fails on this rank
                extra[name] = {"error": repr(e)}
            torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = set_host_threads()
        host = make_corpus(0, args.rows, dev).cpu().numpy().astype(np.float32)
        qh = queries_host.numpy().astype(np.float32)
        _, bits_all = make_mask_words(0, args.rows, args.mask_p)
        probe_q = args.cpu_queries or 0
        if probe_q == 0:
            v, dt = cpu_reference_qps(host, qh, bits_all, k, 2)
            probe_q = int(min(64, max(4, 15.0 / (dt / 2))))
        v, dt = cpu_reference_qps(host, qh, bits_all, k, probe_q)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{probe_q} full-corpus queries of the same workload in {dt:.1f} s (numpy restatement "
                                  "of qdrant-client local mode, oracle/dense.py; parity unpinned)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 in / f32 accumulate", "data": "synthetic",
            "config": config_dict(args, world),
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": nq * DIM * 2, "d2h_bytes_per_step": nq * k * 12,
                    "steps": e2e_steps, "calls_per_step": 1,
                    "per_request": {"value": per_request_qps, "unit": UNIT, "calls_per_step": nq,
                                    "steps": per_request_steps}},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "parity_spot_check": check, "exchange": cx.exchange, "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        cx.barrier()
        if eng.comm_world:
            eng.comm_close()
        dist.destroy_process_group()
    return 0


def committed_traffic(source: str, workload: str):
    """dram__bytes_read + dram__bytes_write of one launch from profiles/traffic.json (written from `ncu --set full`
    captures of scripts/profile_kernels.py) — only while the kernel source it was captured from is unchanged."""
    import hashlib

    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f)[source][workload]
        with open(os.path.join(ROOT, "automative-rag_b200", "csrc", source), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest()[:16] != rec["source_sha16"]:
                return None
        return rec["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None


# ------------------------------------------------------------------------------------------ N = 1 extras
def extra_single_gpu(cx, corpus, queries):
    """Secondary single-GPU numbers: masked scans, k = 100 / 1000, config 3, config 1, config 4a on one GPU."""
    import torch

    from automative_rag_b200 import _ffi

    eng, dev, peaks, hbm_peak = cx.eng, cx.dev, cx.peaks, cx.hbm_peak
    out = {}

    def timed(fn, iters, warm=3):
        return cx.timed(fn, iters, warm)

    n = corpus.shape[0]
    for p in (0.5, 0.1):
        words, bits = make_mask_words(0, n, p)
        m = torch.from_numpy(words).to(dev)
        # the headline's step shape: all of the step's queries in ONE call (a launch per query, chained), here with
        # a shared filter; `one_query_per_call` is the same scan issued one Python call per query (the ~12 us of
        # launch / prologue / merge that chaining hides are then exposed on every query)
        nq_m = queries.shape[0]
        ms = timed(lambda: eng.dense_topk(corpus, queries, TOPK, mask=m), 8) / nq_m
        ms1 = timed(lambda: eng.dense_topk(corpus, queries[:1], TOPK, mask=m), 50)
        alg = int(bits.sum()) * DIM * 2 + n // 8 + DIM * 2 + TOPK * 12
        out[f"dense_mask_p{p}"] = {"ms_per_query": ms, "queries_per_s": 1e3 / ms, "achieved_gbs": alg / ms / 1e6,
                                   "frac_of_hbm": alg / ms / 1e6 / hbm_peak, "queries_per_call": nq_m,
                                   "passing_rows": int(bits.sum()),
                                   "one_query_per_call": {"ms_per_query": ms1, "frac_of_hbm": alg / ms1 / 1e6 / hbm_peak}}
    for kk in (100, 1000):
        ms = timed(lambda: eng.dense_topk(corpus, queries[:1], kk), 30)
        out[f"dense_k{kk}"] = {"ms_per_query": ms, "queries_per_s": 1e3 / ms}

    # small batches over the headline corpus: ONE pass of the tcgen05 kernel for the whole batch (AUTO dispatch), its
    # roof the corpus bytes at the HBM peak (the scan loop costs one such pass PER QUERY)
    try:
        eng.set_dense_impl(_ffi.RS_DENSE_AUTO)
        for nq_b in (2, 8, 32, 128):
            qb = queries[:nq_b] if queries.shape[0] >= nq_b else queries.repeat((nq_b + queries.shape[0] - 1) // queries.shape[0], 1)[:nq_b]
            ms = timed(lambda: eng.dense_topk(corpus, qb, TOPK), 30)
            assert eng.last_dense_impl == _ffi.RS_DENSE_TCGEN05, "AUTO did not pick the batched kernel"
            alg = n * DIM * 2
            out[f"dense_batch_{nq_b}q"] = {"ms_per_batch": ms, "queries_per_s": nq_b / ms * 1e3, "k": TOPK,
                                           "achieved_gbs": alg / ms / 1e6, "frac_of_hbm": alg / ms / 1e6 / hbm_peak}
    except Exception as e:  # noqa: BLE001
        out["dense_batch_small"] = {"error": str(e)}
    finally:
        eng.set_dense_impl(_ffi.RS_DENSE_SCAN)

    # BASELINE config 3: 1024 queries x 10M x 1024 bf16, top-100 — tcgen05 GEMM + fused per-query top-k
    try:
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
        out["dense_batch_config3"] = {
            "ms_per_batch": ms, "queries_per_s": nq3 / ms * 1e3, "rows": n3, "queries": nq3, "k": k3,
            "roofline": {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": cx.tf_sustained, "unit": "TFLOP/s",
                         "frac": fl / ms / 1e9 / cx.tf_sustained, "traffic": committed_traffic("dense_tc5.cu", "config3"),
                         "peak_source": "measured sustained" if "bf16_tflops_sustained" in peaks else "fallback"}}
        if not cx.args.no_parity:
            out["dense_batch_config3"].update(parity_config3(eng, c3, q3, k3))
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
        threads = set_host_threads()
        omaxsim.maxsim_scores(q1, d1)
        t0 = _time.perf_counter()
        for _ in range(20):
            omaxsim.maxsim_scores(q1, d1)
        dt = (_time.perf_counter() - t0) / 20
        out["maxsim_config1_cpu_port"] = {"ms_per_query": dt * 1e3, "queries_per_s": 1.0 / dt, "cores": threads,
                                          "kind": "port"}
    except Exception as e:  # noqa: BLE001
        out["maxsim_config1"] = {"error": str(e)}

    # MaxSim config 4a on the legacy mma.sync kernel, for reference next to the tcgen05 number in config4_sharded
    try:
        nq, lq, d, nd, ld = 256, 32, 128, 1000, 300
        q = torch.randn(nq, lq, d, generator=torch.Generator(device=dev).manual_seed(6), device=dev).bfloat16()
        toks = torch.randn(nd * ld, d, generator=torch.Generator(device=dev).manual_seed(7), device=dev).bfloat16()
        off = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
        eng.set_maxsim_impl(_ffi.RS_MAXSIM_MMA)
        ms = timed(lambda: eng.maxsim(q, toks, off), 5)
        fl = 2.0 * nq * lq * nd * ld * d
        out["maxsim_4a_mma_sync"] = {"ms_per_batch": ms, "queries_per_s": nq / ms * 1e3, "tflops": fl / ms / 1e9,
                                     "frac_of_burst": fl / ms / 1e9 / cx.tf_burst}
    except Exception as e:  # noqa: BLE001
        out["maxsim_4a_mma_sync"] = {"error": str(e)}
    finally:
        eng.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    return out


# ------------------------------------------------------------------------------------------ every N: config 4 sharded
def extra_config4_sharded(cx):
    """BASELINE config 4 with the candidates sharded by rank (SURVEY §8d row "4 sharded"): 256 queries x 32 tokens,
    d = 128, bf16.
      4a  1000 SHARED candidates x 300 tokens, rank r scores documents shard_bounds(1000, G, r) for every query,
          one exchange of the [256, 1000/G] blocks.  Tensor-bound: 0.629 TFLOP per batch over G GPUs.
      4b  every query has its OWN 1000 candidates out of a 20000-document pool whose token embeddings are owned
          round-robin (doc % G); a rank scores the candidates it owns, one max-exchange of the [256, 1000] blocks.
          HBM-bound: 19.66 GB of token reads per batch over G GPUs.
    Every rank checks its sharded result bit for bit against the unsharded computation on its own GPU; rank 0 also
    checks a sample against the CPU oracle."""
    import numpy as np
    import torch

    from automative_rag_b200 import _ffi
    from automative_rag_b200.distributed import ShardedCandidateMaxSim, ShardedMaxSim, shard_bounds
    from oracle import maxsim as omaxsim

    eng, dev, world, rank = cx.eng, cx.dev, cx.world, cx.rank
    nq, lq, d, nd, ld, pool, nc = 256, 32, 128, 1000, 300, 20_000, 1000
    out = {}
    q = torch.randn(nq, lq, d, generator=torch.Generator(device=dev).manual_seed(6), device=dev).bfloat16()
    toks = torch.randn(nd * ld, d, generator=torch.Generator(device=dev).manual_seed(7), device=dev).bfloat16()
    offs = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
    lo, hi = shard_bounds(nd, world, rank)
    loc_toks = toks[lo * ld: hi * ld].contiguous()
    loc_off = (torch.arange(hi - lo + 1, dtype=torch.int32) * ld).to(dev)
    sharded = ShardedMaxSim(loc_toks, loc_off, nd, engine=eng)
    ms_a = cx.timed(lambda: sharded.scores(q), 20)
    impl_a = eng.last_maxsim_impl
    got = sharded.scores(q)
    full = eng.maxsim(q, toks, offs)
    same_a = bool(torch.equal(got, full))
    sel = [0, 131]
    want = omaxsim.maxsim_scores_packed(q[sel].cpu(), None, toks.cpu(), offs.cpu().numpy())
    err_a = float(np.abs(got[sel].cpu().numpy() - want).max() / np.abs(want).max())
    fl = 2.0 * nq * lq * nd * ld * d
    out["maxsim_4a"] = {
        "ms_per_batch": ms_a, "queries_per_s": nq / ms_a * 1e3, "candidates_per_gpu": hi - lo,
        "kernel": {_ffi.RS_MAXSIM_TCGEN05: "maxsim_tc5_kernel (tcgen05, CTA pairs)", _ffi.RS_MAXSIM_TCGEN05_CAND: "maxsim_cand_tc5_kernel",
                   _ffi.RS_MAXSIM_MMA: "maxsim_mma_kernel"}.get(impl_a, str(impl_a)),
        "roofline": {"bound": "tensor", "achieved": fl / ms_a / 1e9 / world, "peak": cx.tf_burst, "unit": "TFLOP/s per GPU",
                     "frac": fl / ms_a / 1e9 / world / cx.tf_burst,
                     "traffic": committed_traffic("maxsim_tc5.cu", "config4a") if world == 1 else None,
                     "peak_source": "measured burst" if "bf16_tflops" in cx.peaks else "fallback",
                     "note": "includes the exchange and the torch glue of the sharded step at N > 1"},
        "parity": {"sharded_equals_unsharded_bitwise": cx.all_ok(same_a), "max_rel_err_vs_cpu_oracle_2_queries": err_a,
                   "ok": cx.all_ok(same_a and err_a < 1e-3)}}
    del toks, loc_toks, full, got

    ptoks = torch.randn(pool * ld, d, generator=torch.Generator(device=dev).manual_seed(8), device=dev).bfloat16()
    poff = (torch.arange(pool + 1, dtype=torch.int32) * ld).to(dev)
    cand = torch.randint(0, pool, (nq, nc), generator=torch.Generator(device=dev).manual_seed(9), device=dev, dtype=torch.int32)
    if world > 1:
        own = torch.arange(rank, pool, world, device=dev)
        loc_pool = ptoks.view(pool, ld * d)[own].reshape(-1, d).contiguous()
        loc_poff = (torch.arange(own.numel() + 1, dtype=torch.int32) * ld).to(dev)
    else:
        loc_pool, loc_poff = ptoks, poff
    cs = ShardedCandidateMaxSim(loc_pool, loc_poff, engine=eng)
    ms_b = cx.timed(lambda: cs.scores(q, cand), 8, warm=2)
    impl_b = eng.last_maxsim_impl
    got_b = cs.scores(q, cand)
    full_b = eng.maxsim(q, ptoks, poff, cand=cand)
    same_b = bool(torch.equal(got_b, full_b))
    cs_cpu = cand[sel].cpu()
    uniq, inv = torch.unique(cs_cpu.reshape(-1).long(), return_inverse=True)
    sub = ptoks.view(pool, ld, d)[uniq.to(dev)].reshape(-1, d).cpu()
    sub_off = np.arange(len(uniq) + 1, dtype=np.int64) * ld
    want_b = omaxsim.maxsim_scores_packed(q[sel].cpu(), None, sub, sub_off, cand=inv.reshape(len(sel), nc).numpy())
    err_b = float(np.abs(got_b[sel].cpu().numpy() - want_b).max() / np.abs(want_b).max())
    nbytes = float(nq) * nc * ld * d * 2
    out["maxsim_4b"] = {
        "ms_per_batch": ms_b, "queries_per_s": nq / ms_b * 1e3,
        "kernel": {_ffi.RS_MAXSIM_TCGEN05_CAND: "maxsim_cand_tc5_kernel (tcgen05, document-streaming)",
                   _ffi.RS_MAXSIM_MMA: "maxsim_mma_kernel"}.get(impl_b, str(impl_b)),
        "roofline": {"bound": "hbm", "achieved": nbytes / ms_b / 1e6 / world, "peak": cx.hbm_peak, "unit": "GB/s per GPU",
                     "frac": nbytes / ms_b / 1e6 / world / cx.hbm_peak,
                     "traffic": committed_traffic("maxsim_cand_tc5.cu", "config4b") if world == 1 else None,
                     "peak_source": cx.peak_src,
                     "note": "owner lookup of the candidate lists and the exchange are inside the timed step"},
        "parity": {"sharded_equals_unsharded_bitwise": cx.all_ok(same_b), "max_rel_err_vs_cpu_oracle_2_queries": err_b,
                   "ok": cx.all_ok(same_b and err_b < 1e-3)}}
    return out


def parity_config3(eng, c3, q3, k3, n_check=2):
    """SURVEY §8d protocol for the batched kernel at config 3's FULL size (10M rows), `n_check` of the 1024 queries,
    CPU side in numpy fp32 on the bf16-rounded values: (i) every returned score recomputed from its row; (ii) no row
    out of 1M sampled non-returned rows beats the k-th returned score beyond tolerance; (iii) the kernel's top-k over
    the first 1M rows (all 1024 queries in the launch) against the full CPU oracle of that slice, tie-aware."""
    import time

    import numpy as np
    import torch

    from tests._parity import assert_scores_close, assert_topk_matches

    t0 = time.perf_counter()
    dev = c3.device
    n3 = c3.shape[0]
    detail, ok = {}, True
    try:
        s, i = eng.dense_topk(c3, q3, k3)
        torch.cuda.synchronize()
        which = [0, q3.shape[0] - 1][:n_check]
        qf = q3[which].float().cpu().numpy()
        qf = qf / np.linalg.norm(qf, axis=1, keepdims=True)  # the kernel scales by 1/|q| (cosine, unit corpus rows)

        def scores_of(rows_dev):  # [m, d] bf16 on the device -> fp32 scores [m, n_check] on the CPU, chunked
            outv = np.empty((rows_dev.shape[0], len(which)), dtype=np.float32)
            for a in range(0, rows_dev.shape[0], 262144):
                outv[a: a + 262144] = rows_dev[a: a + 262144].float().cpu().numpy() @ qf.T
            return outv

        worst = 0.0
        for j, qi in enumerate(which):
            ids, got = i[qi], s[qi].cpu().numpy()
            want = scores_of(c3[ids])[:, j]
            assert_scores_close(got, want, what="config3 (i) returned scores")
            worst = max(worst, float(np.abs(got - want).max()))
        detail["i_returned_scores_max_abs_err"] = worst
        nblk, blk_rows = 1024, 1024
        rng = np.random.default_rng(77)
        starts = np.sort(rng.choice((n3 - blk_rows) // blk_rows, size=nblk, replace=False)) * blk_rows
        idx = (torch.from_numpy(starts).to(dev)[:, None] + torch.arange(blk_rows, device=dev)[None, :]).reshape(-1)
        sc = scores_of(c3[idx])
        beat = 0
        for j, qi in enumerate(which):
            ids, got = i[qi].cpu().numpy(), s[qi].cpu().numpy()
            kth = float(got.min())
            rest = sc[~np.isin(idx.cpu().numpy(), ids), j]
            beat += int((rest > kth + 1e-3 * max(abs(kth), float(np.abs(rest).max())) + 1e-6).sum())
        detail["ii_threshold_sampled_rows"] = int(idx.numel())
        detail["ii_rows_beating_kth"] = beat
        ok &= beat == 0
        n_slice = min(1_000_000, n3)
        s_sl, i_sl = eng.dense_topk(c3[:n_slice], q3, k3)
        all_sc = scores_of(c3[:n_slice])
        for j, qi in enumerate(which):
            assert_topk_matches(s_sl[qi].cpu().numpy(), i_sl[qi].cpu().numpy(), all_sc[:, j], np.ones(n_slice, bool), k3)
        detail["iii_oracle_slice_rows"] = n_slice
    except AssertionError as e:
        ok = False
        detail["failure"] = str(e)[:300]
    return {"parity": bool(ok), "parity_detail": {**detail, "queries_checked": n_check, "seconds": round(time.perf_counter() - t0, 1),
                                                  "protocol": "SURVEY §8d (i)-(iii) at 10M rows; CPU numpy fp32"}}


# ------------------------------------------------------------------------------------------ every N: config 5
def cpu_scores(rows_f16, queries_f32_unit):
    """fp32 CPU scores [n, nq] of fp16-rounded rows (torch CPU tensor or numpy) against unit queries, chunked."""
    import numpy as np

    rows = rows_f16.numpy() if hasattr(rows_f16, "numpy") else rows_f16
    out = np.empty((rows.shape[0], queries_f32_unit.shape[0]), dtype=np.float32)
    for a in range(0, rows.shape[0], 131072):
        out[a: a + 131072] = rows[a: a + 131072].astype(np.float32) @ queries_f32_unit.T
    return out


def extra_config5(cx):
    """BASELINE config 5, weak-scaled: 12.5M x 1024 fp16 rows PER GPU (seed 100 + rank), so 100M rows = 204.8 GB at
    8 GPUs; one query at a time: exact cosine top-1000 per shard (single-query scan) -> ONE exchange (push + merge) ->
    ColBERT MaxSim of the query's 32 tokens against the 1000 winners' token embeddings (synthetic pool of 125k
    documents x 300 tokens x 128 bf16 per GPU, doc id -> slot id % P, owner slot % G — a benchmark-design stand-in
    for the per-query BERT re-encoding of the reference, rerankers.py:371, which is out of scope) -> ONE max-exchange
    -> stable top-10 (rs_rerank_postprocess).  Then the same scan strong-scaled: 12.5M rows TOTAL over the G GPUs.

    Parity block, run here for `parity_queries` queries on every rank, all on host cores with numpy /
    torch-CPU: (i) the returned stage-1 scores recomputed from the fp16 rows; (ii) threshold check: no row out of
    >= 1M sampled non-returned rows per shard beats the k-th returned score beyond tolerance; (iii) the engine's
    top-1000 over a 1M-row slice of every shard against the full CPU oracle of that slice (tie-aware, tests/_parity);
    (iv) oracle MaxSim of the winners each rank owns + the oracle's stable top-10 against the returned top-10."""
    import numpy as np
    import torch

    from automative_rag_b200 import _ffi
    from automative_rag_b200.distributed import ShardedCandidateMaxSim, ShardedDenseIndex
    from oracle import maxsim as omaxsim
    from tests._parity import assert_scores_close, assert_topk_matches

    args, eng, dev, world, rank = cx.args, cx.eng, cx.dev, cx.world, cx.rank
    D, DT, LQ, LD, k1, k2 = DIM, 128, 32, 300, 1000, 10
    n_local, p_local, nqs = args.c5_rows_per_gpu, args.c5_pool_docs_per_gpu, args.c5_queries
    P = p_local * world
    lo = rank * n_local
    eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
    corpus = torch.empty(n_local, D, dtype=torch.float16, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for a in range(0, n_local, 500_000):
        m = min(500_000, n_local - a)
        blk = torch.randn(m, D, generator=g, device=dev)
        corpus[a:a + m] = (blk / blk.norm(dim=1, keepdim=True)).half()
    del blk
    pool = torch.empty(p_local * LD, DT, dtype=torch.bfloat16, device=dev)  # slots rank, rank + G, ...
    g = torch.Generator(device=dev).manual_seed(200 + rank)
    for a in range(0, p_local * LD, 4_000_000):
        m = min(4_000_000, p_local * LD - a)
        pool[a:a + m] = torch.randn(m, DT, generator=g, device=dev).bfloat16()
    pool_off = (torch.arange(p_local + 1, dtype=torch.int64) * LD).to(torch.int32).to(dev)
    gq = torch.Generator().manual_seed(2)
    queries = torch.randn(nqs, D, generator=gq)
    queries = (queries / queries.norm(dim=1, keepdim=True)).half()
    qtok_h = torch.randn(nqs, LQ, DT, generator=gq).bfloat16()
    queries_d, qtok = queries.to(dev), qtok_h.to(dev)
    index = ShardedDenseIndex(corpus, lo, engine=eng, metric=_ffi.RS_METRIC_COSINE)
    rerank = ShardedCandidateMaxSim(pool, pool_off, engine=eng)

    def one_query(j):
        s1, ids = index.search(queries_d[j:j + 1], k1)                      # [1, k1] global ids, same on every rank
        sc = rerank.scores(qtok[j:j + 1], ids, pool=P)                       # document slot = id % P; owners' scores exchanged
        top_idx, top_sc = eng.rerank_postprocess(sc, None, k2)               # stable order, [:k2]
        return ids[0][top_idx[0].long()], top_sc[0], ids[0], s1[0], sc[0]

    # stage breakdown (device time, max over ranks)
    ms_total = cx.timed(lambda: [one_query(j) for j in range(nqs)], 1, warm=1) / nqs
    ms_scan = cx.timed(lambda: [eng.dense_topk(corpus, queries_d[j:j + 1], k1, id_base=lo) for j in range(nqs)], 1, warm=0) / nqs
    scan_bytes = n_local * D * 2 + D * 2 + k1 * 12
    res = {
        "workload": "config5 (weak scaling)", "rows_per_gpu": n_local, "rows_total": n_local * world, "corpus_gb_total": n_local * world * D * 2 / 1e9,
        "pool_docs": P, "k1": k1, "k2": k2, "queries": nqs, "ms_per_query": ms_total, "queries_per_s": 1e3 / ms_total,
        "row_scans_per_s": n_local * world * 1e3 / ms_total, "stage1_scan_ms": ms_scan,
        "exchange_and_rerank_ms": ms_total - ms_scan, "exchange": cx.exchange,
        "roofline": {"bound": "hbm", "kernel": "dense_scan_kernel (k = 1000)", "achieved": scan_bytes / ms_total / 1e6,
                     "peak": cx.hbm_peak, "unit": "GB/s per GPU", "frac": scan_bytes / ms_total / 1e6 / cx.hbm_peak, "traffic": None,
                     "peak_source": cx.peak_src,
                     "note": "whole query (scan + exchange + MaxSim + exchange + top-10) against the bytes of the scan alone"},
    }

    # ---- parity block
    if not args.no_parity:
        t0 = time.perf_counter()
        npq = min(2, nqs)
        ok, detail = True, {}
        try:
            qf = queries[:npq].float().numpy()
            qf = qf / np.linalg.norm(qf, axis=1, keepdims=True)
            # clones: at N = 1 the stage-1 lists are views into the index's wire buffer, which the next search overwrites
            outs = [tuple(t.clone() for t in one_query(j)) for j in range(npq)]
            torch.cuda.synchronize()
            # (i) returned stage-1 scores of the ids this rank owns, recomputed on the CPU
            worst_i = 0.0
            for j in range(npq):
                ids, s1 = outs[j][2].cpu(), outs[j][3].cpu().numpy()
                mine = (ids >= lo) & (ids < lo + n_local)
                rows = corpus[(ids[mine] - lo).to(dev)].cpu()
                want = cpu_scores(rows, qf[j:j + 1])[:, 0]
                assert_scores_close(s1[mine.numpy()], want, what="config5 (i) returned scores")
                if len(want):
                    worst_i = max(worst_i, float(np.abs(s1[mine.numpy()] - want).max()))
            detail["i_returned_scores_max_abs_err"] = cx.max_over_ranks(worst_i)
            # (ii) threshold check on >= 1M sampled non-returned rows of this shard (1024 random blocks of 1024 rows)
            nblk, blk_rows = 1024, 1024
            rng = np.random.default_rng(1234 + rank)
            population = max(1, (n_local - blk_rows) // blk_rows)
            starts = np.sort(rng.choice(population, size=min(nblk, population), replace=False)) * blk_rows
            idx = (torch.from_numpy(starts).to(dev)[:, None] + torch.arange(blk_rows, device=dev)[None, :]).reshape(-1)
            sample = corpus[idx].cpu()
            sc_cpu = cpu_scores(sample, qf)
            beat = 0
            for j in range(npq):
                ids, s1 = outs[j][2].cpu().numpy(), outs[j][3].cpu().numpy()
                kth = float(s1[ids >= 0].min())
                returned = np.isin(idx.cpu().numpy() + lo, ids)
                rest = sc_cpu[~returned, j]
                tol = 1e-3 * max(abs(kth), float(np.abs(rest).max())) + 1e-6
                beat += int((rest > kth + tol).sum())
            detail["ii_threshold_sampled_rows_per_shard"] = int(idx.numel())
            detail["ii_rows_beating_kth"] = int(cx.max_over_ranks(float(beat)))
            ok &= beat == 0
            # (iii) the engine's top-k1 over a 1M-row slice of this shard against the full CPU oracle of the slice
            n_slice = min(1_000_000, n_local)
            sl = corpus[:n_slice]
            s_sl, i_sl = eng.dense_topk(sl, queries_d[:npq], k1, id_base=lo)
            all_sc = cpu_scores(sl.cpu(), qf)
            for j in range(npq):
                assert_topk_matches(s_sl[j].cpu().numpy(), i_sl[j].cpu().numpy(), all_sc[:, j], np.ones(n_slice, bool), k1, id_base=lo)
            detail["iii_oracle_slice_rows"] = n_slice
            # (iv) oracle MaxSim of the winners this rank owns, then the oracle's stable top-k2 over the merged scores
            worst_iv = 0.0
            for j in range(npq):
                top_ids, top_sc, ids, _, sc = [t.cpu() for t in outs[j]]
                slot = ids % P
                mine = (ids >= 0) & (slot % world == rank)
                li = (slot[mine] // world).to(dev)
                docs = pool.view(p_local, LD, DT)[li].cpu()
                want = omaxsim.maxsim_scores(qtok_h[j], [docs[t] for t in range(docs.shape[0])])
                assert_scores_close(sc[mine].numpy(), want, atol=1e-3, what="config5 (iv) MaxSim of owned winners")
                if len(want):
                    worst_iv = max(worst_iv, float(np.abs(sc[mine].numpy() - want).max() / np.abs(want).max()))
                order = omaxsim.hybrid_rerank(sc.numpy(), None, top_k=k2)  # stable sort of the merged scores
                ok &= [int(ids[i]) for i, _ in order] == top_ids.tolist()
                ok &= bool(np.allclose([v for _, v in order], top_sc.numpy(), rtol=0, atol=0))
            detail["iv_maxsim_max_rel_err"] = cx.max_over_ranks(worst_iv)
        except AssertionError as e:
            ok = False
            detail["failure"] = str(e)[:300]
        res["parity"] = cx.all_ok(ok)
        res["parity_detail"] = {**detail, "queries_checked": npq, "seconds": round(time.perf_counter() - t0, 1),
                                "protocol": "SURVEY §8d (i)-(iii) + oracle MaxSim/rerank of the winners; CPU numpy / torch-CPU"}

    # ---- the same pipeline with the step's queries TOGETHER (stage 1 is "tensor for large B"):
    # ONE batched tcgen05 pass over the shard for all queries — k1 = 1000 through the two-level lists of dense_tc5.cu —
    # one exchange, ONE candidate-MaxSim launch, one max-exchange, one rerank tail.  Checked against the per-query path:
    # stage-1 id lists may differ only in entries tied with the k1-th score, the top-10 must be identical whenever the
    # stage-1 lists are.
    def batch_all():
        s1, ids = index.search(queries_d, k1)
        sc = rerank.scores(qtok, ids, pool=P)
        top_idx, top_sc = eng.rerank_postprocess(sc, None, k2)
        return torch.gather(ids, 1, top_idx.long()), top_sc, ids, s1

    try:
        eng.set_dense_impl(_ffi.RS_DENSE_AUTO)
        ms_batch = cx.timed(batch_all, 2, warm=1)
        impl_b, redo_b = eng.last_dense_impl, eng.last_dense_redo
        ms_stage1_b = cx.timed(lambda: eng.dense_topk(corpus, queries_d, k1, id_base=lo), 2, warm=0)
        top_b, tsc_b, ids_b, s1_b = [t.clone() for t in batch_all()]
        eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
        ok_b, same_lists, same_top = True, 0, 0
        for j in range(nqs):
            top_j, _, ids_j, s1_j, _ = [t.clone() for t in one_query(j)]
            kth = float(s1_j[k1 - 1])
            extra = ~torch.isin(ids_b[j], ids_j)
            if bool(extra.any()):
                ok_b &= bool(((s1_b[j][extra] - kth).abs() <= 1e-3 * abs(kth) + 1e-6).all())
            else:
                same_lists += 1
                ok_b &= bool(torch.equal(top_j, top_b[j]))
            same_top += int(torch.equal(top_j, top_b[j]))
        batch_bytes = n_local * D * 2
        res["batched"] = {
            "queries_per_batch": nqs, "ms_per_batch": ms_batch, "queries_per_s": nqs * 1e3 / ms_batch, "stage1_ms": ms_stage1_b,
            "stage1_kernel": "dense_tc5_kernel (k1 = 1000 via per-range lists of <= 128 + exactness check)"
                             if impl_b == _ffi.RS_DENSE_TCGEN05 else "dense_scan_kernel loop",
            "queries_rerun_through_scan": redo_b,
            "roofline": {"bound": "hbm", "achieved": batch_bytes / ms_batch / 1e6, "peak": cx.hbm_peak, "unit": "GB/s per GPU",
                         "frac": batch_bytes / ms_batch / 1e6 / cx.hbm_peak,
                         "traffic": committed_traffic("dense_tc5.cu", "config5_batched_16q")
                         if (n_local == 12_500_000 and nqs == 16 and impl_b == _ffi.RS_DENSE_TCGEN05) else None,
                         "peak_source": cx.peak_src,
                         "note": "whole batch (one pass over the shard for all queries + exchanges + MaxSim + top-10) "
                                 "against the bytes of ONE pass over the shard"},
            "parity_vs_per_query_path": {"ok": cx.all_ok(ok_b), "identical_stage1_lists": same_lists, "identical_top10": same_top,
                                         "queries": nqs},
        }
    except Exception as e:  # noqa: BLE001
        res["batched"] = {"error": str(e)[:300]}
    finally:
        eng.set_dense_impl(_ffi.RS_DENSE_SCAN)

    # ---- the same scan strong-scaled: 12.5M rows in total, every rank scans the first 12.5M / G rows of its shard
    n_strong = n_local // world
    strong = ShardedDenseIndex(corpus[:n_strong], rank * n_strong, engine=eng, metric=_ffi.RS_METRIC_COSINE)
    ms_strong = cx.timed(lambda: [strong.search(queries_d[j:j + 1], TOPK) for j in range(nqs)], 2, warm=1) / nqs
    sb = n_strong * D * 2
    res["strong_12p5m"] = {"rows_total": n_strong * world, "rows_per_gpu": n_strong, "k": TOPK, "ms_per_query": ms_strong,
                           "queries_per_s": 1e3 / ms_strong,
                           "roofline": {"bound": "hbm", "achieved": sb / ms_strong / 1e6, "peak": cx.hbm_peak, "unit": "GB/s per GPU",
                                        "frac": sb / ms_strong / 1e6 / cx.hbm_peak, "traffic": None, "peak_source": cx.peak_src}}
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
