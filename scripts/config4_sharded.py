"""BASELINE config 4 with the candidates sharded over G GPUs (torchrun, one rank per GPU):
  4a  256 queries x 32 tokens vs 1000 SHARED candidates x 300 tokens, bf16: rank r scores documents
      shard_bounds(1000, G, r) for every query, ONE all-gather of the [256, 1000/G] score blocks (ShardedMaxSim).
  4b  every query has its OWN 1000 candidates out of a 20000-document pool whose token embeddings are partitioned
      by owner (doc % G): rank r scores, for every query, the candidates it owns; ONE all-gather of the
      [256, 1000] score matrices, element-wise max (each candidate has exactly one owner).
Prints one JSON line (rank 0): ms per 256-query batch, max over ranks, CUDA events.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29566 scripts/config4_sharded.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import automative_rag_b200 as rag
from automative_rag_b200.distributed import ShardedCandidateMaxSim, ShardedMaxSim, shard_bounds

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = rag.get_engine(local)
nq, lq, d, nd, ld, pool, nc = 256, 32, 128, 1000, 300, 20_000, 1000


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# ---- 4a: identical inputs on every rank (seeded), each rank keeps its slice of the shared candidates
g = torch.Generator(device=dev).manual_seed(6)
q = torch.randn(nq, lq, d, generator=g, device=dev).bfloat16()
toks = torch.randn(nd * ld, d, generator=torch.Generator(device=dev).manual_seed(7), device=dev).bfloat16()
lo, hi = shard_bounds(nd, world, rank)
loc_toks = toks[lo * ld: hi * ld].contiguous()
loc_off = (torch.arange(hi - lo + 1, dtype=torch.int32) * ld).to(dev)
sharded = ShardedMaxSim(loc_toks, loc_off, nd, engine=eng)
ms_a = timed(lambda: sharded.scores(q))
full = eng.maxsim(q, toks, (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev))
same_a = bool(torch.equal(sharded.scores(q), full))

# ---- 4b: the pool partitioned by owner; a rank scores the candidates it owns (others point at its document 0)
ptoks = torch.randn(pool * ld, d, generator=torch.Generator(device=dev).manual_seed(8), device=dev).bfloat16()
cand = torch.randint(0, pool, (nq, nc), generator=torch.Generator(device=dev).manual_seed(9), device=dev, dtype=torch.int32)
own = torch.arange(rank, pool, world, device=dev)                      # documents this rank owns
loc_pool = ptoks.view(pool, ld * d)[own].reshape(-1, d).contiguous()
loc_poff = (torch.arange(own.numel() + 1, dtype=torch.int32) * ld).to(dev)
cand_sharded = ShardedCandidateMaxSim(loc_pool, loc_poff, engine=eng)


def step_b():
    # the partition of the step's candidate lists (owned candidates first, -1 padding) is part of the step
    return cand_sharded.scores(q, cand)


ms_b = timed(step_b, iters=10)
full_b = eng.maxsim(q, ptoks, (torch.arange(pool + 1, dtype=torch.int32) * ld).to(dev), cand=cand)
same_b = bool(torch.equal(step_b(), full_b))
if rank == 0:
    print(json.dumps({"workload": "config4 sharded", "n_gpus": world,
                      "4a_ms_per_batch": ms_a, "4a_queries_per_s": nq / ms_a * 1e3, "4a_equals_single_gpu": same_a,
                      "4b_ms_per_batch": ms_b, "4b_queries_per_s": nq / ms_b * 1e3, "4b_equals_single_gpu": same_b,
                      "note": "4b: each rank scores the candidates it owns (lists padded with -1 to the step's widest), "
                              "the partition of the candidate lists is inside the timed step"}))
if world > 1:
    dist.destroy_process_group()
