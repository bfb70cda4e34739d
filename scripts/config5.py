"""BASELINE config 5 — end-to-end retrieve-then-rerank over a row-sharded corpus (torchrun, one rank per GPU).

  stage 1  exact cosine top-k1 over the sharded fp16 corpus (scan kernel per rank), ONE all-gather of the
           k1 (id, score) pairs, rs_topk_merge -> global top-k1 on every rank
  stage 2  ColBERT MaxSim of the query against the k1 winners.  Candidate token embeddings come from a
           synthetic pool of P documents x Ld tokens x 128 bf16 (doc id -> slot id % P, owner rank slot % G):
           the reference re-encodes candidates with BERT per query (rerankers.py:371), which is out of scope,
           so the pool is a benchmark-design stand-in.  Every rank scores the candidates it
           owns, ONE all-gather of the scores, rs_rerank_postprocess -> top-k2.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29544 \
        scripts/config5.py --rows-per-gpu 12500000 --pool-docs 1000000 --queries 20
  --check  small sizes + comparison with a single-process CPU oracle on rank 0
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.distributed import ShardedDenseIndex

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=12_500_000)
ap.add_argument("--pool-docs", type=int, default=1_000_000)
ap.add_argument("--doc-tokens", type=int, default=300)
ap.add_argument("--queries", type=int, default=20)
ap.add_argument("--k1", type=int, default=1000)
ap.add_argument("--k2", type=int, default=10)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
if args.check:  # the oracle gathers corpus and pool on every rank: keep --check small whatever else was passed
    args.rows_per_gpu = min(args.rows_per_gpu, 200_000)
    args.pool_docs = min(args.pool_docs, 20_000)
    args.queries = min(args.queries, 6)

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = rag.get_engine(local)
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
D, DT, LQ = 1024, 128, 32
n_local, P, LD, k1, k2 = args.rows_per_gpu, args.pool_docs, args.doc_tokens, args.k1, args.k2
lo = rank * n_local

# ---- corpus shard (seed 100 + rank), pool shard (seed 200 + rank): generated on the device in chunks
corpus = torch.empty(n_local, D, dtype=torch.float16, device=dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
for a in range(0, n_local, 500_000):
    m = min(500_000, n_local - a)
    blk = torch.randn(m, D, generator=g, device=dev)
    corpus[a:a + m] = (blk / blk.norm(dim=1, keepdim=True)).half()
del blk
p_local = (P - rank + world - 1) // world  # slots rank, rank + G, ...
pool = torch.empty(p_local * LD, DT, dtype=torch.bfloat16, device=dev)
g = torch.Generator(device=dev).manual_seed(200 + rank)
for a in range(0, p_local * LD, 4_000_000):
    m = min(4_000_000, p_local * LD - a)
    pool[a:a + m] = torch.randn(m, DT, generator=g, device=dev).bfloat16()
pool_off = (torch.arange(p_local + 1, dtype=torch.int64) * LD).to(torch.int32).to(dev)
gq = torch.Generator().manual_seed(2)
queries = torch.randn(args.queries, D, generator=gq)
queries = (queries / queries.norm(dim=1, keepdim=True)).half().to(dev)
qtok = torch.randn(args.queries, LQ, DT, generator=gq).bfloat16().to(dev)
index = ShardedDenseIndex(corpus, lo, engine=eng, metric=_ffi.RS_METRIC_COSINE)
neg_inf = torch.full((1, k1), float("-inf"), device=dev)


def one_query(j):
    s1, ids = index.search(queries[j:j + 1], k1)                       # [1, k1] global ids, same on every rank
    slot = ids[0] % P
    mine = (slot % world == rank) & (ids[0] >= 0)
    cand = torch.where(mine, slot // world, torch.zeros_like(slot)).to(torch.int32).unsqueeze(0)  # [1, k1]
    sc = eng.maxsim(qtok[j:j + 1], pool, pool_off, cand=cand)          # [1, k1]; non-owned slots score doc 0
    sc = torch.where(mine.unsqueeze(0), sc, neg_inf)
    if world > 1:
        allsc = torch.empty(world, k1, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allsc, sc)
        sc = allsc.max(dim=0, keepdim=True).values                      # each candidate has exactly one owner
    top_idx, top_sc = eng.rerank_postprocess(sc.contiguous(), None, k2)  # stable order, [:k2]
    return ids[0][top_idx[0].long()], top_sc[0], ids[0], s1[0]


for j in range(min(3, args.queries)):
    one_query(j)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
res = [one_query(j) for j in range(args.queries)]
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / args.queries
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

ok = None
if args.check:
    # every rank sends its shard to rank 0's host; a numpy/torch-CPU oracle redoes both stages
    from oracle import dense as odense
    from oracle import maxsim as omaxsim
    parts_c = [torch.empty_like(corpus) for _ in range(world)] if world > 1 else [corpus]
    parts_p = [torch.empty_like(pool) for _ in range(world)] if world > 1 else [pool]
    if world > 1:
        dist.all_gather(parts_c, corpus)
        dist.all_gather(parts_p, pool)      # --check uses P % G == 0 so the shards have equal size
    if rank == 0:
        from tests._parity import assert_topk_matches
        full = torch.cat([c.cpu() for c in parts_c]).float().numpy()
        ok = True
        for j in range(min(4, args.queries)):
            top_ids, top_sc, ids, s1 = res[j]
            # stage 1: the merged global top-k1 against the oracle's full score vector (tie-aware)
            all_scores = odense.scores_f32(full, queries[j].float().cpu().numpy())
            assert_topk_matches(s1.cpu().numpy(), ids.cpu().numpy(), all_scores, np.ones(len(full), bool), k1)
            # stage 2: oracle MaxSim + stable rerank over the same k1 candidates
            docs = []
            for gid in ids.cpu().tolist():
                slot = gid % P
                r, li = slot % world, slot // world
                docs.append(parts_p[r][li * LD:(li + 1) * LD].cpu())
            sc = omaxsim.maxsim_scores(qtok[j].cpu(), docs)
            want = omaxsim.hybrid_rerank(sc, None, top_k=k2)
            want_ids = [ids[i].item() for i, _ in want]
            ok &= want_ids == top_ids.cpu().tolist()
            ok &= bool(np.allclose([v for _, v in want], top_sc.cpu().numpy(), rtol=1e-3, atol=1e-3))
            if not ok:
                print("MISMATCH q", j, want_ids, top_ids.cpu().tolist(), [v for _, v in want], top_sc.cpu().tolist(), flush=True)
if rank == 0:
    scan_bytes = n_local * D * 2
    print(json.dumps({"workload": "config5", "n_gpus": world, "rows_per_gpu": n_local, "rows_total": n_local * world,
                      "pool_docs": P, "k1": k1, "k2": k2, "queries": args.queries, "ms_per_query": ms,
                      "queries_per_s": 1e3 / ms, "stage1_scan_floor_ms": scan_bytes / 6545.9e6,
                      "oracle_check": ok}))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok in (None, True) else 1)
