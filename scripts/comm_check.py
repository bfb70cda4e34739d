"""The engine's own exchange over NVLink peer memory (comm.cu) on G real GPUs (torchrun, one rank per GPU):
correctness against NCCL + rs_topk_merge (bit for bit), a soak of back-to-back collectives, and the latency of
one exchange / one sharded request next to the NCCL chain it replaces.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29577 \
        scripts/comm_check.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.distributed import (ShardedCandidateMaxSim, ShardedDenseIndex, ShardedMaxSim, all_gather_topk,
                                             gathered_views, shard_bounds, wire_views, wire_words)

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = rag.get_engine(local)
eng.comm_init(slot_bytes=8 << 20)
assert eng.comm_world == world
out = {"world": world}
ok = True


def nccl_merge(scores, ids, k):
    nq, kin = scores.shape
    wire = torch.empty(wire_words(nq, kin), dtype=torch.int32, device=dev)
    ws, wi = wire_views(wire, nq, kin)
    ws.copy_(scores); wi.copy_(ids)
    g = all_gather_topk(wire)
    gs, gi = gathered_views(g, world, nq, kin)
    return eng.topk_merge(gs, gi, k)


# ---- 1. rs_allgather_topk == all_gather + rs_topk_merge, bit for bit (incl. exact ties across ranks and padding)
for (nq, kin, kout) in ((1, 10, 10), (64, 10, 10), (1, 1000, 1000), (7, 100, 37), (256, 100, 100), (1024, 100, 100)):
    g = torch.Generator(device=dev).manual_seed(1000 * rank + nq + kin)
    sc = torch.randn(nq, kin, generator=g, device=dev)
    sc = (sc * 4).round() / 4 if nq == 7 else sc      # coarse scores: many exact ties between ranks
    sc, _ = sc.sort(dim=1, descending=True)
    ids = torch.randint(0, 1 << 40, (nq, kin), generator=g, device=dev, dtype=torch.int64) * world + rank
    if nq == 7:
        sc[:, kin - 5:] = float("-inf"); ids[:, kin - 5:] = -1   # padding entries
    want_s, want_i = nccl_merge(sc, ids, kout)
    got_s, got_i = eng.allgather_topk(sc, ids, kout)
    same = bool(torch.equal(want_i, got_i) and torch.equal(want_s, got_s))
    ok &= same
    if rank == 0:
        out[f"allgather_topk_{nq}x{kin}->{kout}"] = same

# ---- 2. rs_allgather / rs_allreduce_max_f32
x = torch.arange(4096, dtype=torch.float32, device=dev) + 10000 * rank
got = eng.allgather(x)
want = torch.stack([torch.arange(4096, dtype=torch.float32, device=dev) + 10000 * r for r in range(world)])
ok &= bool(torch.equal(got, want))
y = torch.full((300, 1000), float("-inf"), device=dev)
y[:, rank::world] = torch.randn(300, len(range(rank, 1000, world)), generator=torch.Generator(device=dev).manual_seed(rank), device=dev)
red = eng.allreduce_max(y)
ref = y.clone(); dist.all_reduce(ref, op=dist.ReduceOp.MAX)
ok &= bool(torch.equal(red, ref))
if rank == 0:
    out["allgather_ok"] = bool(torch.equal(got, want)); out["allreduce_max_ok"] = bool(torch.equal(red, ref))

# ---- 3. soak: 2000 back-to-back exchanges of changing sizes, every result checked on the device at the end
bad = torch.zeros((), dtype=torch.int64, device=dev)
for it in range(2000):
    nq = 1 + it % 5
    sc = torch.full((nq, 16), float(it), device=dev) - torch.arange(16, device=dev) - 0.01 * rank
    ids = (torch.arange(16, device=dev, dtype=torch.int64) * world + rank).repeat(nq, 1) + (it << 20)
    s, i = eng.allgather_topk(sc, ids, 16)
    # winners: rank 0's positions 0.. interleaved with the others: score it - j - 0.01 r  -> order (j, r)
    want_i = (torch.arange(16, device=dev) // world) * world + (torch.arange(16, device=dev) % world) + (it << 20)
    bad += (i != want_i).sum()
torch.cuda.synchronize()
ok &= int(bad.item()) == 0
if rank == 0:
    out["soak_2000_mismatches"] = int(bad.item())


def timed(fn, iters=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- 4. latency of one exchange (device time per call, back to back)
for (nq, k) in ((1, 10), (64, 10), (1, 1000), (1024, 100)):
    sc = torch.randn(nq, k, device=dev).sort(dim=1, descending=True).values
    ids = torch.randint(0, 1 << 40, (nq, k), device=dev, dtype=torch.int64)
    us_peer = timed(lambda: eng.allgather_topk(sc, ids, k))
    us_nccl = timed(lambda: nccl_merge(sc, ids, k))
    if rank == 0:
        out[f"exchange_us_{nq}x{k}"] = {"peer_fused": round(us_peer, 2), "nccl_allgather_plus_merge": round(us_nccl, 2)}

# ---- 5. one sharded request end to end (host query in, merged host result out): C entry point vs the torch chain
n_total, d, k = 1_000_000, 1024, 10
lo, hi = shard_bounds(n_total, world, rank)
g = torch.Generator(device=dev).manual_seed(77 + rank)
corpus = torch.randn(hi - lo, d, generator=g, device=dev)
corpus = (corpus / corpus.norm(dim=1, keepdim=True)).half()
qh = torch.randn(64, d, generator=torch.Generator().manual_seed(5)).half().pin_memory()
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)
index = ShardedDenseIndex(corpus, lo, engine=eng)
os_h = torch.empty(1, k).pin_memory(); oi_h = torch.empty(1, k, dtype=torch.int64).pin_memory()
qd = torch.empty(1, d, dtype=torch.float16, device=dev)
res_c, res_t = [], []


def per_request_c():
    for j in range(64):
        eng.dense_topk_sharded_host(corpus, qh[j], k, id_base=lo, out_scores=os_h, out_ids=oi_h)


def per_request_torch():
    for j in range(64):
        qd.copy_(qh[j:j + 1], non_blocking=True)
        s, i = index.search(qd, k)
        os_h.copy_(s, non_blocking=True); oi_h.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()


for j in range(4):
    eng.dense_topk_sharded_host(corpus, qh[j], k, id_base=lo, out_scores=os_h, out_ids=oi_h)
    a = (os_h.clone(), oi_h.clone())
    qd.copy_(qh[j:j + 1]); s, i = index.search(qd, k)
    ok &= bool(torch.equal(a[1], i.cpu()) and torch.equal(a[0], s.cpu()))
for name, fn in (("c_entry_peer", per_request_c), ("torch_chain_peer", per_request_torch)):
    fn(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = torch.tensor([(time.perf_counter() - t0) / (5 * 64) * 1e6], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        out[f"request_us_{name}"] = round(float(dt.item()), 1)

# ---- 6. the sharded MaxSim classes over the peer exchange == single-GPU results
nq, lq, dd, nd, ld = 32, 32, 128, 96, 150
gq = torch.Generator(device=dev).manual_seed(6)
q = torch.randn(nq, lq, dd, generator=gq, device=dev).bfloat16()
toks = torch.randn(nd * ld, dd, generator=torch.Generator(device=dev).manual_seed(7), device=dev).bfloat16()
offs = (torch.arange(nd + 1, dtype=torch.int32) * ld).to(dev)
full = eng.maxsim(q, toks, offs)
dlo, dhi = shard_bounds(nd, world, rank)
sm = ShardedMaxSim(toks[dlo * ld: dhi * ld].contiguous(), (torch.arange(dhi - dlo + 1, dtype=torch.int32) * ld).to(dev), nd, engine=eng)
ok_a = bool(torch.equal(sm.scores(q), full))
cand = torch.randint(0, nd, (nq, 40), generator=torch.Generator(device=dev).manual_seed(9), device=dev, dtype=torch.int32)
cand[0, :3] = -1
own = torch.arange(rank, nd, world, device=dev)
loc_pool = toks.view(nd, ld * dd)[own].reshape(-1, dd).contiguous()
cs = ShardedCandidateMaxSim(loc_pool, (torch.arange(own.numel() + 1, dtype=torch.int32) * ld).to(dev), engine=eng)
full_b = eng.maxsim(q, toks, offs, cand=cand)
ok_b = bool(torch.equal(cs.scores(q, cand), full_b))
ok &= ok_a and ok_b
if rank == 0:
    out["sharded_maxsim_shared_ok"] = ok_a; out["sharded_maxsim_candidates_ok"] = ok_b

flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    out["all_ranks_ok"] = bool(flag.item())
    print(json.dumps(out))
dist.barrier()
eng.comm_close()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
