"""world_size-2 gloo run of the multi-GPU plumbing on CPU: shard -> local top-k -> ONE all-gather ->
merge.  The local scan and the merge are CPU stand-ins (the oracle) because there is no GPU here;
what is under test is the sharding, the wire format and the collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense as odense
from tests._cases import bernoulli_mask, make_dense_case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, k, nq, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from automative_rag_b200.distributed import ShardedDenseIndex, ShardedMaxSim, shard_bounds

        corpus, _ = make_dense_case(7, n, d)
        g = torch.Generator().manual_seed(99)
        queries = torch.randn(nq, d, generator=g).to(torch.float16)
        mask = bernoulli_mask(5, n, 0.6)
        lo, hi = shard_bounds(n, world, rank)

        def local_search(q, kk, m, out_s, out_i):
            for j in range(q.shape[0]):
                s, i = odense.topk(corpus[lo:hi].numpy(), q[j].numpy(), kk, mask[lo:hi], id_base=lo)
                out_s[j] = torch.from_numpy(s)
                out_i[j] = torch.from_numpy(i)

        def merge(s, i, kk):
            ms, mi = odense.merge_topk(s.numpy(), i.numpy(), kk)
            return torch.from_numpy(ms), torch.from_numpy(mi)

        idx = ShardedDenseIndex(corpus[lo:hi], lo, local_search=local_search, merge=merge)
        s, i = idx.search(queries, k)
        np.savez(os.path.join(out_dir, f"dense_{rank}.npz"), s=s.numpy(), i=i.numpy())

        # MaxSim: 7 candidate docs split over 2 ranks, scores gathered in global order
        nd = 7
        full = torch.arange(nq * nd, dtype=torch.float32).view(nq, nd)
        dlo, dhi = shard_bounds(nd, world, rank)
        sm = ShardedMaxSim(None, None, nd, local_score=lambda q, w: full[:, dlo:dhi].clone())
        got = sm.scores(queries)
        np.save(os.path.join(out_dir, f"maxsim_{rank}.npy"), got.numpy())

        # per-query candidate lists: documents owned round-robin, each rank scores what it owns
        from automative_rag_b200.distributed import ShardedCandidateMaxSim

        pool, nc = 23, 9
        gc = torch.Generator().manual_seed(3)
        cand = torch.stack([torch.randperm(pool, generator=gc)[:nc] for _ in range(nq)]).to(torch.int32)
        doc_score = torch.arange(pool, dtype=torch.float32) * 1.5 + 0.25          # stand-in for MaxSim(q, doc)
        own = torch.arange(rank, pool, world)                                      # global ids this rank owns

        def local_score(q, loc_cand, w):
            loc = loc_cand.long()
            glob = own[loc.clamp_min(0).clamp_max(len(own) - 1)]
            out = doc_score[glob] + torch.arange(q.shape[0], dtype=torch.float32)[:, None]
            return torch.where(loc >= 0, out, torch.full_like(out, float("-inf")))

        sc = ShardedCandidateMaxSim(None, None, local_score=local_score).scores(queries, cand)
        np.save(os.path.join(out_dir, f"cand_{rank}.npy"), sc.numpy())
        # config 5's mapping: the candidates are corpus row ids, the document of a row is id % pool (-1 stays padding)
        rows = cand.long() + pool * torch.arange(nq)[:, None] * 7
        rows[0, 0] = -1
        sc_pool = ShardedCandidateMaxSim(None, None, local_score=local_score).scores(queries, rows, pool=pool)
        assert torch.isneginf(sc_pool[0, 0]) and torch.equal(sc_pool.flatten()[1:], sc.flatten()[1:])
        np.save(os.path.join(out_dir, f"cand_want_{rank}.npy"),
                (doc_score[cand.long()] + torch.arange(nq, dtype=torch.float32)[:, None]).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_dense_search_two_ranks_gloo(tmp_path):
    n, d, k, nq, world = 1001, 64, 10, 3, 2
    mp.spawn(_worker, args=(world, _free_port(), n, d, k, nq, str(tmp_path)), nprocs=world, join=True)
    corpus, _ = make_dense_case(7, n, d)
    g = torch.Generator().manual_seed(99)
    queries = torch.randn(nq, d, generator=g).to(torch.float16)
    mask = bernoulli_mask(5, n, 0.6)
    r0 = np.load(tmp_path / "dense_0.npz")
    r1 = np.load(tmp_path / "dense_1.npz")
    assert (r0["i"] == r1["i"]).all() and (r0["s"] == r1["s"]).all()  # identical on every rank
    for j in range(nq):
        s, i = odense.topk(corpus.numpy(), queries[j].numpy(), k, mask)
        assert r0["i"][j].tolist() == i.tolist()      # sharded == unsharded, bit for bit on ids
        np.testing.assert_array_equal(r0["s"][j], s)
    full = np.arange(nq * 7, dtype=np.float32).reshape(nq, 7)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"maxsim_{r}.npy"), full)
        # per-query candidates: every rank ends with the score of every candidate, in the caller's order
        np.testing.assert_array_equal(np.load(tmp_path / f"cand_{r}.npy"), np.load(tmp_path / f"cand_want_{r}.npy"))


def test_partition_candidates_single_process():
    from automative_rag_b200.distributed import partition_candidates

    cand = torch.tensor([[5, 8, 2, 11, 4], [1, 3, 7, 9, 13], [0, 2, 4, 6, 8]], dtype=torch.int32)
    seen = torch.zeros_like(cand, dtype=torch.int32)
    for rank in range(3):
        slots, loc = partition_candidates(cand, 3, rank)
        assert slots.shape == loc.shape and loc.dtype == torch.int32
        for q in range(cand.shape[0]):
            owned = [(j, int(c)) for j, c in enumerate(cand[q].tolist()) if c % 3 == rank]
            n_owned = len(owned)
            assert slots[q, :n_owned].tolist() == [j for j, _ in owned]            # owned first, original order
            assert loc[q, :n_owned].tolist() == [c // 3 for _, c in owned]         # local index of a round-robin owner
            assert (loc[q, n_owned:] == -1).all()                                  # padding = empty document
            seen[q, slots[q, :n_owned]] += 1
    assert (seen == 1).all()                                                       # every candidate has one owner
    # padding entries (-1: fewer than k results upstream) map to the empty document on whichever rank they land
    pad = torch.tensor([[4, -1, -1]], dtype=torch.int32)
    for rank in range(3):
        slots, loc = partition_candidates(pad, 3, rank)
        for j, c in zip(slots[0].tolist(), loc[0].tolist()):
            assert c == (4 // 3 if (pad[0, j] == 4 and rank == 4 % 3) else -1)
