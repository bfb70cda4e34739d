"""Multi-GPU sharding of the two stages: one process per GPU, `torch.distributed`
for the plumbing (rendezvous, handle exchange, barriers), ONE exchange step per stage.

The exchange itself is the engine's own kernel over NVLink peer memory whenever the engine has an
open exchange (`Engine.comm_init(group)`; `rs_allgather_topk` = push + flag + k-way merge fused in
one launch, `rs_allgather`, `rs_allreduce_max_f32`): no NCCL call and no torch op sits between the
local scoring kernel and the merged result.  Without it (gloo on CPU in the tests, or before
`comm_init`) the same steps run as `all_gather_into_tensor` + `rs_topk_merge`.

Dense: the corpus is row-partitioned (contiguous block per rank, global id = shard offset + local
row; the filter mask is partitioned the same way); every rank holds every query, runs its local
top-k, contributes k (id, score) pairs per query to a single all-gather, and merges the G lists
with `rs_topk_merge` — so every rank ends with the identical global top-k.
MaxSim: candidate documents are partitioned by rank; local scores, the same all-gather, then the
rerank tail on every rank.

The reference has no multi-GPU code; correctness here means identical results for G in {1,2,4,8}.
The plumbing (shard bounds, wire format, gather) is separated from the engine calls so it can be
exercised with the gloo backend on CPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of n rows: rank r owns [lo, hi); sizes differ by at most 1."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def wire_words(nq: int, k: int) -> int:
    """int32 words of one rank's contribution: nq*k int64 ids followed by nq*k fp32 scores,
    rounded up to an even count so every rank's block stays 8-byte aligned in the gather output."""
    return (3 * nq * k + 1) // 2 * 2


def wire_views(buf: torch.Tensor, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scores fp32 [nq, k], ids int64 [nq, k]) views into one rank's int32 wire buffer."""
    ids = buf[: 2 * nq * k].view(torch.int64).view(nq, k)
    scores = buf[2 * nq * k: 3 * nq * k].view(torch.float32).view(nq, k)
    return scores, ids


def gathered_views(gathered: torch.Tensor, world_size: int, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Views [G, nq, k] of scores and ids inside the all-gather output (int32 [G * 3*nq*k])."""
    g = gathered.view(world_size, wire_words(nq, k))
    ids = g[:, : 2 * nq * k].view(torch.int64).view(world_size, nq, k)
    scores = g[:, 2 * nq * k: 3 * nq * k].view(torch.float32).view(world_size, nq, k)
    return scores, ids


def all_gather_topk(buf: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of the stage: every rank's wire buffer -> [G * words] on every rank."""
    world = dist.get_world_size(group)
    out = torch.empty(world * buf.numel(), dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out


class ShardedDenseIndex:
    """Row shard of a dense corpus on this rank's GPU + the all-gather merge.

    `local_search(queries, k, mask, out_scores, out_ids)` and `merge(scores [G,nq,k], ids [G,nq,k], k)`
    default to the engine (`rs_dense_topk`, `rs_topk_merge`); tests inject CPU stand-ins to run the
    plumbing under gloo.
    """

    def __init__(self, local_corpus: torch.Tensor, id_base: int, *, engine=None, inv_norm: Optional[torch.Tensor] = None,
                 metric: int = 1, group=None, local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.corpus = local_corpus
        self.id_base = int(id_base)
        self.engine = engine
        self.inv_norm = inv_norm
        self.metric = metric
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._local_search = local_search or self._engine_search
        self._merge = merge or self._engine_merge
        self._wire: Optional[torch.Tensor] = None
        # the engine's own exchange over peer memory, when it spans exactly this group
        self._peer = (local_search is None and merge is None and engine is not None and self.world > 1
                      and engine.comm_world == self.world)

    def _engine_search(self, queries, k, mask, out_scores, out_ids):
        self.engine.dense_topk(self.corpus, queries, k, mask=mask, inv_norm=self.inv_norm, metric=self.metric,
                               id_base=self.id_base, out_scores=out_scores, out_ids=out_ids)

    def _engine_merge(self, scores, ids, k):
        return self.engine.topk_merge(scores, ids, k)

    def search(self, queries: torch.Tensor, k: int, mask: Optional[torch.Tensor] = None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries [nq, d] replicated on every rank; mask covers this rank's rows.  Returns the global
        (scores [nq, k], ids [nq, k]), identical on every rank."""
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        nq = queries.shape[0]
        words = wire_words(nq, k)
        if self._wire is None or self._wire.numel() != words or self._wire.device != queries.device:
            self._wire = torch.empty(words, dtype=torch.int32, device=queries.device)
        scores, ids = wire_views(self._wire, nq, k)
        self._local_search(queries, k, mask, scores, ids)
        if self.world == 1:
            return scores, ids
        if self._peer:  # push to every peer + flag + merge, one launch
            return self.engine.allgather_topk(scores, ids, k)
        gathered = all_gather_topk(self._wire, self.group)
        g_scores, g_ids = gathered_views(gathered, self.world, nq, k)
        return self._merge(g_scores, g_ids, k)


class ShardedMaxSim:
    """Candidate documents partitioned by rank; queries replicated; one all-gather of the scores."""

    def __init__(self, local_tokens: torch.Tensor, local_offsets: torch.Tensor, nd_total: int, *, engine=None,
                 group=None, local_score: Optional[Callable] = None):
        self.tokens, self.offsets, self.nd_total = local_tokens, local_offsets, nd_total
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_score = local_score or (lambda q, w: self.engine.maxsim(q, self.tokens, self.offsets, q_weight=w))
        self._peer = local_score is None and engine is not None and self.world > 1 and engine.comm_world == self.world

    def scores(self, q: torch.Tensor, q_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[nq, nd_total] fp32 on every rank, columns in global document order."""
        local = self._local_score(q, q_weight)  # [nq, nd_local]
        if self.world == 1:
            return local
        nq = local.shape[0]
        per = (self.nd_total + self.world - 1) // self.world  # pad every rank's block to the largest
        per = (per + 3) // 4 * 4                               # ... and to whole 16-byte units
        if local.shape[1] == per:
            send = local
        else:
            send = torch.full((nq, per), float("-inf"), dtype=torch.float32, device=local.device)
            send[:, : local.shape[1]] = local
        if self._peer:
            blocks = self.engine.allgather(send)               # [G, nq, per], the engine's own exchange kernel
        else:
            out = torch.empty(self.world * nq * per, dtype=torch.float32, device=local.device)
            dist.all_gather_into_tensor(out, send.reshape(-1), group=self.group)
            blocks = out.view(self.world, nq, per)
        cols = []
        for r in range(self.world):
            lo, hi = shard_bounds(self.nd_total, self.world, r)
            cols.append(blocks[r, :, : hi - lo])
        return torch.cat(cols, dim=1)


def partition_candidates(cand: torch.Tensor, world_size: int, rank: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-query candidate lists `cand` [nq, nc] of GLOBAL document ids, documents owned round-robin
    (owner = id % world_size, local index = id // world_size).  Returns
      slots     int64 [nq, width]  positions in `cand` of the candidates this rank owns, owned ones first
                                   (original order kept), width = the widest owned count of any query
      loc_cand  int32 [nq, width]  local document indices for `rs_maxsim`, -1 where a query owns fewer than
                                   `width` (an index outside the collection is an empty document: score -inf).
    """
    mine = (cand % world_size) == rank
    order = torch.argsort((~mine).to(torch.int8), dim=1, stable=True)
    width = int(mine.sum(dim=1).max().item()) if cand.numel() else 0
    slots = order[:, :width]
    owned = torch.gather(mine, 1, slots)
    loc = torch.div(torch.gather(cand, 1, slots), world_size, rounding_mode="floor")
    loc_cand = torch.where(owned, loc, torch.full_like(loc, -1)).to(torch.int32).contiguous()
    return slots, loc_cand


def owned_candidates(cand: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    """Per-query candidate lists `cand` [nq, nc] of GLOBAL document ids (documents owned round-robin: owner =
    id % world_size, local index = id // world_size) -> int32 [nq, nc] local indices for `rs_maxsim`, -1 where the
    candidate belongs to another rank or is padding (-1 upstream): an index outside the collection is an empty
    document, which the kernel skips and scores -inf.  Element-wise, no host synchronisation."""
    mine = (cand >= 0) & ((cand % world_size) == rank)
    loc = torch.div(cand, world_size, rounding_mode="floor")
    return torch.where(mine, loc, torch.full_like(loc, -1)).to(torch.int32).contiguous()


class ShardedCandidateMaxSim:
    """Retrieve-then-rerank shape (BASELINE config 4b / stage 2 of config 5): every query has its own candidate
    list; a candidate's token embeddings live on the rank that owns the document (id % G).  Each rank scores the
    candidates it owns (the others are -1 = empty documents the kernel skips), ONE exchange of the [nq, nc] score
    blocks as an element-wise max (every candidate has exactly one owner, all other ranks contribute -inf)."""

    def __init__(self, local_tokens: torch.Tensor, local_offsets: torch.Tensor, *, engine=None, group=None,
                 local_score: Optional[Callable] = None):
        self.tokens, self.offsets = local_tokens, local_offsets
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_score = local_score or (
            lambda q, loc_cand, w: self.engine.maxsim(q, self.tokens, self.offsets, q_weight=w, cand=loc_cand))
        self._peer = local_score is None and engine is not None and self.world > 1 and engine.comm_world == self.world

    def scores(self, q: torch.Tensor, cand: torch.Tensor, q_weight: Optional[torch.Tensor] = None, pool: int = 0
               ) -> torch.Tensor:
        """[nq, nc] fp32 on every rank, column j = score of cand[:, j] (-inf for padding ids).  `pool` > 0: the
        document of a candidate is cand % pool (a corpus row id mapped onto a pool of documents)."""
        nq, nc = cand.shape
        if self.engine is not None and cand.is_cuda and cand.is_contiguous() and cand.dtype in (torch.int32, torch.int64):
            loc = self.engine.owned_candidates(cand, self.world, self.rank, pool)  # one launch
        else:
            c = cand if pool <= 0 else torch.where(cand >= 0, cand % pool, cand)
            loc = owned_candidates(c, self.world, self.rank)
        local = self._local_score(q, loc, q_weight)  # -inf where not owned
        if self.world == 1:
            return local
        if self._peer:
            return self.engine.allreduce_max(local)
        out = torch.empty(self.world * nq * nc, dtype=torch.float32, device=cand.device)
        dist.all_gather_into_tensor(out, local.reshape(-1).contiguous(), group=self.group)
        return out.view(self.world, nq, nc).max(dim=0).values
