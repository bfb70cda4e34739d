"""Host-side logic that needs no GPU: packing, shard bounds, the all-gather wire format."""
import numpy as np
import pytest
import torch

from automative_rag_b200 import distributed as D
from automative_rag_b200.rerankers import pack_documents


def test_shard_bounds_partition_rows():
    for n in (0, 1, 7, 8, 1000, 12_500_001):
        for w in (1, 2, 4, 8):
            spans = [D.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("nq,k", [(1, 1), (1, 10), (3, 7), (4, 1000)])
def test_wire_views_roundtrip(nq, k):
    words = D.wire_words(nq, k)
    assert words % 2 == 0 and words >= 3 * nq * k
    world = 3
    bufs = []
    for r in range(world):
        buf = torch.zeros(words, dtype=torch.int32)
        s, i = D.wire_views(buf, nq, k)
        s.copy_(torch.arange(nq * k, dtype=torch.float32).view(nq, k) + 0.5 + r)
        i.copy_(torch.arange(nq * k, dtype=torch.int64).view(nq, k) + (r << 40))
        bufs.append(buf)
    gs, gi = D.gathered_views(torch.cat(bufs), world, nq, k)
    assert gs.shape == gi.shape == (world, nq, k)
    for r in range(world):
        assert gs[r, nq - 1, k - 1].item() == nq * k - 1 + 0.5 + r
        assert gi[r, 0, 0].item() == (r << 40)
    # every [nq, k] list is dense, lists are strided: the layout rs_topk_merge consumes without a copy
    assert gs.stride(2) == 1 and gs.stride(1) == k and gi.stride(2) == 1 and gi.stride(1) == k
    assert gs.stride(0) == words and gi.stride(0) == words // 2


def test_pack_documents():
    docs = [torch.randn(3, 8), torch.randn(1, 5, 8), torch.randn(2, 8)]
    toks, off = pack_documents(docs, torch.device("cpu"), torch.float16)
    assert toks.shape == (10, 8) and toks.dtype == torch.float16
    assert off.tolist() == [0, 3, 8, 10] and off.dtype == torch.int32
    with pytest.raises(ValueError):
        pack_documents([torch.randn(0, 8)], torch.device("cpu"), torch.float16)
