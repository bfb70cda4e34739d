"""The engine's own exchange over peer memory (comm.cu) and the per-call statistics.

One GPU: a one-rank exchange (export -> open with the single handle) runs the same push / flag / wait / merge kernels
against this GPU's own wire block.  Two or more GPUs: scripts/comm_check.py under torchrun (bit-for-bit against NCCL
all-gather + rs_topk_merge, a 2000-call soak, the sharded MaxSim classes)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from automative_rag_b200 import _ffi
from oracle import dense as odense

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def solo_exchange(engine):
    blob = C.create_string_buffer(_ffi.RS_COMM_HANDLE_BYTES)
    engine._check(engine._lib.rs_comm_export(engine._h, 1, 0, 1 << 20, blob), "rs_comm_export")
    engine._check(engine._lib.rs_comm_open(engine._h, blob), "rs_comm_open")
    assert engine.comm_world == 1
    yield engine
    engine.comm_close()
    assert engine.comm_world == 0


@pytest.mark.parametrize("nq,k_in,k_out", [(1, 10, 10), (64, 10, 10), (3, 1000, 1000), (200, 100, 37), (1, 2048, 2048)])
def test_one_rank_allgather_topk_is_the_merge_of_its_own_list(solo_exchange, nq, k_in, k_out):
    eng = solo_exchange
    rng = np.random.default_rng(nq + k_in)
    scores = np.sort(np.round(rng.standard_normal((nq, k_in)).astype(np.float32), 1), axis=1)[:, ::-1].copy()  # exact ties
    ids = rng.permutation(nq * k_in).astype(np.int64).reshape(nq, k_in) * 5
    scores[:, -2:] = -np.inf
    ids[:, -2:] = -1
    for _ in range(3):  # both parities of the wire block, repeatedly
        s, i = eng.allgather_topk(torch.from_numpy(scores).to(eng.device), torch.from_numpy(ids).to(eng.device), k_out)
        ws, wi = odense.merge_topk(scores[None], ids[None], k_out)
        np.testing.assert_array_equal(i.cpu().numpy(), wi)
        np.testing.assert_array_equal(s.cpu().numpy(), ws)


def test_one_rank_allgather_and_allreduce_max(solo_exchange):
    eng = solo_exchange
    x = torch.randn(7, 1000, device=eng.device)
    assert torch.equal(eng.allgather(x)[0], x)
    assert torch.equal(eng.allreduce_max(x), x)
    with pytest.raises(ValueError, match="exceed the exchange's slot"):
        eng.allreduce_max(torch.zeros(1 << 20, device=eng.device))


def test_exchange_calls_without_an_open_exchange_fail_loudly(engine):
    s = torch.zeros(1, 4, device=engine.device)
    i = torch.zeros(1, 4, dtype=torch.int64, device=engine.device)
    with pytest.raises(ValueError, match="no exchange is open"):
        engine.allgather_topk(s, i, 4)


def test_per_call_statistics(engine):
    """rs_set_profiling / rs_last_call_stats: the engine-side counterpart of the reference's per-request debug
    timing fields (system_service.py:336-372)."""
    dev = engine.device
    c = torch.randn(200_000, 256, device=dev).half()
    q = torch.randn(3, 256, device=dev).half()
    engine.set_profiling(True)
    try:
        engine.set_dense_impl(_ffi.RS_DENSE_SCAN)
        engine.dense_topk(c, q, 10)
        st = engine.last_call_stats()
        assert st["entry"] == _ffi.RS_CALL_DENSE_TOPK and st["kernel_family"] == _ffi.RS_DENSE_SCAN
        assert st["launches"] == 3 and st["queries"] == 3 and st["bytes_scanned"] == 3 * 200_000 * 256 * 2
        assert 0 < st["device_ms"] < 50 and st["gb_per_s"] > 100
        engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
        qq = torch.randn(8, 32, 128, device=dev).bfloat16()
        toks = torch.randn(40 * 100, 128, device=dev).bfloat16()
        off = (torch.arange(41, dtype=torch.int32) * 100).to(dev)
        engine.maxsim(qq, toks, off)
        st = engine.last_call_stats()
        assert st["entry"] == _ffi.RS_CALL_MAXSIM and st["kernel_family"] == _ffi.RS_MAXSIM_TCGEN05
        assert st["flops"] == 2.0 * 8 * 32 * 4000 * 128 and st["device_ms"] > 0
    finally:
        engine.set_profiling(False)
        engine.set_dense_impl(_ffi.RS_DENSE_AUTO)
    with pytest.raises(ValueError, match="no profiled call"):
        engine.last_call_stats()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
@pytest.mark.timeout(600)
def test_peer_exchange_two_ranks():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "scripts", "comm_check.py")],
                         capture_output=True, text=True, timeout=550)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert '"all_ranks_ok": true' in out.stdout
