"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300,
                         env={**os.environ, "BENCH_VERBOSE": "0", "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "config2" and d["gpu_launches"] == 0


def test_reference_arm_nonzero_rank_exits_quietly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000",
                          "--steps", "1", "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=120,
                         env={**os.environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "BENCH_VERBOSE": "0"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_corpus_is_independent_of_the_shard_layout():
    sys.path.insert(0, ROOT)
    import torch

    import bench

    full = bench.make_corpus(0, 250_000, torch.device("cpu"))
    a = bench.make_corpus(0, 130_000, torch.device("cpu"))
    b = bench.make_corpus(130_000, 250_000, torch.device("cpu"))
    assert torch.equal(full, torch.cat([a, b]))
