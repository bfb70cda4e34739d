"""The C-ABI library loads and exports every symbol include/rag_b200.h declares (CPU: no compute)."""
import ctypes
import os
import re

import pytest
import torch

import automative_rag_b200 as rag
from automative_rag_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 15
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in rag_b200.h but not exported by librag_b200.so"
        assert s in _ffi.SIGNATURES, f"{s} has no ctypes signature in _ffi.py"
    assert sorted(_ffi.SIGNATURES) == syms


def test_header_enums_match_the_python_constants():
    """Every enumerator of rag_b200.h (status codes, dtypes, metrics, kernel families) has the same value in _ffi.py."""
    text = open(os.path.join(ROOT, "include", "rag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    found = {}
    for body in re.findall(r"enum\s*\w*\s*\{(.*?)\}", text, flags=re.S):
        for name, value in re.findall(r"(RS_[A-Z0-9_]+)\s*=\s*(-?\d+)", body):
            found[name] = int(value)
    assert {"RS_OK", "RS_F16", "RS_METRIC_COSINE", "RS_MAXSIM_TCGEN05_CAND", "RS_DENSE_TCGEN05"} <= set(found)
    for name, value in found.items():
        assert hasattr(_ffi, name), f"{name} of rag_b200.h is missing in _ffi.py"
        assert getattr(_ffi, name) == value, f"{name}: header {value}, _ffi.py {getattr(_ffi, name)}"


def test_scan_shared_memory_plan_invariants():
    """Host-side sizing of the scan for every supported row length and a spread of k: the top-k buffer can never
    overflow between two checks, a compaction always gets below the high-water mark, every ring slot has exactly one
    owning consumer warp (one or two slots each), and everything fits the 227 KB a CTA may use."""
    lib = rag.load_library()
    out = (ctypes.c_int64 * 7)()
    ks = [1, 2, 10, 31, 32, 33, 100, 128, 129, 500, 1000, 1024, 1025, 1365, 1366, 2000, 2048]
    for d in range(8, 4097, 8):
        for k in ks:
            assert lib.rs_scan_plan(d, k, out) == 0
            tile_rows, consumers, stages, cap, hw, rounds, smem = list(out)
            assert 1 <= tile_rows <= 32 and tile_rows & (tile_rows - 1) == 0
            assert tile_rows * d * 2 <= 8192 or tile_rows == 1
            assert 2 <= consumers <= 13 and stages in (consumers, 2 * consumers)
            assert cap & (cap - 1) == 0 and cap >= 512
            slack = cap - hw
            assert slack >= consumers * tile_rows * rounds            # appends between two checks fit above high water
            assert slack >= 1 and hw >= k + k // 2 >= k               # a compaction (k survivors) frees >= k / 2 slots
            assert rounds >= 1
            ring = stages * tile_rows * d * 2
            assert smem >= ring + cap * 8 and smem <= 227 * 1024
    assert lib.rs_scan_plan(12, 10, out) != 0 and lib.rs_scan_plan(1024, 0, out) != 0 and lib.rs_scan_plan(1024, 4096, out) != 0


def test_chained_scan_plan_fits_two_ctas_per_sm():
    """Launches inside a multi-query call are sized for HALF an SM where that keeps a 12-slot ring (rs_scan_plan_chained):
    2 x (dynamic + 1 KB reserved) <= 228 KB of shared memory, and the d <= 1024 instantiations of the kernel must stay
    within the 73 registers per thread that two 448-thread CTAs leave each other (read from the built library)."""
    import ctypes as C
    import subprocess

    lib = rag.load_library()
    lone, out = (C.c_int64 * 7)(), (C.c_int64 * 8)()
    seen = set()
    for d in range(8, 4097, 8):
        for k in (1, 10, 100, 128, 170, 171, 500, 1000, 2048):
            assert lib.rs_scan_plan(d, k, lone) == 0 and lib.rs_scan_plan_chained(d, k, out) == 0
            co = out[7]
            assert co in (0, 1)
            seen.add(co)
            if co:
                assert d <= 1024 and out[2] >= 12 and 2 * (out[6] + 1024) <= 228 * 1024
                assert out[1] <= out[2] and list(out)[3:6] == list(lone)[3:6]   # same top-k buffer as the lone plan
            else:
                assert list(out)[:7] == list(lone)
    assert seen == {0, 1}
    assert lib.rs_scan_plan_chained(1024, 10, out) == 0 and out[7] == 1         # the headline shape is co-resident
    assert lib.rs_scan_plan_chained(1024, 1000, out) == 0 and out[7] == 0       # config 5 stage 1 keeps the full SM
    so = os.path.join(ROOT, "automative-rag_b200", "lib", "librag_b200.so")
    usage = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
    regs = {}
    for fn, reg in re.findall(r"Function (\S*dense_scan_kernel\S*):\s*\n\s*REG:(\d+)", usage):
        nch = int(re.search(r"Li(\d+)E", fn).group(1))
        regs[nch] = max(regs.get(nch, 0), int(reg))
    assert set(regs) == {1, 2, 4, 8, 16}, regs
    assert all(regs[n] <= 73 for n in (1, 2, 4)), regs


def test_library_loads_and_reports_abi_version():
    lib = rag.load_library()
    assert lib.rs_abi_version() == 2
    m = re.search(r"#define RS_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "rag_b200.h")).read())
    assert int(m.group(1)) == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_engine_fails_loudly_without_a_gpu():
    with pytest.raises(rag.EngineError, match="no CPU fallback"):
        rag.Engine(0)
    with pytest.raises(rag.EngineError):
        rag.B200ColBERTReranker(device="cuda:0")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "automative-rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"


def test_header_is_plain_c_and_the_c_client_links(tmp_path):
    """include/rag_b200.h compiles as C99 (-pedantic) and examples/c_client.c builds against the library with gcc;
    without a GPU the client reports the scan plan and rs_create's NO_DEVICE error and exits 0."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c",
                        os.path.join(inc, "rag_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cudart = "/usr/local/cuda/lib64"
    if not os.path.exists(os.path.join(cudart, "libcudart.so")):
        pytest.skip("no CUDA runtime to link the example against")
    exe = str(tmp_path / "c_client")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-I", inc, os.path.join(ROOT, "examples", "c_client.c"), "-o", exe,
                        "-L", libdir, "-lrag_b200", "-L", cudart, "-lcudart", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{cudart}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        return  # the device part is exercised on the GPU box (profiles/r01_c_client.txt)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ABI version 2" in r.stdout and "no CPU fallback" in r.stdout


def test_sass_of_the_hot_kernels_uses_the_blackwell_units():
    """The shipped library is sm_100a code that really uses TMA bulk copies, TMA tensor loads, tcgen05 MMAs (single-CTA
    and CTA-pair), TMEM loads and the mixed-precision FMA — per kernel, from `cuobjdump -sass` (no GPU needed)."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-sass", _ffi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out
    per_kernel = {}
    name = None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            per_kernel[name] = []
        elif name is not None:
            per_kernel[name].append(line)

    def kernels(fragment):
        hits = {k: "\n".join(v) for k, v in per_kernel.items() if fragment in k}
        assert hits, f"no kernel named *{fragment}* in the library"
        return hits

    for k, sass in kernels("dense_scan_kernel").items():
        assert "UBLKCP" in sass and "FHFMA" in sass and "SYNCS" in sass, k      # TMA bulk copy, 16-bit x 16-bit -> fp32 FMA
        assert "UTMALDG.2D.GATHER4" in sass, k                                  # sparse filters: four rows per TMA instruction
    for k, sass in kernels("maxsim_tc5_kernel").items():
        assert "UTMALDG" in sass and "UTCHMMA.2CTA" in sass and "LDTM" in sass, k    # TMA tensor load, pair tcgen05.mma, tcgen05.ld
        assert "UTCBAR.2CTA.MULTICAST" in sass and "ELECT" in sass, k                # multicast commit, elect.sync issue block
        # Round-2 lessons kept as invariants: the TMEM hand-off must not compile to a GPU-scope memory barrier (25 % of
        # the epilogue's stall samples, profiles/r02_maxsim_tc5_v5_ncu.txt), a tcgen05.mma must cost a handful of issue
        # instructions, not ten (the single issuing thread paced the kernel), and the kernel must stay small enough to
        # keep its hot loops in the instruction cache.
        assert sass.count("MEMBAR.ALL.GPU") <= 2, k                                  # only the two cluster barriers (start, end)
        lines = [ln for ln in sass.splitlines() if ln.strip().startswith("/*") and ";" in ln]
        mma = [i for i, ln in enumerate(lines) if "UTCHMMA" in ln]
        gaps = sorted(b - a for a, b in zip(mma, mma[1:]))
        assert gaps[len(gaps) // 2] <= 5, (k, gaps)                                  # median distance between MMAs
        assert len(lines) < 4500, (k, len(lines))
    for k, sass in kernels("maxsim_cand_tc5_kernel").items():
        assert "UTMALDG" in sass and "UTCHMMA" in sass and "LDTM" in sass, k
    pair = [s for k, s in kernels("dense_tc5_kernel").items() if "UTCHMMA.2CTA" in s]
    single = [s for k, s in kernels("dense_tc5_kernel").items() if "UTCHMMA.2CTA" not in s]
    assert len(pair) == 2 and len(single) == 4                                   # fp16 / bf16 x pair / single-CTA (two tiles, one tile + deeper ring)
    for sass in pair:
        assert "UTMALDG.2D.2CTA" in sass and "UTCBAR.2CTA.MULTICAST" in sass and "UCGABAR" in sass
        assert sass.count("MEMBAR.ALL.GPU") <= 2
    for k, sass in kernels("comm_allgather_topk_kernel").items():
        assert "MEMBAR" in sass, k                                               # the system-scope fence before the flags
    for k, sass in kernels("maxsim_mma_kernel").items():
        assert "HMMA.16816" in sass, k                                           # the legacy mma.sync fallback
