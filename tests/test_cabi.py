"""CPU: the C-ABI library loads, exports every symbol include/colbwt_b200.h declares, refuses to compute without a
GPU (no CPU fallback), and its host-only entry points behave."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "colbwt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(colbwt_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import col_bwt_b200 as cb
    lib = C.CDLL(cb.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/colbwt_b200.h but not exported"
    assert set(syms) == set(cb.EXPORTS), "python binding and header disagree"


def test_no_oracle_in_product():
    """The product must not link, import or call anything under oracle/."""
    import col_bwt_b200 as cb
    out = subprocess.run(["ldd", cb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ref_harness" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "col_bwt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_format_stats_is_the_reference_text(golden_dir):
    import col_bwt_b200 as cb
    assert cb.format_stats("q0", np.array([6, 5, 4, 3, 2, 1, 0, 2, 1, 0], np.uint16)) == b">q0 \n6 5 4 3 2 1 0 2 1 0 \n"
    assert cb.format_stats("e", np.zeros(0, np.uint8)) == b">e \n\n"
    assert cb.format_stats("x", np.array([0, 9, 10, 99, 100, 65535, 4294967295], np.uint32)) == b">x \n0 9 10 99 100 65535 4294967295 \n"
    assert open(os.path.join(golden_dir, "toy_reads.fa.pml"), "rb").read().startswith(cb.format_stats("q0", np.array([6, 5, 4, 3, 2, 1, 0, 2, 1, 0], np.uint8)))


def test_compute_calls_fail_loudly_without_gpu(golden_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import col_bwt_b200 as cb
    with pytest.raises(cb.ColBwtError) as e:
        cb.ColPml.load(os.path.join(golden_dir, "toy.col_pml"))
    assert e.value.code == -4 and "no CPU path" in str(e.value)
    cli = os.path.join(ROOT, "col_bwt_b200", "bin", "pml_query_b200")
    r = subprocess.run([cli, os.path.join(golden_dir, "toy"), "-p", os.path.join(golden_dir, "toy_reads.fa")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """`bench.py --impl reference` (the CPU arm of the driver contract) on the toy workload: exactly one JSON line on
    stdout, whatever libraries print elsewhere, with the keys the contract names."""
    import json
    import sys
    env = dict(os.environ, COLBWT_BENCH_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "0.5"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "bases/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
