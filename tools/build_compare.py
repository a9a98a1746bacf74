#!/usr/bin/env python
"""The in-tree part of `col-bwt build` (scripts/col-bwt.py:94-189: build_FL -> col_split -> build_col_bwt) on one set
of primaries, timed both ways and compared byte for byte:

  reference (CPU, oracle/_ref binaries compiled from the reference's own sources): build_FL, col_split, build_col_bwt
  this repo (B200): col_split_b200 (FL table + walks on the GPU), build_col_bwt_b200 (colbwt_index_from_primaries + save)

Development / measurement tool; prints one JSON object.
Usage: python tools/build_compare.py [--H 32] [--G 1000000] [--rate 10] [--mode tunnels]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthdata import formats as F, pangenome as P, pipeline as PL  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(ROOT, "col_bwt_b200", "bin")
PRIMARIES = (".bwt.heads", ".bwt.len", ".thr_pos", ".col_mums")


def timed(*cmd):
    t0 = time.time()
    r = subprocess.run(list(cmd), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd[0]} failed: {r.stderr[-400:]}")
    return time.time() - t0


def same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=32)
    ap.add_argument("--G", type=int, default=1_000_000)
    ap.add_argument("--rate", type=int, default=10)
    ap.add_argument("--mode", default="tunnels", choices=["tunnels", "all"])
    ap.add_argument("--ref-only", action="store_true", help="CPU self-test of this script: reference chain only, synthesis on the CPU")
    a = ap.parse_args()
    import col_bwt_b200 as cb

    t0 = time.time()
    haps = P.make_haplotypes(a.G, a.H, snp=9e-4, indel=1e-4, seed=1)
    idx = PL.build_index(haps, with_revcomp=True, split_rate=a.rate, min_mum=20, device="cpu" if a.ref_only else "cuda")
    synth_s = time.time() - t0
    out = {"text": f"{a.H} haplotypes x {a.G} bp + reverse complements", "n": int(idx["columns"]["n"]), "bwt_runs": int(idx["heads"].size),
           "multi_mums": int(idx["mum_len"].size), "mode": a.mode, "split_rate": a.rate, "synthesis_s": round(synth_s, 2),
           "host_cores": os.cpu_count()}
    with tempfile.TemporaryDirectory() as tmp:
        ref_p, our_p, api_p = (os.path.join(tmp, d, "x.fa") for d in ("ref", "ours", "api"))
        for p in (ref_p, our_p, api_p):
            os.makedirs(os.path.dirname(p))
        PL.write_reference_inputs(ref_p, idx)
        for ext in PRIMARIES:
            shutil.copy(ref_p + ext, our_p + ext)
            shutil.copy(ref_p + ext, api_p + ext)

        # ---- reference chain (single-threaded CPU tools) ----------------------------------------------------------------
        ref = {"build_FL_s": timed(os.path.join(REF, "build_FL"), ref_p),
               "col_split_s": timed(os.path.join(REF, "col_split"), ref_p, "-m", a.mode, "-s", str(a.rate))}
        shutil.copy(ref_p + ".col_runs", ref_p + ".col_runs.plain")
        n_bits, pos = F.read_bit_vector(ref_p + ".col_runs")      # col_split writes a plain bit_vector,
        F.write_shim_sd_vector(ref_p + ".col_runs", n_bits, pos)  # build_col_bwt loads an sd_vector; not timed
        ref["build_col_bwt_s"] = timed(os.path.join(REF, "build_col_bwt"), ref_p)
        ref["total_s"] = ref["build_FL_s"] + ref["col_split_s"] + ref["build_col_bwt_s"]
        out["reference_cpu"] = {k: round(v, 3) for k, v in ref.items()}
        if a.ref_only:
            print(json.dumps(out))
            return 0

        # ---- this repo, as processes (each pays CUDA start-up) ------------------------------------------------------------
        ours = {"col_split_b200_s": timed(os.path.join(BIN, "col_split_b200"), our_p, "-m", a.mode, "-s", str(a.rate)),
                "build_col_bwt_b200_s": timed(os.path.join(BIN, "build_col_bwt_b200"), our_p)}
        ours["total_s"] = ours["col_split_b200_s"] + ours["build_col_bwt_b200_s"]
        out["b200_cli"] = {k: round(v, 3) for k, v in ours.items()}

        # ---- this repo, through the C-ABI in a warm process ---------------------------------------------------------------
        t = time.time()
        bits, marked = cb.col_split(api_p, a.mode, a.rate)
        t_split = time.time() - t
        t = time.time()
        tbl = cb.ColPml.from_primaries(api_p)
        t_build = time.time() - t
        t = time.time()
        tbl.save(api_p + ".col_pml")
        t_save = time.time() - t
        out["b200_api"] = {"colbwt_col_split_s": round(t_split, 3), "colbwt_index_from_primaries_s": round(t_build, 3),
                           "colbwt_index_save_s": round(t_save, 3), "total_s": round(t_split + t_build + t_save, 3)}
        out["rows"] = int(tbl.r)
        out["set_bits"], out["marked"] = int(bits), int(marked)
        tbl.close()

        out["identical"] = {
            "col_runs": same(ref_p + ".col_runs.plain", our_p + ".col_runs") and same(ref_p + ".col_runs.plain", api_p + ".col_runs"),
            "col_ids": same(ref_p + ".col_ids", our_p + ".col_ids") and same(ref_p + ".col_ids", api_p + ".col_ids"),
            "col_pml": same(ref_p + ".col_pml", our_p + ".col_pml") and same(ref_p + ".col_pml", api_p + ".col_pml"),
        }
        out["speedup_cli"] = round(ref["total_s"] / ours["total_s"], 1)
        out["speedup_api"] = round(ref["total_s"] / (t_split + t_build + t_save), 1)
    print(json.dumps(out))
    return 0 if all(out["identical"].values()) else 1


if __name__ == "__main__":
    sys.exit(main())
