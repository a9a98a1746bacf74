#!/usr/bin/env python
"""Regenerates tests/golden/* by running the reference's OWN sources (compiled verbatim into oracle/_ref by
oracle/Makefile) -- needs /root/reference, so it runs in the build container only; the fixtures are committed.

Cases
  toy   : the known-answer vector of SURVEY.md section 4.3 (hand-chosen splits, thresholds from the text).
  pan4all: the same text with `col_split -m all -s 2` (marks every multi-MUM column range, overlaps resolved).
  pan4  : 4 haplotypes x 6 kbp (+reverse complements), 0.5 % SNPs, indels; primaries from synthdata; marks from the
          reference's build_FL + col_split -m tunnels -s 4; table from the reference's build_col_bwt; expected
          PML/CID text from the reference's pml_query exactly as shipped (MULTI_THREAD on).
Usage: python tests/golden/make_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from synthdata import formats as F, pangenome as P, pipeline as PL  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(ROOT, "oracle", "_ref")


def run(*cmd):
    subprocess.run(list(cmd), check=True, stdout=subprocess.DEVNULL)


def toy(tmp):
    bwt = b"AAGCTTTCTTCCCGGGGGATAAAACAAC\x01AGATTTTTAA\x01AAA"
    heads, lens = [], []
    for c in bwt:
        if heads and heads[-1] == c:
            lens[-1] += 1
        else:
            heads.append(c)
            lens.append(1)
    thr = [0, 0, 0, 0, 4, 8, 9, 3, 2, 18, 19, 18, 25, 25, 0, 27, 18, 31, 25, 32, 32, 39]
    splits = [(0, 1), (2, 0), (3, 0), (4, 2), (6, 0), (7, 0), (8, 3), (10, 4), (12, 0), (13, 5), (15, 0), (18, 0), (19, 0),
              (20, 6), (22, 0), (24, 0), (25, 7), (27, 0), (28, 0), (29, 0), (30, 0), (31, 0), (32, 8), (34, 0), (37, 9),
              (39, 0), (40, 10), (42, 0)]
    p = os.path.join(tmp, "toy.fa")
    F.write_primaries(p, np.array(heads, np.uint8), np.array(lens), np.array(thr))
    F.write_shim_sd_vector(p + ".col_runs", len(bwt), np.array([s[0] for s in splits]))
    np.array([s[1] for s in splits], np.uint8).tofile(p + ".col_ids")
    run(os.path.join(REF, "build_col_bwt"), p)
    reads = os.path.join(tmp, "toy_reads.fa")
    with open(reads, "w") as f:
        f.write(">q0\nGATTACAGAT\n>q1 x\nTTAGGATNACA\n>q2\nCCCC\n")
    run(os.path.join(REF, "pml_query"), p, "-p", reads)
    for src, dst in ((p + ".col_pml", "toy.col_pml"), (reads, "toy_reads.fa"), (reads + ".pml", "toy_reads.fa.pml"),
                     (reads + ".cid", "toy_reads.fa.cid")):
        shutil.copy(src, os.path.join(OUT, dst))


def pan4(tmp, name="pan4", mode="tunnels", rate="4"):
    haps = P.make_haplotypes(6000, 4, snp=5e-3, indel=5e-4, seed=42)
    idx = PL.build_index(haps, split_rate=4)
    p = os.path.join(tmp, name + ".fa")
    PL.write_reference_inputs(p, idx)
    run(os.path.join(REF, "build_FL"), p)
    run(os.path.join(REF, "col_split"), p, "-m", mode, "-s", rate)
    # the primaries as the reference tools leave them: inputs of colbwt_index_from_primaries
    if name == "pan4":   # plain BWT as the reference's rlbwt_to_bwt expands it (checker of colbwt_rlbwt_to_bwt)
        run(os.path.join(REF, "rlbwt_to_bwt"), p)
        shutil.copy(p + ".bwt", os.path.join(OUT, name + ".fa.bwt"))
    for ext in (".bwt.heads", ".bwt.len", ".thr_pos", ".col_runs", ".col_ids", ".col_mums"):
        shutil.copy(p + ext, os.path.join(OUT, name + ".fa" + ext))
    n, pos = F.read_bit_vector(p + ".col_runs")           # col_split writes a plain bit_vector ...
    F.write_shim_sd_vector(p + ".col_runs", n, pos)        # ... build_col_bwt loads an sd_vector
    run(os.path.join(REF, "build_col_bwt"), p)
    seqs, off = P.sample_reads(idx["text"], idx["seq_starts"], 150, 100, sub=0.02, seed=9)
    reads = [bytes(seqs[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    reads += [b"N" * 30, b"acgtacgtnnACGT", b"A", b"ACGTNACGTTTGACNNNNACGATCGATCGATCGACTGACTAGCTAGCTAGC",
              bytes(haps[1][500:900]), bytes(haps[2][10:70]).lower(), b"G" * 120, b"\x01ACGT\x01", b"TTTTTTTTTTNTTTTTTTTT"]
    rp = os.path.join(tmp, name + "_reads.fa")
    names = [f"r{i}" for i in range(len(reads))]
    F.write_fasta(rp, reads, names=names)
    with open(rp, "ab") as f:   # a zero-length record and a multi-line record with a comment
        f.write(b">empty\n\n>multi some comment\nACGTAC\nGTTGCA\nAC\n")
    run(os.path.join(REF, "pml_query"), p, "-p", rp)
    for src, dst in ((p + ".col_pml", name + ".col_pml"), (rp, name + "_reads.fa"), (rp + ".pml", name + "_reads.fa.pml"),
                     (rp + ".cid", name + "_reads.fa.cid")):
        shutil.copy(src, os.path.join(OUT, dst))


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        toy(tmp)
        pan4(tmp)
        pan4(tmp, name="pan4all", mode="all", rate="2")   # BASELINE configs[3]: non-tunnel marking, denser sub-sampling
    print("golden fixtures written to", OUT)
