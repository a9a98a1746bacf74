#!/usr/bin/env python
"""Randomised parity run: random batches (ragged lengths, empty reads, reads with N / lower case / terminator bytes,
long reads cut into chunk tasks with random chunk geometry, every legal PML width, both the streaming call and the
device-resident batch) through the C-ABI on the GPU, each compared with the CPU oracle on every base.

Development tool; `tests/test_gpu_parity.py::test_randomised_batches` runs a short round of the same generator.
Usage: python tools/fuzz_parity.py [--seconds 120] [--seed 1]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_batch(rng, text, seq_starts, haps):
    """Returns (seqs u8, offsets u64, longest read)."""
    from synthdata import pangenome as P
    kind = rng.integers(0, 4)
    n = int(rng.integers(1, 400))
    if kind == 0:     # short ragged reads
        lens = rng.integers(0, 260, n)
    elif kind == 1:   # mixed, some beyond 255 / 65535
        lens = rng.choice([1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 300, 1000, 5000, 70000], n,
                          p=[0.06] * 16 + [0.02, 0.015, 0.005])
    elif kind == 2:   # a few long reads (chunk tasks)
        n = int(rng.integers(1, 12))
        lens = rng.integers(2000, 40000, n)
    else:             # exactly block-aligned lengths: output blocks of 16 / 32 / 64 positions end on read edges
        lens = rng.choice([16, 32, 48, 64, 128, 192], n)
    reads = []
    total_len = len(text)
    for m in lens:
        m = int(m)
        if m == 0:
            reads.append(b"")
            continue
        s = int(rng.integers(0, max(1, total_len - m)))
        r = np.array(text[s:s + m], dtype=np.uint8)
        r = r[(r != 0) & (r != 1)] if rng.random() < 0.9 else r          # mostly drop terminators; sometimes keep them
        if r.size == 0:
            reads.append(b"")
            continue
        mut = rng.random(r.size) < rng.choice([0.0, 0.01, 0.05, 0.3])
        r[mut] = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(mut.sum()))
        u = rng.random()
        if u < 0.08:
            r[rng.integers(0, r.size)] = ord("N")
        elif u < 0.12:
            r = np.frombuffer(bytes(r).lower(), np.uint8).copy()
        elif u < 0.14:
            r[rng.integers(0, r.size)] = rng.integers(0, 256)
        reads.append(bytes(r))
    off = np.zeros(len(reads) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in reads])
    return np.frombuffer(b"".join(reads), np.uint8), off, max((len(x) for x in reads), default=0)


def one_round(cb, tbl, orc, rng, text, seq_starts, haps):
    seqs, off, longest = random_batch(rng, text, seq_starts, haps)
    widths = [w for w, lim in ((1, 256), (2, 65536), (4, 1 << 32)) if longest < lim]
    width = int(rng.choice(widths))
    if rng.random() < 0.5:   # chunk geometry of the long-read path (read per call by the library)
        chunk = int(rng.choice([64, 256, 1024, 4096]))
        os.environ.update(COLBWT_SPLIT="1", COLBWT_SPLIT_CHUNK=str(chunk), COLBWT_SPLIT_WARM=str(int(rng.choice([16, 128, 512]))),
                          COLBWT_SPLIT_MIN=str(2 * chunk))
    else:
        for k in ("COLBWT_SPLIT", "COLBWT_SPLIT_CHUNK", "COLBWT_SPLIT_WARM", "COLBWT_SPLIT_MIN"):
            os.environ.pop(k, None)
    want_p, want_c = orc.query_batch(seqs, off)
    if rng.random() < 0.5:
        got_p, got_c = tbl.query(seqs, off, width)
        how = "query"
    else:
        b = tbl.batch(seqs, off, width)
        b.run(1)
        got_p, got_c = b.download()
        b.close()
        how = "batch"
    ok = np.array_equal(got_p.astype(np.uint32), want_p) and np.array_equal(got_c, want_c)
    return ok, dict(how=how, width=width, reads=len(off) - 1, bases=int(off[-1]), longest=longest,
                    split={k: os.environ.get(k) for k in ("COLBWT_SPLIT", "COLBWT_SPLIT_CHUNK", "COLBWT_SPLIT_WARM")})


def make_index(tmp):
    from synthdata import pangenome as P, pipeline as PL
    haps = P.make_haplotypes(30000, 4, snp=2e-3, indel=2e-4, seed=11)
    idx = PL.build_index(haps, split_rate=5)
    path = os.path.join(tmp, "fuzz.fa.col_pml")
    PL.write_col_pml(path, idx["columns"])
    return path, idx, haps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import col_bwt_b200 as cb
    import oracle
    rng = np.random.default_rng(a.seed)
    with tempfile.TemporaryDirectory() as tmp:
        path, idx, haps = make_index(tmp)
        tbl = cb.ColPml.load(path)
        orc = oracle.Oracle(path)
        t0 = time.time()
        rounds = bases = 0
        while time.time() - t0 < a.seconds:
            ok, info = one_round(cb, tbl, orc, rng, idx["text"], idx["seq_starts"], haps)
            rounds += 1
            bases += info["bases"]
            if not ok:
                print("MISMATCH", info, flush=True)
                return 1
        print(f"{rounds} random batches, {bases} bases, all equal to the oracle (seed {a.seed})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
