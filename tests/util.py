"""Shared test helpers."""
from __future__ import annotations

import numpy as np


def parse_fastx(path: str):
    """kseq-compatible FASTA/FASTQ parse (see oracle/shim/kseq.h): returns (ids, seqs u8, offsets u64)."""
    data = open(path, "rb").read()
    ids, chunks = [], []
    lines = data.split(b"\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln[:1] in (b">", b"@"):
            fastq = ln[:1] == b"@"
            ids.append(ln[1:].split()[0].decode() if ln[1:].split() else "")
            i += 1
            seq = []
            while i < len(lines) and lines[i][:1] not in (b">", b"@", b"+"):
                s = lines[i]
                if len(s) > 1 and s.endswith(b"\r"):
                    s = s[:-1]
                seq.append(s)
                i += 1
            chunks.append(b"".join(seq))
            if fastq and i < len(lines) and lines[i][:1] == b"+":
                i += 1
                q = 0
                while i < len(lines) and q < len(chunks[-1]):
                    q += len(lines[i])
                    i += 1
        else:
            i += 1
    off = np.zeros(len(chunks) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in chunks])
    seqs = np.frombuffer(b"".join(chunks), np.uint8)
    return ids, seqs, off


def adversarial_reads(haps=None):
    """Reads exercising SURVEY.md section 4.2(4): all-N, lower case, length 0/1, absent chars, terminator bytes."""
    r = [b"N" * 40, b"acgtacgtnnACGT", b"A", b"C", b"", b"ACGTNACGTTTGACNNNNACGATCGATCGATCGACTGACTAGCTAGCTAGCTAGCTGATCG",
         b"\x01ACGT\x00ACGT\x01", b"G" * 300, b"T" * 17, b"ACGT" * 40 + b"n", b"*" * 5 + b"ACGTACGT", b"\xff\xfeAC"]
    if haps is not None:
        r += [bytes(haps[0][100:400]), bytes(haps[-1][5:21]), bytes(haps[1][1000:1033]).lower()]
    return r


def concat_reads(reads):
    off = np.zeros(len(reads) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in reads])
    return np.frombuffer(b"".join(reads), np.uint8), off


def check_pml_properties(pml, off):
    """Oracle-free invariants: PML[j] <= m-j and PML[j] in {0, PML[j+1]+1}."""
    pml = pml.astype(np.int64)
    for i in range(len(off) - 1):
        a, b = int(off[i]), int(off[i + 1])
        if b == a:
            continue
        p = pml[a:b]
        m = b - a
        assert (p <= m - np.arange(m)).all()
        nxt = np.concatenate((p[1:], [0]))
        assert ((p == 0) | (p == nxt + 1)).all()
