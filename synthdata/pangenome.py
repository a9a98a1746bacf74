"""Seeded synthetic pangenomes and reads of the shapes BASELINE.json names.

Tooling for tests and bench only.  Everything is numpy (host); sizes up to ~1e9 bases are fine.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b

SEP = 1   # document separator byte
END = 0   # final terminator byte


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def _mutate(rng: np.random.Generator, seq: np.ndarray, snp: float, indel: float) -> np.ndarray:
    """Independent SNPs (rate snp) and 1-3 bp indels (rate indel, half insertions) on a copy of seq."""
    out = seq.copy()
    if snp > 0:
        k = rng.binomial(out.size, snp)
        pos = rng.integers(0, out.size, k)
        # substitute by a *different* base: rotate within ACGT by 1..3
        code = np.searchsorted(ACGT, out[pos])
        out[pos] = ACGT[(code + rng.integers(1, 4, k)) & 3]
    if indel > 0:
        k = rng.binomial(out.size, indel)
        if k:
            pos = np.sort(rng.integers(1, out.size - 4, k))
            pos = pos[np.concatenate(([True], np.diff(pos) > 8))]
            ln = rng.integers(1, 4, pos.size)
            ins = rng.random(pos.size) < 0.5
            pieces, prev = [], 0
            for p, l, i in zip(pos.tolist(), ln.tolist(), ins.tolist()):
                pieces.append(out[prev:p])
                if i:
                    pieces.append(ACGT[rng.integers(0, 4, l)])
                    prev = p
                else:
                    prev = p + l
            pieces.append(out[prev:])
            out = np.concatenate(pieces)
    return out


def make_haplotypes(genome_len: int, n_hap: int, *, snp: float = 1e-3, indel: float = 0.0,
                    seed: int = 1, tree: bool = False) -> list[np.ndarray]:
    """n_hap haplotypes of one random ACGT genome.

    tree=False: star phylogeny, each haplotype = base + independent variants (configs 1, 2).
    tree=True : HPRC-like shared variants: a balanced binary tree whose every edge adds variants at
                rate snp/depth, so variants are shared by clades (config 3).
    """
    rng = np.random.default_rng(seed)
    base = ACGT[rng.integers(0, 4, genome_len)]
    if not tree:
        return [_mutate(rng, base, snp, indel) for _ in range(n_hap)]
    depth = max(1, int(np.ceil(np.log2(max(2, n_hap)))))
    level = [base]
    for _ in range(depth):
        nxt = []
        for g in level:
            nxt.append(_mutate(rng, g, snp / depth, indel / depth))
            nxt.append(_mutate(rng, g, snp / depth, indel / depth))
        level = nxt
    return level[:n_hap]


def build_text(haps: list[np.ndarray], *, with_revcomp: bool = True):
    """Concatenate sequences with SEP between and END last (SURVEY.md section 4.3 layout).

    Returns (text u8, seq_starts int64 [n_seq+1], doc_of_seq int32 [n_seq]).  With revcomp, each
    document contributes two sequences (forward, reverse complement) as mumemto -r does.
    """
    seqs, doc = [], []
    for d, h in enumerate(haps):
        seqs.append(h)
        doc.append(d)
        if with_revcomp:
            seqs.append(revcomp(h))
            doc.append(d)
    total = sum(s.size + 1 for s in seqs)
    text = np.empty(total, dtype=np.uint8)
    starts = np.zeros(len(seqs) + 1, dtype=np.int64)
    p = 0
    for i, s in enumerate(seqs):
        starts[i] = p
        text[p:p + s.size] = s
        p += s.size
        text[p] = SEP
        p += 1
    text[-1] = END
    starts[-1] = p
    return text, starts, np.asarray(doc, dtype=np.int32)


def sample_reads(text: np.ndarray, seq_starts: np.ndarray, n_reads: int, read_len: int, *,
                 sub: float = 0.01, ins: float = 0.0, dele: float = 0.0, seed: int = 2,
                 len_jitter: float = 0.0):
    """Reads drawn uniformly from the sequences of `text` with errors.

    Returns (seqs u8 concatenated, offsets u64 [n_reads+1]).  Substitution-only reads keep exactly
    read_len bases; with indels (nanopore-like) lengths vary.  len_jitter>0 draws the pre-error length
    from a log-normal around read_len (long-read length spread).
    """
    rng = np.random.default_rng(seed)
    n_seq = len(seq_starts) - 1
    seq_len = np.diff(seq_starts) - 1
    if len_jitter > 0:
        lens = np.clip((read_len * rng.lognormal(0.0, len_jitter, n_reads)).astype(np.int64), 16, int(seq_len.min()))
    else:
        lens = np.full(n_reads, min(read_len, int(seq_len.min())), dtype=np.int64)
    which = rng.integers(0, n_seq, n_reads)
    start = seq_starts[which] + (rng.random(n_reads) * (seq_len[which] - lens + 1)).astype(np.int64)
    if ins == 0 and dele == 0:
        offsets = np.zeros(n_reads + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(lens)
        total = int(offsets[-1])
        # gather: index = start[read] + within-read offset
        read_id = np.repeat(np.arange(n_reads, dtype=np.int64), lens)
        within = np.arange(total, dtype=np.int64) - offsets[:-1].astype(np.int64)[read_id]
        seqs = text[start[read_id] + within]
        if sub > 0:
            k = rng.binomial(total, sub)
            pos = rng.integers(0, total, k)
            code = np.searchsorted(ACGT, seqs[pos])
            seqs[pos] = ACGT[(code + rng.integers(1, 4, k)) & 3]
        return seqs, offsets
    # indel model: per base decide keep / substitute / delete, and insert before with prob ins
    out, offsets = [], np.zeros(n_reads + 1, dtype=np.uint64)
    for i in range(n_reads):
        src = text[start[i]:start[i] + lens[i]].copy()
        u = rng.random(src.size)
        subm = u < sub
        code = np.searchsorted(ACGT, src[subm])
        src[subm] = ACGT[(code + rng.integers(1, 4, int(subm.sum()))) & 3]
        keep = ~((u >= sub) & (u < sub + dele))
        insm = rng.random(src.size) < ins
        reps = keep.astype(np.int64) + insm
        rd = np.repeat(src, reps)
        # positions that are inserted copies get a random base
        first = np.cumsum(reps) - reps
        ins_pos = first[insm]
        rd[ins_pos] = ACGT[rng.integers(0, 4, ins_pos.size)]
        out.append(rd)
        offsets[i + 1] = offsets[i] + np.uint64(rd.size)
    return (np.concatenate(out) if out else np.zeros(0, np.uint8)), offsets


def sample_reads_device(text, seq_starts: np.ndarray, n_reads: int, read_len: int, *, sub: float = 0.01, ins: float = 0.0,
                        dele: float = 0.0, seed: int = 2, len_sigma: float = 0.0, device="cuda", chunk_bases: int = 1 << 28,
                        return_origin: bool = False):
    """torch version of sample_reads for bench-scale batches (1e9+ bases), generated on `device` in chunks.

    Error model: every output base advances the source position by 1 (plain), 0 (insertion: the base is random) or
    2 (deletion: one source base skipped); substitutions replace the base by a different one.  len_sigma > 0 draws
    read lengths from a log-normal around read_len.  Returns host numpy (seqs u8, offsets u64)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    txt = torch.as_tensor(text, device=dev)
    ss = torch.as_tensor(seq_starts, device=dev)
    n_seq = ss.numel() - 1
    seq_len = ss[1:] - ss[:-1] - 1
    min_seq = int(seq_len.min())
    if len_sigma > 0:
        z = torch.randn(n_reads, generator=g, device=dev)
        lens = (read_len * torch.exp(len_sigma * z - 0.5 * len_sigma * len_sigma)).long().clamp(32, int(min_seq / 1.2))
    else:
        lens = torch.full((n_reads,), min(read_len, int(min_seq / 1.2)), dtype=torch.long, device=dev)
    offsets = torch.zeros(n_reads + 1, dtype=torch.long, device=dev)
    offsets[1:] = torch.cumsum(lens, 0)
    total = int(offsets[-1])
    out = np.empty(total, dtype=np.uint8)
    acgt = torch.as_tensor(ACGT, device=dev)
    code_of = torch.zeros(256, dtype=torch.long, device=dev)
    code_of[acgt.long()] = torch.arange(4, device=dev)
    off_host = offsets.cpu().numpy()
    origin_seq = np.empty(n_reads, dtype=np.int64)
    origin_pos = np.empty(n_reads, dtype=np.int64)
    r0 = 0
    while r0 < n_reads:
        r1 = int(np.searchsorted(off_host, off_host[r0] + chunk_bases, side="right")) - 1
        r1 = min(max(r1, r0 + 1), n_reads)
        ln = lens[r0:r1]
        nb = int(off_host[r1] - off_host[r0])
        rid = torch.repeat_interleave(torch.arange(r1 - r0, device=dev), ln)
        which = torch.randint(0, n_seq, (r1 - r0,), generator=g, device=dev)
        span = (ln.double() * (1.0 + 1.5 * dele) + 8).long()          # source bases a read may consume
        room = (seq_len[which] - span).clamp(min=1)
        start = ss[which] + (torch.rand(r1 - r0, generator=g, device=dev, dtype=torch.float64) * room.double()).long()
        u = torch.rand(nb, generator=g, device=dev)
        step = torch.ones(nb, dtype=torch.long, device=dev)
        is_ins = u < ins
        step[is_ins] = 0
        step[(u >= ins) & (u < ins + dele)] = 2
        cs = torch.cumsum(step, 0)
        first = (offsets[r0:r1] - offsets[r0])                          # chunk-local first base of each read
        base_cs = cs[first] - step[first]
        src = start[rid] + (cs - step) - base_cs[rid]
        last_ok = (ss[which + 1] - 2)[rid]
        b = txt[torch.minimum(src, last_ok)]
        rnd = acgt[torch.randint(0, 4, (nb,), generator=g, device=dev)]
        b = torch.where(is_ins, rnd, b)
        is_sub = torch.rand(nb, generator=g, device=dev) < sub
        rot = torch.randint(1, 4, (nb,), generator=g, device=dev)
        b = torch.where(is_sub, acgt[(code_of[b.long()] + rot) & 3], b)
        out[off_host[r0]:off_host[r1]] = b.cpu().numpy()
        origin_seq[r0:r1] = which.cpu().numpy()
        origin_pos[r0:r1] = (start - ss[which]).cpu().numpy()
        r0 = r1
    if return_origin:   # (sequence index, offset of the read's first base inside it): experiments on read ordering
        return out, off_host.astype(np.uint64), origin_seq, origin_pos
    return out, off_host.astype(np.uint64)
