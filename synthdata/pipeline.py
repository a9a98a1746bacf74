"""End-to-end synthetic index construction: haplotypes -> text -> SA/BWT/LCP -> thresholds, multi-MUMs ->
tunnel marks -> move-table columns -> `.col_pml` (the file `col_pml::load` reads, col_bwt.hpp:375-380).

Stands in for `col-bwt build -r -m tunnels -s S` (scripts/col-bwt.py:94-189), whose mumemto and Movi steps
are un-vendored.  Tooling for tests and bench only.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import bwtbuild, formats, pangenome, table


def build_index(haps, *, with_revcomp=True, split_rate=10, min_mum=20, device="cpu", marks=True, verbose=False):
    """Returns dict: columns (ch, idx, interval, offset, col_id, thr, n, bwt_r), primaries (heads, lens,
    thr_run, split_pos, split_ids, mum_len, mum_pos), text, seq_starts, doc_of_seq, num_docs."""
    t0 = time.time()
    if str(device) == "cpu":
        torch.set_num_threads(1)  # torch's CPU sort is far slower when its threads oversubscribe a small cgroup
    text, seq_starts, doc_of_seq = pangenome.build_text(haps, with_revcomp=with_revcomp)
    num_docs = len(haps)
    # large texts: no doubling levels (4 bytes x n per round), LCP from the irreducible entries instead -- same array
    big = text.size > 700_000_000
    si = bwtbuild.SuffixIndex(text, device=device, keep_levels=not big)
    bwt = si.bwt()
    lcp = bwtbuild.lcp_irreducible(si, bwt) if big else si.lcp()
    si.levels = []  # free
    if big:
        si.kmer = None
        if str(device) != "cpu":
            torch.cuda.empty_cache()
    heads, starts, lens = bwtbuild.bwt_runs(bwt)
    thr = bwtbuild.thresholds(bwt, lcp, heads, starts)
    if verbose:
        print(f"[synth] n={si.n} bwt_r={heads.numel()} sa+lcp+thr {time.time() - t0:.1f}s", flush=True)
    run_starts = starts.cpu().numpy()
    if marks and num_docs >= 2:
        mum_len, mum_pos = bwtbuild.multi_mums(si, lcp, bwt, seq_starts, doc_of_seq, num_docs, min_mum)
        m_start, m_id = table.mark_tunnels(si.sa, si.isa, starts, mum_len, mum_pos, num_docs, split_rate)
        res = table.resolve_marks_fast(si.n, run_starts, m_start, m_id, num_docs)
        if res is None:
            res = table.resolve_marks(si.n, run_starts, m_start, m_id, np.full(m_start.size, num_docs))
        split_pos, split_ids = res
    else:
        mum_len = mum_pos = np.zeros(0, np.uint64)
        split_pos, split_ids = run_starts.astype(np.uint64), np.zeros(run_starts.size, np.uint8)
    heads_np, lens_np, thr_np = heads.cpu().numpy(), lens.cpu().numpy(), thr.cpu().numpy()
    cols = table.build_columns(heads_np, lens_np, thr_np, split_pos, split_ids)
    if verbose:
        print(f"[synth] rows={cols['ch'].size} mums={mum_len.size} marked_rows={(cols['col_id'] > 0).sum()} "
              f"total {time.time() - t0:.1f}s", flush=True)
    return {
        "columns": cols, "heads": heads_np, "lens": lens_np, "thr_run": thr_np, "split_pos": split_pos,
        "split_ids": split_ids, "mum_len": mum_len, "mum_pos": mum_pos, "text": text, "seq_starts": seq_starts,
        "doc_of_seq": doc_of_seq, "num_docs": num_docs,
    }


def write_col_pml(path: str, cols: dict) -> None:
    rows = formats.rows_from_columns(cols["ch"], cols["idx"], cols["interval"], cols["offset"], cols["col_id"], cols["thr"])
    formats.write_col_pml(path, cols["bwt_r"], cols["n"], rows)


def write_reference_inputs(prefix: str, idx: dict) -> None:
    """The files the reference's own build tools read: prefix.{bwt.heads,bwt.len,thr_pos,col_mums}."""
    formats.write_primaries(prefix, idx["heads"], idx["lens"], idx["thr_run"])
    formats.write_col_mums(prefix + ".col_mums", idx["num_docs"], idx["mum_len"], idx["mum_pos"])


def synth_move_table(r: int, *, mean_len: float = 16.0, mark_frac: float = 0.08, seed: int = 5, device="cuda", snap: float = 0.0):
    """Directly synthesised valid move table with r rows (BASELINE configs[4]: index scaled to an HPRC-like run count
    without building a text): random run characters (adjacent runs differ) and lengths, LF columns from the stable
    character sort exactly as LF_table::compute_table derives them, thresholds uniform in the legal gap
    (end of previous run of the character, start of this run] -- or, with probability `snap`, on a run START inside
    that gap, which is where a real index keeps most of them: the minimum LCP of the gap usually sits where the BWT
    character changes (0.854 of the thresholds of a 16-haplotype tree-structured pangenome built by this package,
    profiles/r2/threshold_fit.log); uniform thresholds put two distinct in-row flip offsets on 19 % of the rows, a real
    table on ~0.01 % -- random chain ids on a fraction of the rows, and one terminator row.  Returns (rows_u8 [r,18] torch tensor on `device` in the `.col_pml` row layout, n, dict of
    LF columns for walking).  Everything stays on `device`."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dev = torch.device(device)
    # characters: codes 0..3 with c[i] != c[i-1]; row r//2 becomes the terminator (byte 1)
    inc = torch.randint(1, 4, (r,), generator=g, device=dev, dtype=torch.int64)
    code = torch.cumsum(inc, 0) % 4
    acgt = torch.tensor([ord(x) for x in "ACGT"], dtype=torch.uint8, device=dev)
    ch = acgt[code]
    term = r // 2
    ch[term] = 1
    # lengths: 1 + geometric, capped (the reference's 16-bit offsets need rows < 65536)
    u = torch.rand(r, generator=g, device=dev, dtype=torch.float64)
    lens = (1 + torch.floor(torch.log(1 - u) / np.log(1 - 1.0 / mean_len))).clamp(1, 1000).to(torch.int64)
    lens[term] = 1
    idx = torch.cumsum(lens, 0) - lens
    n = int(lens.sum())
    # compute_table: stable order by character byte
    _, order = torch.sort(ch, stable=True)
    ls = lens[order]
    fpos = torch.cumsum(ls, 0) - ls
    dest_s = torch.searchsorted(idx, fpos, right=True) - 1
    doff_s = fpos - idx[dest_s]
    dest = torch.empty(r, dtype=torch.int64, device=dev)
    doff = torch.empty(r, dtype=torch.int64, device=dev)
    dest[order] = dest_s
    doff[order] = doff_s
    del dest_s, doff_s, fpos, ls
    # thresholds: previous row of the same character is the previous entry of the same character group in `order`
    chs = ch[order]
    prev = torch.roll(order, 1)
    same = torch.ones(r, dtype=torch.bool, device=dev)
    same[0] = False
    same[1:] = chs[1:] == chs[:-1]
    lo = idx[prev] + lens[prev]                       # first position after the previous run of the character
    hi = idx[order]                                   # start of this run
    uu = torch.rand(r, generator=g, device=dev, dtype=torch.float64)
    t = lo + torch.floor(uu * (hi - lo + 1).to(torch.float64)).to(torch.int64)
    if snap > 0:
        # a run start inside the gap: the rows after the previous run of the character up to this run are prev+1 .. order
        u2 = torch.rand(r, generator=g, device=dev, dtype=torch.float64)
        j = prev + 1 + torch.floor(u2 * (order - prev).clamp(min=1).to(torch.float64)).to(torch.int64)
        j = torch.minimum(j.clamp(min=0), order)
        on_start = torch.rand(r, generator=g, device=dev) < snap
        t = torch.where(on_start & same, idx[j], t)
        del u2, j, on_start
    thr = torch.zeros(r, dtype=torch.int64, device=dev)
    thr[order] = torch.where(same, t, torch.zeros_like(t))
    del chs, prev, same, lo, hi, uu, t, order
    cid = torch.where(torch.rand(r, generator=g, device=dev) < mark_frac,
                      torch.randint(1, 256, (r,), generator=g, device=dev), torch.zeros(r, dtype=torch.int64, device=dev)).to(torch.uint8)
    rows = torch.empty((r, 18), dtype=torch.uint8, device=dev)
    rows[:, 0] = ch
    for b in range(5):
        rows[:, 1 + b] = ((idx >> (8 * b)) & 255).to(torch.uint8)
        rows[:, 13 + b] = ((thr >> (8 * b)) & 255).to(torch.uint8)
    for b in range(4):
        rows[:, 6 + b] = ((dest >> (8 * b)) & 255).to(torch.uint8)
    rows[:, 10] = (doff & 255).to(torch.uint8)
    rows[:, 11] = ((doff >> 8) & 255).to(torch.uint8)
    rows[:, 12] = cid
    return rows, n, {"ch": ch, "lens": lens, "dest": dest, "doff": doff}


def walk_reads(cols: dict, n_reads: int, read_len: int, *, sub: float = 0.01, seed: int = 6, chunk_reads: int = 1 << 21):
    """Reads that spell LF walks of a move table (so they match it except at the injected substitutions): start at a
    random (row, offset), emit the row character, step LF (LF_table.hpp:251-262), repeat; the walk runs right to left,
    so the emitted characters are reversed.  Walks that hit the terminator row keep going (its byte 1 becomes 'A').
    Returns host numpy (seqs u8, offsets u64)."""
    ch, lens, dest, doff = cols["ch"], cols["lens"], cols["dest"], cols["doff"]
    dev = ch.device
    r = ch.numel()
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = np.empty(n_reads * read_len, dtype=np.uint8)
    acgt = torch.tensor([ord(x) for x in "ACGT"], dtype=torch.uint8, device=dev)
    for a in range(0, n_reads, chunk_reads):
        m = min(chunk_reads, n_reads - a)
        k = torch.randint(0, r, (m,), generator=g, device=dev)
        o = (torch.rand(m, generator=g, device=dev, dtype=torch.float64) * lens[k].to(torch.float64)).to(torch.int64)
        buf = torch.empty((m, read_len), dtype=torch.uint8, device=dev)
        for j in range(read_len - 1, -1, -1):
            c = ch[k]
            buf[:, j] = torch.where(c <= 1, acgt[0].expand_as(c), c)
            o = doff[k] + o
            k = dest[k]
            while True:
                over = o >= lens[k]
                if not bool(over.any()):
                    break
                o = torch.where(over, o - lens[k], o)
                k = torch.where(over, k + 1, k)
        if sub > 0:
            hit = torch.rand((m, read_len), generator=g, device=dev) < sub
            rot = torch.randint(1, 4, (m, read_len), generator=g, device=dev)
            code = ((buf >> 1) & 3).to(torch.int64)                     # A0 C1 T2 G3
            lut = torch.tensor([ord(x) for x in "ACTG"], dtype=torch.uint8, device=dev)
            buf = torch.where(hit, lut[(code + rot) & 3], buf)
        out[a * read_len:(a + m) * read_len] = buf.reshape(-1).cpu().numpy()
    offsets = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    return out, offsets
