"""Suffix array / BWT / LCP / thresholds / multi-MUMs for synthetic texts, in torch (CPU or CUDA).

Stand-in for the un-vendored `mumemto mum -K -R -T [-r]` step of `col-bwt build`
(/root/reference/scripts/col-bwt.py:120-148): it produces the same artefacts the in-tree readers
consume -- BWT run heads/lengths, per-run threshold positions, multi-MUM (len, bwt_pos) pairs.
Tooling for tests and bench only (index construction is out of the hot-path scope, SURVEY.md
section 8f); torch library ops are used freely here.

Algorithms
  suffix array : prefix doubling over packed k-mer keys, one torch.sort per round; the per-round
                 rank arrays are kept ("levels") ...
  LCP          : ... so LCP(SA[i-1], SA[i]) is exact by binary lifting over the levels plus a final
                 symbol-wise comparison of the packed k-mer keys.
  thresholds   : per BWT run of char c, the position of the minimum LCP in (end of previous c-run,
                 start of this c-run], leftmost on ties; 0 for the first run of a char (the
                 convention visible in SURVEY.md section 4.3's thr_pos vector).
  multi-MUMs   : windows of exactly N consecutive SA rows, N distinct documents, common prefix
                 >= min_len, not extendable right (neighbour LCPs smaller) nor left (BWT chars differ).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _codes(text: torch.Tensor):
    """Order-preserving dense recoding of bytes; returns (codes int64, sigma, bits)."""
    present = torch.zeros(256, dtype=torch.bool, device=text.device)
    present[text.long().unique()] = True
    lut = torch.cumsum(present.long(), 0) - 1
    sigma = int(present.sum())
    bits = max(1, math.ceil(math.log2(max(2, sigma))))
    return lut[text.long()], sigma, bits


def _shift(v: torch.Tensor, h: int, fill: int = 0) -> torch.Tensor:
    """v[i+h], `fill` past the end."""
    out = torch.full_like(v, fill)
    if h < v.numel():
        out[: v.numel() - h] = v[h:]
    return out


def _dense_rank(sorted_keys: torch.Tensor) -> torch.Tensor:
    r = torch.zeros_like(sorted_keys)
    r[1:] = torch.cumsum((sorted_keys[1:] != sorted_keys[:-1]).to(sorted_keys.dtype), 0)
    return r


class SuffixIndex:
    """SA, ISA, packed k-mer keys and doubling levels of one text (all on `device`)."""

    def __init__(self, text_u8: np.ndarray | torch.Tensor, device: str | torch.device = "cpu", keep_levels: bool = True):
        text = torch.as_tensor(text_u8, device=device)
        assert text.dtype == torch.uint8
        n = self.n = text.numel()
        assert n < 2**31, "rank pairs are packed into one int64 key"
        assert int(text[-1]) == 0 and int((text == 0).sum()) == 1, "text must end with a unique 0 byte"
        self.text = text
        codes, self.sigma, bits = _codes(text)
        self.bits = bits
        k = self.k = max(1, min(20, 60 // bits))
        key = torch.zeros(n, dtype=torch.int64, device=text.device)
        for j in range(k):
            key = (key << bits) | _shift(codes, j)
        del codes
        self.kmer = key
        sk, sa = torch.sort(key)
        rank = torch.empty_like(sa)
        rank[sa] = _dense_rank(sk)
        del sk
        self.levels: list[tuple[int, torch.Tensor]] = []
        h = k
        while True:
            if keep_levels:
                self.levels.append((h, rank.to(torch.int32)))
            if int(rank.max()) == n - 1:
                break
            key2 = (rank << 32) | _shift(rank + 1, h)
            sk, sa = torch.sort(key2)
            del key2
            rank = torch.empty_like(sa)
            rank[sa] = _dense_rank(sk)
            del sk
            h *= 2
        self.sa = sa
        self.isa = rank

    # -- derived arrays -------------------------------------------------------------------------
    def bwt(self) -> torch.Tensor:
        """BWT[i] = text[SA[i]-1] (cyclic: the row of suffix 0 gets the final terminator)."""
        return self.text[(self.sa - 1) % self.n]

    def lcp(self) -> torch.Tensor:
        """LCP[i] = lcp(suffix SA[i-1], suffix SA[i]) for i>=1, LCP[0]=0.  int64, exact."""
        n = self.n
        a, b = self.sa[1:], self.sa[:-1]
        acc = torch.zeros_like(a)
        for h, rk in reversed(self.levels):
            ai, bi = a + acc, b + acc
            ok = (ai < n) & (bi < n)
            eq = ok & (rk[ai.clamp(max=n - 1)] == rk[bi.clamp(max=n - 1)])
            acc += eq.to(acc.dtype) * h
        # remainder (< k symbols) from the packed k-mer keys
        ai, bi = (a + acc).clamp(max=n - 1), (b + acc).clamp(max=n - 1)
        x = self.kmer[ai] ^ self.kmer[bi]
        rem = torch.zeros_like(acc)
        alive = torch.ones_like(acc, dtype=torch.bool)
        mask = (1 << self.bits) - 1
        for j in range(self.k):
            sh = (self.k - 1 - j) * self.bits
            alive &= ((x >> sh) & mask) == 0
            rem += alive.to(rem.dtype)
        acc += rem
        out = torch.zeros(n, dtype=torch.int64, device=a.device)
        out[1:] = acc
        return out


def lcp_irreducible(si: "SuffixIndex", bwt: torch.Tensor) -> torch.Tensor:
    """The same LCP array without the doubling levels (which cost 4 bytes x n x ~14 rounds: the limit on text size).

    LCP[i] is *reducible* when both suffixes are preceded by the same text character (BWT[i] == BWT[i-1]); then
    PLCP[SA[i]] = PLCP[SA[i]-1] - 1 (Karkkainen, Manzini & Puglisi 2009), so in text order PLCP[j] + j is non-decreasing and
    constant over reducible positions.  Only the irreducible entries -- about one per BWT run -- are computed, by comparing
    the packed k-mer keys k symbols at a time; a running maximum in text order gives the rest.  int64, exact."""
    n, dev, k, bits = si.n, si.sa.device, si.k, si.bits
    sa = si.sa
    irr = torch.ones(n, dtype=torch.bool, device=dev)
    irr[1:] = (bwt[1:] != bwt[:-1]) | (sa[1:] == 0) | (sa[:-1] == 0)
    pos = torch.nonzero(irr).flatten()
    pos = pos[pos > 0]
    a, b = sa[pos], sa[pos - 1]
    acc = torch.zeros_like(a)
    alive = torch.arange(a.numel(), device=dev)
    mask = (1 << bits) - 1
    while alive.numel():
        ai = (a[alive] + acc[alive]).clamp(max=n - 1)
        bi = (b[alive] + acc[alive]).clamp(max=n - 1)
        x = si.kmer[ai] ^ si.kmer[bi]
        same = x == 0
        # a mismatch inside this k-mer: count the matching leading symbols
        xs = x[~same]
        rem = torch.zeros_like(xs)
        going = torch.ones_like(xs, dtype=torch.bool)
        for j in range(k):
            going &= ((xs >> ((k - 1 - j) * bits)) & mask) == 0
            rem += going.to(rem.dtype)
        acc[alive[~same]] += rem
        acc[alive[same]] += k
        alive = alive[same]
    # text order: A[j] = PLCP[j] + j at irreducible j, running maximum elsewhere (SA row 0 has LCP 0 by definition)
    A = torch.full((n,), -1, dtype=torch.int64, device=dev)
    A[a] = acc + a
    A[sa[0]] = sa[0]
    del acc, alive, a, b, pos, irr
    A = torch.cummax(A, 0).values
    out = A[sa] - sa
    out[0] = 0
    return out


def bwt_runs(bwt: torch.Tensor):
    """Run-length encode the BWT with terminator bytes (<=1) folded to 1 (col_bwt.hpp:171).
    Returns (heads u8, starts int64, lens int64)."""
    b = torch.where(bwt <= 1, torch.ones_like(bwt), bwt)
    n = b.numel()
    head = torch.ones(n, dtype=torch.bool, device=b.device)
    head[1:] = b[1:] != b[:-1]
    starts = torch.nonzero(head).flatten()
    lens = torch.diff(starts, append=torch.tensor([n], device=b.device))
    return b[starts], starts, lens


def thresholds(bwt: torch.Tensor, lcp: torch.Tensor, heads: torch.Tensor, starts: torch.Tensor) -> torch.Tensor:
    """Per-run threshold positions (`.thr_pos`, col_bwt.hpp:440-457)."""
    n = bwt.numel()
    dev = bwt.device
    b = torch.where(bwt <= 1, torch.ones_like(bwt), bwt)
    thr = torch.zeros(heads.numel(), dtype=torch.int64, device=dev)
    pos = torch.arange(n, dtype=torch.int64, device=dev)
    prev_same = torch.zeros(n, dtype=torch.bool, device=dev)
    prev_same[1:] = b[1:] == b[:-1]
    key = lcp * n + pos  # value-major, position-minor: amin gives the leftmost minimum
    big = torch.iinfo(torch.int64).max
    for c in heads.unique().tolist():
        is_c = b == c
        run_sel = torch.nonzero(heads == c).flatten()          # BWT runs of char c, in order
        c_start = torch.zeros(n, dtype=torch.int64, device=dev)
        c_start[starts[run_sel]] = 1
        # segment k (k>=1) = positions after the end of c-run k-1 up to and including the start of c-run k
        seg = torch.cumsum(c_start, 0) - c_start               # number of c-run starts strictly before p
        inside = is_c & prev_same                               # interior of a c-run: belongs to no segment
        seg = torch.where(inside, torch.full_like(seg, run_sel.numel()), seg)
        best = torch.full((run_sel.numel() + 1,), big, dtype=torch.int64, device=dev)
        best.scatter_reduce_(0, seg, key, reduce="amin", include_self=True)
        t = best[: run_sel.numel()] % n
        t[0] = 0                                                # first run of a char: convention 0
        thr[run_sel] = t
    return thr


def multi_mums(si: SuffixIndex, lcp: torch.Tensor, bwt: torch.Tensor, seq_starts: np.ndarray, doc_of_seq: np.ndarray,
               num_docs: int, min_len: int = 20):
    """Multi-MUMs as (len, first SA row) sorted by row -- the `.col_mums` payload (col_split.cpp:90-106)."""
    n, dev, N = si.n, si.sa.device, int(num_docs)
    if N < 2:
        raise ValueError("need >= 2 documents")
    ss = torch.as_tensor(seq_starts, device=dev)
    # clip LCPs so that matches never run across a separator: distance from SA[i] to its sequence end
    seq_of = torch.searchsorted(ss, si.sa, right=True) - 1
    room = ss[seq_of + 1] - 1 - si.sa
    room_prev = torch.zeros_like(room)
    room_prev[1:] = room[:-1]
    l = torch.minimum(lcp, torch.minimum(room, room_prev))
    del room, room_prev
    # sliding minimum of l over rows i+1 .. i+N-1  (window of N-1 values)
    w = N - 1
    big = torch.iinfo(torch.int64).max
    lpad = torch.cat([l, torch.zeros(N + 1, dtype=l.dtype, device=dev)])
    m = lpad.clone()
    span = 1
    while span * 2 <= w:
        m = torch.minimum(m, _shift(m, span, big))
        span *= 2
    if span < w:
        m = torch.minimum(m, _shift(m, w - span, big))
    inner = m[1: n + 1]                      # inner[i] = min l[i+1 .. i+N-1]
    left = l                                 # l[i]   : lcp with the row above the window
    right = lpad[N: n + N]                   # l[i+N] : lcp with the row below the window (0 past the end)
    sel = inner >= min_len
    sel &= left < inner
    sel &= right < inner
    cand = torch.nonzero(sel).flatten()
    del sel
    cand = cand[cand + N <= n]
    if cand.numel() == 0:
        z = np.zeros(0, dtype=np.uint64)
        return z, z
    rows = cand[:, None] + torch.arange(N, device=dev)[None, :]
    docs = torch.as_tensor(doc_of_seq, device=dev).long()[seq_of[rows]]
    ds, _ = torch.sort(docs, dim=1)
    distinct = (ds[:, 1:] != ds[:, :-1]).all(dim=1)
    bw = bwt[rows]
    left_max = (bw != bw[:, :1]).any(dim=1) | (bw <= 1).any(dim=1)
    keep = distinct & left_max
    cand = cand[keep]
    return inner[cand].cpu().numpy().astype(np.uint64), cand.cpu().numpy().astype(np.uint64)
