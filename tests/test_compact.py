"""Compact result form (include/colbwt_b200.h: colbwt_compact_*): the host-side expander against a numpy restatement
of the format, on CPU.  The GPU producer (compact.cu) is checked in test_gpu_parity.py."""
import numpy as np
import pytest

import col_bwt_b200 as cb

MAGIC = 0x31504d4354574243


def dense_from_match(match: np.ndarray, off: np.ndarray) -> np.ndarray:
    """PML as the reference computes it (col_bwt.hpp:510-528): right to left, ++ on a match, 0 on a mismatch."""
    pml = np.zeros(match.size, np.uint32)
    for i in range(off.size - 1):
        run = 0
        for j in range(int(off[i + 1]) - 1, int(off[i]) - 1, -1):
            run = run + 1 if match[j] else 0
            pml[j] = run
    return pml


def bits_to_words(bits: np.ndarray) -> np.ndarray:
    n_words = (bits.size + 31) // 32
    padded = np.zeros(n_words * 32, np.uint8)
    padded[: bits.size] = bits
    return np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view("<u4")


def build_compact(match: np.ndarray, cid: np.ndarray, off: np.ndarray, cuts) -> np.ndarray:
    """Numpy restatement of the buffer layout: header | directory | per segment fixed part | values."""
    segs, blobs = [], []
    at = (64 + 72 * (len(cuts) - 1) + 15) & ~15
    fixed = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        b0, b1 = int(off[a] - off[0]), int(off[b] - off[0])
        nb = b1 - b0
        n_words = (nb + 31) // 32
        n_groups = (n_words + 63) // 64
        wb = (n_words * 4 + 15) & ~15
        c = cid[b0:b1]
        mw, cw = bits_to_words(match[b0:b1]), bits_to_words(c != 0)
        nz = (c != 0).astype(np.int64)
        prefix = np.zeros(n_groups + 1, "<u4")
        csum = np.concatenate([[0], np.cumsum(nz)])
        for g in range(n_groups + 1):
            prefix[g] = csum[min(nb, 2048 * g)]
        blob = np.zeros(2 * wb + (((n_groups + 1) * 4 + 15) & ~15) if nb else 0, np.uint8)
        if nb:
            blob[: n_words * 4] = mw.view(np.uint8)
            blob[wb: wb + n_words * 4] = cw.view(np.uint8)
            blob[2 * wb: 2 * wb + (n_groups + 1) * 4] = prefix.view(np.uint8)
        segs.append([a, b - a, b0, nb, at, at + wb, at + 2 * wb, 0, int(nz.sum())])
        fixed.append(blob)
        at += blob.size
    for s, (a, b) in zip(segs, zip(cuts[:-1], cuts[1:])):
        b0, b1 = int(off[a] - off[0]), int(off[b] - off[0])
        c = cid[b0:b1]
        vals = c[c != 0]
        s[7] = at
        blobs.append(vals)
        at += (vals.size + 15) & ~15
    buf = np.zeros(at, np.uint8)
    hdr = np.array([MAGIC, len(segs), off.size - 1, int(off[-1] - off[0]), at, 0, 0, 0], "<u8")
    buf[:64] = hdr.view(np.uint8)
    buf[64: 64 + 72 * len(segs)] = np.array(segs, "<u8").reshape(-1).view(np.uint8)
    for s, f, v in zip(segs, fixed, blobs):
        buf[s[4]: s[4] + f.size] = f
        buf[s[7]: s[7] + v.size] = v
    return buf


def dense_from_match_fast(match: np.ndarray, off: np.ndarray) -> np.ndarray:
    """Same recurrence, vectorised: distance to the next stop (a mismatch, or the position after the read's last base)."""
    n = match.size
    stop = np.full(n + 1, n, np.int64)
    mis = np.flatnonzero(match == 0)
    stop[mis] = mis
    ends = off[1:][np.diff(off) > 0].astype(np.int64)          # position after the last base of every non-empty read
    is_end = np.zeros(n + 1, bool)
    is_end[ends] = True
    # a read end at e stops the runs of the bases left of it, but base e itself (first base of the next read) keeps its own value
    nxt_mis = np.minimum.accumulate(stop[::-1])[::-1][:n]
    e_pos = np.where(is_end, np.arange(n + 1), n + 1)
    nxt_end = np.minimum.accumulate(e_pos[::-1])[::-1]
    nxt_end = nxt_end[1:][:n]                                   # first read end strictly right of the base
    return (np.minimum(nxt_mis, nxt_end) - np.arange(n)).astype(np.uint32)


def test_expand_into_unaligned_and_skewed_arrays():
    """The expander's AVX-512 path writes whole 64-byte lines when the PML and chain-id arrays are line-aligned at the same
    positions, and goes through a window otherwise: result arrays at odd addresses, and skewed against each other."""
    import ctypes as C
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 400, size=500)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n = int(off[-1])
    match = (rng.random(n) < 0.8).astype(np.uint8)
    cid = np.where(rng.random(n) < 0.2, rng.integers(1, 256, size=n), 0).astype(np.uint8)
    want = dense_from_match(match, off)
    buf = build_compact(match, cid, off, [0, 7, 250, 500])
    for width, dt in ((1, np.uint8), (2, np.uint16), (4, np.uint32)):
        if width == 1:
            lens1 = np.minimum(lens, 255)
            off1 = np.concatenate([[0], np.cumsum(lens1)]).astype(np.uint64)
            m1, c1 = match[: int(off1[-1])], cid[: int(off1[-1])]
            cases = (off1, m1, c1, dense_from_match(m1, off1), build_compact(m1, c1, off1, [0, 7, 250, 500]))
        else:
            cases = (off, match, cid, want, buf)
        o, m, c, w, b = cases
        for skew_p, skew_c in ((0, 0), (width, 1), (0, 16), (3 * width, 3), (32, 0)):
            pml = cb._aligned_empty(int(o[-1]), dt, skew=skew_p)
            out_c = cb._aligned_empty(int(o[-1]), np.uint8, skew=skew_c)
            pml[:] = 0xAB
            out_c[:] = 0xEE
            rc = cb._L.colbwt_compact_expand(b.ctypes.data, o.ctypes.data, o.size - 1, pml.ctypes.data, width, out_c.ctypes.data)
            assert rc == 0
            assert np.array_equal(pml.astype(np.uint32), w) and np.array_equal(out_c, c), (width, skew_p, skew_c)


@pytest.mark.parametrize("no_avx512", ["0", "1"])
def test_expand_block_edges_and_long_runs(monkeypatch, no_avx512):
    """Reads that end exactly on 64-base block edges, runs longer than 65535, empty reads between them; both code paths of
    expand.cpp (AVX-512 where the CPU has it, and the portable table-driven sweep)."""
    monkeypatch.setenv("COLBWT_NO_AVX512", no_avx512)
    rng = np.random.default_rng(11)
    lens = [64, 64, 128, 1, 63, 0, 0, 65, 70000, 0, 4096, 8192, 191, 1, 1, 62, 100000, 3, 64 * 7, 5]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n = int(off[-1])
    match = (rng.random(n) < 0.9).astype(np.uint8)
    match[int(off[8]): int(off[9])] = 1                                   # 70000 matches in a row
    match[int(off[16]): int(off[16]) + 99000] = 1
    match[int(off[2]): int(off[3])] = 1
    cid = np.where(rng.random(n) < 0.5, rng.integers(1, 256, size=n), 0).astype(np.uint8)
    cid[int(off[10]): int(off[11])] = 7                                   # 4096 non-zero ids in a row (full expand masks)
    want = dense_from_match_fast(match, off)
    assert np.array_equal(want[: int(off[8])], dense_from_match(match[: int(off[8])], off[:9]))   # the fast restatement against the loop
    for cuts in ([0, len(lens)], [0, 3, 4, 9, 17, len(lens)]):
        buf = build_compact(match, cid, off, cuts)
        pml, got = cb.compact_expand(buf, off, 4)
        assert np.array_equal(pml, want) and np.array_equal(got, cid)
        assert int(pml.max()) >= 99000
        _, only = cb.compact_expand(buf, off, 4, cid_only=True)
        assert np.array_equal(only, cid)


@pytest.mark.parametrize("no_avx512", ["0", "1"])
@pytest.mark.parametrize("width", [1, 2, 4])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_expand_matches_reference_recurrence(width, seed, no_avx512, monkeypatch):
    monkeypatch.setenv("COLBWT_NO_AVX512", no_avx512)
    rng = np.random.default_rng(seed)
    max_len = 250 if width == 1 else 5000
    lens = rng.integers(0, max_len, size=300)
    lens[::17] = 0                                  # empty reads
    lens[5] = 1
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n = int(off[-1])
    p_match = [0.99, 0.64, 0.5][seed]
    match = (rng.random(n) < p_match).astype(np.uint8)
    match[int(off[40]): int(off[41])] = 1           # a read that matches end to end
    match[int(off[41]): int(off[42])] = 0           # and one that never does
    cid = np.where(rng.random(n) < [0.07, 0.6, 0.0][seed], rng.integers(1, 256, size=n), 0).astype(np.uint8)
    cuts = sorted(set([0, 1, 77, 78, 200, 300]))    # ragged segments, none aligned to 32 bases
    buf = build_compact(match, cid, off, cuts)
    pml, got_cid = cb.compact_expand(buf, off, width)
    assert np.array_equal(pml.astype(np.uint32), dense_from_match(match, off))
    assert np.array_equal(got_cid, cid)
    _, only_cid = cb.compact_expand(buf, off, width, cid_only=True)   # the group-wise path the chain-id transport uses (streamed stores)
    assert np.array_equal(only_cid, cid)
    assert [s["n_reads"] for s in cb.compact_segments(buf)] == [b - a for a, b in zip(cuts[:-1], cuts[1:])]


def test_expand_rejects_foreign_buffers():
    off = np.array([0, 10, 20], np.uint64)
    buf = build_compact(np.ones(20, np.uint8), np.zeros(20, np.uint8), off, [0, 2])
    with pytest.raises(cb.ColBwtError):
        cb.compact_expand(buf, np.array([0, 10, 21], np.uint64))       # other reads
    bad = buf.copy()
    bad[0] ^= 1
    with pytest.raises(cb.ColBwtError):
        cb.compact_expand(bad, off)
    long_off = np.array([0, 300], np.uint64)
    lb = build_compact(np.ones(300, np.uint8), np.zeros(300, np.uint8), long_off, [0, 1])
    with pytest.raises(cb.ColBwtError):
        cb.compact_expand(lb, long_off, 1)                               # 300 bases do not fit 8-bit PML
    pml, _ = cb.compact_expand(lb, long_off, 2)
    assert pml[0] == 300 and pml[-1] == 1


def test_bound_covers_the_worst_case():
    import ctypes as C
    off = np.arange(0, 151 * 1000, 150, dtype=np.uint64)
    n = int(off[-1])
    bound = cb._L.colbwt_compact_bound(off.ctypes.data, off.size - 1)
    assert bound >= 64 + 72 + n // 4 + n                                  # header, one segment, two bit arrays, every id non-zero
    assert bound < 64 + 72 * 4 + n // 4 + n + 4096
    assert cb._L.colbwt_compact_bound(np.array([5, 3], np.uint64).ctypes.data, 1) == 0
