"""CPU: the product's device logic (colbwt_core.cuh row builder + lane state machine, pack.cpp), run lane by lane
on the host by tests/native/emulate.cpp, against the oracle.  This is how kernel logic is debugged without a GPU;
the GPU parity tests proper are in test_gpu_parity.py."""
import os

import numpy as np
import pytest

import oracle
from emu import Emu
from synthdata import formats as F, pangenome as P, pipeline as PL
from util import adversarial_reads, check_pml_properties, concat_reads, parse_fastx


def cols_from_file(path):
    meta, rows = F.read_col_pml(path)
    return {"ch": rows["ch"], "idx": F.u40_unpack(rows["idx"]), "interval": rows["interval"], "offset": rows["offset"],
            "col_id": rows["col_id"], "thr": F.u40_unpack(rows["thr"]), "n": meta["n"], "bwt_r": meta["bwt_r"]}


@pytest.mark.parametrize("case", ["toy", "pan4", "pan4all"])
@pytest.mark.parametrize("width", [2, 4])
@pytest.mark.parametrize("force_bytes", [False, True])
@pytest.mark.parametrize("narrow", [False, True])
def test_emulated_kernel_logic_on_golden(golden_dir, case, width, force_bytes, narrow):
    path = os.path.join(golden_dir, f"{case}.col_pml")
    orc = oracle.Oracle(path)
    emu = Emu(cols_from_file(path))
    ids, seqs, off = parse_fastx(os.path.join(golden_dir, f"{case}_reads.fa"))
    p0, c0 = orc.query_batch(seqs, off)
    p1, c1 = emu.query(seqs, off, pml_width=width, force_bytes=force_bytes, narrow=narrow)
    assert np.array_equal(p0, p1) and np.array_equal(c0, c1)


@pytest.mark.parametrize("width", [1, 2, 4])
@pytest.mark.parametrize("narrow", [False, True])
def test_emulated_cooperative_flush(small_index, golden_dir, width, narrow):
    """k_traverse's output path: flush requests served word by word (flush_word), ragged read edges included."""
    orc = oracle.Oracle(small_index["path"])
    emu = Emu(small_index["cols"])
    sets = [(small_index["seqs"], small_index["off"])]
    if width > 1:
        sets.append(concat_reads(adversarial_reads(small_index["haps"])))
    for seqs, off in sets:
        p0, c0 = orc.query_batch(seqs, off)
        for fb in (False, True):
            p1, c1 = emu.query(seqs, off, pml_width=width, force_bytes=fb, narrow=narrow, defer=True)
            assert np.array_equal(p0, p1) and np.array_equal(c0, c1)
    if width > 1:
        path = os.path.join(golden_dir, "pan4.col_pml")
        ids, seqs, off = parse_fastx(os.path.join(golden_dir, "pan4_reads.fa"))
        p0, c0 = oracle.Oracle(path).query_batch(seqs, off)
        p1, c1 = Emu(cols_from_file(path)).query(seqs, off, pml_width=width, narrow=narrow, defer=True)
        assert np.array_equal(p0, p1) and np.array_equal(c0, c1)


@pytest.mark.parametrize("width", [1, 2, 4])
def test_emulated_in_trip_variant(small_index, golden_dir, width):
    """k_traverse<..., INTRIP> (launched for tables of 2 GiB and more): fast-forward neighbours and reposition targets in the
    gathered row's own 128-byte line are resolved inside the trip.  Same results, fewer lane iterations."""
    orc = oracle.Oracle(small_index["path"])
    emu = Emu(small_index["cols"])
    sets = [(small_index["seqs"], small_index["off"])]
    if width > 1:
        sets.append(concat_reads(adversarial_reads(small_index["haps"])))
    for seqs, off in sets:
        p0, c0 = orc.query_batch(seqs, off)
        for fb in (False, True):
            p1, c1 = emu.query(seqs, off, pml_width=width, force_bytes=fb, defer=True)
            base_iters = emu.iters
            p2, c2 = emu.query(seqs, off, pml_width=width, force_bytes=fb, intrip=True)
            assert np.array_equal(p0, p2) and np.array_equal(c0, c2) and np.array_equal(p1, p2)
            assert emu.iters < base_iters
    if width > 1:
        path = os.path.join(golden_dir, "pan4.col_pml")
        ids, seqs, off = parse_fastx(os.path.join(golden_dir, "pan4_reads.fa"))
        p0, c0 = oracle.Oracle(path).query_batch(seqs, off)
        p1, c1 = Emu(cols_from_file(path)).query(seqs, off, pml_width=width, intrip=True)
        assert np.array_equal(p0, p1) and np.array_equal(c0, c1)


def test_emulated_kernel_logic_on_synthetic(small_index):
    orc = oracle.Oracle(small_index["path"])
    emu = Emu(small_index["cols"])
    assert emu.flags & 1 == 0
    extra, eoff = concat_reads(adversarial_reads(small_index["haps"]))
    for seqs, off in ((small_index["seqs"], small_index["off"]), (extra, eoff)):
        p0, c0 = orc.query_batch(seqs, off)
        for fb in (False, True):
            for narrow in (False, True):
                p1, c1 = emu.query(seqs, off, force_bytes=fb, narrow=narrow)
                assert np.array_equal(p0, p1) and np.array_equal(c0, c1)
        check_pml_properties(p1, off)
    # 8-bit PML: legal when every read is shorter than 256 bases
    p0, c0 = orc.query_batch(small_index["seqs"], small_index["off"])
    p1, c1 = emu.query(small_index["seqs"], small_index["off"], pml_width=1)
    assert np.array_equal(p0, p1) and np.array_equal(c0, c1)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emulated_random_tables_small_alphabet_quirks(seed):
    """Degenerate tables: tiny genomes, high divergence, a haplotype with N runs and lower case, no marks / dense marks."""
    rng = np.random.default_rng(seed)
    haps = P.make_haplotypes(int(rng.integers(300, 3000)), int(rng.integers(2, 5)), snp=0.03, indel=0.005, seed=seed)
    if seed == 2:   # N run and lower-case stretch inside the text: rows with "other" characters
        h = haps[0].copy()
        h[50:60] = ord("N")
        h[100:120] = np.frombuffer(bytes(h[100:120]).lower(), np.uint8)
        haps[0] = h
    idx = PL.build_index(haps, split_rate=int(rng.integers(1, 4)), min_mum=8)
    cols = idx["columns"]
    orc = oracle.Oracle(columns=cols)
    emu = Emu(cols)
    seqs, off = P.sample_reads(idx["text"], idx["seq_starts"], 300, 80, sub=0.05, seed=seed)
    extra, eoff = concat_reads(adversarial_reads(haps))
    for s, o in ((seqs, off), (extra, eoff)):
        p0, c0 = orc.query_batch(s, o)
        for narrow in (False, True):
            p1, c1 = emu.query(s, o, narrow=narrow)
            assert np.array_equal(p0, p1) and np.array_equal(c0, c1)


def test_far_reposition_targets_take_exact_search():
    """A table where the next row of some character is > 31 rows away (distance fields overflow)."""
    rng = np.random.default_rng(5)
    # text: long stretch alternating A/C only, then G/T region => BWT has regions without G rows
    a = np.frombuffer(b"AC", np.uint8)[rng.integers(0, 2, 4000)]
    b = P.ACGT[rng.integers(0, 4, 500)]
    idx = PL.build_index([np.concatenate((a, b)), np.concatenate((a[::-1], b))], with_revcomp=False, marks=False)
    cols = idx["columns"]
    emu = Emu(cols)
    assert emu.slow_rows > 10
    orc = oracle.Oracle(columns=cols)
    reads = [bytes(a[100:200]) + b"G" + bytes(a[300:330]), b"GGGG" + bytes(a[5:50]) + b"T", bytes(b[10:90])]
    s, o = concat_reads(reads)
    p0, c0 = orc.query_batch(s, o)
    for narrow in (False, True):
        p1, c1 = emu.query(s, o, narrow=narrow)
        assert np.array_equal(p0, p1) and np.array_equal(c0, c1)


def test_row_builder_flags_overlong_rows():
    # one run of 70000 A's: a row longer than the 16-bit offset field can address
    text = np.concatenate((np.full(70000, ord("A"), np.uint8), np.frombuffer(b"CGT\x00", np.uint8)))
    from synthdata import bwtbuild as B, table as T
    si = B.SuffixIndex(text)
    bwt = si.bwt()
    heads, starts, lens = B.bwt_runs(bwt)
    thr = B.thresholds(bwt, si.lcp(), heads, starts)
    cols = T.build_columns(heads.numpy(), lens.numpy(), thr.numpy(), starts.numpy(), np.zeros(starts.numel(), np.uint8), strict=False)
    assert Emu(cols).flags & 1


def test_host_packer_matches_scalar_definition():
    import ctypes as C
    from emu import SO, build
    build()
    L = C.CDLL(SO)
    L.emu_pack.restype = C.c_int
    L.emu_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    rng = np.random.default_rng(0)
    for n in (1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 150, 1000):
        s = P.ACGT[rng.integers(0, 4, n)]
        w = np.zeros((n + 15) // 16 + 2, np.uint32)
        assert L.emu_pack(s.ctypes.data, n, w.ctypes.data) == 1
        code = (s >> 1) & 3
        for j in range(n):
            assert (int(w[j >> 4]) >> (2 * (j & 15))) & 3 == code[j]
        for bad in (b"N", b"a", b"\x00", b"U"):
            for pos in (0, n // 2, n - 1):
                t = s.copy()
                t[pos] = bad[0]
                assert L.emu_pack(t.ctypes.data, n, w.ctypes.data) == 0


def test_slice_packer_overreads_safely_and_flags_irregular_reads():
    import ctypes as C
    from emu import SO, build
    build()
    L = C.CDLL(SO)
    L.emu_pack_slice.restype = C.c_uint64
    L.emu_pack_slice.argtypes = [C.c_void_p] * 2 + [C.c_uint64] + [C.c_void_p] * 3
    rng = np.random.default_rng(1)
    lens = [150, 1, 0, 16, 17, 31, 32, 33, 64, 149, 3, 150, 150, 7]      # the last reads sit at the very end of the buffer
    reads = [bytes(P.ACGT[rng.integers(0, 4, n)]) for n in lens]
    reads[4] = reads[4][:5] + b"N" + reads[4][6:]
    reads[9] = reads[9][:-1] + b"a"                                        # irregular in the last byte
    reads[12] = b"n" + reads[12][1:]
    # the byte right AFTER a regular read being irregular must not make that read irregular (over-read masking)
    reads[1] = b"C"
    seqs, off = concat_reads(reads)
    seqs = np.concatenate((seqs, np.zeros(0, np.uint8)))
    n = len(reads)
    words = np.zeros(int(((np.diff(off).astype(np.int64) + 15) // 16).sum()) + 4, np.uint32)
    meta = np.zeros(4 * n, np.uint64)
    irr = np.zeros(n, np.uint64)
    k = L.emu_pack_slice(seqs.ctypes.data, off.ctypes.data, n, words.ctypes.data, meta.ctypes.data, irr.ctypes.data)
    assert sorted(irr[:k].tolist()) == [4, 9, 12]
    meta = meta.reshape(n, 4)
    w = 0
    for i, r in enumerate(reads):
        assert meta[i, 0] == off[i] and meta[i, 2] == w
        assert meta[i, 1] == (0 if i in (4, 9, 12) else len(r))
        if i not in (4, 9, 12):
            for j, c in enumerate(r):
                assert (int(words[w + (j >> 4)]) >> (2 * (j & 15))) & 3 == (c >> 1) & 3
        w += (len(r) + 15) // 16


def test_synthetic_move_table_with_many_threshold_conflicts():
    """Directly synthesised table (random thresholds): ~20 % of rows hold two thresholds, i.e. mode-3 slots whose
    neighbours are known -- the exact-compare shortcut.  Reads are LF walks with substitutions."""
    rows, n, cols = PL.synth_move_table(30000, mean_len=8, device="cpu", seed=9)
    r = np.frombuffer(rows.numpy().tobytes(), dtype=F.ROW_DTYPE)
    cd = {"ch": r["ch"], "idx": F.u40_unpack(r["idx"]), "interval": r["interval"], "offset": r["offset"], "col_id": r["col_id"],
          "thr": F.u40_unpack(r["thr"]), "n": n, "bwt_r": len(r)}
    emu = Emu(cd)
    assert emu.slow_rows > 1000
    orc = oracle.Oracle(columns=cd)
    seqs, off = PL.walk_reads(cols, 400, 150, sub=0.02)
    p0, c0 = orc.query_batch(seqs, off)
    assert 0.05 < (p0 == 0).mean() < 0.5
    for narrow in (False, True):
        p1, c1 = emu.query(seqs, off, narrow=narrow)
        assert np.array_equal(p0, p1) and np.array_equal(c0, c1)


@pytest.mark.parametrize("errors", [(0.02, 0.015, 0.015), (0.002, 0.0, 0.0)])
def test_long_reads_chunk_tasks_with_chain_verification_are_exact(small_index, errors):
    """Long reads are cut into speculative chunk tasks (warm-up from the initial state) and repaired by fixup_chain;
    whatever the chunk / warm-up lengths -- including warm-up 0, where almost every chunk must be re-traversed -- the
    result equals the serial traversal.  Tasks are run in reverse scheduling order to show order independence."""
    idx = small_index["idx"]
    sub, ins, dele = errors
    seqs, off = P.sample_reads(idx["text"], idx["seq_starts"], 30, 3000, sub=sub, ins=ins, dele=dele, seed=6, len_jitter=0.4)
    rd = [bytes(seqs[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    rd[3] = rd[3][:700] + b"NNNN" + rd[3][704:]          # long irregular reads take the byte path
    rd[5] = rd[5].lower()
    rd += adversarial_reads(small_index["haps"])
    seqs, off = concat_reads(rd)
    orc = oracle.Oracle(small_index["path"])
    emu = Emu(small_index["cols"])
    p0, c0 = orc.query_batch(seqs, off)
    redone = {}
    for chunk, warm in ((64, 0), (64, 16), (100, 64), (256, 256), (500, 1000)):
        for narrow in (False, True):
            for width in (2, 4):
                p1, c1 = emu.query_split(seqs, off, chunk, warm, pml_width=width, narrow=narrow)
                assert np.array_equal(p0, p1) and np.array_equal(c0, c1), (chunk, warm, narrow, width)
        redone[(chunk, warm)] = emu.redone / max(1, emu.tasks)
    assert redone[(64, 0)] > 0.9 and redone[(256, 256)] < 0.5 * redone[(64, 16)]


@pytest.mark.parametrize("length,chunk,warm", [(8192, 4096, 512), (8193, 4096, 512), (100000, 4096, 512), (70000, 64, 16), (129, 64, 0), (5000, 4096, 9999)])
def test_chunk_planner_tiles_the_read(length, chunk, warm):
    """tasks.h: chunks are listed top first, tile [0, len) exactly, are balanced, and never warm up past the read."""
    import ctypes as C
    from emu import SO, build
    build()
    L = C.CDLL(SO)
    L.emu_plan.restype = C.c_uint32
    L.emu_plan.argtypes = [C.c_uint32] * 3 + [C.c_void_p, C.c_uint32]
    out = np.zeros(4 * 4096, np.uint32)
    n = L.emu_plan(length, chunk, warm, out.ctypes.data, 4096)
    t = out[: 4 * n].reshape(n, 4)
    assert n == -(-length // chunk)
    assert t[0, 1] == length and t[0, 2] == length                       # top chunk: no warm-up, starts from the true state
    assert t[-1, 0] == 0
    assert (t[1:, 1] == t[:-1, 0]).all()                                 # contiguous, descending
    assert (t[:, 1] > t[:, 0]).all() and (t[:, 1] - t[:, 0]).max() - (t[:, 1] - t[:, 0]).min() <= n
    assert (t[1:, 2] == np.minimum(length, t[1:, 1] + warm)).all()
    assert (t[:, 3] == np.arange(n)).all()


def test_chunk_geometry_follows_the_batch():
    """tasks.h: SplitParams::adapted -- about six tasks per lane, chunk within [1024, 4096], shorter warm-up with shorter
    chunks, reads of 1.5 chunks or more are cut, shorter ones above chunk/8 are scheduled whole among the tasks."""
    import ctypes as C
    from emu import SO, build
    build()
    L = C.CDLL(SO)
    L.emu_adapted.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    out = np.zeros(4, np.uint32)

    def geo(bases, lanes):
        L.emu_adapted(bases, lanes, out.ctypes.data)
        return tuple(int(x) for x in out)
    lanes = 148 * 1024
    assert geo(4_000_000_000, lanes) == (4096, 512, 6144, 512)          # c3small: plenty of work per lane
    assert geo(1_245_000_000, lanes) == (1536, 256, 2304, 192)          # configs[2] per GPU: 8.2 kbases per lane
    assert geo(300_000_000, lanes) == (1024, 256, 1536, 128)            # small batch: floor
    c, w, m, wm = geo(2_000_000_000, lanes)
    assert c % 256 == 0 and 2048 <= c <= 2560 and w == 256 and m == c + c // 2


def test_streaming_chunks_follow_the_reads(monkeypatch):
    """tasks.h: chunk_geometry + plan_chunks -- 96-Mbase chunks for short reads; where a chunk would hold fewer than 32 Ki reads
    (long reads, also behind millions of short ones in the same batch) it grows, up to 512 Mbases; chunks tile the batch."""
    import ctypes as C
    from emu import SO, build
    build()
    L = C.CDLL(SO)
    L.emu_plan_chunks.restype = C.c_uint64
    L.emu_plan_chunks.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    monkeypatch.delenv("COLBWT_CHUNK_BASES", raising=False)

    def plan(lens, staged=0):
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        out, caps = np.zeros(4096, np.uint64), np.zeros(2, np.uint64)
        n = int(L.emu_plan_chunks(off.ctypes.data, off.size - 1, staged, out.ctypes.data, out.size, caps.ctypes.data))
        cuts = out[: n + 1].astype(np.int64)
        assert cuts[0] == 0 and cuts[-1] == off.size - 1 and (np.diff(cuts) > 0).all()      # tiles the batch, no empty chunk
        bases = np.diff(off[cuts].astype(np.int64))
        assert bases.max() == int(caps[0]) and np.diff(cuts).max() == int(caps[1])
        return cuts, bases
    M = 1 << 20
    cuts, bases = plan(np.full(3_000_000, 150))                              # short reads: 96 Mbases per chunk
    assert len(bases) == 5 and (bases[:-1] <= 96 * M).all() and (bases[:-1] > 96 * M - 150).all()
    cuts, bases = plan(np.full(100_000, 10_000))                             # long reads: 32 Ki reads per chunk
    assert (np.diff(cuts)[:-1] == 32768).all() and bases.max() == 327_680_000
    cuts, bases = plan(np.full(5_000, 200_000))                              # very long reads: capped at 512 Mbases
    assert bases.max() <= 512 * M and bases[:-1].min() > 512 * M - 200_000
    lens = np.concatenate([np.full(2_000_000, 150), np.full(40_000, 10_000)])   # mixed batch (configs[4]): the long tail gets long-read chunks
    cuts, bases = plan(lens)
    assert (bases[:2] <= 96 * M).all() and bases[-2:].sum() > 300 * M and len(bases) <= 6
    cuts, bases = plan(np.full(100_000, 10_000), staged=5)                   # pageable results: staging bounded to 256 MB per slot
    assert bases.max() * 5 <= 256 * M
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "50000")                        # tests pin tiny chunks: no growth
    cuts, bases = plan(np.full(1000, 3000))
    assert bases.max() <= 51_000 and len(bases) >= 59


def test_mode_choice_is_measured_then_kept():
    """tasks.h: choose_mode -- the rule first, every other allowed mode once, then the fastest (the rule keeps its place
    within 5 %).  Modes of colbwt_query: bit 0 = reads packed on the device, bit 1 = compact transport."""
    import ctypes as C
    from emu import build, SO
    build()
    L = C.CDLL(SO)
    L.emu_choose_mode.argtypes = [C.c_int, C.c_uint32, C.POINTER(C.c_double), C.c_int, C.c_int]

    def f(rule, rates, allowed=0b1111, large=1):
        arr = (C.c_double * 4)(*rates)
        return L.emu_choose_mode(rule, allowed, arr, 4, large)
    for rule in (0, 1):
        assert f(rule, [0, 0, 0, 0]) == rule                      # nothing measured: the rule
        assert f(rule, [0, 0, 0, 0], large=0) == rule
        known = [0, 0, 0, 0]
        known[rule] = 20e9
        first_other = 1 - rule
        assert f(rule, known) == first_other                      # the rule was measured: probe the next unmeasured mode
        assert f(rule, known, large=0) == rule                    # small calls never probe
        known[first_other] = 5e9
        assert f(rule, known) == 2 and f(rule, known, allowed=0b0011) == rule   # only allowed modes are probed
        assert f(rule, [20e9, 17e9, 30e9, 12e9]) == 2             # all measured: the fastest
        r = [20e9] * 4
        r[3] = 20.9e9
        assert f(rule, r) == rule                                 # within 5 %: stay with the rule
        r[3] = 22e9
        assert f(rule, r) == 3 and f(rule, r, large=0) == 3       # small calls follow what the large ones found
    # a rule this call cannot use (no pinned input -> no device packing): the host-packing modes are what is left
    assert f(1, [0, 0, 0, 0], allowed=0b0101) == 0
    assert f(1, [9e9, 0, 0, 0], allowed=0b0101) == 2
    assert f(1, [9e9, 0, 12e9, 0], allowed=0b0101) == 2 and f(1, [9e9, 0, 9.2e9, 0], allowed=0b0101) == 0

    # six modes (mode = packing bit + 2 x transport); few host threads start from mode 3 = device packing + compact transport;
    # a call that cannot pack on the device (long reads, pageable input) keeps the rule's TRANSPORT: mode 2
    def g(rule, rates, allowed):
        arr = (C.c_double * 6)(*rates)
        return L.emu_choose_mode(rule, allowed, arr, 6, 1)
    assert g(3, [0] * 6, 0b111111) == 3
    assert g(3, [0] * 6, 0b010101) == 2
    assert g(3, [0, 0, 8e9, 0, 0, 0], 0b010101) == 0              # then the other allowed modes, once each
    assert g(3, [7e9, 0, 8e9, 0, 0, 0], 0b010101) == 4
    assert g(3, [7e9, 0, 8e9, 0, 8.3e9, 0], 0b010101) == 2        # 4 % better is not enough to displace the rule's stand-in
