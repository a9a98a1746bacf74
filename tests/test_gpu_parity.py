"""GPU (-m gpu): parity of the CUDA path with the oracle, through the C-ABI (col_bwt_b200 -> libcolbwt_b200.so).

Bit-exact is the bar: PML and chain ids are integers.  Reads nothing from /root/reference."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from synthdata import formats as F, pangenome as P, pipeline as PL
from util import adversarial_reads, check_pml_properties, concat_reads, parse_fastx

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cb():
    import col_bwt_b200
    return col_bwt_b200


@pytest.mark.parametrize("case", ["toy", "pan4", "pan4all"])
def test_golden_text_matches_reference_pml_query(cb, golden_dir, case):
    """Golden .pml/.cid are the bytes the reference's pml_query wrote; format through the C-ABI and compare bytes."""
    tbl = cb.ColPml.load(os.path.join(golden_dir, f"{case}.col_pml"))
    ids, seqs, off = parse_fastx(os.path.join(golden_dir, f"{case}_reads.fa"))
    for width in (cb.PML_U16, cb.PML_U32):
        pml, cid = tbl.query(seqs, off, width)
        txt_p = b"".join(cb.format_stats(ids[i], pml[int(off[i]):int(off[i + 1])]) for i in range(len(ids)))
        txt_c = b"".join(cb.format_stats(ids[i], cid[int(off[i]):int(off[i + 1])]) for i in range(len(ids)))
        assert txt_p == open(os.path.join(golden_dir, f"{case}_reads.fa.pml"), "rb").read()
        assert txt_c == open(os.path.join(golden_dir, f"{case}_reads.fa.cid"), "rb").read()


def test_query_pml_single_read_api(cb, golden_dir):
    tbl = cb.ColPml.load(os.path.join(golden_dir, "toy"))      # prefix form, as pml_query.cpp:110-111
    assert (tbl.n, tbl.r, tbl.bwt_r) == (43, 28, 22)
    pml, cid = tbl.query_pml("GATTACAGAT")
    assert pml.tolist() == [6, 5, 4, 3, 2, 1, 0, 2, 1, 0] and cid.tolist() == [0, 10, 0, 0, 6, 4, 0, 0, 0, 0]
    pml, cid = tbl.query_pml(b"TTAGGATNACA")
    assert pml.tolist() == [1, 0, 1, 0, 3, 2, 1, 0, 1, 0, 1] and cid.tolist() == [0, 0, 0, 7, 5, 9, 0, 0, 0, 0, 0]
    pml, cid = tbl.query_pml(b"")
    assert pml.size == 0 and cid.size == 0


def test_synthetic_index_vs_oracle_all_paths(cb, small_index):
    orc = oracle.Oracle(small_index["path"])
    tbl = cb.ColPml.load(small_index["path"])
    assert tbl.r == len(small_index["cols"]["ch"])
    extra, eoff = concat_reads(adversarial_reads(small_index["haps"]))
    for seqs, off in ((small_index["seqs"], small_index["off"]), (extra, eoff)):
        want_p, want_c = orc.query_batch(seqs, off)
        widths = (cb.PML_U8, cb.PML_U16, cb.PML_U32) if int(np.diff(off).max()) < 256 else (cb.PML_U16, cb.PML_U32)
        for width in widths:
            pml, cid = tbl.query(seqs, off, width)
            assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
            b = tbl.batch(seqs, off, width)
            b.run(2)
            p2, c2 = b.download()
            assert np.array_equal(p2.astype(np.uint32), want_p) and np.array_equal(c2, want_c)
            b.close()
        check_pml_properties(pml, off)


def test_mixed_batch_regular_and_irregular_reads_keep_input_order(cb, small_index):
    rng = np.random.default_rng(4)
    seqs = small_index["seqs"].copy()
    off = small_index["off"]
    # sprinkle N / lower case into ~10 % of the reads
    for i in rng.choice(len(off) - 1, (len(off) - 1) // 10, replace=False):
        a, b = int(off[i]), int(off[i + 1])
        seqs[a + int(rng.integers(0, b - a))] = ord("N") if rng.random() < 0.5 else ord("a")
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"])
    pml, cid = tbl.query(seqs, off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


def test_streaming_chunks_and_pinned_buffers(cb, small_index, monkeypatch):
    """Many small chunks through the multi-slot pipeline, pageable and pinned output buffers."""
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "20000")
    seqs, off = small_index["seqs"], small_index["off"]
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"])
    pml, cid = tbl.query(seqs, off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
    pp, pc = cb.PinnedArray(seqs.size, np.uint16), cb.PinnedArray(seqs.size, np.uint8)
    tbl.query(seqs, off, out=(pp.array, pc.array))
    assert np.array_equal(pp.array.astype(np.uint32), want_p) and np.array_equal(pc.array, want_c)
    # a second call reuses the cached pipeline
    pml, cid = tbl.query(seqs, off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


def test_variable_lengths_and_long_reads_u32(cb, small_index):
    """Reads from 1 base to > 65535 bases in one batch (PML needs 32 bits)."""
    idx = small_index["idx"]
    rng = np.random.default_rng(8)
    hap = np.concatenate([small_index["haps"][k % 4] for k in range(3)])
    long_read = hap[:70000].copy()
    pos = rng.integers(0, long_read.size, 700)
    long_read[pos] = P.ACGT[rng.integers(0, 4, 700)]
    reads = [bytes(long_read)]
    s, o = P.sample_reads(idx["text"], idx["seq_starts"], 400, 600, sub=0.03, seed=5, len_jitter=0.8)
    reads += [bytes(s[int(o[i]):int(o[i + 1])]) for i in range(len(o) - 1)]
    reads += [b"A", b"", b"CG"]
    seqs, off = concat_reads(reads)
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"])
    pml, cid = tbl.query(seqs, off, cb.PML_U32)
    assert np.array_equal(pml, want_p) and np.array_equal(cid, want_c)
    for too_narrow in (cb.PML_U16, cb.PML_U8):
        with pytest.raises(cb.ColBwtError) as e:
            tbl.query(seqs, off, too_narrow)
        assert e.value.code == -5


def test_nanopore_like_reads_with_indels(cb, small_index):
    idx = small_index["idx"]
    seqs, off = P.sample_reads(idx["text"], idx["seq_starts"], 60, 4000, sub=0.02, ins=0.015, dele=0.015, seed=6, len_jitter=0.3)
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"])
    pml, cid = tbl.query(seqs, off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


def test_table_with_other_characters_and_far_targets(cb, tmp_path):
    """Rows carrying N / lower case, and regions where a character's next row is > 31 rows away."""
    rng = np.random.default_rng(5)
    a = np.frombuffer(b"AC", np.uint8)[rng.integers(0, 2, 4000)]
    b = P.ACGT[rng.integers(0, 4, 500)].copy()
    b[100:110] = ord("N")
    b[200:230] = np.frombuffer(bytes(b[200:230]).lower(), np.uint8)
    idx = PL.build_index([np.concatenate((a, b)), np.concatenate((a[::-1], b))], with_revcomp=False, marks=False)
    path = str(tmp_path / "odd.col_pml")
    PL.write_col_pml(path, idx["columns"])
    reads = [bytes(a[100:200]) + b"G" + bytes(a[300:330]), b"GGGG" + bytes(a[5:50]) + b"T", bytes(b[10:300]), bytes(b[90:120]),
             b"NNNNNNNNNNNN", bytes(b[195:240])] + adversarial_reads()
    seqs, off = concat_reads(reads)
    want_p, want_c = oracle.Oracle(path).query_batch(seqs, off)
    tbl = cb.ColPml.load(path)
    assert tbl.stats.slow_rows > 10
    pml, cid = tbl.query(seqs, off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


def test_from_rows_equals_load(cb, small_index):
    meta, rows = F.read_col_pml(small_index["path"])
    t1 = cb.ColPml.from_rows(rows, meta["bwt_r"], meta["n"])
    t2 = cb.ColPml.load(small_index["path"])
    seqs, off = small_index["seqs"][:30000], small_index["off"][:201]
    p1, c1 = t1.query(seqs, off)
    p2, c2 = t2.query(seqs, off)
    assert np.array_equal(p1, p2) and np.array_equal(c1, c2)
    assert t1.stats.marked_rows == int((small_index["cols"]["col_id"] > 0).sum())


def test_error_codes(cb, tmp_path, golden_dir):
    with pytest.raises(cb.ColBwtError) as e:
        cb.ColPml.load(str(tmp_path / "missing"))
    assert e.value.code == -1
    data = open(os.path.join(golden_dir, "toy.col_pml"), "rb").read()
    (tmp_path / "short.col_pml").write_bytes(data[:200])
    with pytest.raises(cb.ColBwtError) as e:
        cb.ColPml.load(str(tmp_path / "short.col_pml"))
    assert e.value.code == -1
    hdr = np.frombuffer(data[:32], "<u8").copy()
    hdr[3] += 1                                                   # size != r
    (tmp_path / "bad.col_pml").write_bytes(hdr.tobytes() + data[32:])
    with pytest.raises(cb.ColBwtError) as e:
        cb.ColPml.load(str(tmp_path / "bad.col_pml"))
    assert e.value.code == -2
    # a row of 70000 symbols cannot be addressed by the reference's 16-bit offset field
    from synthdata import bwtbuild as B, table as T
    text = np.concatenate((np.full(70000, ord("A"), np.uint8), np.frombuffer(b"CGT\x00", np.uint8)))
    si = B.SuffixIndex(text)
    bwt = si.bwt()
    heads, starts, lens = B.bwt_runs(bwt)
    thr = B.thresholds(bwt, si.lcp(), heads, starts)
    cols = T.build_columns(heads.numpy(), lens.numpy(), thr.numpy(), starts.numpy(), np.zeros(starts.numel(), np.uint8), strict=False)
    rows = F.rows_from_columns(cols["ch"], cols["idx"], cols["interval"], cols["offset"], cols["col_id"], cols["thr"])
    with pytest.raises(cb.ColBwtError) as e:
        cb.ColPml.from_rows(rows, cols["bwt_r"], cols["n"])
    assert e.value.code == -3
    # malformed batches are refused, not traversed
    tbl = cb.ColPml.load(os.path.join(golden_dir, "toy"))
    seqs = np.frombuffer(b"ACGTACGTAC", np.uint8)
    for call in (lambda: tbl.query(seqs, np.array([0, 6, 4, 10], np.uint64)),
                 lambda: tbl.batch(seqs, np.array([0, 6, 4, 10], np.uint64))):
        with pytest.raises(cb.ColBwtError) as e:
            call()
        assert e.value.code == -5 and "non-decreasing" in str(e.value)
    p, c = tbl.query(seqs, np.array([0, 4, 4, 10], np.uint64))    # an empty read in the middle is fine
    assert p.size == 10


def test_cli_is_a_drop_in_for_pml_query(golden_dir, tmp_path):
    """pml_query_b200 <prefix> -p <reads> writes the same two text files as the reference's pml_query."""
    import shutil
    for f in ("pan4.col_pml", "pan4_reads.fa"):
        shutil.copy(os.path.join(golden_dir, f), tmp_path / f)
    cli = os.path.join(ROOT, "col_bwt_b200", "bin", "pml_query_b200")
    r = subprocess.run([cli, str(tmp_path / "pan4"), "-p", str(tmp_path / "pan4_reads.fa"), "-v"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "pan4_reads.fa.pml").read_bytes() == open(os.path.join(golden_dir, "pan4_reads.fa.pml"), "rb").read()
    assert (tmp_path / "pan4_reads.fa.cid").read_bytes() == open(os.path.join(golden_dir, "pan4_reads.fa.cid"), "rb").read()
    # gzip input, FASTQ records
    import gzip
    ids, seqs, off = parse_fastx(os.path.join(golden_dir, "pan4_reads.fa"))
    with gzip.open(tmp_path / "r.fq.gz", "wb") as f:
        for i, name in enumerate(ids[:40]):
            s = bytes(seqs[int(off[i]):int(off[i + 1])])
            f.write(b"@" + name.encode() + b" c\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    r = subprocess.run([cli, str(tmp_path / "pan4"), "-p", str(tmp_path / "r.fq.gz")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    golden = open(os.path.join(golden_dir, "pan4_reads.fa.pml"), "rb").read().split(b"\n")
    assert (tmp_path / "r.fq.gz.pml").read_bytes().split(b"\n")[:80] == golden[:80]
    # failure: non-zero exit code and no output files (col-bwt.py:70-77 behaviour)
    r = subprocess.run([cli, str(tmp_path / "nope"), "-p", str(tmp_path / "pan4_reads.fa")], capture_output=True, text=True)
    assert r.returncode != 0


def test_config1_scale_properties_and_oracle_sample(cb, tmp_path):
    """BASELINE.json configs[0] shape scaled to test time: 4 haplotypes x 250 kbp (+revcomp), tunnels -s 10,
    100k x 150 bp reads; full-size invariants on every base plus oracle parity on a read sample."""
    haps = P.make_haplotypes(250000, 4, snp=1e-3, seed=1)
    idx = PL.build_index(haps, split_rate=10)
    path = str(tmp_path / "c1.col_pml")
    PL.write_col_pml(path, idx["columns"])
    seqs, off = P.sample_reads(idx["text"], idx["seq_starts"], 100000, 150, sub=0.01, seed=2)
    tbl = cb.ColPml.load(path)
    pml, cid = tbl.query(seqs, off)
    p2d = pml.reshape(-1, 150).astype(np.int64)
    assert (p2d <= 150 - np.arange(150)[None, :]).all()
    nxt = np.concatenate((p2d[:, 1:], np.zeros((p2d.shape[0], 1), np.int64)), axis=1)
    assert ((p2d == 0) | (p2d == nxt + 1)).all()
    assert cid.max() > 0 and set(np.unique(cid)) <= set(np.unique(idx["columns"]["col_id"]))
    k = 20000
    want_p, want_c = oracle.Oracle(path).query_batch(seqs[: k * 150], off[: k + 1])
    assert np.array_equal(pml[: k * 150].astype(np.uint32), want_p) and np.array_equal(cid[: k * 150], want_c)


def test_two_gpus_in_process_replicas(cb, small_index, monkeypatch):
    """colbwt_index_load on 2 devices: chunks are dealt round-robin to the replicas, results land in input order."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "30000")
    seqs, off = small_index["seqs"], small_index["off"]
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"], devices=2)
    assert tbl.stats.n_devices == 2
    pml, cid = tbl.query(seqs, off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
    b = tbl.batch(seqs, off, device_slot=1)
    b.run(1)
    p2, c2 = b.download()
    assert np.array_equal(p2.astype(np.uint32), want_p) and np.array_equal(c2, want_c)


def test_narrow_layout_opt_in(cb, small_index, monkeypatch):
    """COLBWT_NARROW=1 builds the 8-byte hot/cold layout; results must not change."""
    monkeypatch.setenv("COLBWT_NARROW", "1")
    tbl = cb.ColPml.load(small_index["path"])
    extra, eoff = concat_reads(adversarial_reads(small_index["haps"]))
    orc = oracle.Oracle(small_index["path"])
    for seqs, off in ((small_index["seqs"], small_index["off"]), (extra, eoff)):
        want_p, want_c = orc.query_batch(seqs, off)
        pml, cid = tbl.query(seqs, off)
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


@pytest.mark.parametrize("case", ["pan4", "pan4all"])
def test_index_from_primaries_is_byte_identical_to_reference_build_col_bwt(cb, golden_dir, tmp_path, case):
    """GPU table construction from .bwt.heads/.bwt.len/.thr_pos/.col_runs/.col_ids (as the reference's tools left
    them) == the `.col_pml` the reference's build_col_bwt wrote, byte for byte; and it answers queries identically."""
    tbl = cb.ColPml.from_primaries(os.path.join(golden_dir, case + ".fa"))
    out = tmp_path / "rebuilt.col_pml"
    tbl.save(str(out))
    assert out.read_bytes() == open(os.path.join(golden_dir, case + ".col_pml"), "rb").read()
    ids, seqs, off = parse_fastx(os.path.join(golden_dir, f"{case}_reads.fa"))
    pml, cid = tbl.query(seqs, off, cb.PML_U32)
    txt_p = b"".join(cb.format_stats(ids[i], pml[int(off[i]):int(off[i + 1])]) for i in range(len(ids)))
    txt_c = b"".join(cb.format_stats(ids[i], cid[int(off[i]):int(off[i + 1])]) for i in range(len(ids)))
    assert txt_p == open(os.path.join(golden_dir, f"{case}_reads.fa.pml"), "rb").read()
    assert txt_c == open(os.path.join(golden_dir, f"{case}_reads.fa.cid"), "rb").read()


def test_index_from_primaries_matches_tooling_at_scale(cb, small_index, tmp_path):
    """Same on the 240 kbp synthetic index: rows == synthdata.table.build_columns (itself pinned to build_col_bwt)."""
    idx = small_index["idx"]
    p = str(tmp_path / "s.fa")
    F.write_primaries(p, idx["heads"], idx["lens"], idx["thr_run"])
    F.write_bit_vector(p + ".col_runs", small_index["cols"]["n"], idx["split_pos"])
    np.asarray(idx["split_ids"], np.uint8).tofile(p + ".col_ids")
    tbl = cb.ColPml.from_primaries(p)
    tbl.save(p + ".col_pml")
    assert open(p + ".col_pml", "rb").read() == open(small_index["path"], "rb").read()
    # load -> save round trip
    cb.ColPml.load(small_index["path"]).save(p + ".2")
    assert open(p + ".2", "rb").read() == open(small_index["path"], "rb").read()


def test_index_from_primaries_errors_and_cli(cb, golden_dir, tmp_path):
    import shutil
    with pytest.raises(cb.ColBwtError) as e:
        cb.ColPml.from_primaries(str(tmp_path / "nope.fa"))
    assert e.value.code == -1
    for ext in (".bwt.heads", ".bwt.len", ".thr_pos", ".col_runs", ".col_ids"):
        shutil.copy(os.path.join(golden_dir, "pan4.fa" + ext), tmp_path / ("x.fa" + ext))
    ids = (tmp_path / "x.fa.col_ids").read_bytes()
    (tmp_path / "x.fa.col_ids").write_bytes(ids[:-3])                 # fewer ids than set bits
    with pytest.raises(cb.ColBwtError) as e:
        cb.ColPml.from_primaries(str(tmp_path / "x.fa"))
    assert e.value.code == -2
    (tmp_path / "x.fa.col_ids").write_bytes(ids)
    cli = os.path.join(ROOT, "col_bwt_b200", "bin", "build_col_bwt_b200")
    r = subprocess.run([cli, str(tmp_path / "x.fa")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "x.fa.col_pml").read_bytes() == open(os.path.join(golden_dir, "pan4.col_pml"), "rb").read()


def test_synthetic_move_table_from_rows(cb):
    """configs[4] style: a directly synthesised move table (no text), LF-walk reads; many rows hold two thresholds."""
    rows, n, cols = PL.synth_move_table(200000, mean_len=12, device="cuda", seed=3)
    raw = rows.cpu().numpy()
    seqs, off = PL.walk_reads(cols, 3000, 150, sub=0.02)
    r = np.frombuffer(raw.tobytes(), dtype=F.ROW_DTYPE)
    cd = {"ch": r["ch"], "idx": F.u40_unpack(r["idx"]), "interval": r["interval"], "offset": r["offset"], "col_id": r["col_id"],
          "thr": F.u40_unpack(r["thr"]), "n": n, "bwt_r": len(r)}
    want_p, want_c = oracle.Oracle(columns=cd).query_batch(seqs, off)
    tbl = cb.ColPml.from_rows(raw, len(r), n)
    assert tbl.stats.slow_rows > 10000
    pml, cid = tbl.query(seqs, off, cb.PML_U8)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


@pytest.mark.parametrize("split_env", [{"COLBWT_SPLIT": "1", "COLBWT_SPLIT_CHUNK": "64", "COLBWT_SPLIT_WARM": "16", "COLBWT_SPLIT_MIN": "128"},
                                       {"COLBWT_SPLIT": "1", "COLBWT_SPLIT_CHUNK": "256", "COLBWT_SPLIT_WARM": "256", "COLBWT_SPLIT_MIN": "512"},
                                       {"COLBWT_SPLIT": "1", "COLBWT_SPLIT_CHUNK": "128", "COLBWT_SPLIT_WARM": "0", "COLBWT_SPLIT_MIN": "256"},
                                       {"COLBWT_SPLIT": "1"}])
def test_long_reads_split_into_chunk_tasks(cb, small_index, monkeypatch, split_env):
    """The long-read path (speculative chunk tasks + k_fixup) on the GPU: tiny chunks with a short warm-up force many
    re-traversals, the default parameters almost none; both must reproduce the serial result exactly."""
    for k, v in split_env.items():
        monkeypatch.setenv(k, v)
    idx = small_index["idx"]
    seqs, off = P.sample_reads(idx["text"], idx["seq_starts"], 50, 12000, sub=0.02, ins=0.015, dele=0.015, seed=6, len_jitter=0.5)
    rd = [bytes(seqs[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    rd[3] = rd[3][:700] + b"NNNN" + rd[3][704:]
    rd += adversarial_reads(small_index["haps"])
    s2, o2 = P.sample_reads(idx["text"], idx["seq_starts"], 500, 150, sub=0.01, seed=1)
    rd += [bytes(s2[int(o2[i]):int(o2[i + 1])]) for i in range(len(o2) - 1)]
    seqs, off = concat_reads(rd)
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"])
    pml, cid = tbl.query(seqs, off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
    b = tbl.batch(seqs, off, cb.PML_U32)
    assert b.launches == 3          # packed pass, byte pass, chain fix-up
    b.run(2)
    p2, c2 = b.download()
    assert np.array_equal(p2, want_p) and np.array_equal(c2, want_c)
    # bookkeeping of the long-read path (colbwt_batch_counters): the cut reads, their chunk tasks, the whole reads scheduled
    # among them, and how many chunks the last traversal had to redo because their speculative start had not converged
    k = b.counters
    assert k["split_reads"] >= 40 and k["chunk_tasks"] >= 2 * k["split_reads"] and k["scheduled"] >= k["chunk_tasks"]
    if split_env.get("COLBWT_SPLIT_WARM") == "0":
        assert k["retraversed"] > 0.5 * (k["chunk_tasks"] - k["split_reads"])    # no warm-up: nearly every speculative chunk is redone
    else:
        assert k["retraversed"] <= k["chunk_tasks"] - k["split_reads"]


def test_device_side_packing_with_pinned_input(cb, small_index, monkeypatch):
    """When the sequence buffer is pinned, raw bytes are copied and packed on the device (k_pack_reads); irregular reads
    are found there and routed to the byte pass.  Same results as host packing."""
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "50000")
    rng = np.random.default_rng(4)
    seqs = small_index["seqs"].copy()
    off = small_index["off"]
    for i in rng.choice(len(off) - 1, (len(off) - 1) // 10, replace=False):
        a, b = int(off[i]), int(off[i + 1])
        seqs[a + int(rng.integers(0, b - a))] = ord("N") if rng.random() < 0.5 else ord("a")
    extra, eoff = concat_reads(adversarial_reads(small_index["haps"]))
    seqs = np.concatenate((seqs, extra))
    off = np.concatenate((off, eoff[1:] + off[-1]))
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"])
    hs = cb.PinnedArray(seqs.size, np.uint8)
    hs.array[:] = seqs
    for env in ("1", "0"):
        monkeypatch.setenv("COLBWT_DEVICE_PACK", env)
        pml, cid = tbl.query(hs.array, off, cb.PML_U16)
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c), env


def test_degenerate_batches_and_offset_base(cb, small_index):
    """Empty batch, all-empty reads, a batch whose offsets do not start at 0, repeated calls on one handle."""
    tbl = cb.ColPml.load(small_index["path"])
    pml, cid = tbl.query(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert pml.size == 0 and cid.size == 0
    pml, cid = tbl.query(np.zeros(0, np.uint8), np.zeros(6, np.uint64))          # five empty reads
    assert pml.size == 0 and cid.size == 0
    seqs, off = small_index["seqs"], small_index["off"]
    orc = oracle.Oracle(small_index["path"])
    # offsets starting at read 100 of the same buffer: outputs are relative to off[0]
    sub_off = off[100:301].copy()
    want_p, want_c = orc.query_batch(seqs[int(sub_off[0]):int(sub_off[-1])], sub_off - sub_off[0])
    pml, cid = tbl.query(seqs, sub_off)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
    b = tbl.batch(seqs, sub_off)
    b.run(1)
    p2, c2 = b.download()
    assert np.array_equal(p2.astype(np.uint32), want_p) and np.array_equal(c2, want_c)
    # a batch that is one base long
    pml, cid = tbl.query(np.frombuffer(b"G", np.uint8), np.array([0, 1], np.uint64), cb.PML_U8)
    w = orc.query(b"G")
    assert pml.tolist() == w[0].tolist() and cid.tolist() == w[1].tolist()


def test_two_threads_share_one_index(cb, small_index):
    import threading
    tbl = cb.ColPml.load(small_index["path"])
    seqs, off = small_index["seqs"], small_index["off"]
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    results = [None, None]

    def work(i):
        for _ in range(3):
            results[i] = tbl.query(seqs, off)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for pml, cid in results:
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


@pytest.mark.parametrize("case,mode,rate", [("pan4", "tunnels", 4), ("pan4all", "all", 2)])
def test_col_split_on_gpu_matches_reference_col_split(cb, golden_dir, tmp_path, case, mode, rate):
    """colbwt_col_split (GPU FL walks + host overlap sweep) writes the same .col_runs / .col_ids, byte for byte, as the
    reference's build_FL + col_split wrote for the same .bwt.heads/.bwt.len/.col_mums (tests/golden/make_golden.py)."""
    import shutil
    for ext in (".bwt.heads", ".bwt.len", ".col_mums"):
        shutil.copy(os.path.join(golden_dir, case + ".fa" + ext), tmp_path / ("x.fa" + ext))
    bits, marked = cb.col_split(str(tmp_path / "x.fa"), mode, rate)
    for ext in (".col_runs", ".col_ids"):
        assert (tmp_path / ("x.fa" + ext)).read_bytes() == open(os.path.join(golden_dir, case + ".fa" + ext), "rb").read(), ext
    ids = np.fromfile(os.path.join(golden_dir, case + ".fa.col_ids"), np.uint8)
    assert bits == ids.size and marked == int((ids > 0).sum())
    # the CLI form
    os.remove(tmp_path / "x.fa.col_runs")
    r = subprocess.run([os.path.join(ROOT, "col_bwt_b200", "bin", "col_split_b200"), str(tmp_path / "x.fa"), "-m", mode, "-s", str(rate)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "x.fa.col_runs").read_bytes() == open(os.path.join(golden_dir, case + ".fa.col_runs"), "rb").read()


def test_col_split_on_gpu_at_scale_and_whole_build_chain(cb, small_index, tmp_path):
    """primaries + multi-MUMs -> colbwt_col_split -> colbwt_index_from_primaries == the tooling's table (itself pinned to
    the reference tools), i.e. the in-tree part of `col-bwt build` end to end on the GPU."""
    idx = small_index["idx"]
    p = str(tmp_path / "s.fa")
    PL.write_reference_inputs(p, idx)                      # .bwt.heads .bwt.len .thr_pos .col_mums
    bits, marked = cb.col_split(p, "tunnels", 5)
    n, pos = F.read_bit_vector(p + ".col_runs")
    ids = np.fromfile(p + ".col_ids", np.uint8)
    assert np.array_equal(pos, idx["split_pos"]) and np.array_equal(ids, idx["split_ids"])
    tbl = cb.ColPml.from_primaries(p)
    tbl.save(p + ".col_pml")
    assert open(p + ".col_pml", "rb").read() == open(small_index["path"], "rb").read()
    with pytest.raises(cb.ColBwtError):
        cb.col_split(str(tmp_path / "missing.fa"))


def test_rlbwt_to_bwt_on_gpu(cb, golden_dir, small_index, tmp_path):
    """colbwt_rlbwt_to_bwt / rlbwt_to_bwt_b200 write the file the reference's rlbwt_to_bwt writes (golden pan4.fa.bwt), and
    agree with the pinned restatement (synthdata.formats.expand_rlbwt) on edge cases and on a multi-chunk-sized input."""
    import shutil
    p = str(tmp_path / "x.fa")
    for ext in (".bwt.heads", ".bwt.len"):
        shutil.copy(os.path.join(golden_dir, "pan4.fa" + ext), p + ext)
    assert cb.rlbwt_to_bwt(p) == os.path.getsize(os.path.join(golden_dir, "pan4.fa.bwt"))
    assert open(p + ".bwt", "rb").read() == open(os.path.join(golden_dir, "pan4.fa.bwt"), "rb").read()
    os.remove(p + ".bwt")
    r = subprocess.run([os.path.join(ROOT, "col_bwt_b200", "bin", "rlbwt_to_bwt_b200"), p], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(p + ".bwt", "rb").read() == open(os.path.join(golden_dir, "pan4.fa.bwt"), "rb").read()
    # terminators kept as they are, zero-length runs, a run longer than 65535, surplus length records, lengths not multiple of 16
    q = str(tmp_path / "e.fa")
    np.frombuffer(b"A\x00C\x01GTA", np.uint8).tofile(q + ".bwt.heads")
    F.write_u40(q + ".bwt.len", np.array([3, 1, 0, 2, 70000, 1, 5, 9, 9]))
    assert cb.rlbwt_to_bwt(q) == 70012
    assert open(q + ".bwt", "rb").read() == F.expand_rlbwt(q).tobytes()
    # the synthetic index's runs, and 300 M characters in 40 runs (two 256 MB chunks)
    idx = small_index["idx"]
    s = str(tmp_path / "s.fa")
    F.write_primaries(s, idx["heads"], idx["lens"], idx["thr_run"])
    assert cb.rlbwt_to_bwt(s) == small_index["cols"]["n"]
    assert open(s + ".bwt", "rb").read() == F.expand_rlbwt(s).tobytes()
    b = str(tmp_path / "b.fa")
    rng = np.random.default_rng(5)
    np.frombuffer(b"ACGT" * 10, np.uint8).tofile(b + ".bwt.heads")
    F.write_u40(b + ".bwt.len", rng.integers(1, 15_000_000, 40))
    n = cb.rlbwt_to_bwt(b)
    got = np.fromfile(b + ".bwt", np.uint8)
    assert got.size == n and np.array_equal(got, F.expand_rlbwt(b))
    # empty and missing inputs
    e = str(tmp_path / "z.fa")
    open(e + ".bwt.heads", "wb").close()
    open(e + ".bwt.len", "wb").close()
    assert cb.rlbwt_to_bwt(e) == 0 and os.path.getsize(e + ".bwt") == 0
    with pytest.raises(cb.ColBwtError) as err:
        cb.rlbwt_to_bwt(str(tmp_path / "missing.fa"))
    assert err.value.code == -1


def test_randomised_batches(cb, small_index, monkeypatch):
    """A short round of tools/fuzz_parity.py: random ragged batches, every legal PML width, random chunk geometry for the
    long-read path, streaming call and device-resident batch alike -- equal to the oracle on every base."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_parity as fz
    for k in ("COLBWT_SPLIT", "COLBWT_SPLIT_CHUNK", "COLBWT_SPLIT_WARM", "COLBWT_SPLIT_MIN"):
        monkeypatch.delenv(k, raising=False)      # restored after the test; one_round sets them freely
    tbl = cb.ColPml.load(small_index["path"])
    orc = oracle.Oracle(small_index["path"])
    rng = np.random.default_rng(7)
    idx = small_index["idx"]
    for _ in range(30):
        ok, info = fz.one_round(cb, tbl, orc, rng, idx["text"], idx["seq_starts"], small_index["haps"])
        assert ok, info


# ---- compact result form (compact.cu + expand.cpp) -------------------------------------------------------------------
def _compact_cases(small_index):
    idx = small_index["idx"]
    extra, eoff = concat_reads(adversarial_reads(small_index["haps"]))
    s, o = P.sample_reads(idx["text"], idx["seq_starts"], 120, 3000, sub=0.02, ins=0.015, dele=0.015, seed=16, len_jitter=0.5)
    return [(small_index["seqs"], small_index["off"]), (extra, eoff), (s, o)]


@pytest.mark.parametrize("chunk", [None, "20000"])
def test_compact_result_expands_to_the_dense_arrays(cb, small_index, monkeypatch, chunk):
    """colbwt_query_compact + colbwt_compact_expand == the oracle's dense arrays: one segment and many, pageable and pinned
    result buffers, irregular and long reads."""
    if chunk:
        monkeypatch.setenv("COLBWT_CHUNK_BASES", chunk)
    orc = oracle.Oracle(small_index["path"])
    tbl = cb.ColPml.load(small_index["path"])
    for seqs, off in _compact_cases(small_index):
        want_p, want_c = orc.query_batch(seqs, off)
        width = cb.PML_U16 if int(np.diff(off).max()) < 65536 else cb.PML_U32
        res = tbl.query_compact(seqs, off)
        segs = cb.compact_segments(res)
        assert sum(s["n_reads"] for s in segs) == off.size - 1 and sum(s["n_bases"] for s in segs) == seqs.size
        if chunk:
            assert len(segs) > 1 or seqs.size <= 20000
        assert sum(s["n_values"] for s in segs) == int((want_c != 0).sum())
        pml, cid = cb.compact_expand(res, off, width)
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
        # pinned result buffer: the GPU writes into it directly
        bound = cb._L.colbwt_compact_bound(off.ctypes.data, off.size - 1)
        pin = cb.PinnedArray(bound, np.uint8)
        res2 = tbl.query_compact(seqs, off, out=pin.array)
        pml, cid = cb.compact_expand(res2, off, width)
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
        assert res2.size <= bound and res2.size < seqs.size + 4096      # far below the dense 3 bytes per base


def test_compact_result_buffer_too_small_is_an_error(cb, small_index):
    tbl = cb.ColPml.load(small_index["path"])
    seqs, off = small_index["seqs"], small_index["off"]
    with pytest.raises(cb.ColBwtError) as e:
        tbl.query_compact(seqs, off, out=np.empty(seqs.size // 64, np.uint8))
    assert e.value.code == -6
    # the index is usable afterwards
    res = tbl.query_compact(seqs, off)
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    pml, cid = cb.compact_expand(res, off, cb.PML_U16)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


@pytest.mark.parametrize("device_pack", ["0", "1"])
def test_dense_query_over_compact_transport(cb, small_index, monkeypatch, device_pack):
    """COLBWT_COMPACT_D2H=1: dense results for the caller, compact form over the link, expanded by the host threads."""
    monkeypatch.setenv("COLBWT_COMPACT_D2H", "1")
    monkeypatch.setenv("COLBWT_DEVICE_PACK", device_pack)
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "50000")
    orc = oracle.Oracle(small_index["path"])
    tbl = cb.ColPml.load(small_index["path"])
    for seqs, off in _compact_cases(small_index):
        want_p, want_c = orc.query_batch(seqs, off)
        widths = (cb.PML_U8, cb.PML_U16, cb.PML_U32) if int(np.diff(off).max()) < 256 else (cb.PML_U16, cb.PML_U32)
        for width in widths:
            pml, cid = tbl.query(seqs, off, width)                       # pageable buffers
            assert tbl.last_transport == "compact"
            assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
        h_seqs = cb.PinnedArray(seqs.size, np.uint8)                     # pinned input: device packing possible
        h_seqs.array[:] = seqs
        pp, pc = cb.PinnedArray(seqs.size, np.uint16), cb.PinnedArray(seqs.size, np.uint8)
        tbl.query(h_seqs.array, off, cb.PML_U16, out=(pp.array, pc.array))
        assert tbl.last_packing == ("device" if device_pack == "1" and int(np.diff(off).max()) < 8192 else "host")
        assert np.array_equal(pp.array.astype(np.uint32), want_p) and np.array_equal(pc.array, want_c)
    monkeypatch.setenv("COLBWT_COMPACT_D2H", "0")
    pml, cid = tbl.query(seqs, off, cb.PML_U16)
    assert tbl.last_transport == "dense"
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


def test_in_trip_kernel_variant(cb, small_index, golden_dir, monkeypatch):
    """COLBWT_INTRIP=1 forces the kernel variant the launcher picks for tables of 2 GiB and more (same-line neighbours
    resolved inside the trip): identical results on short, irregular and split long reads, all PML widths."""
    monkeypatch.setenv("COLBWT_INTRIP", "1")
    orc = oracle.Oracle(small_index["path"])
    tbl = cb.ColPml.load(small_index["path"])
    for seqs, off in _compact_cases(small_index):
        want_p, want_c = orc.query_batch(seqs, off)
        widths = (cb.PML_U8, cb.PML_U16, cb.PML_U32) if int(np.diff(off).max()) < 256 else (cb.PML_U16, cb.PML_U32)
        for width in widths:
            b = tbl.batch(seqs, off, width)
            b.run(1)
            pml, cid = b.download()
            b.close()
            assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
    monkeypatch.setenv("COLBWT_SPLIT", "1")
    monkeypatch.setenv("COLBWT_SPLIT_CHUNK", "300")
    monkeypatch.setenv("COLBWT_SPLIT_WARM", "100")
    seqs, off = _compact_cases(small_index)[2]
    want_p, want_c = orc.query_batch(seqs, off)
    pml, cid = tbl.query(seqs, off, cb.PML_U16)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


@pytest.mark.parametrize("device_pack", ["0", "1"])
def test_dense_query_with_compact_chain_ids(cb, small_index, monkeypatch, device_pack):
    """COLBWT_COMPACT_D2H=2: PML is copied into the caller's pinned array as it is, the chain ids cross the link as bit
    words + non-zero values and are rebuilt group by group on the host.  Pageable result arrays cannot take this route
    (the copy engine writes the PML in place) and fall back to another transport; results are the same either way."""
    monkeypatch.setenv("COLBWT_COMPACT_D2H", "2")
    monkeypatch.setenv("COLBWT_DEVICE_PACK", device_pack)
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "70000")
    orc = oracle.Oracle(small_index["path"])
    tbl = cb.ColPml.load(small_index["path"])
    for seqs, off in _compact_cases(small_index):
        want_p, want_c = orc.query_batch(seqs, off)
        h_seqs = cb.PinnedArray(seqs.size, np.uint8)
        h_seqs.array[:] = seqs
        widths = (cb.PML_U8, cb.PML_U16, cb.PML_U32) if int(np.diff(off).max()) < 256 else (cb.PML_U16, cb.PML_U32)
        for width in widths:
            dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]
            pp, pc = cb.PinnedArray(seqs.size, dt), cb.PinnedArray(seqs.size, np.uint8)
            pc.array[:] = 0xEE                                            # every byte must be rewritten, zeros included
            tbl.query(h_seqs.array, off, width, out=(pp.array, pc.array))
            assert tbl.last_transport == "pml-dense+cid-compact"
            assert np.array_equal(pp.array.astype(np.uint32), want_p) and np.array_equal(pc.array, want_c)
            h2d, d2h = tbl.last_bytes
            assert h2d > 0 and seqs.size * width < d2h < seqs.size * (width + 1) + 4096 * (1 + seqs.size // 70000)
        pml, cid = tbl.query(seqs, off, cb.PML_U16)                      # pageable results: another transport serves them
        assert tbl.last_transport != "pml-dense+cid-compact"
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


@pytest.mark.parametrize("device_pack", ["0", "1"])
def test_dense_query_with_alternating_transports(cb, small_index, monkeypatch, device_pack):
    """COLBWT_COMPACT_D2H=3: chunks alternate between plain copies and the compact form expanded by the host threads, so
    that the copy engine and the host cores fill the caller's (pinned) arrays together; many small chunks here."""
    monkeypatch.setenv("COLBWT_COMPACT_D2H", "3")
    monkeypatch.setenv("COLBWT_DEVICE_PACK", device_pack)
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "40000")
    orc = oracle.Oracle(small_index["path"])
    tbl = cb.ColPml.load(small_index["path"])
    for seqs, off in _compact_cases(small_index):
        want_p, want_c = orc.query_batch(seqs, off)
        h_seqs = cb.PinnedArray(seqs.size, np.uint8)
        h_seqs.array[:] = seqs
        widths = (cb.PML_U8, cb.PML_U16, cb.PML_U32) if int(np.diff(off).max()) < 256 else (cb.PML_U16, cb.PML_U32)
        for width in widths:
            dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width]
            pp, pc = cb.PinnedArray(seqs.size, dt), cb.PinnedArray(seqs.size, np.uint8)
            pp.array[:] = 0xAB
            pc.array[:] = 0xEE
            for _ in range(2):                                            # second call: every slot has served both kinds of chunk
                tbl.query(h_seqs.array, off, width, out=(pp.array, pc.array))
                assert tbl.last_transport == "mixed"
                assert np.array_equal(pp.array.astype(np.uint32), want_p) and np.array_equal(pc.array, want_c)
            h2d, d2h = tbl.last_bytes
            if seqs.size > 200000:                                        # several chunks: some went dense, some compact
                assert seqs.size * 0.2 * (width + 1) < d2h < seqs.size * 0.9 * (width + 1)
        pml, cid = tbl.query(seqs, off, cb.PML_U16)                      # pageable results: another transport serves them
        assert tbl.last_transport != "mixed"
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


@pytest.mark.parametrize("compact", ["0", "1", "2", "3"])
def test_two_replicas_on_one_gpu_feeder_threads(cb, small_index, monkeypatch, compact):
    """Two replicas of the table on the SAME device: exercises the per-device feeder threads of colbwt_query (one per
    replica, chunks taken from a shared counter) on a single-GPU box; results must land in input order."""
    monkeypatch.setenv("COLBWT_CHUNK_BASES", "15000")
    monkeypatch.setenv("COLBWT_COMPACT_D2H", compact)
    seqs, off = small_index["seqs"], small_index["off"]
    want_p, want_c = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    tbl = cb.ColPml.load(small_index["path"], devices=[0, 0, 0])
    assert tbl.stats.n_devices == 3
    for _ in range(3):
        pml, cid = tbl.query(seqs, off)
        assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)
    h_seqs = cb.PinnedArray(seqs.size, np.uint8)
    h_seqs.array[:] = seqs
    pp, pc = cb.PinnedArray(seqs.size, np.uint16), cb.PinnedArray(seqs.size, np.uint8)
    tbl.query(h_seqs.array, off, cb.PML_U16, out=(pp.array, pc.array))   # pinned results: the chain-id transport is possible
    assert np.array_equal(pp.array.astype(np.uint32), want_p) and np.array_equal(pc.array, want_c)
    res = tbl.query_compact(seqs, off)
    pml, cid = cb.compact_expand(res, off, cb.PML_U16)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


def test_fitted_synthetic_table_from_device_rows(cb):
    """configs[4] as bench.py builds it: thresholds snapped to run starts like a real index (few exact-search rows), the
    18-byte rows handed over in GPU memory (colbwt_index_from_rows with a device pointer), mixed short and long walks."""
    import torch
    rows, n, cols = PL.synth_move_table(300000, mean_len=16, device="cuda", seed=4, snap=0.854)
    torch.cuda.synchronize()
    s1, o1 = PL.walk_reads(cols, 2000, 150, sub=0.01, seed=7)
    s2, o2 = PL.walk_reads(cols, 6, 9000, sub=0.05, seed=8)
    seqs = np.concatenate([s1, s2])
    off = np.concatenate([o1, o2[1:] + o1[-1]]).astype(np.uint64)
    raw = rows.cpu().numpy()
    r = np.frombuffer(raw.tobytes(), dtype=F.ROW_DTYPE)
    cd = {"ch": r["ch"], "idx": F.u40_unpack(r["idx"]), "interval": r["interval"], "offset": r["offset"], "col_id": r["col_id"],
          "thr": F.u40_unpack(r["thr"]), "n": n, "bwt_r": len(r)}
    want_p, want_c = oracle.Oracle(columns=cd).query_batch(seqs, off)
    tbl = cb.ColPml.from_device_rows(rows.data_ptr(), len(r), len(r), n)
    t_host = cb.ColPml.from_rows(raw, len(r), n)
    assert tbl.stats.slow_rows == t_host.stats.slow_rows
    assert tbl.stats.slow_rows < 0.03 * len(r)          # uniform thresholds: ~19 % of the rows
    pml, cid = tbl.query(seqs, off, cb.PML_U16)
    assert np.array_equal(pml.astype(np.uint32), want_p) and np.array_equal(cid, want_c)


def test_col_split_overlapping_marks_against_the_reference_binary(cb, tmp_path):
    """`-m all` / `-m tunnels` on a small divergent pangenome with many overlapping multi-MUM ranges: the device-side resolution
    (sorts + prefix sums, col_split.cu) must write, byte for byte, what the reference's priority-queue sweep writes -- the
    reference's own build_FL + col_split (oracle/_ref, compiled from the reference sources) run on the same inputs."""
    import shutil
    if not oracle.have_ref() or not os.path.exists(oracle.ref_bin("col_split")):
        pytest.skip("oracle/_ref not built")
    haps = P.make_haplotypes(3000, 5, snp=4e-3, indel=5e-4, seed=21, tree=True)
    idx = PL.build_index(haps, split_rate=2, min_mum=8)
    p, q = str(tmp_path / "o.fa"), str(tmp_path / "ref.fa")
    PL.write_reference_inputs(p, idx)
    for ext in (".bwt.heads", ".bwt.len", ".thr_pos", ".col_mums"):
        shutil.copy(p + ext, q + ext)
    subprocess.run([oracle.ref_bin("build_FL"), q], check=True, capture_output=True)
    for mode, rate in (("all", 1), ("all", 3), ("tunnels", 1), ("tunnels", 7)):
        subprocess.run([oracle.ref_bin("col_split"), q, "-m", mode, "-s", str(rate)], check=True, capture_output=True)
        bits, marked = cb.col_split(p, mode, rate)
        for ext in (".col_runs", ".col_ids"):
            assert open(p + ext, "rb").read() == open(q + ext, "rb").read(), (mode, rate, ext)
        ids = np.fromfile(p + ".col_ids", np.uint8)
        assert bits == ids.size and marked == int((ids > 0).sum()) and marked > 0
    with open(p + ".col_mums", "rb") as f:
        raw = bytearray(f.read())
    assert len(raw) >= 25
    raw[5:15], raw[15:25] = raw[15:25], raw[5:15]                                   # swap the first two (len, pos) pairs: positions descend
    open(p + ".col_mums", "wb").write(bytes(raw))
    with pytest.raises(cb.ColBwtError) as e:
        cb.col_split(p, "all", 1)
    assert e.value.code == -2
