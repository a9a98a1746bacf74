"""CPU: the oracle (oracle/colbwt_oracle.c) pinned against the reference's own outputs."""
import os

import numpy as np
import pytest

import oracle
from util import adversarial_reads, check_pml_properties, concat_reads, parse_fastx


@pytest.mark.parametrize("case", ["toy", "pan4", "pan4all"])
def test_oracle_matches_reference_golden_text(golden_dir, case):
    """Golden .pml/.cid were written by the reference's pml_query as shipped (tests/golden/make_golden.py)."""
    orc = oracle.Oracle(os.path.join(golden_dir, f"{case}.col_pml"))
    ids, seqs, off = parse_fastx(os.path.join(golden_dir, f"{case}_reads.fa"))
    pml, cid = orc.query_batch(seqs, off)
    assert orc.format_text(ids, off, pml) == open(os.path.join(golden_dir, f"{case}_reads.fa.pml"), "rb").read()
    assert orc.format_text(ids, off, cid) == open(os.path.join(golden_dir, f"{case}_reads.fa.cid"), "rb").read()


def test_known_answer_vector_survey_4_3(golden_dir):
    orc = oracle.Oracle(os.path.join(golden_dir, "toy.col_pml"))
    pml, cid = orc.query(b"GATTACAGAT")
    assert pml.tolist() == [6, 5, 4, 3, 2, 1, 0, 2, 1, 0] and cid.tolist() == [0, 10, 0, 0, 6, 4, 0, 0, 0, 0]
    pml, cid = orc.query(b"TTAGGATNACA")
    assert pml.tolist() == [1, 0, 1, 0, 3, 2, 1, 0, 1, 0, 1] and cid.tolist() == [0, 0, 0, 7, 5, 9, 0, 0, 0, 0, 0]
    pml, cid = orc.query(b"CCCC")
    assert pml.tolist() == [0, 0, 1, 0] and cid.tolist() == [0, 0, 0, 0]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_in_process(small_index):
    """Same table, same reads: plain-C restatement vs the reference's header-only col_pml::query_pml."""
    orc = oracle.Oracle(small_index["path"])
    ref = oracle.Reference(small_index["path"])
    extra, eoff = concat_reads(adversarial_reads(small_index["haps"]))
    for seqs, off in ((small_index["seqs"], small_index["off"]), (extra, eoff)):
        p0, c0 = orc.query_batch(seqs, off)
        p1, c1 = ref.query_batch(seqs, off, threads=2)
        assert np.array_equal(p0, p1) and np.array_equal(c0, c1)
        check_pml_properties(p0, off)


def test_oracle_load_rejects_missing_and_short(tmp_path, golden_dir):
    with pytest.raises(OSError):
        oracle.Oracle(str(tmp_path / "nope.col_pml"))
    data = open(os.path.join(golden_dir, "toy.col_pml"), "rb").read()
    p = tmp_path / "short.col_pml"
    p.write_bytes(data[:100])
    with pytest.raises(OSError):
        oracle.Oracle(str(p))


@pytest.mark.parametrize("case", ["toy_rl", "pan4"])
def test_rlbwt_expansion_matches_reference_rlbwt_to_bwt(golden_dir, tmp_path, case):
    """synthdata.formats.expand_rlbwt (the checker of colbwt_rlbwt_to_bwt) == the reference's rlbwt_to_bwt, byte for byte:
    against the committed pan4.fa.bwt fixture, and against the compiled reference tool where it is present."""
    import shutil
    import subprocess
    from synthdata import formats as F
    p = str(tmp_path / "x.fa")
    if case == "pan4":
        for ext in (".bwt.heads", ".bwt.len"):
            shutil.copy(os.path.join(golden_dir, "pan4.fa" + ext), p + ext)
        want = open(os.path.join(golden_dir, "pan4.fa.bwt"), "rb").read()
        assert F.expand_rlbwt(p).tobytes() == want
    else:   # terminators (0 and 1) are written as they are, zero-length runs write nothing, extra length records are ignored
        heads = np.frombuffer(b"A\x00C\x01GTA", np.uint8)
        np.asarray(heads).tofile(p + ".bwt.heads")
        F.write_u40(p + ".bwt.len", np.array([3, 1, 0, 2, 70000, 1, 5, 9, 9]))
        assert F.expand_rlbwt(p).tobytes() == b"AAA\x00\x01\x01" + b"G" * 70000 + b"T" + b"A" * 5
    if oracle.have_ref():
        subprocess.run([oracle.ref_bin("rlbwt_to_bwt"), p], check=True, stdout=subprocess.DEVNULL)
        assert open(p + ".bwt", "rb").read() == F.expand_rlbwt(p).tobytes()
