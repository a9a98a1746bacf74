"""CPU: the synthetic-index tooling (synthdata/) against the reference's own build tools (oracle/_ref)."""
import subprocess

import numpy as np
import pytest

import oracle
from synthdata import bwtbuild as B, formats as F, pangenome as P, pipeline as PL, table as T


def test_suffix_array_bwt_lcp_thresholds_on_kat_text():
    text = np.frombuffer(b"GATTACAGATTACCGATTACA\x01TTACAGATCACAGGATTAGA\x00", np.uint8).copy()
    si = B.SuffixIndex(text)
    tb = bytes(text)
    sa = sorted(range(len(tb)), key=lambda i: tb[i:])
    assert si.sa.tolist() == sa
    bwt = si.bwt()
    assert bytes(bwt.numpy()).replace(b"\x00", b"$").replace(b"\x01", b"$") == b"AAGCTTTCTTCCCGGGGGATAAAACAAC$AGATTTTTAA$AAA"
    heads, starts, lens = B.bwt_runs(bwt)
    thr = B.thresholds(bwt, si.lcp(), heads, starts)
    assert thr.tolist() == [0, 0, 0, 0, 4, 8, 9, 3, 2, 18, 19, 18, 25, 25, 0, 27, 18, 31, 25, 32, 32, 39]


def test_lcp_exact_on_repetitive_text():
    haps = P.make_haplotypes(1500, 3, snp=0.01, indel=0.002, seed=5)
    text, _, _ = P.build_text(haps)
    si = B.SuffixIndex(text)
    tb = bytes(text)
    sa = si.sa.tolist()

    def nl(a, b):
        k = 0
        while a + k < len(tb) and b + k < len(tb) and tb[a + k] == tb[b + k]:
            k += 1
        return k
    assert si.lcp().tolist() == [0] + [nl(sa[i - 1], sa[i]) for i in range(1, len(sa))]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("split_rate", [1, 7])
def test_marks_and_table_match_reference_tools(tmp_path, split_rate):
    """mark_tunnels/resolve_marks == col_split -m tunnels; build_columns == build_col_bwt, byte for byte."""
    haps = P.make_haplotypes(12000, 4, snp=3e-3, indel=3e-4, seed=21)
    idx = PL.build_index(haps, split_rate=split_rate)
    p = str(tmp_path / "x.fa")
    PL.write_reference_inputs(p, idx)
    run = lambda *a: subprocess.run(list(a), check=True, stdout=subprocess.DEVNULL)
    run(oracle.ref_bin("build_FL"), p)
    run(oracle.ref_bin("col_split"), p, "-m", "tunnels", "-s", str(split_rate))
    n, pos = F.read_bit_vector(p + ".col_runs")
    ids = np.fromfile(p + ".col_ids", dtype=np.uint8)
    assert np.array_equal(pos, idx["split_pos"]) and np.array_equal(ids, idx["split_ids"])
    # sequential restatement of find_col_runs agrees with the vectorised one
    m_start, m_id = T.mark_tunnels(*_sa_isa(idx), idx["mum_len"], idx["mum_pos"], idx["num_docs"], split_rate)
    run_starts = np.cumsum(idx["lens"]) - idx["lens"]
    pos2, ids2 = T.resolve_marks(n, run_starts, m_start, m_id, np.full(m_start.size, idx["num_docs"]))
    assert np.array_equal(pos2, pos) and np.array_equal(ids2, ids)
    F.write_shim_sd_vector(p + ".col_runs", n, pos)
    run(oracle.ref_bin("build_col_bwt"), p)
    meta, rows = F.read_col_pml(p + ".col_pml")
    c = idx["columns"]
    mine = F.rows_from_columns(c["ch"], c["idx"], c["interval"], c["offset"], c["col_id"], c["thr"])
    assert meta == {"bwt_r": c["bwt_r"], "n": c["n"], "r": len(mine), "size": len(mine)}
    assert np.array_equal(rows.view(np.uint8), mine.view(np.uint8))


def _sa_isa(idx):
    import torch
    si = B.SuffixIndex(idx["text"], keep_levels=False)
    run_starts = torch.as_tensor(np.cumsum(idx["lens"]) - idx["lens"])
    return si.sa, si.isa, run_starts


def test_lcp_from_irreducible_entries_equals_lcp_from_levels():
    """bwtbuild.lcp_irreducible (used for texts too large to keep the doubling levels) against SuffixIndex.lcp."""
    import torch
    from synthdata import bwtbuild
    torch.set_num_threads(1)
    rng = np.random.default_rng(3)
    haps = P.make_haplotypes(4000, 5, snp=2e-3, indel=2e-4, seed=9, tree=True)
    text, _, _ = P.build_text(haps, with_revcomp=True)
    cases = [text] + [np.concatenate([rng.choice(np.frombuffer(b"AC\x01", np.uint8), size=int(rng.integers(3, 150))), [0]]).astype(np.uint8) for _ in range(15)]
    for t in cases:
        si = bwtbuild.SuffixIndex(t)
        assert bool((si.lcp() == bwtbuild.lcp_irreducible(si, si.bwt())).all())
