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
    si = bwtbuild.SuffixIndex(text, device=device)
    bwt = si.bwt()
    lcp = si.lcp()
    si.levels = []  # free
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
