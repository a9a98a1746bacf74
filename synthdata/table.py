"""Sub-run marking and move-table columns for synthetic indexes.

Own restatement (numpy/torch, vectorised) of two build-side steps, so that indexes of the named shapes
can be produced at scale on the GPU box where /root/reference and its CPU tools do not exist:

  mark_tunnels    <->  col_split<>::split in Mode::Tunneled   /root/reference/include/col_split.hpp:54-136
  resolve_marks   <->  col_split<>::find_col_runs             col_split.hpp:258-338
  build_columns   <->  col_bwt ctor + compute_table + read_thresholds
                       /root/reference/include/col_bwt.hpp:124-230,440-457, LF_table.hpp:365-387

tests/test_synth.py checks all three against the reference's own binaries (oracle/_ref/col_split,
build_col_bwt) on small inputs.  Tooling only: index construction is outside the hot-path scope.
"""
from __future__ import annotations

import heapq

import numpy as np
import torch

ID_MAX = 256  # bit_max(ID_BITS), common.hpp:47,302-304


def bin_id(i: np.ndarray) -> np.ndarray:
    """col_split.hpp:222-224 / col_bwt.hpp:48-50: ids >= 256 fold to (id % 255) + 1."""
    i = np.asarray(i, dtype=np.int64)
    return np.where(i >= ID_MAX, (i % (ID_MAX - 1)) + 1, i)


def mark_tunnels(sa: torch.Tensor, isa: torch.Tensor, run_starts: torch.Tensor, mum_len: np.ndarray,
                 mum_pos: np.ndarray, num_docs: int, split_rate: int):
    """Marked BWT ranges of `col_split -m tunnels -s split_rate`.

    For MUM k (rows p..p+N-1, length L) let R_t be the rows of its suffixes advanced by t symbols.
    The reference walks FL from R_0; step t succeeds while R_t lies inside one F-run, i.e. while
    R_{t+1}'s first and last row fall in the same BWT run; column j is marked at R_{j+1} when
    j % split_rate == 0 and steps 0..j all succeeded (FL_loop, col_split.hpp:65-103).

    Returns (start int64[], id int64[] (1-based MUM number, unbinned)), height is always num_docs.
    """
    dev = sa.device
    N = int(num_docs)
    L = torch.as_tensor(np.asarray(mum_len, dtype=np.int64), device=dev)
    p = torch.as_tensor(np.asarray(mum_pos, dtype=np.int64), device=dev)
    if L.numel() == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    first_txt, last_txt = sa[p], sa[p + N - 1]
    mum_of = torch.repeat_interleave(torch.arange(L.numel(), device=dev), L)
    col0 = torch.cumsum(L, 0) - L
    t = torch.arange(int(L.sum()), device=dev) - col0[mum_of]          # column j within its MUM
    f = isa[first_txt[mum_of] + t + 1]                                   # R_{j+1}.first
    l = isa[last_txt[mum_of] + t + 1]                                    # R_{j+1}.last
    run_f = torch.searchsorted(run_starts, f, right=True)
    run_l = torch.searchsorted(run_starts, l, right=True)
    ok = (run_f == run_l) & (l - f == N - 1)
    # first failing column per MUM (L if none)
    fail = torch.where(ok, L[mum_of], t)
    first_fail = L.clone()
    first_fail.scatter_reduce_(0, mum_of, fail, reduce="amin", include_self=True)
    keep = (t < first_fail[mum_of]) & (t % split_rate == 0)
    return f[keep].cpu().numpy(), (mum_of[keep] + 1).cpu().numpy()


def resolve_marks(n: int, run_starts: np.ndarray, starts: np.ndarray, ids: np.ndarray, heights: np.ndarray,
                  mode_all: bool = False):
    """Sequential restatement of col_split.hpp:105-132 (second pass: one (id,height) per distinct start)
    + find_col_runs (col_split.hpp:258-338).  Returns (positions of set bits of col_runs, id per bit)."""
    # second FL pass, collect_ids (col_split.hpp:114-127): later marks overwrite earlier ones at the same
    # start in tunnel mode; in all mode the taller (earlier on ties) wins.
    slot: dict[int, tuple[int, int]] = {}
    for s, i, h in zip(starts.tolist(), bin_id(ids).tolist(), heights.tolist()):
        if mode_all:
            ei, eh = slot.get(s, (0, 0))
            slot[s] = (ei if eh >= h else i, max(eh, h))
        else:
            slot[s] = (i, h)
    order = sorted(slot)
    out_pos: list[int] = []
    out_id: list[int] = []
    heads = run_starts.tolist() + [None]   # bwt_run_select(r+1) is past the end
    cur = {"run": 1, "last_id": 0}         # run_cursor starts at 1: curr_bwt_pos = select(1) = 0

    def bwt_pos():
        return heads[cur["run"] - 1] if cur["run"] <= len(run_starts) else None

    def update_bwt_pos(idx: int, new_id: int = 0):
        while cur["run"] <= len(run_starts) and bwt_pos() is not None and bwt_pos() < idx:
            out_pos.append(bwt_pos())
            out_id.append(cur["last_id"])
            cur["run"] += 1
        if bwt_pos() is not None and bwt_pos() == idx:
            cur["run"] += 1
        cur["last_id"] = new_id

    heap: list[tuple[int, int, int]] = []   # (end, start, id), min-heap on (end, start)

    def update_col_ranges(idx: int):
        while heap and heap[0][0] <= idx:
            e_end, _e_start, _e_id = heapq.heappop(heap)
            if len(heap) == 1 and heap[0][0] > e_end:
                update_bwt_pos(e_end, heap[0][2])
                out_pos.append(e_end)
                out_id.append(heap[0][2])
            elif not heap and e_end < idx:
                update_bwt_pos(e_end, 0)
                out_pos.append(e_end)
                out_id.append(0)

    for s in order:
        cid, h = slot[s]
        update_col_ranges(s)
        heapq.heappush(heap, (s + h, s, cid))
        if len(heap) == 1 and cid > 0:
            update_bwt_pos(s, cid)
            out_pos.append(s)
            out_id.append(cid)
    update_col_ranges(n)
    update_bwt_pos(n, 0)
    return np.asarray(out_pos, dtype=np.uint64), np.asarray(out_id, dtype=np.uint8)


def resolve_marks_fast(n: int, run_starts: np.ndarray, starts: np.ndarray, ids: np.ndarray, height: int):
    """Vectorised resolve_marks for the common tunnel-mode situation: all marks have one height, lie
    inside a single BWT run and are pairwise disjoint.  Returns None when that does not hold (caller
    falls back to resolve_marks)."""
    if starts.size == 0:
        return run_starts.astype(np.uint64), np.zeros(run_starts.size, dtype=np.uint8)
    # later duplicates of a start overwrite earlier ones
    order = np.argsort(starts, kind="stable")
    s, i = starts[order], bin_id(ids[order])
    last = np.concatenate((s[1:] != s[:-1], [True]))
    s, i = s[last], i[last]
    e = s + height
    if (s[1:] < e[:-1]).any() or e[-1] > n:
        return None
    run_of_s = np.searchsorted(run_starts, s, side="right")
    run_of_e = np.searchsorted(run_starts, e - 1, side="right")
    if (run_of_s != run_of_e).any():
        return None
    # bits: run heads (id = 0 unless a mark starts there), mark starts (id), mark ends (0 unless next mark starts there)
    pos = np.concatenate((run_starts.astype(np.int64), s, e[e < n]))
    idv = np.concatenate((np.zeros(run_starts.size, np.int64), i, np.zeros(int((e < n).sum()), np.int64)))
    prio = np.concatenate((np.zeros(run_starts.size, np.int8), np.full(s.size, 2, np.int8), np.ones(int((e < n).sum()), np.int8)))
    o = np.lexsort((prio, pos))
    pos, idv = pos[o], idv[o]
    lastp = np.concatenate((pos[1:] != pos[:-1], [True]))   # highest priority (mark start) wins a shared position
    return pos[lastp].astype(np.uint64), idv[lastp].astype(np.uint8)


def build_columns(heads: np.ndarray, lens: np.ndarray, thr: np.ndarray, split_pos: np.ndarray, split_ids: np.ndarray,
                  strict: bool = True):
    """Columns of the `.col_pml` table from the primaries, as the reference constructs them.

    heads/lens/thr: per BWT run.  split_pos/split_ids: set bits of `.col_runs` (must include every run head,
    as col_split writes them) and the id byte of each.  Returns dict(ch, idx, interval, offset, col_id, thr,
    n, bwt_r)."""
    heads = np.asarray(heads, dtype=np.uint8)
    lens = np.asarray(lens, dtype=np.int64)
    assert (heads < 128).all(), "reference reads heads into a signed char (col_bwt.hpp:156,165)"
    run_start = np.cumsum(lens) - lens
    n = int(lens.sum())
    split_pos = np.asarray(split_pos, dtype=np.int64)
    assert np.isin(run_start, split_pos).all(), "col_runs must mark every BWT run head"
    idx = split_pos                                   # every set bit starts a row (col_bwt.hpp:177-213)
    run_of = np.searchsorted(run_start, idx, side="right") - 1
    ch = np.where(heads <= 1, 1, heads)[run_of].astype(np.uint8)     # col_bwt.hpp:171
    col_id = np.asarray(split_ids, dtype=np.uint8)
    row_len = np.diff(idx, append=n)
    # compute_table (LF_table.hpp:365-387): rows in stable char order tile F
    order = np.argsort(ch, kind="stable")
    fpos = np.zeros(idx.size, dtype=np.int64)
    fpos[order] = np.cumsum(row_len[order]) - row_len[order]
    interval = np.searchsorted(idx, fpos, side="right") - 1
    offset = fpos - idx[interval]
    assert not strict or (offset.max(initial=0) < 65536 and idx.size < 2**32), "16-bit offset / 32-bit interval fields would wrap"
    return {
        "ch": ch, "idx": idx.astype(np.uint64), "interval": interval.astype(np.uint32),
        "offset": (offset & 0xFFFF).astype(np.uint16), "col_id": col_id, "thr": np.asarray(thr, dtype=np.uint64)[run_of],
        "n": n, "bwt_r": int(heads.size),
    }
