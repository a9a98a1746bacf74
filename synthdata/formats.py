"""On-disk formats that sit either side of the col-bwt query path.

Every layout here is pinned by an in-tree *reader* of the reference:
  .bwt.heads / .bwt.len   LF_table.hpp:107-111, col_bwt.hpp:167-171   (1 B char, 5 B LE length per BWT run)
  .thr_pos                col_bwt.hpp:446-448                         (5 B LE absolute BWT position per BWT run)
  .col_ids                col_bwt.hpp:178,199                         (1 B per set bit of .col_runs)
  .col_runs               col_split.hpp:384-386 (sdsl bit_vector: u64 bit count + u64 words, LSB first)
  .col_mums               col_split.cpp:90-106                        (5 B num_docs, then 5 B (len,pos) pairs)
  .col_pml                col_bwt.hpp:360-380 + LF_table.hpp:325-357  (4 x u64 header + 18 B packed rows)
  P.pml / P.cid           pml_query.cpp:65-90                         (text: ">id \n" then "v " * m then "\n")

Tooling for tests and bench only; the product reads `.col_pml` in C++ (col_bwt_b200/csrc/index_io.cpp).
"""
from __future__ import annotations

import numpy as np

RW_BYTES = 5  # common.hpp:46

#: packed 18-byte row of the reference's col_thr (LF_table.hpp:33-84, col_bwt.hpp:40-115)
ROW_DTYPE = np.dtype(
    [("ch", "u1"), ("idx", "u1", (5,)), ("interval", "<u4"), ("offset", "<u2"), ("col_id", "u1"), ("thr", "u1", (5,))]
)
assert ROW_DTYPE.itemsize == 18


def u40_pack(v: np.ndarray) -> np.ndarray:
    """uint64 array -> (len, 5) little-endian bytes."""
    v = np.ascontiguousarray(v, dtype="<u8")
    return v.view(np.uint8).reshape(-1, 8)[:, :5].copy()


def u40_unpack(b: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 5)
    out = np.zeros((b.shape[0], 8), dtype=np.uint8)
    out[:, :5] = b
    return out.view("<u8").reshape(-1)


def write_u40(path: str, v: np.ndarray) -> None:
    u40_pack(v).tofile(path)


def read_u40(path: str) -> np.ndarray:
    return u40_unpack(np.fromfile(path, dtype=np.uint8))


def write_primaries(prefix: str, heads: np.ndarray, lens: np.ndarray, thr: np.ndarray) -> None:
    """prefix is the reference's `<out>.fa` stem: writes prefix.bwt.heads/.bwt.len/.thr_pos."""
    np.asarray(heads, dtype=np.uint8).tofile(prefix + ".bwt.heads")
    write_u40(prefix + ".bwt.len", lens)
    write_u40(prefix + ".thr_pos", thr)


def expand_rlbwt(prefix: str) -> np.ndarray:
    """What the reference's rlbwt_to_bwt writes for prefix.bwt.heads / prefix.bwt.len (src/rlbwt_to_bwt.cpp:24-27): each head
    byte as it is, `len` times, over the records both files hold."""
    heads = np.fromfile(prefix + ".bwt.heads", dtype=np.uint8)
    raw = np.fromfile(prefix + ".bwt.len", dtype=np.uint8)
    r = min(heads.size, raw.size // 5)
    lens = u40_unpack(raw[:r * 5].reshape(r, 5))
    return np.repeat(heads[:r], lens.astype(np.int64))


def write_bit_vector(path: str, n: int, positions: np.ndarray) -> None:
    """sdsl::bit_vector layout (u64 bit count, ceil(n/64) u64 words, LSB first)."""
    words = np.zeros((n + 63) // 64, dtype="<u8")
    positions = np.asarray(positions, dtype=np.uint64)
    np.bitwise_or.at(words, (positions >> np.uint64(6)).astype(np.int64), np.uint64(1) << (positions & np.uint64(63)))
    with open(path, "wb") as f:
        f.write(np.uint64(n).tobytes())
        words.tofile(f)


def read_bit_vector(path: str) -> tuple[int, np.ndarray]:
    raw = np.fromfile(path, dtype="<u8")
    n = int(raw[0])
    bits = np.unpackbits(raw[1:].view(np.uint8), bitorder="little")[:n]
    return n, np.flatnonzero(bits).astype(np.uint64)


def write_shim_sd_vector(path: str, n: int, positions: np.ndarray) -> None:
    """The *oracle shim's* private sd_vector layout (oracle/shim/sdsl/sd_vector.hpp), which is what
    oracle/_ref/build_col_bwt expects for `.col_runs` (build_col_bwt.cpp:19-25 loads an sd_vector)."""
    positions = np.asarray(positions, dtype="<u8")
    with open(path, "wb") as f:
        f.write(np.uint64(n).tobytes())
        f.write(np.uint64(positions.size).tobytes())
        positions.tofile(f)


def write_col_mums(path: str, num_docs: int, lens: np.ndarray, pos: np.ndarray) -> None:
    vals = np.empty(1 + 2 * len(lens), dtype=np.uint64)
    vals[0] = num_docs
    vals[1::2] = lens
    vals[2::2] = pos
    write_u40(path, vals)


def read_col_pml(path: str) -> tuple[dict, np.ndarray]:
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(32), dtype="<u8")
        rows = np.frombuffer(f.read(), dtype=ROW_DTYPE)
    meta = {"bwt_r": int(hdr[0]), "n": int(hdr[1]), "r": int(hdr[2]), "size": int(hdr[3])}
    return meta, rows


def write_col_pml(path: str, bwt_r: int, n: int, rows: np.ndarray) -> None:
    assert rows.dtype == ROW_DTYPE
    with open(path, "wb") as f:
        f.write(np.array([bwt_r, n, len(rows), len(rows)], dtype="<u8").tobytes())
        rows.tofile(f)


def rows_from_columns(ch, idx, interval, offset, col_id, thr) -> np.ndarray:
    rows = np.zeros(len(ch), dtype=ROW_DTYPE)
    rows["ch"] = ch
    rows["idx"] = u40_pack(np.asarray(idx, dtype=np.uint64))
    rows["interval"] = interval
    rows["offset"] = offset
    rows["col_id"] = col_id
    rows["thr"] = u40_pack(np.asarray(thr, dtype=np.uint64))
    return rows


def write_fasta(path: str, reads: list[bytes] | None = None, *, seqs: np.ndarray | None = None,
                offsets: np.ndarray | None = None, names: list[str] | None = None) -> None:
    with open(path, "wb") as f:
        if reads is not None:
            for i, r in enumerate(reads):
                nm = names[i] if names else f"q{i}"
                f.write(b">" + nm.encode() + b"\n" + bytes(r) + b"\n")
        else:
            buf = seqs.tobytes()
            for i in range(len(offsets) - 1):
                nm = names[i] if names else f"q{i}"
                f.write(b">" + nm.encode() + b"\n" + buf[int(offsets[i]):int(offsets[i + 1])] + b"\n")


def parse_stat_text(path: str) -> tuple[list[str], list[np.ndarray]]:
    """Parse the reference's PATTERN.pml / PATTERN.cid text (pml_query.cpp:65-90)."""
    names, vals = [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    while i + 1 < len(lines):
        if lines[i].startswith(b">"):
            names.append(lines[i][1:].rstrip(b" ").decode())
            vals.append(np.array(lines[i + 1].split(), dtype=np.uint64))
            i += 2
        else:
            i += 1
    return names, vals
