"""col_bwt_b200 -- B200-native query path of col-bwt (PML + chain ids), Python host mirror over the C-ABI.

The compute lives in `libcolbwt_b200.so` (hand-written sm_100a CUDA, see csrc/); this module is a thin ctypes
binding whose class `ColPml` mirrors the reference's `col_pml` (include/col_bwt.hpp:386-575):

    reference                                   here
    col_pml tbl; tbl.load(ifstream)             tbl = ColPml.load(path)
    tbl.query_pml(pattern, m) -> (pml, cid)     tbl.query_pml(pattern) -> (pml, cid)      (one read)
    per-read loop of pml_query.cpp:74-86        tbl.query(seqs, offsets) -> (pml, cid)    (whole batch)
    tbl.size() / runs() / bwt_runs()            tbl.n / tbl.r / tbl.bwt_r

There is no CPU fallback: importing works without a GPU (so symbols can be checked), every compute call raises
`ColBwtError` if no CUDA device is present, and a missing shared library raises at import.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COLBWT_LIB") or os.path.join(_HERE, "libcolbwt_b200.so")   # COLBWT_LIB: development builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C col_bwt_b200/csrc`). col_bwt_b200 has no CPU fallback."
    )

_L = C.CDLL(LIB_PATH)


class ColBwtError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        super().__init__(f"{where}: status {code}: {_L.colbwt_last_error().decode(errors='replace')}")


class Stats(C.Structure):
    _fields_ = [("n", C.c_uint64), ("r", C.c_uint64), ("bwt_r", C.c_uint64), ("marked_rows", C.c_uint64),
                ("slow_rows", C.c_uint64), ("device_bytes", C.c_uint64), ("max_row_len", C.c_uint32), ("n_devices", C.c_int32)]


# every symbol include/colbwt_b200.h declares
EXPORTS = {
    "colbwt_index_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "colbwt_index_from_rows": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "colbwt_index_from_primaries": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "colbwt_index_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "colbwt_col_split": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "colbwt_rlbwt_to_bwt": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]),
    "colbwt_index_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "colbwt_index_free": (None, [C.c_void_p]),
    "colbwt_query": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]),
    "colbwt_index_last_packing": (C.c_int, [C.c_void_p]),
    "colbwt_index_last_transport": (C.c_int, [C.c_void_p]),
    "colbwt_index_last_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "colbwt_compact_bound": (C.c_size_t, [C.c_void_p, C.c_uint64]),
    "colbwt_query_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "colbwt_compact_expand": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]),
    "colbwt_host_alloc": (C.c_void_p, [C.c_size_t]),
    "colbwt_host_free": (None, [C.c_void_p]),
    "colbwt_batch_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "colbwt_batch_run": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "colbwt_batch_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "colbwt_batch_device_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "colbwt_batch_launches": (C.c_int, [C.c_void_p]),
    "colbwt_batch_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "colbwt_batch_free": (None, [C.c_void_p]),
    "colbwt_format_stats": (C.c_size_t, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.c_uint64]),
    "colbwt_gather_bench": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "colbwt_last_error": (C.c_char_p, []),
    "colbwt_version": (C.c_char_p, []),
}
for _name, (_res, _args) in EXPORTS.items():
    _f = getattr(_L, _name)   # AttributeError here = the library does not export what the header declares
    _f.restype = _res
    _f.argtypes = _args

PML_U8, PML_U16, PML_U32 = 1, 2, 4
_PML_DTYPE = {1: np.uint8, 2: np.uint16, 4: np.uint32}


def version() -> str:
    return _L.colbwt_version().decode()


def _check(rc: int, where: str) -> None:
    if rc != 0:
        raise ColBwtError(rc, where)


def _dev_array(devices):
    if devices is None:
        return None, 1
    if isinstance(devices, int):
        return None, devices
    arr = (C.c_int * len(devices))(*devices)
    return arr, len(devices)


def _as_batch(seqs, offsets):
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if offsets.ndim != 1 or offsets.size < 1:
        raise ValueError("offsets must hold n_reads+1 entries")
    return seqs, offsets


class PinnedArray:
    """numpy view over pinned host memory from colbwt_host_alloc (freed with the object)."""

    def __init__(self, n: int, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = max(1, n * self.dtype.itemsize)
        self.ptr = _L.colbwt_host_alloc(self.nbytes)
        if not self.ptr:
            raise ColBwtError(-6, "colbwt_host_alloc")
        self.array = np.frombuffer((C.c_uint8 * self.nbytes).from_address(self.ptr), dtype=self.dtype, count=n)

    def __del__(self):
        if getattr(self, "ptr", None) and _L is not None:
            self.array = None
            _L.colbwt_host_free(self.ptr)
            self.ptr = None


class Batch:
    """A batch of reads resident on one GPU (colbwt_batch_*): upload once, traverse many times."""

    def __init__(self, index: "ColPml", seqs, offsets, pml_width: int = PML_U16, device_slot: int = 0):
        self.seqs, self.offsets = _as_batch(seqs, offsets)
        self.index = index
        self.pml_width = pml_width
        self.n_bases = int(self.offsets[-1] - self.offsets[0])
        h = C.c_void_p()
        _check(_L.colbwt_batch_upload(index._h, device_slot, self.seqs.ctypes.data, self.offsets.ctypes.data,
                                      self.offsets.size - 1, pml_width, C.byref(h)), "colbwt_batch_upload")
        self._h = h

    def run(self, iters: int = 1) -> float:
        """Traverse `iters` times; returns CUDA-event milliseconds per traversal."""
        ms = C.c_float()
        _check(_L.colbwt_batch_run(self._h, iters, C.byref(ms)), "colbwt_batch_run")
        return float(ms.value)

    @property
    def launches(self) -> int:
        return _L.colbwt_batch_launches(self._h)

    @property
    def counters(self) -> dict:
        """Long-read bookkeeping: work items scheduled longest-first, reads cut into chunk tasks, chunk tasks re-traversed by
        the last run (speculative start did not converge), chunk tasks."""
        out = (C.c_uint64 * 4)()
        _check(_L.colbwt_batch_counters(self._h, out), "colbwt_batch_counters")
        return {"scheduled": int(out[0]), "split_reads": int(out[1]), "retraversed": int(out[2]), "chunk_tasks": int(out[3])}

    def download(self):
        pml = np.zeros(self.n_bases, _PML_DTYPE[self.pml_width])
        cid = np.zeros(self.n_bases, np.uint8)
        _check(_L.colbwt_batch_download(self._h, pml.ctypes.data, cid.ctypes.data), "colbwt_batch_download")
        return pml, cid

    def device_ptrs(self):
        p, c, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(_L.colbwt_batch_device_ptrs(self._h, C.byref(p), C.byref(c), C.byref(n)), "colbwt_batch_device_ptrs")
        return p.value, c.value, n.value

    def close(self):
        if getattr(self, "_h", None) and _L is not None:   # _L is None during interpreter shutdown
            _L.colbwt_batch_free(self._h)
            self._h = None

    __del__ = close


class ColPml:
    """Mirror of the reference's `col_pml` (include/col_bwt.hpp:386-575) with the table in GPU memory."""

    def __init__(self, handle):
        self._h = handle
        st = Stats()
        _check(_L.colbwt_index_stats(self._h, C.byref(st)), "colbwt_index_stats")
        self.stats = st
        self.n, self.r, self.bwt_r = st.n, st.r, st.bwt_r

    @classmethod
    def load(cls, path: str, devices=None) -> "ColPml":
        """col_pml::load (col_bwt.hpp:375-380). `path` = `<prefix>.col_pml` or the prefix itself.
        devices: None (GPU 0), an int (that many GPUs, 0..k-1) or a list of CUDA ordinals."""
        arr, k = _dev_array(devices)
        h = C.c_void_p()
        _check(_L.colbwt_index_load(os.fsencode(path), arr, k, C.byref(h)), "colbwt_index_load")
        return cls(h)

    @classmethod
    def from_rows(cls, rows: np.ndarray, bwt_r: int, n: int, devices=None) -> "ColPml":
        """From memory: `rows` = the packed 18-byte col_thr rows as they sit in the file."""
        raw = np.ascontiguousarray(rows).view(np.uint8).reshape(-1)
        if raw.size % 18:
            raise ValueError("rows must be a whole number of 18-byte records")
        arr, k = _dev_array(devices)
        h = C.c_void_p()
        _check(_L.colbwt_index_from_rows(raw.ctypes.data, bwt_r, n, raw.size // 18, arr, k, C.byref(h)), "colbwt_index_from_rows")
        return cls(h)

    @classmethod
    def from_device_rows(cls, ptr: int, r: int, bwt_r: int, n: int, devices=None) -> "ColPml":
        """Same as from_rows with the 18-byte rows already in GPU memory (`ptr` = device address of r * 18 bytes)."""
        arr, k = _dev_array(devices)
        h = C.c_void_p()
        _check(_L.colbwt_index_from_rows(C.c_void_p(ptr), bwt_r, n, r, arr, k, C.byref(h)), "colbwt_index_from_rows")
        return cls(h)

    @classmethod
    def from_primaries(cls, prefix: str, devices=None) -> "ColPml":
        """col_pml(heads, lengths, col_ids, thresholds, splits) (col_bwt.hpp:391-395) built on the GPU from
        PREFIX.{bwt.heads,bwt.len,thr_pos,col_runs,col_ids} -- what src/build_col_bwt.cpp does on the CPU."""
        arr, k = _dev_array(devices)
        h = C.c_void_p()
        _check(_L.colbwt_index_from_primaries(os.fsencode(prefix), arr, k, C.byref(h)), "colbwt_index_from_primaries")
        return cls(h)

    def save(self, path: str) -> None:
        """col_bwt::serialize (col_bwt.hpp:360-370): writes the `.col_pml` file."""
        _check(_L.colbwt_index_save(self._h, os.fsencode(path)), "colbwt_index_save")

    def size(self) -> int:          # LF_table::size
        return self.n

    def runs(self) -> int:          # LF_table::runs
        return self.r

    def bwt_runs(self) -> int:      # col_bwt::bwt_runs
        return self.bwt_r

    def query(self, seqs, offsets, pml_width: int = PML_U16, out=None):
        """The per-read loop of pml_query.cpp:74-86 for a whole batch (host buffers in, host buffers out).
        out = (pml_array, cid_array) to write into caller (e.g. pinned) buffers."""
        seqs, offsets = _as_batch(seqs, offsets)
        total = int(offsets[-1] - offsets[0])
        if out is None:
            pml = np.empty(total, _PML_DTYPE[pml_width])   # every base of every read is written by the library
            cid = np.empty(total, np.uint8)
        else:
            pml, cid = out
            assert pml.dtype.itemsize == pml_width and pml.size >= total and cid.size >= total
        _check(_L.colbwt_query(self._h, seqs.ctypes.data, offsets.ctypes.data, offsets.size - 1, pml.ctypes.data, pml_width,
                               cid.ctypes.data), "colbwt_query")
        return pml, cid

    def query_compact(self, seqs, offsets, out: np.ndarray | None = None, capacity: int | None = None):
        """Same traversal, compact result (one match bit per base + the non-zero chain ids; include/colbwt_b200.h).
        Returns the used part of the result buffer (uint8 view); `out` = caller (e.g. pinned) uint8 buffer."""
        seqs, offsets = _as_batch(seqs, offsets)
        if out is None:
            cap = capacity or int(_L.colbwt_compact_bound(offsets.ctypes.data, offsets.size - 1))
            out = np.empty(max(cap, 64), np.uint8)
        used = C.c_size_t()
        _check(_L.colbwt_query_compact(self._h, seqs.ctypes.data, offsets.ctypes.data, offsets.size - 1, out.ctypes.data, out.size,
                                       C.byref(used)), "colbwt_query_compact")
        return out[: used.value]

    @property
    def last_transport(self) -> str:
        """How the dense results of the last query() crossed the link: "dense", "compact" (expanded on the host) or
        "pml-dense+cid-compact" (PML copied as it is, only the sparse chain ids in compact form) or "mixed" (chunks alternate
        between "dense" and "compact": the copy engine and the host threads fill the arrays together)."""
        return {0: "dense", 1: "compact", 2: "pml-dense+cid-compact", 3: "mixed"}.get(_L.colbwt_index_last_transport(self._h), "?")

    @property
    def last_bytes(self) -> tuple:
        """(host-to-device, device-to-host) bytes the last query() / query_compact() asked the copy engines to move."""
        a, b = C.c_uint64(), C.c_uint64()
        _check(_L.colbwt_index_last_bytes(self._h, C.byref(a), C.byref(b)), "colbwt_index_last_bytes")
        return int(a.value), int(b.value)

    @property
    def last_packing(self) -> str:
        """Where the last query() packed the reads: "host" or "device" (the library measures both and keeps the faster)."""
        return "device" if _L.colbwt_index_last_packing(self._h) == 1 else "host"

    def query_pml(self, pattern: bytes | str):
        """col_pml::query_pml(pattern) (col_bwt.hpp:403-412): (PML lengths, chain ids), indexed by read position."""
        if isinstance(pattern, str):
            pattern = pattern.encode()
        seq = np.frombuffer(bytes(pattern), dtype=np.uint8)
        width = PML_U16 if seq.size < 65536 else PML_U32
        pml, cid = self.query(seq, np.array([0, seq.size], np.uint64), width)
        return pml.astype(np.uint64), cid.astype(np.uint64)

    def batch(self, seqs, offsets, pml_width: int = PML_U16, device_slot: int = 0) -> Batch:
        return Batch(self, seqs, offsets, pml_width, device_slot)

    def close(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.colbwt_index_free(self._h)
            self._h = None

    __del__ = close


def _aligned_empty(n: int, dtype, align: int = 64, skew: int = 0) -> np.ndarray:
    """n elements whose first byte sits `skew` bytes after a multiple of `align` (line-aligned arrays let the expander
    write whole lines with streaming stores; the skew is for tests of the other path)."""
    item = np.dtype(dtype).itemsize
    raw = np.empty(n * item + align + skew + item, np.uint8)
    start = (-raw.ctypes.data) % align + skew
    return raw[start: start + n * item].view(dtype)


def compact_expand(result: np.ndarray, offsets, pml_width: int = PML_U16, cid_only: bool = False):
    """Dense (pml, cid) arrays from a compact result -- host only, no GPU needed.  cid_only: rebuild the chain ids alone
    (returns (None, cid))."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    result = np.ascontiguousarray(result, dtype=np.uint8)
    total = int(offsets[-1] - offsets[0])
    pml = None if cid_only else _aligned_empty(total, _PML_DTYPE[pml_width])
    cid = _aligned_empty(total, np.uint8)
    cid[:] = 0xEE
    _check(_L.colbwt_compact_expand(result.ctypes.data, offsets.ctypes.data, offsets.size - 1, None if cid_only else pml.ctypes.data, pml_width,
                                    cid.ctypes.data), "colbwt_compact_expand")
    return pml, cid


def compact_segments(result: np.ndarray):
    """The directory of a compact result as a list of dicts (offsets in bytes into `result`)."""
    hdr = np.frombuffer(result[:64].tobytes(), "<u8")
    if hdr[0] != 0x31504d4354574243:
        raise ValueError("not a compact result")
    n = int(hdr[1])
    d = np.frombuffer(result[64: 64 + 72 * n].tobytes(), "<u8").reshape(n, 9)
    keys = ("first_read", "n_reads", "first_base", "n_bases", "match_off", "cid_off", "prefix_off", "values_off", "n_values")
    return [dict(zip(keys, map(int, row))) for row in d]


def col_split(prefix: str, mode: str = "tunnels", split_rate: int = 10, device: int = 0):
    """`col_split PREFIX -m MODE -s RATE` (src/col_split.cpp) on the GPU: writes PREFIX.col_runs / PREFIX.col_ids.
    Returns (set bits of col_runs, bits with a non-zero chain id)."""
    if mode not in ("tunnels", "all"):
        raise ValueError("mode must be 'tunnels' or 'all'")     # col_split.cpp:34-46
    a, b = C.c_uint64(), C.c_uint64()
    _check(_L.colbwt_col_split(os.fsencode(prefix), int(mode == "all"), split_rate, device, C.byref(a), C.byref(b)), "colbwt_col_split")
    return a.value, b.value


def rlbwt_to_bwt(prefix: str, device: int = 0) -> int:
    """`rlbwt_to_bwt PREFIX` (src/rlbwt_to_bwt.cpp:8-34) on the GPU: PREFIX.bwt.heads + PREFIX.bwt.len -> PREFIX.bwt.
    Returns the number of characters written."""
    n = C.c_uint64()
    _check(_L.colbwt_rlbwt_to_bwt(os.fsencode(prefix), device, C.byref(n)), "colbwt_rlbwt_to_bwt")
    return n.value


def format_stats(read_id: str, values: np.ndarray) -> bytes:
    """Text of pml_query.cpp:79-85 for one read."""
    v = np.ascontiguousarray(values)
    if v.dtype.itemsize not in (1, 2, 4):
        v = v.astype(np.uint32)
    rid = read_id.encode()
    need = _L.colbwt_format_stats(None, 0, rid, len(rid), v.ctypes.data, v.dtype.itemsize, v.size)
    buf = C.create_string_buffer(need)
    k = _L.colbwt_format_stats(buf, need, rid, len(rid), v.ctypes.data, v.dtype.itemsize, v.size)
    return buf.raw[:k]


def gather_bench(bytes_: int, loads: int, dependent: bool = False, device: int = 0) -> float:
    """Random 32-byte-sector gather rate (sectors/s) over a buffer of `bytes_` bytes."""
    out = C.c_double()
    _check(_L.colbwt_gather_bench(device, bytes_, loads, int(dependent), C.byref(out)), "colbwt_gather_bench")
    return float(out.value)
