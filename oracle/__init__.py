"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Importers allowed by the repo rules: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
`--impl reference` legs.  The product package (col_bwt_b200) must never import this module.

Two checkers live here:
  * `Oracle`     -- oracle/liboracle.so, the plain-C restatement (oracle/colbwt_oracle.c), kind "port";
  * `Reference`  -- oracle/_ref/libref_harness.so, the reference's OWN header-only col_pml compiled from
                    /root/reference where it lies (oracle/harness/ref_harness.cpp), kind "reference".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref (Makefile decides)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def ref_bin(name: str) -> str:
    return os.path.join(REF_DIR, name)


def have_ref() -> bool:
    return os.path.exists(ref_bin("libref_harness.so")) and os.path.exists(ref_bin("pml_query_nomt"))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _seq_arrays(seqs, offsets):
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    return seqs, offsets


class Oracle:
    """Plain-C restatement (kind "port")."""

    def __init__(self, col_pml_path: str | None = None, *, columns: dict | None = None):
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = self.L = C.CDLL(so)
        L.oracle_load.restype = C.c_void_p
        L.oracle_load.argtypes = [C.c_char_p]
        L.oracle_from_columns.restype = C.c_void_p
        L.oracle_from_columns.argtypes = [C.c_uint64] * 3 + [C.c_void_p] * 6
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_query.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.oracle_query_batch.restype = C.c_uint64
        L.oracle_query_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.oracle_text_bound.restype = C.c_size_t
        L.oracle_text_bound.argtypes = [C.c_size_t, C.c_uint64]
        for f in (L.oracle_format_u32, L.oracle_format_u8):
            f.restype = C.c_size_t
            f.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint64]
        if col_pml_path is not None:
            self.h = L.oracle_load(col_pml_path.encode())
        else:
            c = columns
            cols = [np.ascontiguousarray(c["ch"], np.uint8), np.ascontiguousarray(c["idx"], np.uint64),
                    np.ascontiguousarray(c["interval"], np.uint32), np.ascontiguousarray(c["offset"], np.uint16),
                    np.ascontiguousarray(c["col_id"], np.uint8), np.ascontiguousarray(c["thr"], np.uint64)]
            self.h = L.oracle_from_columns(int(c["bwt_r"]), int(c["n"]), len(cols[0]), *[a.ctypes.data for a in cols])
        if not self.h:
            raise OSError(f"oracle: cannot load {col_pml_path}")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_free(self.h)
            self.h = None

    def query(self, read: bytes | np.ndarray):
        r = np.frombuffer(bytes(read), dtype=np.uint8) if not isinstance(read, np.ndarray) else np.ascontiguousarray(read, np.uint8)
        m = r.size
        pml = np.zeros(m, np.uint32)
        cid = np.zeros(m, np.uint8)
        self.L.oracle_query(self.h, r.ctypes.data, m, pml.ctypes.data, cid.ctypes.data)
        return pml, cid

    def query_batch(self, seqs, offsets, want_output: bool = True):
        seqs, offsets = _seq_arrays(seqs, offsets)
        n_reads = offsets.size - 1
        if want_output:
            pml = np.zeros(seqs.size, np.uint32)
            cid = np.zeros(seqs.size, np.uint8)
            self.L.oracle_query_batch(self.h, seqs.ctypes.data, offsets.ctypes.data, n_reads, pml.ctypes.data, cid.ctypes.data)
            return pml, cid
        return self.L.oracle_query_batch(self.h, seqs.ctypes.data, offsets.ctypes.data, n_reads, None, None)

    def format_text(self, ids: list[str], offsets, values) -> bytes:
        """pml_query.cpp:65-90 text for a whole batch (values: u32 PML or u8 CID array)."""
        out = []
        fn = self.L.oracle_format_u32 if values.dtype == np.uint32 else self.L.oracle_format_u8
        for i, name in enumerate(ids):
            a, b = int(offsets[i]), int(offsets[i + 1])
            v = np.ascontiguousarray(values[a:b])
            nm = name.encode()
            buf = C.create_string_buffer(self.L.oracle_text_bound(len(nm), b - a))
            k = fn(buf, nm, len(nm), v.ctypes.data, b - a)
            out.append(buf.raw[:k])
        return b"".join(out)


class Reference:
    """The reference's own col_pml::query_pml, compiled from /root/reference (kind "reference")."""

    def __init__(self, col_pml_path: str):
        so = ref_bin("libref_harness.so")
        if not os.path.exists(so):
            raise OSError("oracle/_ref/libref_harness.so missing (run `make -C oracle` where /root/reference exists)")
        L = self.L = C.CDLL(so)
        L.ref_load.restype = C.c_void_p
        L.ref_load.argtypes = [C.c_char_p]
        L.ref_free.argtypes = [C.c_void_p]
        for f in (L.ref_n, L.ref_r, L.ref_bwt_r):
            f.restype = C.c_ulong
            f.argtypes = [C.c_void_p]
        L.ref_query.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, C.c_void_p]
        L.ref_query_batch.restype = C.c_ulong
        L.ref_query_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        self.h = L.ref_load(col_pml_path.encode())
        if not self.h:
            raise OSError(f"reference harness: cannot load {col_pml_path}")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_free(self.h)
            self.h = None

    @property
    def n(self):
        return self.L.ref_n(self.h)

    @property
    def r(self):
        return self.L.ref_r(self.h)

    def query(self, read: bytes):
        r = np.frombuffer(bytes(read), dtype=np.uint8)
        pml = np.zeros(r.size, np.uint32)
        cid = np.zeros(r.size, np.uint8)
        self.L.ref_query(self.h, r.ctypes.data, r.size, pml.ctypes.data, cid.ctypes.data)
        return pml, cid

    def query_batch(self, seqs, offsets, threads: int = 1, want_output: bool = True):
        seqs, offsets = _seq_arrays(seqs, offsets)
        n_reads = offsets.size - 1
        if want_output:
            pml = np.zeros(seqs.size, np.uint32)
            cid = np.zeros(seqs.size, np.uint8)
            self.L.ref_query_batch(self.h, seqs.ctypes.data, offsets.ctypes.data, n_reads, pml.ctypes.data, cid.ctypes.data, threads)
            return pml, cid
        return self.L.ref_query_batch(self.h, seqs.ctypes.data, offsets.ctypes.data, n_reads, None, None, threads)
