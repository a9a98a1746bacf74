"""ctypes front-end of tests/native/libemulate.so (host emulation of the product's device logic). Test infrastructure."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "native", "libemulate.so")
SRC = [os.path.join(HERE, "native", "emulate.cpp"), os.path.join(HERE, "..", "col_bwt_b200", "csrc", "pack.cpp")]
HDR = os.path.join(HERE, "..", "col_bwt_b200", "csrc", "colbwt_core.cuh")
HDR2 = os.path.join(HERE, "..", "col_bwt_b200", "csrc", "tasks.h")
HDR3 = os.path.join(HERE, "..", "col_bwt_b200", "csrc", "fastx.h")


EXTRA = os.environ.get("COLBWT_EMU_CXXFLAGS", "").split()   # e.g. -DCOLBWT_STAGE_BLOCK=32 (development variants)
if EXTRA:
    SO = SO[:-3] + "_" + "".join(ch for ch in "".join(EXTRA) if ch.isalnum()) + ".so"


def build():
    newest = max(os.path.getmtime(p) for p in SRC + [HDR, HDR2, HDR3])
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO] + EXTRA + SRC + ["-lz"], check=True)


class Emu:
    def __init__(self, cols: dict):
        build()
        L = self.L = C.CDLL(SO)
        L.emu_build.restype = C.c_void_p
        L.emu_build.argtypes = [C.c_uint64, C.c_uint64] + [C.c_void_p] * 6
        L.emu_free.argtypes = [C.c_void_p]
        L.emu_slow_rows.restype = C.c_uint64
        L.emu_slow_rows.argtypes = [C.c_void_p]
        L.emu_flags.restype = C.c_uint32
        L.emu_flags.argtypes = [C.c_void_p]
        L.emu_rows.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_query.restype = C.c_uint64
        L.emu_query.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.keep = [np.ascontiguousarray(cols["ch"], np.uint8), np.ascontiguousarray(cols["idx"], np.uint64),
                     np.ascontiguousarray(cols["interval"], np.uint32), np.ascontiguousarray(cols["offset"], np.uint16),
                     np.ascontiguousarray(cols["col_id"], np.uint8), np.ascontiguousarray(cols["thr"], np.uint64)]
        self.r = len(self.keep[0])
        self.h = L.emu_build(int(cols["n"]), self.r, *[a.ctypes.data for a in self.keep])

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_free(self.h)
            self.h = None

    @property
    def slow_rows(self):
        return self.L.emu_slow_rows(self.h)

    @property
    def flags(self):
        return self.L.emu_flags(self.h)

    def rows(self):
        out = np.zeros((self.r, 4), np.uint32)
        self.L.emu_rows(self.h, out.ctypes.data)
        return out

    def query_split(self, seqs, offsets, chunk, warm, min_len=0, pml_width=2, narrow=False):
        """Long-read path: chunk tasks + chain verification, as the driver plans them. Sets self.redone / self.tasks."""
        L = self.L
        L.emu_query_split.restype = C.c_uint64
        L.emu_query_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        seqs = np.ascontiguousarray(seqs, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        pml = np.zeros(seqs.size + 8, {1: np.uint8, 2: np.uint16, 4: np.uint32}[pml_width])
        cid = np.zeros(seqs.size + 8, np.uint8)
        nt = C.c_uint64()
        self.redone = L.emu_query_split(self.h, seqs.ctypes.data, offsets.ctypes.data, offsets.size - 1, pml.ctypes.data, pml_width,
                                        cid.ctypes.data, chunk, warm, min_len, int(narrow), C.byref(nt))
        self.tasks = nt.value
        return pml[:seqs.size].astype(np.uint32), cid[:seqs.size]

    def query(self, seqs, offsets, pml_width=2, force_bytes=False, narrow=False, defer=False, intrip=False):
        """defer: post flush requests and serve them with flush_word, as the warps of k_traverse do.
        intrip: the kernel variant that resolves same-line neighbours inside the trip (implies defer)."""
        seqs = np.ascontiguousarray(seqs, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        pml = np.zeros(seqs.size + 8, {1: np.uint8, 2: np.uint16, 4: np.uint32}[pml_width])
        cid = np.zeros(seqs.size + 8, np.uint8)
        self.iters = self.L.emu_query(self.h, seqs.ctypes.data, offsets.ctypes.data, offsets.size - 1, pml.ctypes.data,
                                      pml_width, cid.ctypes.data, int(force_bytes) | (2 if narrow else 0) | (4 if (defer or intrip) else 0) | (8 if intrip else 0))
        return pml[:seqs.size].astype(np.uint32), cid[:seqs.size]
