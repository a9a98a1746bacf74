"""CPU: the CLI's FASTA/FASTQ(.gz) batch reader (col_bwt_b200/csrc/fastx.h, replaces PatternProcessor + kseq,
include/common/io.hpp:6-35) against an independent Python parse of the same kseq rules and against the records the
reference's own pml_query saw (golden headers)."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from emu import SO, build
from util import parse_fastx


def read_with_cli_reader(path, max_bases=1 << 30, max_reads=1 << 30):
    build()
    L = C.CDLL(SO)
    L.emu_fastx.restype = C.c_uint64
    L.emu_fastx.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_void_p]
    cap = max(1024, 4 * os.path.getsize(path) + 1024) if not path.endswith(".gz") else 1 << 24
    seqs = np.zeros(cap, np.uint8)
    off = np.zeros(cap // 2 + 2, np.uint64)
    ids = C.create_string_buffer(cap)
    nb = C.c_uint64()
    n = L.emu_fastx(path.encode(), max_bases, max_reads, seqs.ctypes.data, cap, off.ctypes.data, off.size, ids, cap, C.byref(nb))
    assert n < 2**63, "reader failed"
    names = ids.value.decode().split("\n")[:-1] if n else []
    return names, seqs[: int(off[n])].copy(), off[: n + 1].copy(), nb.value


@pytest.mark.parametrize("case", ["toy", "pan4"])
def test_reader_sees_what_the_reference_saw(golden_dir, case):
    path = os.path.join(golden_dir, f"{case}_reads.fa")
    names, seqs, off, _ = read_with_cli_reader(path)
    ids, pseqs, poff = parse_fastx(path)
    assert names == ids and np.array_equal(seqs, pseqs) and np.array_equal(off, poff)
    # the reference's output lists one header per record, in order, with the id up to the first blank
    golden_ids = [l[1:].rstrip(b" ").decode() for l in open(path + ".pml", "rb").read().split(b"\n") if l.startswith(b">")]
    assert names == golden_ids
    golden_lens = [len(l.split()) for l in open(path + ".pml", "rb").read().split(b"\n")[1::2]][: len(names)]
    assert np.diff(off).tolist() == golden_lens


def test_fastq_gzip_crlf_multiline_and_batching(tmp_path):
    recs = [("r1", b"ACGTACGTAC"), ("r2", b""), ("r3", b"NNNNacgt"), ("r4", b"A" * 1000), ("r5", b"GATTACA")]
    fa = tmp_path / "x.fa"
    with open(fa, "wb") as f:
        f.write(b">r1 first comment\r\nACGTA\r\nCGTAC\r\n>r2\n\n>r3\tx\nNNNN\nacgt\n>r4\n" + b"A" * 400 + b"\n" + b"A" * 600 + b"\n>r5\nGATTACA")
    fq = tmp_path / "x.fq.gz"
    with gzip.open(fq, "wb") as f:
        for name, s in recs:
            f.write(b"@" + name.encode() + b" c\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    for path in (str(fa), str(fq)):
        names, seqs, off, nb = read_with_cli_reader(path)
        assert names == [r[0] for r in recs]
        assert [bytes(seqs[int(off[i]):int(off[i + 1])]) for i in range(len(recs))] == [r[1] for r in recs]
        # tiny batches: same records, more batches
        names2, seqs2, off2, nb2 = read_with_cli_reader(path, max_bases=12, max_reads=2)
        assert names2 == names and np.array_equal(seqs2, seqs) and np.array_equal(off2, off) and nb2 >= 3
    assert read_with_cli_reader(str(fa))[0] == parse_fastx(str(fa))[0]


def test_missing_file_reports_failure(tmp_path):
    build()
    L = C.CDLL(SO)
    L.emu_fastx.restype = C.c_uint64
    L.emu_fastx.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_void_p]
    buf = np.zeros(16, np.uint8)
    off = np.zeros(4, np.uint64)
    ids = C.create_string_buffer(16)
    nb = C.c_uint64()
    # gzopen on a missing file fails
    assert L.emu_fastx(str(tmp_path / "nope.fa").encode(), 10, 10, buf.ctypes.data, 16, off.ctypes.data, 4, ids, 16, C.byref(nb)) == 2**64 - 1
