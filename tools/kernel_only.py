"""Kernel-only traversal of one bench workload (file-based or directly synthesised), for ncu captures:
    ncu ... python tools/kernel_only.py <workload> [reads]
3 warm-up traversals, then 5 timed ones; prints one JSON line (with CRC-32s of the outputs, so that library variants --
COLBWT_LIB -- can be compared for equality)."""
import json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench, col_bwt_b200 as cb
wl = sys.argv[1]
reads = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = bench.WORKLOADS[wl]
if "snap" in w:
    rows, n, seqs, off, _ = bench.synth_workload(wl, 0, "cuda:0", False, False)
    tbl = cb.ColPml.from_device_rows(rows.data_ptr(), int(rows.shape[0]), int(rows.shape[0]), int(n))
    del rows
    torch.cuda.empty_cache()
    if reads:
        seqs, off = seqs[: int(off[reads])], off[: reads + 1]
else:
    path, text, ss, meta = bench.build_workload(wl, "cuda:0", False)
    seqs, off = bench.make_reads(wl, text, ss, 0, reads, "cuda:0")
    tbl = cb.ColPml.load(path)
width = bench.pml_width_for(cb, int(np.diff(off).max()))
b = tbl.batch(seqs, off, width)
for _ in range(3):
    b.run(1)
ms = b.run(5)
pml, cid = b.download()
crc = [zlib.crc32(memoryview(np.ascontiguousarray(pml)).cast("B")), zlib.crc32(memoryview(np.ascontiguousarray(cid)).cast("B"))]
print(json.dumps({"workload": wl, "lib": os.environ.get("COLBWT_LIB", "default"), "crc32_pml_cid": crc, "ms": ms, "gbases_s": seqs.size / ms / 1e6, "pml_bytes": width, "launches": b.launches, "long_reads": b.counters}))
