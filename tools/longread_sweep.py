"""Long-read kernel (chunk tasks + k_fixup): kernel-only rate of one library build (COLBWT_LIB) on a long-read workload for
several chunk / warm-up geometries (COLBWT_SPLIT_CHUNK / COLBWT_SPLIT_WARM are read per batch upload).
Usage: python tools/longread_sweep.py <label> [workload=c3] [reads]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench, col_bwt_b200 as cb
label = sys.argv[1] if len(sys.argv) > 1 else "default"
wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
reads = int(sys.argv[3]) if len(sys.argv) > 3 else None
path, text, ss, meta = bench.build_workload(wl, "cuda:0", False)
seqs, off = bench.make_reads(wl, text, ss, 0, reads, "cuda:0")
tbl = cb.ColPml.load(path)
width = bench.pml_width_for(cb, int(np.diff(off).max()))
geoms = [None, (4096, 512), (2048, 256), (3072, 384), (1536, 256), (1024, 256), (2048, 512)]   # None = the library's own choice (tasks.h: adapted)
ref = None
for g in geoms:
    chunk, warm = g if g else ("auto", "auto")
    for k in ("COLBWT_SPLIT_CHUNK", "COLBWT_SPLIT_WARM"):
        os.environ.pop(k, None)
    if g:
        os.environ["COLBWT_SPLIT_CHUNK"], os.environ["COLBWT_SPLIT_WARM"] = str(chunk), str(warm)
    b = tbl.batch(seqs, off, width)
    for _ in range(2):
        b.run(1)
    ms = b.run(3)
    pml, cid = b.download()
    if ref is None:
        ref = (pml, cid)
    same = bool(np.array_equal(pml, ref[0]) and np.array_equal(cid, ref[1]))
    print(json.dumps({"label": label, "workload": wl, "reads": int(off.size - 1), "chunk": chunk, "warm": warm, "ms": round(ms, 2),
                      "gbases_s": round(seqs.size / ms / 1e6, 2), "same_output_as_first": same, **b.counters}), flush=True)
    b.close()
