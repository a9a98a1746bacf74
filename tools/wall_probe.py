"""Second wall of k_traverse: kernel-only rate of one library build (COLBWT_LIB) on the C2 reads in
input order and in TRUE-ORIGIN order (where DRAM traffic drops ~3.7x, so whatever still limits the kernel is not DRAM).
Usage: COLBWT_LIB=... python tools/wall_probe.py <label> [orders=input,origin] [workload=c2]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench, col_bwt_b200 as cb
from synthdata import pangenome as P
label = sys.argv[1] if len(sys.argv) > 1 else "default"
orders = (sys.argv[2] if len(sys.argv) > 2 else "input,origin").split(",")
wl = sys.argv[3] if len(sys.argv) > 3 else "c2"
w = bench.WORKLOADS[wl]
path, text, ss, meta = bench.build_workload(wl, "cuda:0", False)
seqs, off, oseq, opos = P.sample_reads_device(np.asarray(text), ss, w["reads"], w["read_len"], sub=w["sub"], seed=2, device="cuda:0", return_origin=True)
tbl = cb.ColPml.load(path)
m = w["read_len"]
reads = seqs.reshape(-1, m)
strand = oseq & 1
end = opos + m
coord = np.where(strand == 0, end, -(end))
perm = {"input": None, "origin": np.lexsort((coord, strand))}
out = {"label": label, "lib": os.environ.get("COLBWT_LIB", "default")}
for o in orders:
    s = np.ascontiguousarray(reads[perm[o]]).reshape(-1) if perm[o] is not None else seqs
    b = tbl.batch(s, off, 1)
    for _ in range(2):
        b.run(1)
    ms = b.run(3)
    out[o] = {"ms": round(ms, 3), "gbases_s": round(seqs.size / ms / 1e6, 2)}
    b.close()
print(json.dumps(out), flush=True)
