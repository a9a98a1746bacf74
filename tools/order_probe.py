"""Scratch: how much would a locality-aware read order buy?  Reads of the bench workload are traversed (kernel-only) in
input order, in the order of their TRUE origin (strand, end coordinate) -- an upper bound no real scheduler can reach --
and after a shuffle."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench, col_bwt_b200 as cb
from synthdata import pangenome as P
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[wl]
path, text, ss, meta = bench.build_workload(wl, "cuda:0", False)
seqs, off, oseq, opos = P.sample_reads_device(np.asarray(text), ss, w["reads"], w["read_len"], sub=w["sub"], seed=2, device="cuda:0", return_origin=True)
tbl = cb.ColPml.load(path)
m = w["read_len"]
reads = seqs.reshape(-1, m)
def run(order, name):
    s = np.ascontiguousarray(reads[order]).reshape(-1) if order is not None else seqs
    b = tbl.batch(s, off, 1)
    for _ in range(2): b.run(1)
    ms = b.run(3)
    print(f"{name:28s} {ms:8.2f} ms  {seqs.size / ms / 1e6:7.1f} Gbases/s", flush=True)
    b.close()
run(None, "input order")
# forward strands have even sequence index; the reverse-complement strand runs the other way along the genome
strand = oseq & 1
end = opos + m
coord = np.where(strand == 0, end, -(end))
run(np.lexsort((coord, strand)), "true origin (strand, end)")
run(np.random.default_rng(0).permutation(reads.shape[0]), "shuffled")
