"""Executable-level drop-in check on the GPU box: the reference's pml_query (as shipped, and with MULTI_THREAD off), built
into oracle/_ref, against col_bwt_b200/bin/pml_query_b200 on the same index file and FASTA: wall time and `cmp` of outputs.
Workload = BASELINE configs[0] (4 haplotypes x 1 Mbp + revcomp, tunnels -s 10, 100k x 150 bp reads)."""
import os, subprocess, sys, time, shutil, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from synthdata import pangenome as P, pipeline as PL, formats as F

d = "/tmp/cli_compare"
os.makedirs(d, exist_ok=True)
n_reads = int(sys.argv[sys.argv.index("--reads") + 1]) if "--reads" in sys.argv else 100_000
haps = P.make_haplotypes(1_000_000, 4, snp=1e-3, seed=1)
idx = PL.build_index(haps, split_rate=10, device="cuda", verbose=True)
PL.write_col_pml(f"{d}/c1.fa.col_pml", idx["columns"])
seqs, off = P.sample_reads_device(idx["text"], idx["seq_starts"], n_reads, 150, sub=0.01, seed=2)
F.write_fasta(f"{d}/reads.fa", seqs=seqs, offsets=off)
res = {}
def run(name, exe, skip=False):
    for ext in (".pml", ".cid"):
        if os.path.exists(f"{d}/reads.fa{ext}"): os.remove(f"{d}/reads.fa{ext}")
    t = time.time()
    r = subprocess.run([exe, f"{d}/c1.fa", "-p", f"{d}/reads.fa", "-v"], capture_output=True, text=True)
    if name.startswith("b200"):
        print("   ", [l.strip() for l in r.stdout.split("\n") if "end to end" in l])
    dt = time.time() - t
    assert r.returncode == 0, r.stderr
    for ext in (".pml", ".cid"):
        shutil.move(f"{d}/reads.fa{ext}", f"{d}/{name}{ext}")
    res[name] = dt
    print(f"{name}: {dt:.2f} s wall ({seqs.size / dt / 1e6:.2f} Mbases/s incl. index load, FASTA parse, text output)", flush=True)
run("b200", os.path.join(ROOT, "col_bwt_b200", "bin", "pml_query_b200"))
run("b200_again", os.path.join(ROOT, "col_bwt_b200", "bin", "pml_query_b200"))
run("ref_nomt", os.path.join(ROOT, "oracle", "_ref", "pml_query_nomt"))
if "--shipped" in sys.argv:
    run("ref_shipped", os.path.join(ROOT, "oracle", "_ref", "pml_query"))
for other in [k for k in res if k.startswith("ref")]:
    for ext in (".pml", ".cid"):
        same = open(f"{d}/b200{ext}", "rb").read() == open(f"{d}/{other}{ext}", "rb").read()
        print(f"cmp b200{ext} {other}{ext}: {'identical' if same else 'DIFFERENT'}")
        assert same
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "cli_compare.json"), "w"))
