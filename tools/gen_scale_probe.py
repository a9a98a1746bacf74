"""Scratch probe for the GPU box: time synthetic index generation at growing sizes and a first kernel timing."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synthdata import pangenome as P, pipeline as PL, formats as F
import col_bwt_b200 as cb

out = {}
for H, G, nreads in ((4, 1_000_000, 100_000), (32, 1_000_000, 2_000_000), (32, 10_000_000, 10_000_000)):
    if len(sys.argv) > 1 and G * H > int(sys.argv[1]):
        break
    t0 = time.time()
    haps = P.make_haplotypes(G, H, snp=1e-3, indel=1e-4 if H > 4 else 0.0, seed=1)
    t1 = time.time()
    idx = PL.build_index(haps, split_rate=10, device="cuda", verbose=True)
    torch.cuda.synchronize(); t2 = time.time()
    cols = idx["columns"]
    rows = F.rows_from_columns(cols["ch"], cols["idx"], cols["interval"], cols["offset"], cols["col_id"], cols["thr"])
    t3 = time.time()
    tbl = cb.ColPml.from_rows(rows, cols["bwt_r"], cols["n"])
    t4 = time.time()
    seqs, off = P.sample_reads_device(idx["text"], idx["seq_starts"], nreads, 150, sub=0.01, seed=2)
    t5 = time.time()
    b = tbl.batch(seqs, off)
    t6 = time.time()
    ms = [b.run(1) for _ in range(4)]
    pml, cid = b.download()
    st = tbl.stats
    rec = dict(H=H, G=G, n=int(cols["n"]), r=int(st.r), bwt_r=int(st.bwt_r), marked=int(st.marked_rows), slow=int(st.slow_rows),
               max_len=int(st.max_row_len), t_haps=t1 - t0, t_index=t2 - t1, t_rows=t3 - t2, t_upload=t4 - t3, t_reads=t5 - t4,
               t_batch=t6 - t5, kernel_ms=ms, gbases_s=seqs.size / min(ms) / 1e6, mismatch_frac=float((pml == 0).mean()),
               cid_frac=float((cid > 0).mean()), mean_pml=float(pml.mean()))
    print(json.dumps(rec), flush=True)
    out[f"{H}x{G}"] = rec
    del tbl, b, idx, rows, cols, pml, cid, seqs
    torch.cuda.empty_cache()
for nbytes in (64 << 20, 512 << 20, 4 << 30):
    ind = cb.gather_bench(nbytes, 1 << 28, False)
    dep = cb.gather_bench(nbytes, 1 << 26, True)
    print(f"gather {nbytes >> 20} MiB: independent {ind / 1e9:.1f} Gsectors/s, dependent {dep / 1e9:.1f} Gsectors/s", flush=True)
    out[f"gather_{nbytes >> 20}"] = dict(independent=ind, dependent=dep)
json.dump(out, open("gpurun_out/gen_scale_probe.json", "w"), indent=1)
