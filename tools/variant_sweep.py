"""Scratch: time the traversal kernel variants (COLBWT_VARIANT) on the cached bench workload; one subprocess each."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import bench, col_bwt_b200 as cb
    wl = sys.argv[2]
    path, text, ss, meta = bench.build_workload(wl, "cuda:0", False)
    seqs, off = bench.make_reads(wl, text, ss, 0, int(sys.argv[3]) if len(sys.argv) > 3 else None, "cuda:0")
    tbl = cb.ColPml.load(path)
    b = tbl.batch(seqs, off, int(os.environ.get('PMLW', '2')))
    for _ in range(3): b.run(1)
    ms = b.run(5)
    print(json.dumps({"narrow": os.environ.get("COLBWT_NARROW", ""), "ctas": os.environ.get("COLBWT_CTAS", ""), "ms": ms, "gbases_s": seqs.size / ms / 1e6}))
else:
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "3", "4", "5", "7"]
    for v in variants:
        env = dict(os.environ)
        env["COLBWT_NARROW"], env["COLBWT_CTAS"] = v.split(":")
        r = subprocess.run([sys.executable, __file__, "child", wl] + sys.argv[3:], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-500:], flush=True)
