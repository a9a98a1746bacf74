"""Probe: does the random-gather rate depend on how much of the SM's 256 KB is left to the L1?
Dependent 16-byte gathers over a 366 MiB table (the C2 table size), varying resident lanes per SM and the dynamic
shared memory each CTA reserves.  One subprocess per point (the knobs are read from the environment)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = ("import sys; sys.path.insert(0, %r); import col_bwt_b200 as cb; "
        "print(cb.gather_bench(int(sys.argv[1]), 1 << 27, (int(sys.argv[2]) << 8) | 1))" % ROOT)
names = {0: "nc", 1: "cg", 3: "nc.no_allocate", 8: "plain"}
for nbytes in (366 << 20, 4 << 30):
    for ctas in (3, 4, 5, 6, 8):
        for smem_kb in (0, 17, 33, 49):
            if ctas * (smem_kb + 1) > 227:
                continue
            row = []
            for v in (0, 1, 3):
                env = dict(os.environ, COLBWT_GB_CTAS=str(ctas), COLBWT_GB_SMEM=str(smem_kb * 1024))
                out = subprocess.run([sys.executable, "-c", code, str(nbytes), str(v)], env=env, capture_output=True, text=True)
                try:
                    row.append(f"{names[v]} {float(out.stdout.strip().splitlines()[-1]) / 1e9:6.1f}")
                except Exception:
                    row.append(f"{names[v]} failed: {out.stderr[-200:]}")
            print(f"{nbytes >> 20:5d} MiB  {ctas} CTAs/SM ({ctas * 256:4d} lanes)  smem/CTA {smem_kb:2d} KB (total {ctas * smem_kb:3d} KB)   " + "   ".join(row) + "  G gathers/s", flush=True)
