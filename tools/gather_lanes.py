"""Probe: does the DEPENDENT random-gather rate (one 16-byte load per lane, next address from the loaded value -- the access
pattern of k_traverse) still grow with the lanes resident per SM, and what latency does Little's law give?  Table sizes from
L2-resident to DRAM-resident; 17 KB of dynamic shared memory per CTA keeps the L1 at the size k_traverse runs with.
One subprocess per point (the knobs are read from the environment).  Usage: python tools/gather_lanes.py [MiB,MiB,...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32, 96, 160, 366]
code = ("import sys; sys.path.insert(0, %r); import col_bwt_b200 as cb; "
        "print(cb.gather_bench(int(sys.argv[1]) << 20, 1 << 27, 1))" % ROOT)
for mib in sizes:
    for ctas in (2, 3, 4, 6, 8):
        env = dict(os.environ, COLBWT_GB_CTAS=str(ctas), COLBWT_GB_SMEM=str(17 * 1024))
        out = subprocess.run([sys.executable, "-c", code, str(mib)], env=env, capture_output=True, text=True)
        try:
            rate = float(out.stdout.strip().splitlines()[-1])
            lanes = ctas * 256
            # Little: lanes per SM in flight / (gathers per second per SM) = seconds per dependent gather
            lat_ns = lanes * 148 / rate * 1e9
            print(f"{mib:5d} MiB  {ctas} CTAs/SM ({lanes:4d} lanes)  {rate / 1e9:7.1f} G gathers/s   {lat_ns:7.0f} ns per dependent gather (Little)", flush=True)
        except Exception:
            print(f"{mib:5d} MiB  {ctas} CTAs/SM failed: {out.stderr[-200:]}", flush=True)
