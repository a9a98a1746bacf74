"""Scratch: random-line gather rate vs buffer size (effective L2 capacity curve, and TLB reach at tens of GB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import col_bwt_b200 as cb
sizes = [int(x) for x in sys.argv[1:]] or [32, 64, 96, 128, 160, 192, 256, 366, 512, 1024]
for mb in sizes:
    r = cb.gather_bench(mb << 20, 1 << 28, 0)
    print(f"{mb:6d} MiB {r / 1e9:7.1f} G loads/s", flush=True)
