"""Scratch: random-line gather rate vs buffer size (effective L2 capacity curve)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import col_bwt_b200 as cb
for mb in (32, 64, 96, 128, 160, 192, 256, 366, 512, 1024):
    r = cb.gather_bench(mb << 20, 1 << 28, 0)
    print(f"{mb:5d} MiB {r / 1e9:7.1f} G loads/s", flush=True)
