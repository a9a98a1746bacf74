"""Scratch probe: random-sector gather rate for each load flavour (see traverse.cu: ld_variant)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import col_bwt_b200 as cb
names = {0: "ld.global.nc", 1: "ld.global.cg", 2: "ld.global.cs", 3: "nc.L1::no_allocate", 4: "ld.global.cv", 5: "L1::no_allocate", 6: "ld.global.lu", 7: "nc.L1::evict_last", 8: "ld.global"}
for nbytes in (366 << 20, 4 << 30):
    for v in range(9):
        ind = cb.gather_bench(nbytes, 1 << 28, (v << 8) | 0)
        dep = cb.gather_bench(nbytes, 1 << 27, (v << 8) | 1)
        print(f"{nbytes >> 20:5d} MiB  v{v} {names[v]:20s} independent {ind / 1e9:7.1f} G/s   dependent {dep / 1e9:7.1f} G/s", flush=True)
