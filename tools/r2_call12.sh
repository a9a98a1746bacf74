#!/bin/bash
# round 2, GPU call 12: per-chunk stream timelines (COLBWT_TRACE=2) of the pinned modes on C2 -- where do host packing, the two
# copy directions and the kernels overlap, and where not
set -x
mkdir -p gpurun_out
for mode in "0 0" "1 0" "0 2" "1 2" "0 1" "1 1"; do
  set -- $mode
  COLBWT_TRACE=2 COLBWT_DEVICE_PACK=$1 COLBWT_COMPACT_D2H=$2 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --check-reads 2000 > gpurun_out/r2_tl_dp$1_ct$2.json 2> gpurun_out/r2_tl_dp$1_ct$2.err
  echo "dp=$1 ct=$2 rc=$?"
  grep -A9 "stream time per stage" gpurun_out/r2_tl_dp$1_ct$2.err | tail -22
  grep "colbwt_query\] 15" gpurun_out/r2_tl_dp$1_ct$2.err | tail -2
done
