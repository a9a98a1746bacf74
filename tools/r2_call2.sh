#!/bin/bash
# round 2, GPU call 2: new GPU tests; c2 bench with the auto mode choice; the four pinned e2e modes
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?"
COLBWT_TRACE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_auto.json 2> gpurun_out/r2_bench_c2_auto.err
echo "bench rc=$?"
for dp in 0 1; do for ct in 0 1; do
  COLBWT_TRACE=1 COLBWT_DEVICE_PACK=$dp COLBWT_COMPACT_D2H=$ct python bench.py --steps 5 --cpu-seconds 0 --check-reads 2000 > gpurun_out/r2_bench_c2_dp${dp}_ct${ct}.json 2> gpurun_out/r2_bench_c2_dp${dp}_ct${ct}.err
  echo "dp=$dp ct=$ct rc=$?"
done; done
nproc; free -g | head -2
