#!/bin/bash
# round 2, GPU call 4: new tests; in-process multi-replica mode on one GPU; fitted configs[4] at a quarter size; D2H timeline
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1
echo "pytest rc=$?"
COLBWT_TRACE=2 COLBWT_DEVICE_PACK=0 COLBWT_COMPACT_D2H=0 python bench.py --steps 3 --cpu-seconds 10 > gpurun_out/r2_bench_c2_trace2.json 2> gpurun_out/r2_bench_c2_trace2.err
echo "bench trace2 rc=$?"
COLBWT_BENCH_INPROC_SAME_GPU=1 COLBWT_TRACE=1 python bench.py --inproc 2 --reads 3000000 --steps 3 --cpu-seconds 0 --check-reads 20000 > gpurun_out/r2_bench_c2_inproc2same.json 2> gpurun_out/r2_bench_c2_inproc2same.err
echo "inproc rc=$?"
python bench.py --workload c5mid --steps 3 --cpu-seconds 5 --check-reads 60000 --verbose > gpurun_out/r2_bench_c5mid.json 2> gpurun_out/r2_bench_c5mid.err
echo "c5mid rc=$?"
python tools/pcie_concurrent.py > gpurun_out/r2_pcie_1gpu.log 2>&1
tail -3 gpurun_out/r2_pytest4.log
