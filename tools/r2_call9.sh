#!/bin/bash
# round 2, GPU call 9: new GPU tests (chain-id-compact transport, counters); what the host cores read/write per second; C2 with the
# six (packing x transport) modes traced; mixed batch (configs[4] at a quarter) traced after the per-chunk growth rule; configs[2] again
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest9.log
g++ -O2 -msse2 -pthread -o /tmp/host_membw tools/probe/host_membw.cpp && /tmp/host_membw 256 > gpurun_out/r2_host_membw.log 2>&1
cat gpurun_out/r2_host_membw.log
COLBWT_TRACE=1 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --check-reads 20000 > gpurun_out/r2_bench_c2_modes.json 2> gpurun_out/r2_bench_c2_modes.err
echo "c2 rc=$?"; grep "colbwt_query\] 15" gpurun_out/r2_bench_c2_modes.err | head -14
COLBWT_TRACE=1 timeout 600 python bench.py --workload c5mid --steps 3 --cpu-seconds 0 --check-reads 60000 > gpurun_out/r2_bench_c5mid.json 2> gpurun_out/r2_bench_c5mid.err
echo "c5mid rc=$?"; grep "colbwt_query\]" gpurun_out/r2_bench_c5mid.err | tail -8
COLBWT_TRACE=1 timeout 600 python bench.py --workload c3 --steps 3 --cpu-seconds 0 --check-reads 3000 > gpurun_out/r2_bench_c3_n1b.json 2> gpurun_out/r2_bench_c3_n1b.err
echo "c3 rc=$?"; grep "colbwt_query\]" gpurun_out/r2_bench_c3_n1b.err | tail -6
for wl in c3 c5mid c3small; do python tools/kernel_only.py $wl >> gpurun_out/r2_kernels9.log 2>> gpurun_out/r2_kernels9.err; done
cat gpurun_out/r2_kernels9.log
python tools/gather_lanes.py 32,96,160,366 > gpurun_out/r2_gather_lanes.log 2>&1
cat gpurun_out/r2_gather_lanes.log
