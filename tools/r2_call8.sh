#!/bin/bash
# round 2, GPU call 8: in-trip resolution (COLBWT_INTRIP=1 build) against the default build on every workload class, the
# long-read geometry rule (auto) against pinned geometries, configs[4] at full size on one GPU
set -x
mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/r2_lscpu.log
I=$PWD/col_bwt_b200/_sbintrip/libcolbwt_b200.so
python tools/wall_probe.py base > gpurun_out/r2_intrip_wall.log 2> gpurun_out/r2_intrip_wall.err
COLBWT_LIB=$I python tools/wall_probe.py intrip >> gpurun_out/r2_intrip_wall.log 2>> gpurun_out/r2_intrip_wall.err
cat gpurun_out/r2_intrip_wall.log
for wl in c3small c3 c2small c1 c5mid; do
  python tools/kernel_only.py $wl >> gpurun_out/r2_intrip_kernels.log 2>> gpurun_out/r2_intrip_kernels.err
  COLBWT_LIB=$I python tools/kernel_only.py $wl >> gpurun_out/r2_intrip_kernels.log 2>> gpurun_out/r2_intrip_kernels.err
done
cat gpurun_out/r2_intrip_kernels.log
python tools/longread_sweep.py default c3 > gpurun_out/r2_longread_sweep2.log 2> gpurun_out/r2_longread_sweep2.err
COLBWT_LIB=$I python tools/longread_sweep.py intrip c3 >> gpurun_out/r2_longread_sweep2.log 2>> gpurun_out/r2_longread_sweep2.err
cat gpurun_out/r2_longread_sweep2.log
timeout 700 python bench.py --workload c5 --steps 3 --cpu-seconds 5 --check-reads 100000 --verbose > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err
echo "c5 rc=$?"; tail -3 gpurun_out/r2_bench_c5_n1.err; cat gpurun_out/r2_bench_c5_n1.json | cut -c1-600
