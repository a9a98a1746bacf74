#!/bin/bash
# round 2, GPU call 1: long-read kernel profile on c3small (no code change), launch list, per-stage trace
set -x
mkdir -p gpurun_out
python bench.py --workload c3small --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/r2_c3small_base.json 2> gpurun_out/r2_c3small_base.err
echo "bench rc=$?"
PMLW=4 python tools/variant_sweep.py child c3small > gpurun_out/r2_c3small_child.log 2>&1
echo "child rc=$?"
PMLW=4 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file gpurun_out/r2_launches_c3small.csv python tools/variant_sweep.py child c3small > gpurun_out/r2_ncu_launches.log 2>&1
PMLW=4 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_traverse|k_fixup' -s 6 -c 2 -f -o gpurun_out/r2_c3small_long python tools/variant_sweep.py child c3small > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/
