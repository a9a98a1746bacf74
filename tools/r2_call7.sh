#!/bin/bash
# round 2, GPU call 7 (re-entry: the outputs of calls 4-6 were lost with the container): GPU tests, C2 bench line + trace,
# reference arm, 1-GPU link ceiling, build chain timing at C2 scale, configs[3] -m all -s 100, launch list, configs[2] at n = 1.5e9
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest7.log
COLBWT_TRACE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c2_reference.json 2> gpurun_out/r2_bench_c2_reference.err
echo "reference rc=$?"
python tools/pcie_concurrent.py > gpurun_out/r2_pcie_1gpu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --check-reads 2000 > gpurun_out/r2_ncu_launches_c2.log 2>&1
echo "ncu launches rc=$?"
COLBWT_TRACE=1 timeout 600 python tools/build_compare.py --H 32 --G 10000000 > gpurun_out/r2_build_compare_c2.json 2> gpurun_out/r2_build_compare_c2.err
echo "build_compare c2 rc=$?"
COLBWT_TRACE=1 timeout 600 python tools/build_compare.py --H 32 --G 10000000 --mode all --rate 10 > gpurun_out/r2_build_compare_c2_all.json 2> gpurun_out/r2_build_compare_c2_all.err
echo "build_compare c2 all rc=$?"
timeout 600 python bench.py --workload c4_all_s100 --steps 5 --cpu-seconds 0 --check-reads 20000 > gpurun_out/r2_bench_c4_all_s100.json 2> gpurun_out/r2_bench_c4_all_s100.err
echo "c4_all_s100 rc=$?"
timeout 900 python bench.py --workload c3 --steps 3 --cpu-seconds 5 --check-reads 3000 --verbose > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
echo "c3 rc=$?"; tail -5 gpurun_out/r2_bench_c3_n1.err
timeout 400 python tools/longread_sweep.py default c3 > gpurun_out/r2_longread_sweep.log 2> gpurun_out/r2_longread_sweep.err
cat gpurun_out/r2_longread_sweep.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv; nproc; free -g | head -2
