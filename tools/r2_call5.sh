#!/bin/bash
# round 2, GPU call 5: col_split rewrite (tests + C2-scale timing), ncu of the DRAM-resident synthetic table, configs[3] -m all -s 100, configs[2] at n = 1.5e9
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest5.log
python tools/build_compare.py --H 32 --G 10000000 > gpurun_out/r2_build_compare_c2.json 2> gpurun_out/r2_build_compare_c2.err
echo "build_compare c2 rc=$?"
python tools/build_compare.py --H 32 --G 1000000 --mode all --rate 2 > gpurun_out/r2_build_compare_c2small_all.json 2> gpurun_out/r2_build_compare_c2small_all.err
echo "build_compare all rc=$?"
python tools/kernel_only.py c5mid > gpurun_out/r2_c5mid_kernel.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_traverse' -s 3 -c 1 -f -o gpurun_out/r2_c5mid python tools/kernel_only.py c5mid > gpurun_out/r2_c5mid_ncu.log 2>&1
echo "ncu c5mid rc=$?"
python bench.py --workload c4_all_s100 --steps 5 --cpu-seconds 0 --check-reads 20000 > gpurun_out/r2_bench_c4_all_s100.json 2> gpurun_out/r2_bench_c4_all_s100.err
echo "c4_all_s100 rc=$?"
timeout 1200 python bench.py --workload c3 --steps 3 --cpu-seconds 5 --check-reads 3000 --verbose > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
echo "c3 rc=$?"; tail -5 gpurun_out/r2_bench_c3_n1.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
