#!/bin/bash
# round 2, GPU call 6: long-read geometry and 5 CTAs/SM on configs[2]; more lanes on the DRAM-resident synthetic table; col_split timing trace;
# configs[4] at full size on one GPU; configs[2] at n = 2.05e9
set -x
mkdir -p gpurun_out
python tools/longread_sweep.py default c3 > gpurun_out/r2_longread_sweep.log 2> gpurun_out/r2_longread_sweep.err
COLBWT_LIB=$PWD/col_bwt_b200/_sbc5/libcolbwt_b200.so python tools/longread_sweep.py ctas5 c3 >> gpurun_out/r2_longread_sweep.log 2>> gpurun_out/r2_longread_sweep.err
cat gpurun_out/r2_longread_sweep.log
for v in _sbc5 _sbl0c6; do
  COLBWT_LIB=$PWD/col_bwt_b200/$v/libcolbwt_b200.so python tools/kernel_only.py c5mid >> gpurun_out/r2_c5mid_ctas.log 2>&1
done
grep workload gpurun_out/r2_c5mid_ctas.log
COLBWT_TRACE=1 python tools/build_compare.py --H 32 --G 1000000 --mode all --rate 2 > gpurun_out/r2_build_compare_c2small_all2.json 2> gpurun_out/r2_build_compare_c2small_all2.err
grep colbwt_col_split gpurun_out/r2_build_compare_c2small_all2.err
COLBWT_TRACE=1 python tools/build_compare.py --H 32 --G 10000000 --mode all --rate 10 > gpurun_out/r2_build_compare_c2_all.json 2> gpurun_out/r2_build_compare_c2_all.err
grep colbwt_col_split gpurun_out/r2_build_compare_c2_all.err; cat gpurun_out/r2_build_compare_c2_all.json
timeout 900 python bench.py --workload c5 --steps 3 --cpu-seconds 5 --check-reads 100000 --verbose > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err
echo "c5 rc=$?"; tail -3 gpurun_out/r2_bench_c5_n1.err; free -g | head -2
timeout 900 python bench.py --workload c3big --steps 3 --cpu-seconds 0 --check-reads 2000 --verbose > gpurun_out/r2_bench_c3big_n1.json 2> gpurun_out/r2_bench_c3big_n1.err
echo "c3big rc=$?"; tail -3 gpurun_out/r2_bench_c3big_n1.err
