#!/bin/bash
# round 2, GPU call 3: compact path after the lag fix; row-load flavour x CTAs/SM x L1 size in input and true-origin order
set -x
mkdir -p gpurun_out
COLBWT_TRACE=1 python bench.py --steps 5 --cpu-seconds 0 --check-reads 2000 > gpurun_out/r2_bench_c2_lag.json 2> gpurun_out/r2_bench_c2_lag.err
echo "bench rc=$?"
python tools/wall_probe.py base > gpurun_out/r2_wall_probe.log 2>gpurun_out/r2_wall_probe.err
for v in l1c4 l2c4 l0c5 l1c5 l0c6 l1c6 l2c6; do
  COLBWT_LIB=$PWD/col_bwt_b200/_sb$v/libcolbwt_b200.so python tools/wall_probe.py $v >> gpurun_out/r2_wall_probe.log 2>>gpurun_out/r2_wall_probe.err
done
cat gpurun_out/r2_wall_probe.log
# counters for the baseline and one more-lanes variant, origin order only
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_requests_srcunit_tex.sum,lts__t_tag_requests.sum,l1tex__m_xbar2l1tex_read_sectors.sum,smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warps_issue_stalled_lg_throttle_per_warp_active.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for v in base l1c6; do
  LIB=""; [ $v != base ] && LIB=$PWD/col_bwt_b200/_sb$v/libcolbwt_b200.so
  COLBWT_LIB=$LIB ncu --metrics $M --clock-control none -k regex:k_traverse -s 2 -c 1 --csv --log-file gpurun_out/r2_wall_ncu_${v}_origin.csv python tools/wall_probe.py $v origin > /dev/null 2>&1
  COLBWT_LIB=$LIB ncu --metrics $M --clock-control none -k regex:k_traverse -s 2 -c 1 --csv --log-file gpurun_out/r2_wall_ncu_${v}_input.csv python tools/wall_probe.py $v input > /dev/null 2>&1
done
ls -la gpurun_out
