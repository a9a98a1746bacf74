#!/bin/bash
# round 2, multi-GPU call: bash tools/r2_multi.sh N [parts]   (parts: any of pcie,c2,c3,inproc,c5,ref; default all but c5)
# One torchrun launch per workload gives the N-rank line and, through --sweep-out, the same job with only the first 1, 2, 4
# ranks active (bench.py), i.e. the 1/2/4/8 curve without paying for four launches.
N=${1:-2}
PARTS=${2:-pcie,c2,c3,inproc,ref}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nproc; free -g | head -2; nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r2_topo_n$N.log
if [[ $PARTS == *pcie* ]]; then
  PCIE_MB=256 $TR tools/pcie_concurrent.py > gpurun_out/r2_pcie_n$N.log 2> gpurun_out/r2_pcie_n$N.err
  cat gpurun_out/r2_pcie_n$N.log
fi
if [[ $PARTS == *c2* ]]; then
  rm -f gpurun_out/r2_sweep_c2_n$N.jsonl
  COLBWT_TRACE=1 timeout 900 $TR bench.py --gpus $N --steps 4 --warmup 3 --check-reads 20000 --sweep-out gpurun_out/r2_sweep_c2_n$N.jsonl > gpurun_out/r2_bench_c2_n$N.json 2> gpurun_out/r2_bench_c2_n$N.err
  echo "c2 N=$N rc=$?"; cut -c1-400 gpurun_out/r2_bench_c2_n$N.json
fi
if [[ $PARTS == *c3* ]]; then
  rm -f gpurun_out/r2_sweep_c3_n$N.jsonl
  timeout 900 $TR bench.py --gpus $N --workload c3 --steps 3 --warmup 3 --check-reads 2000 --sweep-out gpurun_out/r2_sweep_c3_n$N.jsonl > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err
  echo "c3 N=$N rc=$?"; cut -c1-400 gpurun_out/r2_bench_c3_n$N.json
fi
if [[ $PARTS == *inproc* ]]; then
  COLBWT_TRACE=1 timeout 600 python bench.py --inproc $N --reads 5000000 --steps 3 --cpu-seconds 0 --check-reads 20000 > gpurun_out/r2_bench_c2_inproc$N.json 2> gpurun_out/r2_bench_c2_inproc$N.err
  echo "inproc N=$N rc=$?"; cut -c1-400 gpurun_out/r2_bench_c2_inproc$N.json
fi
if [[ $PARTS == *c5* ]]; then
  rm -f gpurun_out/r2_sweep_c5_n$N.jsonl
  timeout 900 $TR bench.py --gpus $N --workload c5 --steps 3 --warmup 3 --check-reads 100000 --sweep-out gpurun_out/r2_sweep_c5_n$N.jsonl > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err
  echo "c5 N=$N rc=$?"; cut -c1-400 gpurun_out/r2_bench_c5_n$N.json
fi
if [[ $PARTS == *ref* ]]; then
  timeout 300 python bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_c2_reference_n$N.json 2> gpurun_out/r2_bench_c2_reference_n$N.err
  echo "reference rc=$?"
fi
