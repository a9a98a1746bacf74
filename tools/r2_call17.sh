#!/bin/bash
# round 2, GPU call 17: pool spin on/off on one GPU (C2, compact transport pinned), and the default bench line of the final code
set -x
mkdir -p gpurun_out
for spin in 0 2000; do
  COLBWT_POOL_SPIN=$spin COLBWT_TRACE=1 COLBWT_DEVICE_PACK=0 COLBWT_COMPACT_D2H=1 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --check-reads 2000 > gpurun_out/r2_spin$spin.json 2> gpurun_out/r2_spin$spin.err
  echo "spin=$spin rc=$?"; grep "colbwt_query\] device\|colbwt_query\] 15" gpurun_out/r2_spin$spin.err | grep -B1 "dense, compact" | tail -4
done
python bench.py --steps 10 --warmup 5 > gpurun_out/r2_bench_c2_final.json 2> gpurun_out/r2_bench_c2_final.err
echo "final rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_c2_final.json').read().strip().splitlines()[-1])
print('kernel %.2f'%(d['value']/1e9), 'e2e %.2f'%(d['e2e']['value']/1e9), d['e2e']['transport'], d['e2e']['packing'], 'compact %.2f'%(d['e2e_compact']['value']/1e9), d['parity_vs_oracle'], d['cpu_baseline']['value'])
PY
