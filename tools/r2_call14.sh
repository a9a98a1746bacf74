#!/bin/bash
# round 2, GPU call 14: validation of the host-side changes (line-aligned expander, AVX-512 packer) before the 8-GPU call
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest14.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest14.log
COLBWT_TRACE=1 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 --check-reads 200000 > gpurun_out/r2_bench_c2_v4.json 2> gpurun_out/r2_bench_c2_v4.err
echo "c2 rc=$?"; grep "colbwt_query\]" gpurun_out/r2_bench_c2_v4.err | head -24
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_c2_v4.json').read().strip().splitlines()[-1])
print('kernel %.2f'%(d['value']/1e9), 'e2e %.2f'%(d['e2e']['value']/1e9), d['e2e']['transport'], d['e2e']['packing'], 'compact %.2f'%(d['e2e_compact']['value']/1e9), d['parity_vs_oracle'])
PY
