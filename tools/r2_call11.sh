#!/bin/bash
# round 2, GPU call 11: the AVX-512 expander on the box (GPU tests, C2 / configs[2] / configs[4]-quarter with every transport traced)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest11.log
COLBWT_TRACE=1 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --check-reads 20000 > gpurun_out/r2_bench_c2_modes3.json 2> gpurun_out/r2_bench_c2_modes3.err
echo "c2 rc=$?"; grep "colbwt_query\]" gpurun_out/r2_bench_c2_modes3.err | head -28
COLBWT_TRACE=1 timeout 600 python bench.py --workload c3 --steps 3 --cpu-seconds 0 --check-reads 3000 > gpurun_out/r2_bench_c3_n1d.json 2> gpurun_out/r2_bench_c3_n1d.err
echo "c3 rc=$?"; grep "colbwt_query\]" gpurun_out/r2_bench_c3_n1d.err | head -20
COLBWT_TRACE=1 timeout 600 python bench.py --workload c5mid --steps 3 --cpu-seconds 0 --check-reads 60000 > gpurun_out/r2_bench_c5mid3.json 2> gpurun_out/r2_bench_c5mid3.err
echo "c5mid rc=$?"; grep "colbwt_query\]" gpurun_out/r2_bench_c5mid3.err | head -20
for f in c2_modes3 c3_n1d c5mid3; do python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_$f.json').read().strip().splitlines()[-1])
print('$f', 'kernel %.2f'%(d['value']/1e9), 'e2e %.2f'%(d['e2e']['value']/1e9), d['e2e']['transport'], d['e2e']['packing'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], 'compact %.2f'%(d['e2e_compact']['value']/1e9), d['parity_vs_oracle'])
PY
done
