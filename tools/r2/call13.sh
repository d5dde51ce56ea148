#!/bin/bash
# round-2 GPU call 13 (1 GPU): straight-line pair decode
mkdir -p gpurun_out/r2 gpurun_out/r2/ncu
PT="python -m pytest -q -m gpu --timeout=300"
timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_resetup.py tests/test_gpu_classical.py tests/test_golden_round2.py -x 2>&1 | tail -2 | cut -c1-300
OUT=gpurun_out/r2/sweep_pairs2.jsonl; : > $OUT
timeout 400 python tools/r2/sweep_kernel.py 256 solve 2>/dev/null | tee -a $OUT
timeout 500 python tools/r2/sweep_kernel.py 512 solve 2>/dev/null | tee -a $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_tile_enc_kernel" -s 26 -c 2 -f -o gpurun_out/r2/ncu/enc_jacobi_256_pairs python tools/r2/sweep_kernel.py 256 > gpurun_out/r2/ncu/enc_jacobi_256_pairs.log 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2/bench_default4.json 2>/dev/null; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2/bench_default4.json") if l.startswith('{')][-1]); r=d['roofline']
print("default: it/s", round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'jacobi ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'spmv', round(r['spmv']['ms_per_launch'],4), 'iter frac', round(r['iteration']['frac'],3), 'refgpu', (d.get('reference_gpu') or {}).get('value'))
PY
