#!/bin/bash
# round-2 GPU call 19 (1 GPU): row-pattern coding (one byte per row on stencil tiles): parity suites, then kernel and solve times with / without
mkdir -p gpurun_out/r2/rowpat
F=gpurun_out/r2/rowpat
PT="python -m pytest -q -m gpu --timeout=300"
echo "== parity suites with row patterns (default)"
timeout 900 $PT tests/test_gpu_parity.py tests/test_gpu_resetup.py tests/test_gpu_classical.py tests/test_golden_round2.py tests/test_gpu_smoothers.py tests/test_gpu_krylov.py -x 2>&1 | tail -4 | cut -c1-300
echo "== kernels + solve, 256^3"
for RPAT in 0 1; do
  AMGXB_ENC_ROWPAT=$RPAT AMGXB_COLENC_VERBOSE=1 timeout 600 python tools/r2/sweep_kernel.py 256 solve > $F/sweep256_rowpat$RPAT.json 2> $F/sweep256_rowpat$RPAT.err
  grep -m3 "row patterns" $F/sweep256_rowpat$RPAT.err | cut -c1-300
  tail -1 $F/sweep256_rowpat$RPAT.json | cut -c1-500
done
echo "== 512^3"
for RPAT in 0 1; do
  AMGXB_ENC_ROWPAT=$RPAT timeout 600 python tools/r2/sweep_kernel.py 512 solve 2>/dev/null | tail -1 | tee $F/sweep512_rowpat$RPAT.json | cut -c1-500
done
echo "== bench"
timeout 900 python bench.py --no-cpu-baseline > $F/bench_poisson256.json 2> $F/bench_poisson256.err
python - <<PY
import json
d=json.loads([l for l in open("$F/bench_poisson256.json") if l.startswith('{')][-1]); r=d.get('roofline') or {}
print('it/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'iters', d['config'].get('iterations_per_step'), d['config'].get('solve_status'), 'kernel ms', r.get('ms_per_launch'), 'frac', r.get('frac'), 'refgpu', (d.get('reference_gpu') or {}).get('value'))
PY
