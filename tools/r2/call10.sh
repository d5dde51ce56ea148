#!/bin/bash
# round-2 GPU call 10 (1 GPU): colour-sorted DILU tile kernels
mkdir -p gpurun_out/r2
PT="python -m pytest -q -m gpu --timeout=300"
timeout 600 $PT tests/test_gpu_block_dilu.py tests/test_gpu_parity.py -rf 2>&1 | tail -8 | cut -c1-300
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith('{')][-1]); r=d['roofline']
    print("  $2: it/s", round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'iters', d['config']['iterations_per_step'], d['config']['solve_status'], 'launches/it', round(d['gpu_launches']/d['steps']/d['config']['iterations_per_step']), 'final', d.get('final_relative_residual'), 'spmv frac', round(r['frac'],3))
except Exception as e: print('  $2 failed', e)
PY
}
for T in 1 0; do
AMGXB_DILU_TILES=$T timeout 900 python bench.py --workload block --steps 3 --warmup 2 > gpurun_out/r2/bench_block_dt$T.json 2> gpurun_out/r2/bench_block_dt$T.err; show gpurun_out/r2/bench_block_dt$T.json "dilu tiles=$T dDFI"
done
timeout 900 python bench.py --workload block --mode dDDI --steps 3 --warmup 2 > gpurun_out/r2/bench_block_dt1_dDDI.json 2>/dev/null; show gpurun_out/r2/bench_block_dt1_dDDI.json "dilu tiles=1 dDDI"
echo "== phase timing block"
AMGXB_PHASE_TIMING=1 timeout 600 python bench.py --workload block --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/r2/phase_block160_tiles.txt > /dev/null; grep -A12 "phase timing" gpurun_out/r2/phase_block160_tiles.txt | tail -13
echo "== reference GPU build on the same block problem"
timeout 900 python tools/bench_block_ref.py dDFI 160 160 160 AMG 2>&1 | tail -4 | cut -c1-300
