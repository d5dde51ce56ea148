#!/bin/bash
# round-2 GPU call 11 (1 GPU): pair dictionary of the coded kernels; DILU tile kernel with 4 blocks of gathers in flight
mkdir -p gpurun_out/r2
PT="python -m pytest -q -m gpu --timeout=300"
echo "== parity with pair tables (default) and without"
timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_resetup.py tests/test_gpu_classical.py tests/test_gpu_block_dilu.py tests/test_golden_round2.py -x 2>&1 | tail -3 | cut -c1-300
echo "== coded kernel sweep: pairs on / off"
OUT=gpurun_out/r2/sweep_pairs.jsonl; : > $OUT
for P in 1 0; do
  AMGXB_ENC_PAIRS=$P AMGXB_COLENC_VERBOSE=1 timeout 400 python tools/r2/sweep_kernel.py 256 solve 2>gpurun_out/r2/pairs_$P.err | tee -a $OUT
done
AMGXB_ENC_PAIRS=1 timeout 500 python tools/r2/sweep_kernel.py 512 solve 2>/dev/null | tee -a $OUT
head -4 gpurun_out/r2/pairs_1.err
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith('{')][-1]); r=d['roofline']
    print("  $2: it/s", round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'iters', d['config']['iterations_per_step'], d['config']['solve_status'], 'launches/it', round(d['gpu_launches']/d['steps']/d['config']['iterations_per_step']), 'final', d.get('final_relative_residual'), 'kernel frac', round(r['frac'],3))
except Exception as e: print('  $2 failed', e)
PY
}
echo "== block: DILU tile kernels (short rows only, 4 blocks in flight)"
for T in 1 0; do
AMGXB_DILU_TILES=$T timeout 900 python bench.py --workload block --steps 3 --warmup 2 > gpurun_out/r2/bench_block_e$T.json 2> gpurun_out/r2/bench_block_e$T.err; show gpurun_out/r2/bench_block_e$T.json "dilu tiles=$T dDFI"
AMGXB_DILU_TILES=$T timeout 900 python bench.py --workload block --mode dDDI --steps 3 --warmup 2 > gpurun_out/r2/bench_block_e${T}_dDDI.json 2>/dev/null; show gpurun_out/r2/bench_block_e${T}_dDDI.json "dilu tiles=$T dDDI"
done
AMGXB_PHASE_TIMING=1 timeout 600 python bench.py --workload block --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/r2/phase_block160_tiles2.txt > /dev/null; grep -A5 "phase timing" gpurun_out/r2/phase_block160_tiles2.txt | tail -6
echo "== default bench"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2/bench_default3.json 2>/dev/null; show gpurun_out/r2/bench_default3.json default
