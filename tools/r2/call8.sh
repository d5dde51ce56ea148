#!/bin/bash
# round-2 GPU call 8 (1 GPU): block tile kernel, DILU prefetch + fused small levels, length-sorted rows on the banded matrix, ncu of the Jacobi sweep
mkdir -p gpurun_out/r2 gpurun_out/r2/ncu
PT="python -m pytest -q -m gpu --timeout=300"
echo "== full suite"
timeout 900 $PT tests -rf > gpurun_out/r2/suite4.log 2>&1; tail -6 gpurun_out/r2/suite4.log | cut -c1-300
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith('{')][-1]); r=d['roofline']
    print("  $2: it/s", round(d['value'],1), 'iters', d['config']['iterations_per_step'], d['config']['solve_status'], 'launches/it', round(d['gpu_launches']/d['steps']/d['config']['iterations_per_step']), 'kernel', r['kernel'][:40], 'ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'spmv', (round(r['spmv']['ms_per_launch'],4), round(r['spmv']['frac'],3)) if 'spmv' in r else '')
except Exception as e: print('  $2 failed', e)
PY
}
echo "== block (dDFI 160^3): tile kernel x fused DILU levels"
for T in 1 0; do for F in 1 0; do
AMGXB_BLOCK_TILES=$T AMGXB_DILU_FUSED=$F timeout 900 python bench.py --workload block --steps 3 --warmup 2 > gpurun_out/r2/bench_block_t${T}_f${F}.json 2> gpurun_out/r2/bench_block_t${T}_f${F}.err; show gpurun_out/r2/bench_block_t${T}_f${F}.json "tiles=$T fused=$F"
done; done
timeout 900 python bench.py --workload block --mode dDDI --steps 3 --warmup 2 > gpurun_out/r2/bench_block_dDDI.json 2> gpurun_out/r2/bench_block_dDDI.err; show gpurun_out/r2/bench_block_dDDI.json "dDDI"
echo "== banded 4M: length-sorted rows on / off, coded streams on / off"
for V in "1 3" "0 3" "1 0" "0 0"; do set -- $V
AMGXB_TILE_PERM=$1 AMGXB_COLENC=$2 timeout 900 python bench.py --workload banded --steps 3 --warmup 2 > gpurun_out/r2/bench_banded_p$1_e$2.json 2> gpurun_out/r2/bench_banded_p$1_e$2.err; show gpurun_out/r2/bench_banded_p$1_e$2.json "perm=$1 colenc=$2"
done
echo "== reference GPU on the banded matrix (FGMRES + aggregation)"
timeout 600 python tools/r2/ref_banded.py 2>&1 | tail -3
echo "== ncu --set full: fused Jacobi sweep (coded), block tile kernel, DILU sweep"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_tile_enc_kernel" -s 26 -c 3 -f -o gpurun_out/r2/ncu/enc_jacobi_256 python tools/r2/sweep_kernel.py 256 > gpurun_out/r2/ncu/enc_jacobi_256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"block4_tile_kernel|dilu_sweep_4x4|dilu_level_kernel" -s 60 -c 8 -f -o gpurun_out/r2/ncu/block_96b python bench.py --workload block --grid 96 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2/ncu/block_96b.log 2>&1
echo "== default bench line"
timeout 900 python bench.py > gpurun_out/r2/bench_default2.json 2> gpurun_out/r2/bench_default2.err; show gpurun_out/r2/bench_default2.json default
