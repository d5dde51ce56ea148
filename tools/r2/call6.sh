#!/bin/bash
# round-2 GPU call 6 (1 GPU): suite with the new defaults, bench lines, workloads, ncu launch lists and --set full captures
mkdir -p gpurun_out/r2 gpurun_out/r2/ncu
PT="python -m pytest -q -m gpu --timeout=300"
echo "== full suite, defaults (coded streams on, occupancy plan, fused DILU levels)"
timeout 900 $PT tests -rf > gpurun_out/r2/suite3.log 2>&1; tail -6 gpurun_out/r2/suite3.log | cut -c1-300
echo "== classical parity with the coarse sub-cycle replayed as a graph"
AMGXB_GRAPH_COARSE=1 timeout 600 $PT tests/test_gpu_classical.py tests/test_gpu_cycles.py -x 2>&1 | tail -2
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
echo "== bench default (new format)"
timeout 900 python bench.py > gpurun_out/r2/bench_default.json 2> gpurun_out/r2/bench_default.err; tail -c 2500 gpurun_out/r2/bench_default.json
echo "== bench 512"
timeout 900 python bench.py --grid 512 --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/bench_512.json 2> gpurun_out/r2/bench_512.err; tail -c 1200 gpurun_out/r2/bench_512.json
echo "== bench block (dDFI, 160^3) fused DILU levels on / off"
for F in 32768 0; do
AMGXB_DILU_FUSED_ROWS=$F timeout 900 python bench.py --workload block --steps 3 --warmup 2 > gpurun_out/r2/bench_block_$F.json 2> gpurun_out/r2/bench_block_$F.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench_block_$F.json") if l.startswith('{')][-1]); print("  fused_rows=$F: it/s", round(d['value'],1), 'iters', d['config']['iterations_per_step'], d['config']['solve_status'], 'launches/it', round(d['gpu_launches']/3/d['config']['iterations_per_step']), 'spmv frac', round(d['roofline']['frac'],3))
except Exception as e: print('  block failed', e)
PY
done
echo "== bench banded (4M rows)"
timeout 900 python bench.py --workload banded --steps 3 --warmup 2 > gpurun_out/r2/bench_banded.json 2> gpurun_out/r2/bench_banded.err; tail -c 1500 gpurun_out/r2/bench_banded.json
echo "== ncu launch lists"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/ncu/launches_solve_256.csv python tools/profile_solve.py 256 3 > gpurun_out/r2/ncu/prof_solve.log 2>&1; tail -3 gpurun_out/r2/ncu/prof_solve.log | cut -c1-200
python tools/r2/summarize_launches.py gpurun_out/r2/ncu/launches_solve_256.csv "r02 launch list, PCG+AMG solve (3 iterations), 7-pt Poisson 256^3, defaults" > gpurun_out/r2/ncu/launches_solve_256.md; head -20 gpurun_out/r2/ncu/launches_solve_256.md
echo "== ncu --set full: tile kernels (coded and plain), level-1, transfer"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_tile_enc_kernel" -s 3 -c 4 -f -o gpurun_out/r2/ncu/enc_256 python tools/r2/sweep_kernel.py 256 > gpurun_out/r2/ncu/enc_256.log 2>&1
AMGXB_COLENC=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_tile_kernel" -s 3 -c 4 -f -o gpurun_out/r2/ncu/plain_256 python tools/r2/sweep_kernel.py 256 > gpurun_out/r2/ncu/plain_256.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"reduce_kernel|restrict_kernel|prolong_set_kernel|map_kernel" -c 8 -f -o gpurun_out/r2/ncu/level1_256 python tools/profile_solve.py 256 2 > gpurun_out/r2/ncu/level1_256.log 2>&1
echo "== ncu block / DILU"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"block4_kernel|dilu_sweep_4x4|dilu_level_kernel" -s 40 -c 8 -f -o gpurun_out/r2/ncu/block_96 python bench.py --workload block --grid 96 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2/ncu/block_96.log 2>&1
ls -la gpurun_out/r2/ncu/
