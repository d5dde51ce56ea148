#!/bin/bash
# round-2 GPU call 29 (1 GPU): final suite, final bench lines of every workload, refreshed ncu captures (row patterns, window kernel)
mkdir -p gpurun_out/r2/final gpurun_out/r2/ncu
PT="python -m pytest -q -m gpu --timeout=300"
echo "== full suite"
timeout 900 $PT tests -rf > gpurun_out/r2/final/suite.log 2>&1; tail -4 gpurun_out/r2/final/suite.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
F=gpurun_out/r2/final
timeout 900 python bench.py > $F/bench_poisson256.json 2> $F/bench_poisson256.err
timeout 900 python bench.py --impl reference --steps 2 > $F/bench_poisson256_reference_arm.json 2>/dev/null
timeout 900 python bench.py --grid 512 --steps 2 --warmup 3 --no-cpu-baseline > $F/bench_poisson512.json 2>/dev/null
timeout 900 python bench.py --workload block --steps 3 --warmup 2 > $F/bench_block_dDFI.json 2>/dev/null
timeout 900 python bench.py --workload banded --steps 3 --warmup 2 > $F/bench_banded.json 2>/dev/null
timeout 900 python tools/bench_classical.py 256 512 2>&1 | grep '^{' > $F/classical.jsonl
for f in $F/bench_*.json; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$f") if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print("$f".split('/')[-1], '| it/s', round(d['value'],1), '| e2e', round(d['e2e']['value'],1), '| iters', d.get('config',{}).get('iterations_per_step'), d.get('config',{}).get('solve_status'), '| kernel ms', r.get('ms_per_launch'), 'frac', r.get('frac'), '| spmv', (r.get('spmv') or {}).get('ms_per_launch'), (r.get('spmv') or {}).get('frac'), '| refgpu', (d.get('reference_gpu') or {}).get('value'), '| cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print("$f", 'failed', e)
PY
done
cut -c1-260 $F/classical.jsonl
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_tile_enc_kernel" -s 3 -c 2 -f -o gpurun_out/r2/ncu/enc_spmv_256_final python tools/r2/sweep_kernel.py 256 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_tile_enc_kernel" -s 26 -c 2 -f -o gpurun_out/r2/ncu/enc_jacobi_256_final python tools/r2/sweep_kernel.py 256 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_window_kernel" -s 2 -c 3 -f -o gpurun_out/r2/ncu/window_banded_final python tools/r2/sweep_banded.py > gpurun_out/r2/ncu/window_banded_final.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/ncu/launches_solve_256_final.csv python tools/profile_solve.py 256 3 > /dev/null 2>&1
python tools/r2/summarize_launches.py gpurun_out/r2/ncu/launches_solve_256_final.csv "r02 launch list, PCG+AMG solve (3 iterations), 7-pt Poisson 256^3, final defaults" > gpurun_out/r2/ncu/launches_solve_256_final.md; head -14 gpurun_out/r2/ncu/launches_solve_256_final.md
ls -la gpurun_out/r2/ncu | tail -6
