#!/bin/bash
# round-2 GPU call 9 (1 GPU): A/B of the length-sorted rows on the default workload, phase timing of the block and Poisson solves
mkdir -p gpurun_out/r2
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith('{')][-1]); r=d['roofline']
    print("  $2: it/s", round(d['value'],1), 'iters', d['config']['iterations_per_step'], d['config']['solve_status'], 'launches/it', round(d['gpu_launches']/d['steps']/d['config']['iterations_per_step']), 'ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'spmv', (round(r['spmv']['ms_per_launch'],4), round(r['spmv']['frac'],3)) if 'spmv' in r else '', 'iter frac', round(r.get('iteration',{}).get('frac',0),3))
except Exception as e: print('  $2 failed', e)
PY
}
for V in "-1 0" "0 0" "-1 4" "0 4" "-1 0"; do set -- $V
  PERM=""; [ "$1" != "-1" ] && PERM="AMGXB_TILE_PERM=$1"
  UNR=""; [ "$2" != "0" ] && UNR="AMGXB_TILE_UNROLL=$2"
  env $PERM $UNR timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/bench_ab_$1_$2.json 2>/dev/null; show gpurun_out/r2/bench_ab_$1_$2.json "perm=$1 unroll=$2"
done
echo "== phase timing: Poisson 256^3 (graphs off: small levels show launch latency, large levels GPU time)"
AMGXB_PHASE_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu 2> gpurun_out/r2/phase_poisson256.txt > /dev/null; grep -A60 "phase timing" gpurun_out/r2/phase_poisson256.txt | tail -64 | head -70
echo "== phase timing: block dDFI 160^3"
AMGXB_PHASE_TIMING=1 timeout 600 python bench.py --workload block --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/r2/phase_block160.txt > /dev/null; grep -A60 "phase timing" gpurun_out/r2/phase_block160.txt | tail -64 | head -70
