#!/bin/bash
# round-2 GPU call 20 (1 GPU): producer-side L2 prefetch of the per-row vectors and of the x rows ahead (UBLKPF.L2): off / on
mkdir -p gpurun_out/r2/l2pf gpurun_out/r2/ncu
F=gpurun_out/r2/l2pf
PT="python -m pytest -q -m gpu --timeout=300"
timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_resetup.py tests/test_golden_round2.py -x 2>&1 | tail -2 | cut -c1-300
for PF in 0 1; do
  AMGXB_L2_PREFETCH=$PF timeout 600 python tools/r2/sweep_kernel.py 256 solve 2>/dev/null | tail -1 | tee $F/sweep256_pf$PF.json | cut -c1-400
  AMGXB_L2_PREFETCH=$PF AMGXB_COLENC=0 timeout 600 python tools/r2/sweep_kernel.py 256 solve 2>/dev/null | tail -1 | tee $F/sweep256_plain_pf$PF.json | cut -c1-400
done
for PF in 0 1; do
  AMGXB_L2_PREFETCH=$PF timeout 600 python tools/r2/sweep_kernel.py 512 2>/dev/null | tail -1 | tee $F/sweep512_pf$PF.json | cut -c1-400
done
echo "== banded"
for PF in 0 1; do
  AMGXB_L2_PREFETCH=$PF timeout 600 python bench.py --workload banded --steps 3 --warmup 2 --no-cpu-baseline --no-reference-gpu > $F/banded_pf$PF.json 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open("$F/banded_pf$PF.json") if l.startswith('{')][-1]); r=d.get('roofline') or {}
print('banded pf=$PF it/s', round(d['value'],1), 'kernel ms', r.get('ms_per_launch'), 'frac', r.get('frac'), 'spmv', r.get('spmv'))
PY
done
echo "== ncu banded SpMV"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_tile" -s 3 -c 2 -f -o gpurun_out/r2/ncu/banded_spmv python bench.py --workload banded --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/ncu/banded_spmv.log 2>&1
ls -la gpurun_out/r2/ncu | grep banded
