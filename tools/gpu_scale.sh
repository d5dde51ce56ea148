#!/bin/bash
# weak-scaling sweep of bench.py at 256^3 rows per GPU
mkdir -p gpurun_out
N=$1
for g in $N; do
  if [ "$g" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/scale_$g.json 2> gpurun_out/scale_$g.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$g --master-addr 127.0.0.1 --master-port 2970$g bench.py --gpus $g --steps 3 --warmup 2 > gpurun_out/scale_$g.log 2>&1
    grep '"metric"' gpurun_out/scale_$g.log > gpurun_out/scale_$g.json
  fi
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_$g.json').read().strip().splitlines()[-1])
    print('N=$g', 'value', round(d['value'],2), 'global its/s', round(d['config']['global_iterations_per_sec'],2), 'its/step', d['config']['iterations_per_step'], 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'setup', round(d['config']['setup_seconds'],2))
except Exception as e:
    print('N=$g FAILED', e)
    import subprocess; print(subprocess.run('tail -5 gpurun_out/scale_$g.log gpurun_out/scale_$g.err 2>/dev/null', shell=True, capture_output=True, text=True).stdout[-1500:])
PY
done
