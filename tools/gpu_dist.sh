#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench x1 256 graphs on";  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['iterations_per_step'], d['gpu_launches'], d['e2e']['value'])"
echo "== bench x1 256 graphs off"; AMGXB_GRAPHS=0 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['iterations_per_step'], d['gpu_launches'])"
echo "== bench x1 128 graphs on";  timeout 600 python bench.py --steps 3 --warmup 2 --grid 128 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['iterations_per_step'], d['gpu_launches'])"
echo "== dist worker x2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29701 tests/dist_gpu_worker.py > gpurun_out/dist_worker.log 2>&1
grep -n "AMGX error\|NCCL WARN\|DIST_GPU_OK\|assert\|Error" gpurun_out/dist_worker.log | head -10 | cut -c1-300
for g in 256 128; do
echo "== bench x2 (n=$g)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 3 --warmup 2 --grid $g > gpurun_out/bench2_$g.log 2>&1
grep "metric" gpurun_out/bench2_$g.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['iterations_per_step'], d['gpu_launches'])" || tail -5 gpurun_out/bench2_$g.log
done
