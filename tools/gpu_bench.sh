#!/bin/bash
mkdir -p gpurun_out
echo "== pytest durations"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -25
echo "== bench 256"; timeout 900 python bench.py --steps 3 --warmup 2 2>gpurun_out/bench_err.log | tee gpurun_out/bench_256.json | cut -c1-1500; tail -5 gpurun_out/bench_err.log
echo "== ncu launch list (128^3 solve)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 700 --csv --log-file gpurun_out/launches_128.csv python bench.py --steps 1 --warmup 1 --grid 128 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full (fine-level kernels 256^3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_tile -s 6 -c 3 -o gpurun_out/prof_csr_256 python tools/quick_bench.py 256 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out | head -30
