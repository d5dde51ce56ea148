#!/bin/bash
mkdir -p gpurun_out
echo "== default bench.py (as the driver runs it)"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-600 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-700
echo "== configs"; timeout 1500 python tools/bench_configs.py all 2>&1 | tail -8
