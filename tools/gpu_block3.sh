#!/bin/bash
mkdir -p gpurun_out
for M in dDFI dDDI; do timeout 900 python tools/bench_block_ref.py $M 128 128 80 PCG 2>&1 | tail -1; done
timeout 900 python tools/bench_block_ref.py dDFI 128 128 80 AMG 2>&1 | tail -1
