#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_solve_256.csv python tools/profile_solve.py 256 3 > gpurun_out/prof_solve.log 2>&1; tail -30 gpurun_out/prof_solve.log | cut -c1-200
timeout 300 python tools/profile_solve.py 256 20 2>&1 | head -3
