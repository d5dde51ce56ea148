#!/bin/bash
for T in 0 20000; do
echo "-- TAIL=$T"
AMGXB_TAIL_CHECK=1 AMGXB_TAIL_ROWS=$T timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29744 tools/dbg_tail.py 48 2>&1 | grep "tail-check\|^rank" | cut -c1-330
done
