#!/bin/bash
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -3
echo "== main (chunks 6)"
timeout 200 python scripts/tune.py --steps 100 --repeat 2 --set poll_single=0,1,2 2>&1 | tail -7
echo "== c3"
THK_LIBDIR=lib_c3 timeout 200 python scripts/tune.py --steps 100 --repeat 2 --set poll_single=1,2 2>&1 | tail -5
