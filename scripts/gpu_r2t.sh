#!/bin/bash
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
THK_ATT_FLAGGED=1 timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_kv_f16.py -x -q > $OUT/pytest_r2t.txt 2>&1; echo "pytest (att_flagged=1) rc=$?"; tail -6 $OUT/pytest_r2t.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 2 --set att_flagged=0,1 2>&1 | tail -5
timeout 300 python scripts/tune.py --steps 100 --repeat 1 --ctx 2048 --set att_flagged=0,1 2>&1 | tail -3
