#!/bin/bash
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -x -q > $OUT/pytest_r2t.txt 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_r2t.txt
timeout 200 python scripts/prefill_once.py 2>&1 | tail -3
