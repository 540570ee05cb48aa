#!/bin/bash
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kv_f16.py tests/test_gpu_graph_trace.py tests/test_gpu_decode.py -x -q > $OUT/pytest_r2t.txt 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_r2t.txt
