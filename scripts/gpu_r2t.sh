#!/bin/bash
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > $OUT/pytest_r2t.txt 2>&1; echo "pytest gemm rc=$?"; tail -12 $OUT/pytest_r2t.txt
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -k "prefill or batch" > $OUT/pytest_r2t2.txt 2>&1; echo "pytest prefill rc=$?"; tail -6 $OUT/pytest_r2t2.txt
timeout 200 python scripts/prefill_once.py 2>&1 | tail -3
