#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2o.txt 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r2o.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 2 --set l2_ahead_kb=0,64,128 > $OUT/tune_r2o.txt 2>&1; tail -7 $OUT/tune_r2o.txt
THK_LIBDIR=lib_s4 timeout 300 python scripts/tune.py --steps 100 --repeat 2 --set l2_ahead_kb=0,64,128 > $OUT/tune_r2o_s4.txt 2>&1; tail -7 $OUT/tune_r2o_s4.txt
