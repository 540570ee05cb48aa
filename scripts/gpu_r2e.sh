#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q > $OUT/pytest_r2e.txt 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_r2e.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 2 > $OUT/tune_r2e.txt 2>&1; tail -3 $OUT/tune_r2e.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 1 --set nosync=1 > $OUT/tune_r2e_nosync.txt 2>&1; tail -2 $OUT/tune_r2e_nosync.txt
THK_LIBDIR=lib_prof timeout 300 python scripts/tile_timeline.py --kinds 0,3 > $OUT/tiles_r2e.txt 2>&1; head -32 $OUT/tiles_r2e.txt
THK_LIBDIR=lib_prof timeout 300 python scripts/tune.py --steps 50 --repeat 1 --profile > $OUT/prof_r2e.txt 2>&1; grep -E "WAITS|BEST" $OUT/prof_r2e.txt
