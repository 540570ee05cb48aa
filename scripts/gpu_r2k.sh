#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/sanitize_decode.py --layers 2 --steps 300 --n-past 0 --ctx 320 > $OUT/native_lib.txt 2>&1
echo "lib: steps done $(grep -c '^step' $OUT/native_lib.txt); $(tail -1 $OUT/native_lib.txt | cut -c1-150)"
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -x -q > $OUT/pytest_r2k.txt 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r2k.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 2 > $OUT/tune_r2k.txt 2>&1; tail -3 $OUT/tune_r2k.txt
THK_LIBDIR=lib_s4 timeout 300 python scripts/tune.py --steps 100 --repeat 2 > $OUT/tune_r2k_s4.txt 2>&1; tail -3 $OUT/tune_r2k_s4.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 1 --set nosync=1 > $OUT/tune_r2k_nosync.txt 2>&1; tail -2 $OUT/tune_r2k_nosync.txt
THK_LIBDIR=lib_prof timeout 300 python scripts/tile_timeline.py --kinds 0,3 > $OUT/tiles_r2k.txt 2>&1; head -16 $OUT/tiles_r2k.txt
THK_LIBDIR=lib_prof timeout 300 python scripts/tune.py --steps 50 --repeat 1 --profile > $OUT/prof_r2k.txt 2>&1; grep -E "WAITS|BEST" $OUT/prof_r2k.txt
