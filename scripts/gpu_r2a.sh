#!/bin/bash
# round 2, first GPU call: parity of the v5 decode kernel on one GPU, then its speed and tile timeline
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_r2a.txt
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_ops.py tests/test_gpu_cpp_api.py -m gpu -x -q > $OUT/pytest_r2a.txt 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_r2a.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 2 > $OUT/tune_r2a.txt 2>&1; tail -4 $OUT/tune_r2a.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 1 --set nosync=1 > $OUT/tune_r2a_nosync.txt 2>&1; tail -2 $OUT/tune_r2a_nosync.txt
THK_LIBDIR=lib_prof timeout 300 python scripts/tile_timeline.py --kinds 0,2,3,4 > $OUT/tiles_r2a.txt 2>&1; head -40 $OUT/tiles_r2a.txt
THK_LIBDIR=lib_prof timeout 300 python scripts/tune.py --steps 50 --repeat 1 --profile > $OUT/prof_r2a.txt 2>&1; tail -3 $OUT/prof_r2a.txt | cut -c1-1500
