#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -x -q > $OUT/pytest_r2l.txt 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r2l.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 2 > $OUT/tune_r2l.txt 2>&1; tail -3 $OUT/tune_r2l.txt
THK_LIBDIR=lib_s4 timeout 300 python scripts/tune.py --steps 100 --repeat 2 > $OUT/tune_r2l_s4.txt 2>&1; tail -3 $OUT/tune_r2l_s4.txt
timeout 300 python scripts/tune.py --steps 100 --repeat 1 --set nosync=1 > $OUT/tune_r2l_nosync.txt 2>&1; tail -2 $OUT/tune_r2l_nosync.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 1 -f -o $OUT/prof_decode_v7a \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_v7a.log 2>&1
ls -la $OUT/prof_decode_v7a.ncu-rep
THK_LIBDIR=lib_prof timeout 300 python scripts/tile_timeline.py --kinds 0,3 > $OUT/tiles_r2l.txt 2>&1; head -30 $OUT/tiles_r2l.txt
