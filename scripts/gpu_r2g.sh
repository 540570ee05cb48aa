#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
CUDA_LAUNCH_BLOCKING=1 timeout 300 python scripts/sanitize_decode.py --layers 2 --steps 140 --n-past 0 --ctx 160 > $OUT/native_7b2_steps.txt 2>&1; tail -4 $OUT/native_7b2_steps.txt | cut -c1-300
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_decode.py --layers 2 --steps 140 --n-past 0 --ctx 160 > $OUT/sanitize_memcheck_7b2_long.txt 2>&1; grep -v "^step" $OUT/sanitize_memcheck_7b2_long.txt | head -60 | cut -c1-220
