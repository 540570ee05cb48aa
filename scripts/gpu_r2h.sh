#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in lib_fsw lib_s4 lib_fsw4; do
  THK_LIBDIR=$v CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/sanitize_decode.py --layers 2 --steps 300 --n-past 0 --ctx 320 > $OUT/native_$v.txt 2>&1
  echo "$v: steps done $(grep -c '^step' $OUT/native_$v.txt); $(tail -1 $OUT/native_$v.txt | cut -c1-150)"
done
