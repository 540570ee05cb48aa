#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in lib lib_nosmr lib_nopeek lib_fsw; do
  THK_LIBDIR=$v timeout 120 python scripts/sanitize_decode.py --tiny --steps 40 > $OUT/tiny_$v.txt 2>&1
  echo "$v tiny: steps $(grep -c '^step' $OUT/tiny_$v.txt); $(tail -1 $OUT/tiny_$v.txt | cut -c1-160)"
  THK_LIBDIR=$v timeout 120 python scripts/sanitize_decode.py --layers 2 --steps 100 --n-past 0 --ctx 160 > $OUT/native_$v.txt 2>&1
  echo "$v 7b2: steps $(grep -c '^step' $OUT/native_$v.txt); $(tail -1 $OUT/native_$v.txt | cut -c1-160)"
done
