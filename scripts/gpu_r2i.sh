#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_decode.py --layers 2 --steps 60 --n-past 0 --ctx 160 > $OUT/sanitize_synccheck_7b2.txt 2>&1; grep -v "^step" $OUT/sanitize_synccheck_7b2.txt | head -60 | cut -c1-250; grep -c "^step" $OUT/sanitize_synccheck_7b2.txt
