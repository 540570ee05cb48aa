#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
THK_LIBDIR=lib_dbg timeout 120 python scripts/sanitize_decode.py --layers 1 --steps 2 > $OUT/dbg_flush.txt 2>&1; grep -E "DEBUG|abort|step" $OUT/dbg_flush.txt | cut -c1-250
THK_LIBDIR=lib_dbg timeout 120 python scripts/sanitize_decode.py --tiny --steps 10 > $OUT/dbg_flush_tiny.txt 2>&1; grep -E "DEBUG|abort|step" $OUT/dbg_flush_tiny.txt | cut -c1-250 | tail -5
