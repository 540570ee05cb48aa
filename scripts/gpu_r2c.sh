#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 1 -f -o $OUT/prof_decode_v5a \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_v5a.log 2>&1
tail -3 $OUT/ncu_full_v5a.log; ls -la $OUT/prof_decode_v5a.ncu-rep
