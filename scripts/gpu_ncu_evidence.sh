#!/bin/bash
# evidence run: ncu --set full of one decode_kernel launch, launch lists of the decode bench and of one 128-token prefill
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 1 -f -o $OUT/prof_decode_v7 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-parity > $OUT/ncu_full_v7.log 2>&1
ls -la $OUT/prof_decode_v7.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_v7.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-parity > $OUT/ncu_launches_v7.log 2>&1
wc -l $OUT/launches_v7.csv
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches_prefill_v7.csv \
    python scripts/prefill_once.py > $OUT/ncu_prefill_v7.log 2>&1
wc -l $OUT/launches_prefill_v7.csv; tail -3 $OUT/ncu_prefill_v7.log
timeout 200 python scripts/prefill_once.py 2>&1 | tail -3
