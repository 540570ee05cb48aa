#!/bin/bash
# launch list of the 128-token prefill after the tiled attention matmul + split-K GEMM (8 of 32 layers: the shares are per layer)
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
LAYERS=8 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_prefill_v7b.csv \
    python scripts/prefill_once.py > $OUT/ncu_prefill_v7b.log 2>&1
wc -l $OUT/launches_prefill_v7b.csv; tail -3 $OUT/ncu_prefill_v7b.log
