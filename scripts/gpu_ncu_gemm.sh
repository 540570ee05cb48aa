#!/bin/bash
# ncu --set full of one prefill GEMM launch (wq of the second layer, second pass: M=128, N=4096, K=4096, split-K 4)
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
LAYERS=2 timeout 250 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_tc -s 21 -c 1 -f -o $OUT/prof_gemm_v7 \
    python scripts/prefill_once.py > $OUT/ncu_gemm_v7.log 2>&1
ls -la $OUT/prof_gemm_v7.ncu-rep; tail -2 $OUT/ncu_gemm_v7.log
