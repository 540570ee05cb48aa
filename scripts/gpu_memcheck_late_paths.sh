#!/bin/bash
# memcheck of the code paths added late in round 2: f16 KV (7B width, 2 layers, 500 cached positions), flagged QKV->attention hand-over
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_decode.py --layers 2 --steps 4 --n-past 500 --ctx 512 --kv-f16 > $OUT/sanitize_memcheck_7b2_f16kv.txt 2>&1; tail -3 $OUT/sanitize_memcheck_7b2_f16kv.txt
THK_ATT_FLAGGED=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_decode.py --layers 2 --steps 4 --n-past 500 --ctx 512 > $OUT/sanitize_memcheck_7b2_attflag.txt 2>&1; tail -3 $OUT/sanitize_memcheck_7b2_attflag.txt
