#!/bin/bash
# round 2: full single-GPU parity suite, the bench line (with parity_check / extra), the reference arm, sanitizers
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -rs > $OUT/pytest_r2q.txt 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_r2q.txt
timeout 900 python bench.py --steps 200 --warmup 10 > $OUT/bench_r2q.json 2> $OUT/bench_r2q.err; echo "bench rc=$?"; tail -c 3500 $OUT/bench_r2q.json; tail -3 $OUT/bench_r2q.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 5 --warmup 0 > $OUT/bench_r2q_reference.json 2>&1; tail -c 1200 $OUT/bench_r2q_reference.json
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_decode.py --tiny --steps 6 > $OUT/sanitize_memcheck_tiny_r2.txt 2>&1; tail -3 $OUT/sanitize_memcheck_tiny_r2.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_decode.py --layers 2 --steps 4 --n-past 500 --ctx 512 > $OUT/sanitize_memcheck_7b2_r2.txt 2>&1; tail -3 $OUT/sanitize_memcheck_7b2_r2.txt
