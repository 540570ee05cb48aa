#!/bin/bash
# round 2 final: full single-GPU parity suite, the bench line (parity_check / extra), the reference arm, memcheck of the tiny decode
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/pytest_r2v.txt 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_r2v.txt
timeout 900 python bench.py --steps 200 --warmup 10 > $OUT/bench_r2v.json 2> $OUT/bench_r2v.err; echo "bench rc=$?"; tail -c 4500 $OUT/bench_r2v.json; tail -3 $OUT/bench_r2v.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 3 --warmup 0 > $OUT/bench_r2v_reference.json 2>&1; tail -c 900 $OUT/bench_r2v_reference.json
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_decode.py --tiny --steps 6 > $OUT/sanitize_memcheck_tiny_r2v.txt 2>&1; tail -3 $OUT/sanitize_memcheck_tiny_r2v.txt
