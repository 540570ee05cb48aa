#!/bin/bash
# One gpurun call: bench line, ctx=2048, reference arm, ncu launch list and one ncu --set full capture.
# (The in-kernel timeline needs the profiling build: THK_LIBDIR=lib_prof python scripts/tune.py --profile)
# Usage (from the repo root on the GPU box): bash scripts/gpu_bench_profile.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
echo "== bench ctx=512 =="
python bench.py --steps 200 --warmup 10 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
tail -c 3000 $OUT/bench_${TAG}.json
echo "== bench ctx=2048 =="
python bench.py --steps 100 --warmup 5 --ctx 2048 --no-cpu-baseline > $OUT/bench_${TAG}_ctx2048.json 2> $OUT/bench_${TAG}_ctx2048.err
tail -c 1500 $OUT/bench_${TAG}_ctx2048.json
echo "== reference arm =="
python bench.py --impl reference --steps 2 --warmup 0 > $OUT/bench_${TAG}_reference.json 2>&1
tail -c 1200 $OUT/bench_${TAG}_reference.json
echo "== ncu launch list =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/launches_${TAG}.log 2>&1
grep -c decode_kernel $OUT/launches_${TAG}.csv
grep decode_kernel $OUT/launches_${TAG}.csv | tail -3
echo "== ncu full (decode_kernel) =="
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 4 -c 1 -f -o $OUT/prof_decode_${TAG} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_${TAG}.log 2>&1
tail -5 $OUT/ncu_full_${TAG}.log
ls -la $OUT
