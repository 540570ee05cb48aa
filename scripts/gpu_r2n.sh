#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in lib lib_fsw lib_nopeek; do
  for rep in 1 2 3; do
  THK_LIBDIR=$v timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -x -q -k "greedy_ids" > $OUT/pytest_r2n_$v.txt 2>&1; echo "$v rep$rep greedy: $(tail -1 $OUT/pytest_r2n_$v.txt)"
  done
  THK_LIBDIR=$v timeout 300 python scripts/tune.py --steps 100 --repeat 2 > $OUT/tune_r2n_$v.txt 2>&1; tail -1 $OUT/tune_r2n_$v.txt
done
