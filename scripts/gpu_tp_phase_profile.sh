#!/bin/bash
# 8-GPU box: per-phase in-kernel timeline of the tensor-parallel decode at N=8 and N=4 (profiling build), then the product build's numbers
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
run() { # n libdir extra
  THK_LIBDIR=$2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 100 --warmup 5 --no-parity --no-e2e --no-cpu-baseline $3
}
run 8 lib_prof --phase-profile > $OUT/tp8_prof.out 2> $OUT/tp8_prof.err; grep -E "^PHASES|^WAITS" $OUT/tp8_prof.err; tail -1 $OUT/tp8_prof.out | cut -c1-200
run 4 lib_prof --phase-profile > $OUT/tp4_prof.out 2> $OUT/tp4_prof.err; grep -E "^PHASES|^WAITS" $OUT/tp4_prof.err; tail -1 $OUT/tp4_prof.out | cut -c1-200
run 8 lib "" > $OUT/tp8.out 2> $OUT/tp8.err; tail -1 $OUT/tp8.out | cut -c1-200
