#!/bin/bash
# tensor-parallel parity + bench on N GPUs of one box (N = number of visible GPUs)
OUT=gpurun_out; mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -rs > $OUT/pytest_r2_tp_${N}gpu.txt 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_r2_tp_${N}gpu.txt
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 > $OUT/bench_r2_tp$n.json 2> $OUT/bench_r2_tp$n.err; echo "bench tp$n rc=$?"; tail -c 1800 $OUT/bench_r2_tp$n.json; tail -3 $OUT/bench_r2_tp$n.err | cut -c1-300
  fi
done
