#!/usr/bin/env python
"""scripts/prefill_once.py -- BASELINE configs[2] once, for a profiler: LLaMA-7B (synthetic), one warm-up and one measured
128-token batched prefill through th_eval_gpu (op graph + tcgen05 GEMM path), then one fused decode step.  Prints wall times.
    ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_prefill.csv python scripts/prefill_once.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import token_hawk_b200 as th
    torch.cuda.set_device(0)
    dev = th.Device(0, stream=torch.cuda.current_stream().cuda_stream)
    m = th.LlamaModel.synthetic(dev, 32000, 4096, 256, 32, int(os.environ.get("LAYERS", "32")), 256)
    prompt = (np.arange(128, dtype=np.int64) * 7919 % 32000).astype(np.int32).tolist()
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tok, _ = m.eval(prompt, 0)
        torch.cuda.synchronize()
        print(f"prefill128 rep{rep}: {(time.perf_counter() - t0) * 1e3:.2f} ms, launches {getattr(m, 'last_launches', '?')}", flush=True)
    t0 = time.perf_counter()
    m.eval([tok], 128)
    print(f"decode after prefill: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
    m.close()


if __name__ == "__main__":
    main()
