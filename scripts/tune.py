#!/usr/bin/env python
"""scripts/tune.py -- A/B the persistent decode kernel's tuning knobs on ONE model instance (7B synthetic, ctx 512).

Creates the model once, then for every knob setting times `--steps` launches between CUDA events (after a
short warm-up) and, with --profile, prints the per-phase breakdown of one launch from the in-kernel
timeline.  Output: one line per setting, best first at the end.  Not a bench: bench.py is the contract.

    python scripts/tune.py [--steps 100] [--ctx 512] [--profile] [--set l2_ahead_kb=0,64,128]
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

E, H, V, L, NMULT = 4096, 32, 32000, 32, 256


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--ctx", type=int, default=512)
    ap.add_argument("--layers", type=int, default=L)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--set", action="append", default=[], help="key=v1,v2,...")
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    import torch
    import token_hawk_b200 as th
    import analyze_timeline

    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()
    dev = th.Device(0, stream=stream.cuda_stream)
    model = th.LlamaModel.synthetic(dev, V, E, NMULT, H, args.layers, args.ctx)
    n_past = args.ctx - 1
    model.fill_kv(n_past)
    model.set_token(1)
    keys, vals = [], []
    for s in args.set:
        k, v = s.split("=")
        keys.append(k)
        vals.append([int(x) for x in v.split(",")])
    results = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(args.repeat):
        for combo in itertools.product(*vals) if keys else [()]:
            for k, v in zip(keys, combo):
                model.tune(k, v)
            for _ in range(5):
                model.step_async(n_past)
            torch.cuda.synchronize()
            ev0.record(stream)
            for _ in range(args.steps):
                model.step_async(n_past)
            ev1.record(stream)
            torch.cuda.synchronize()
            model.check()
            ms = ev0.elapsed_time(ev1) / args.steps
            tag = " ".join(f"{k}={v}" for k, v in zip(keys, combo)) or "default"
            print(f"rep{rep} {tag:40s} {ms:8.4f} ms/token  {1e3/ms:7.1f} tok/s", flush=True)
            results.append((ms, tag))
            if args.profile and rep == args.repeat - 1:
                model.profile(True)
                model.step_async(n_past)
                marks, prod = model.profile(True, fetch=True)
                model.profile(False)
                s = analyze_timeline.summarize(marks, prod, args.layers, args.ctx)
                brief = {k: {kk: round(vv, 2) for kk, vv in v.items() if kk in ("us", "arrive_skew_us", "barrier_latency_us", "prologue_us", "first_tile_wait_us", "tile_span_med_us", "tile_span_max_us")}
                         for k, v in s.items() if isinstance(v, dict)}
                print("PHASES", tag, json.dumps(brief), "kernel_us", round(s["kernel_us"], 1), flush=True)
                w = model.last_wait_cycles.astype(float)
                life = w[:, :, 3].clip(min=1)
                import numpy as np
                names = ["ring", "handoff", "poll"]
                for role, sl in (("math", slice(0, 8)), ("producer", slice(8, 9)), ("epilogue", slice(9, 10)), ("reduce", slice(10, 12))):
                    fr = [float(np.median(w[:, sl, i] / life[:, sl])) for i in range(3)]
                    print(f"WAITS {tag} {role:9s} lifetime {np.median(life[:, sl])/1e6:6.3f} Mcycles; share waiting: "
                          + ", ".join(f"{n} {100*f:5.1f}%" for n, f in zip(names, fr)), flush=True)
    results.sort()
    print("BEST", results[0][1], f"{results[0][0]:.4f} ms  {1e3/results[0][0]:.1f} tok/s")
    model.close()


if __name__ == "__main__":
    main()
