#!/usr/bin/env python
"""scripts/tile_timeline.py -- per-tile dynamics of one phase of the persistent decode kernel (7B synthetic, ctx 512).

For the chosen phases prints, per tile index, the median over CTAs of: when the producer issued the tile and when the
math warps retired it, both relative to the earliest phase start, plus the inter-tile gaps.  Shows where a phase loses
time: ring drain at the start, pipeline restart bubble, steady state, tail."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
E, H, V, L, NMULT = 4096, 32, 32000, 32, 256
NAMES = ["qkv", "att", "wo", "w13", "w2"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", type=int, default=10)
    ap.add_argument("--kinds", default="0,2,3,4")
    ap.add_argument("--set", action="append", default=[])
    args = ap.parse_args()
    import torch
    import token_hawk_b200 as th
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()
    dev = th.Device(0, stream=stream.cuda_stream)
    model = th.LlamaModel.synthetic(dev, V, E, NMULT, H, L, 512)
    model.fill_kv(511)
    model.set_token(1)
    for s in args.set:
        k, v = s.split("=")
        model.tune(k, int(v))
    for _ in range(5):
        model.step_async(511)
    for kind in [int(x) for x in args.kinds.split(",")]:
        ph = 5 * args.layer + kind
        model.tune("prof_phase", ph)
        model.profile(True)
        model.step_async(511)
        marks, prod = model.profile(True, fetch=True)
        model.profile(False)
        done, issue = model.last_tile_times
        start = marks[:, ph, 0]
        t0 = start.min()
        pro = marks[:, ph, 1] - t0
        arrive = marks[:, ph, 4] - t0
        nxt = marks[:, ph + 1, 0] - t0
        prev_arrive = marks[:, ph - 1, 4] - t0
        nt = int((done > 0).sum(1).max())
        print(f"== layer {args.layer} {NAMES[kind]} (phase {ph}): tiles/CTA up to {nt}; prev arrive med {np.median(prev_arrive)/1e3:.2f} max {prev_arrive.max()/1e3:.2f}; "
              f"start med {np.median(start - t0)/1e3:.2f}; prologue end med {np.median(pro)/1e3:.2f}; arrive med {np.median(arrive)/1e3:.2f} max {arrive.max()/1e3:.2f}; next start min {nxt.min()/1e3:.2f}")
        print(" tile  issue_med  retire_med  retire_p90  gap_med")
        prev = None
        for t in range(min(nt, 64)):
            ok = done[:, t] > 0
            d = (done[ok, t] - t0) / 1e3
            i = (issue[ok, t] - t0) / 1e3
            gap = "" if prev is None else f"{np.median(d) - prev:7.2f}"
            print(f" {t:4d}  {np.median(i):9.2f}  {np.median(d):10.2f}  {np.percentile(d, 90):10.2f}  {gap}")
            prev = np.median(d)
    model.close()


if __name__ == "__main__":
    main()
