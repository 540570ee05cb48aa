#!/usr/bin/env python
"""Small fused-decode run for compute-sanitizer (memcheck / racecheck / synccheck): a model with LLaMA-7B tensor shapes but
1-2 layers, a few steps at growing n_past, compared with nothing -- the sanitizer's report is the result.
    compute-sanitizer --tool memcheck python scripts/sanitize_decode.py [--tiny] [--layers 1] [--steps 3] [--n-past 0]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--n-past", type=int, default=0)
    ap.add_argument("--ctx", type=int, default=160)
    ap.add_argument("--kv-f16", action="store_true", help="f16 KV cache option of the fused path")
    args = ap.parse_args()
    import token_hawk_b200 as th
    dev = th.Device(0)
    if args.tiny:
        m = th.LlamaModel.synthetic(dev, 512, 512, 256, 8, 2, 64, kv_f16=args.kv_f16)
    else:
        m = th.LlamaModel.synthetic(dev, 32000, 4096, 256, 32, args.layers, args.ctx, kv_f16=args.kv_f16)
    if args.n_past:
        m.fill_kv(args.n_past)
    tok = 1
    for i in range(args.steps):
        tok, logits = m.eval([tok], args.n_past + i)
        print("step", i, "token", tok, "logit0", float(logits[0]), flush=True)
    m.close()
    dev.close()


if __name__ == "__main__":
    main()
