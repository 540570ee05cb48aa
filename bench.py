#!/usr/bin/env python
"""bench.py -- LLaMA-7B f16 single-token decode at ctx=512 on B200 (BASELINE.json configs[1]).

One "step" = one decode step (one token through all 32 layers + logits) with the KV cache already
holding 511 positions, i.e. attention length 512.  Weights (13.2 GB) and KV (0.5 GB) stream from HBM
every step -- far larger than the 126 MB L2, so no L2 flush is needed between iterations.

  value : tokens/s with the token id resident on the device (K timed launches of the persistent
          decode kernel between CUDA events on the launching stream)
  e2e   : the same metric through the public call th_eval_gpu (token id from host memory in, logits
          to pinned host memory out, greedy sample on the host) -- H2D/D2H inside the timed region
  roofline     : algorithmic HBM bytes per launch / measured launch time vs the measured HBM peak
  cpu_baseline : the oracle (CPU port of the reference's arithmetic) on the box's host cores

`--impl reference` times the CPU implementation only (the reference has no CPU path of its own and
its WebGPU path cannot run here -- SURVEY.md 0; the oracle port is the reference arm).
Launch: python bench.py [--gpus N --steps K --warmup W]; N>1 via torch.distributed.run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# LLaMA-7B
E, H, F, V, L, NMULT = 4096, 32, 11008, 32000, 32, 256
W_BYTES = L * 2 * (4 * E * E + 3 * E * F) + 2 * V * E + (2 * L + 1) * 4 * E + 2 * E   # 13,215,2xx,xxx


METRIC = "tokens/sec LLaMA-7B f16 single-token decode; achieved HBM GB/s vs roofline"      # BASELINE.json


def workload_name(ctx, n_gpus):
    return f"LLaMA-7B f16, 1-token decode, ctx={ctx} (n_past={ctx - 1}), {n_gpus}xB200"


def kv_bytes(n_ctx_len, n_layer=L):
    return 2 * n_ctx_len * E * 4 * n_layer + 2 * E * 4 * n_layer


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(n_ctx_len, sample_layers=4, steps=2):
    """Oracle (CPU port) on a bounded sample: `sample_layers` of the 32 layers at 7B dimensions plus the
    output projection, attention length n_ctx_len, all host threads; extrapolated to 32 layers."""
    from oracle import oracle as o
    o.build()
    cfg = o.Config(n_layer=sample_layers, n_ctx=n_ctx_len)
    m = o.Model.synthetic(cfg)
    m.fill_kv_synthetic(n_ctx_len - 1)
    m.eval([1], n_ctx_len - 1)                      # warm-up (page in weights)
    t0 = time.perf_counter()
    for _ in range(steps):
        m.eval([1], n_ctx_len - 1)
    t_model = (time.perf_counter() - t0) / steps
    x = np.ones(E, np.float32)
    Wout = m.tensor("output.weight")
    o.matvec_f16(x, Wout)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.matvec_f16(x, Wout)
    t_out = (time.perf_counter() - t0) / steps
    t_layer = max(t_model - t_out, 1e-9) / sample_layers
    t_token = L * t_layer + t_out
    return {"value": 1.0 / t_token, "unit": "tokens/s", "cores": o.num_threads(), "kind": "port",
            "sample": f"{sample_layers} of {L} layers at 7B dims + output projection, attention length {n_ctx_len}, "
                      f"{steps} steps after 1 warm-up, extrapolated to {L} layers ({t_layer*1e3:.1f} ms/layer, {t_out*1e3:.1f} ms logits)",
            "ms_per_token": t_token * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, vals = max(1, args.steps), []
    base = None
    t_all = time.perf_counter()
    for i in range(min(steps, 3)):                 # each step = one bounded sample; capped to keep the run short
        base = cpu_baseline(args.ctx, sample_layers=2, steps=1)
        vals.append(base["value"])
        if time.perf_counter() - t_all > 150:
            break
    v = float(np.median(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.ctx, max(1, args.gpus)), "n_layer": L,
                       "arithmetic": "f32 activations/accumulate x f16 weights (reference arithmetic)",
                       "arm": "CPU port of the reference arithmetic (oracle) on the host cores; bounded sample per step"},
            "cpu_baseline": base, "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference has no CPU path and its WebGPU/Dawn path cannot be built here; this arm is the oracle port on host cores"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ctx", type=int, default=512, help="attention length of the measured step (n_past = ctx-1)")
    ap.add_argument("--layers", type=int, default=L, help="debug: fewer layers (invalid as a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--phase-profile", action="store_true", help="print per-phase time from in-kernel timestamps (stderr)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import token_hawk_b200 as th

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective; the driver reads ONE JSON line from stdout,
        # so stdout is parked on stderr until the result line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    else:
        dist = None
        saved_stdout = None
    stream = torch.cuda.current_stream()
    dev = th.Device(local_rank, stream=stream.cuda_stream)

    n_past = args.ctx - 1
    model = th.LlamaModel.synthetic(dev, V, E, NMULT, H, args.layers, args.ctx, tp_rank=rank, tp_size=world)
    if world > 1:
        from token_hawk_b200 import tp as tpmod
        tpmod.wire_distributed(model, rank, world)      # peers' exchange regions via CUDA IPC
    model.fill_kv(n_past)
    model.set_token(1)
    for _ in range(max(3, args.warmup)):
        model.step_async(n_past)
    torch.cuda.synchronize()
    model.check()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident token, K launches between events on the launching stream ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        model.step_async(n_past)
    ev1.record(stream)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop()
    model.check()
    ms_step = ms_total / args.steps
    tok_s = 1e3 / ms_step

    # ---- e2e: th_eval_gpu with host token in / host logits out, every step ----
    e2e = None
    if not args.no_e2e:
        for _ in range(3):
            model.eval([1], n_past)
        ke = max(10, min(args.steps, 100))
        barrier()
        ev0.record(stream)
        t0 = time.perf_counter()
        for _ in range(ke):
            tok, logits = model.eval([1], n_past)
        ev1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) / ke
        ms_e2e = max_over_ranks(max(ev0.elapsed_time(ev1) / ke, wall * 1e3))
        e2e = {"value": 1e3 / ms_e2e, "unit": "tokens/s", "h2d_bytes_per_step": 4 * world, "d2h_bytes_per_step": V * 4 + (4 * world if world > 1 else 0),
               "ms_per_step": ms_e2e, "steps": ke, "api": "th_eval_gpu (capi_eval): host token id -> logits in pinned host memory -> host greedy"}

    if args.phase_profile:      # needs the profiling build of the kernel library: THK_LIBDIR=lib_prof (see token_hawk_b200/build.py)
        model.profile(True)
        model.step_async(n_past)
        marks, prod = model.profile(True, fetch=True)
        model.profile(False)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"timeline_ctx{args.ctx}.npz"), marks=marks, prod=prod,
                            n_layer=args.layers, ctx=args.ctx)
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import analyze_timeline
        print("PHASES " + json.dumps(analyze_timeline.summarize(marks, prod, args.layers, args.ctx)), file=sys.stderr)

    peak, peak_src = load_peaks()
    # per-GPU algorithmic bytes: matrices and KV shard by tp; gains and the embedding row are replicated
    bytes_w = (args.layers * 2 * (4 * E * E + 3 * E * F) + 2 * V * E) // world + (2 * args.layers + 1) * 4 * E + 2 * E
    bytes_total = bytes_w + kv_bytes(args.ctx, args.layers) // world
    achieved = bytes_total / (ms_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "decode_kernel_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roof = {"bound": "hbm", "kernel": "decode_kernel (persistent, 1 launch per token per GPU)", "achieved": achieved, "per_gpu": True, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_total, "weights_only_GBps": bytes_w / (ms_step * 1e-3) / 1e9,
            "frac_of_8TBps_spec": achieved / 8000.0}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.ctx)

    line = {"metric": METRIC, "value": tok_s, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.ctx, world), "n_layer": args.layers,
                       "arithmetic": "f32 activations/accumulate x f16 weights (reference arithmetic)",
                       "l2": "inputs (13.2 GB weights + 0.5 GB KV per step) exceed the 126 MB L2; no flush needed",
                       "kv": "f32, synthetic fill for positions < n_past",
                       "parallelism": "single GPU" if world == 1 else f"tp{world}: row/column-sharded matvecs, in-kernel one-shot all-reduce over NVLink peer memory (2 per layer)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roof, "cpu_baseline": cpu}
    if args.layers != L:
        line["invalid"] = "debug run with fewer layers"
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.barrier()
        model.close()
        dist.destroy_process_group()
    else:
        model.close()


if __name__ == "__main__":
    main()
