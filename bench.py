#!/usr/bin/env python
"""bench.py -- LLaMA-7B f16 single-token decode at ctx=512 on B200 (BASELINE.json configs[1]).

One "step" = one decode step (one token through all 32 layers + logits) with the KV cache already
holding 511 positions, i.e. attention length 512.  Weights (13.2 GB) and KV (0.5 GB) stream from HBM
every step -- far larger than the 126 MB L2, so no L2 flush is needed between iterations.

  value : tokens/s with the token id resident on the device (K timed launches of the persistent
          decode kernel between CUDA events on the launching stream; max over ranks)
  e2e   : the same metric through the public call th_eval_gpu (token id from host memory in, logits
          to pinned host memory out, greedy sample on the host) -- H2D/D2H inside the timed region
  roofline     : algorithmic HBM bytes per launch / measured launch time vs the measured HBM peak
  cpu_baseline : the oracle (CPU port of the reference's arithmetic) on the box's host cores: the FULL 32-layer
                 model, real steps, every core (N=1 only)
  parity_check : before anything is timed, a 2-layer model with the 7B tensor shapes runs 3 decode steps on the same
                 ranks and is compared with the oracle on rank 0 (logits within 1e-3 relative, greedy ids equal);
                 the run fails if it does not match -- a throughput number of a wrong kernel is worthless
  extra        : (N=1) BASELINE configs[3] (ctx=2048 decode) and configs[2] (128-token batched prefill + 1 decode)
  allreduce    : (N>1) per-exchange cost: plain ncclAllReduce of 4096 x f32 (the baseline SURVEY 8e names) next to
                 the in-kernel exchange's cost derived from the measured step time

`--impl reference` times the CPU implementation only (the reference has no CPU path of its own and
its WebGPU path cannot run here -- SURVEY.md 0; the oracle port is the reference arm): same config
keys, full 32-layer model, real steps, every host core whatever OMP_NUM_THREADS says.
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


def workload_name(ctx):
    return f"LLaMA-7B f16, 1-token decode, ctx={ctx} (n_past={ctx - 1})"


def config_dict(ctx, n_layer, world):
    """The same keys for both arms (the driver compares them)."""
    return {"workload": workload_name(ctx), "n_layer": n_layer,
            "arithmetic": "f32 activations/accumulate x f16 weights (reference arithmetic)",
            "l2": "inputs (13.2 GB weights + 0.5 GB KV per step) exceed the 126 MB L2; no flush needed",
            "kv": "f32, synthetic fill for positions < n_past",
            "parallelism": "single GPU" if world == 1 else
                           f"tp{world}: row/column-sharded matvecs, in-kernel one-hop exchange over NVLink peer memory + local reducer warps (2 per layer)"}


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.3)                 # nvidia-smi needs a moment before its first sample
            self.skip = len(self.rows)      # samples taken before the timed region starts do not count
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows[self.skip:]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def oracle_steps(n_ctx_len, max_steps, budget_s, n_layer=L):
    """The oracle (CPU port) on the FULL model: 32 layers at 7B dimensions, attention length n_ctx_len, every host core.
    One untimed warm-up step (pages the 13.5 GB in), then up to max_steps timed steps within budget_s seconds."""
    from oracle import oracle as o
    o.build()
    cores = o.use_all_cores()
    t0 = time.perf_counter()
    m = o.Model.synthetic(o.Config(n_layer=n_layer, n_ctx=n_ctx_len))
    m.fill_kv_synthetic(n_ctx_len - 1)
    t_fill = time.perf_counter() - t0
    m.eval([1], n_ctx_len - 1)
    times = []
    t_all = time.perf_counter()
    while len(times) < max_steps and (len(times) < 3 or time.perf_counter() - t_all < budget_s):
        t1 = time.perf_counter()
        m.eval([1], n_ctx_len - 1)
        times.append(time.perf_counter() - t1)
    t_token = float(np.median(times))
    return {"value": 1.0 / t_token, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"full model: {n_layer} layers at 7B dims + output projection, attention length {n_ctx_len}, {len(times)} real steps "
                      f"after 1 warm-up (median), {cores} OpenMP threads; synthetic fill took {t_fill:.0f} s (untimed)",
            "ms_per_token": t_token * 1e3, "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = oracle_steps(args.ctx, max(3, args.steps), 120.0, args.layers)
    v = base["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": base["steps"], "warmup": 1, "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.ctx, args.layers, max(1, args.gpus)),
            "cpu_baseline": base, "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference has no CPU path and its WebGPU/Dawn path cannot be built here; this arm is the oracle port "
                    "(CPU restatement of the reference's shaders) on the host cores, full model, real steps"}
    if args.layers != L:
        line["invalid"] = "debug run with fewer layers"
    print(json.dumps(line))


def parity_check(th, dev, rank, world, dist, torch):
    """2 layers with the 7B tensor shapes on the SAME ranks / wiring as the benchmark, 3 decode steps, against the
    oracle on rank 0.  Returns the dict for the JSON line; raises if the GPU path does not match."""
    n_ctx = 64
    m = th.LlamaModel.synthetic(dev, V, E, NMULT, H, 2, n_ctx, tp_rank=rank, tp_size=world)
    if world > 1:
        from token_hawk_b200 import tp as tpmod
        tpmod.wire_distributed(m, rank, world)
    toks, got_ids, got_logits = [1, 3000, 31999], [], []
    for i, t in enumerate(toks):
        m.eval_launch([t], i)
        tok, logits = m.eval_finish()
        got_ids.append(int(tok))
        if world > 1:
            parts = [torch.empty(V // world, dtype=torch.float32, device="cuda") for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(logits).cuda())
            logits = torch.cat(parts).cpu().numpy()
        got_logits.append(logits)
    if world > 1:
        dist.barrier()
    m.close()
    res = None
    if rank == 0:
        from oracle import oracle as o
        o.build()
        o.use_all_cores()
        om = o.Model.synthetic(o.Config(n_layer=2, n_ctx=n_ctx))
        errs, same = [], True
        for i, t in enumerate(toks):
            ref = om.eval([t], i)
            errs.append(float(np.abs(got_logits[i] - ref).max() / np.abs(ref).max()))
            same = same and (o.greedy(ref) == got_ids[i])
        res = {"rel_err": max(errs), "greedy_equal": bool(same), "tolerance": 1e-3,
               "what": f"2-layer model with LLaMA-7B tensor shapes, 3 decode steps on the benchmark's {world} rank(s) vs the CPU oracle"}
        if not (res["rel_err"] < 1e-3 and same):
            raise SystemExit("parity check FAILED: " + json.dumps(res))
    return res


def nccl_allreduce_us(torch, dist, n_elems=E, reps=400):
    """The baseline SURVEY 8e names: ncclAllReduce(sum, 4096 x f32) per exchange, back to back on one stream."""
    t = torch.zeros(n_elems, dtype=torch.float32, device="cuda")
    for _ in range(20):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        dist.all_reduce(t)
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / reps
    tt = torch.tensor([us], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


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
    ap.add_argument("--no-extra", action="store_true", help="skip the ctx=2048 and prefill-128 measurements (N=1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-flight parity check (profiling runs only)")
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

    parity = None if args.no_parity else parity_check(th, dev, rank, world, dist, torch)

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

    def timed_steps(m, n_steps, n_past_):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(n_steps):
            m.step_async(n_past_)
        ev1.record(stream)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        m.check()
        return ms / n_steps

    # ---- value: device-resident token, K launches between events on the launching stream ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step = timed_steps(model, args.steps, n_past)
    clocks = sampler.stop()
    tok_s = 1e3 / ms_step

    # ---- e2e: th_eval_gpu with host token in / host logits out, every step ----
    e2e = None
    if not args.no_e2e:
        for _ in range(3):
            model.eval([1], n_past)
        ke = max(10, min(args.steps, 100))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"timeline_ctx{args.ctx}_n{world}_r{rank}.npz"), marks=marks, prod=prod,
                            n_layer=args.layers, ctx=args.ctx)
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import analyze_timeline
        if rank == 0:
            summ = analyze_timeline.summarize(marks, prod, args.layers, args.ctx)
            keys = ("us", "arrive_skew_us", "barrier_latency_us", "prologue_us", "first_tile_wait_us", "tile_span_med_us", "tile_span_max_us")
            brief = {k: {kk: round(vv, 2) for kk, vv in v.items() if kk in keys} for k, v in summ.items() if isinstance(v, dict)}
            print(f"PHASES n_gpus={world} " + json.dumps(brief) + f" kernel_us {summ['kernel_us']:.1f}", file=sys.stderr)
            w = model.last_wait_cycles.astype(float)
            life = w[:, :, 3].clip(min=1)
            for role, sl in (("math", slice(0, 8)), ("producer", slice(8, 9)), ("epilogue", slice(9, 10)), ("reducer", slice(10, 11))):
                fr = [float(np.median(w[:, sl, i] / life[:, sl])) for i in range(3)]
                print(f"WAITS {role:9s} share of lifetime waiting: ring {100*fr[0]:5.1f}%, handoff {100*fr[1]:5.1f}%, poll {100*fr[2]:5.1f}%", file=sys.stderr)

    peak, peak_src = load_peaks()

    def roofline_of(ms, ctx, traffic=None, kv_elem_bytes=4):
        # per-GPU algorithmic bytes: matrices and KV shard by tp; gains and the embedding row are replicated
        bytes_w = (args.layers * 2 * (4 * E * E + 3 * E * F) + 2 * V * E) // world + (2 * args.layers + 1) * 4 * E + 2 * E
        bytes_total = bytes_w + kv_bytes(ctx, args.layers) * kv_elem_bytes // 4 // world
        achieved = bytes_total / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "decode_kernel (persistent, 1 launch per token per GPU)", "achieved": achieved, "per_gpu": True, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_total, "weights_only_GBps": bytes_w / (ms * 1e-3) / 1e9,
                "frac_of_8TBps_spec": achieved / 8000.0}

    # dram__bytes of one launch from the committed ncu capture: measured for 1 GPU, ctx 512, 32 layers only
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "decode_kernel_traffic.json")
    if world == 1 and args.ctx == 512 and args.layers == L and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roof = roofline_of(ms_step, args.ctx, traffic)

    # ---- N>1: what one exchange costs, next to the plain NCCL all-reduce ----
    allreduce = None
    if world > 1:
        nccl_us = nccl_allreduce_us(torch, dist)
        allreduce = {"exchanges_per_token": 2 * args.layers, "nccl_allreduce_4096xf32_us": nccl_us,
                     "nccl_only_ms_per_token": 2 * args.layers * nccl_us * 1e-3,
                     "note": "ncclAllReduce(sum, 4096 x f32) back to back on one stream, max over ranks: the cost a decode step would pay "
                             "per exchange with NCCL on the data path (64 per token); the in-kernel exchange is inside ms_per_step"}

    model.close()

    # ---- N=1: the other single-GPU configurations of BASELINE.json (extra keys, same launch-timing method) ----
    extra = None
    if world == 1 and not args.no_extra and args.layers == L and args.ctx == 512:
        extra = {}
        try:
            m2 = th.LlamaModel.synthetic(dev, V, E, NMULT, H, L, 2048)
            m2.fill_kv(2047)
            m2.set_token(1)
            for _ in range(5):
                m2.step_async(2047)
            ms2 = timed_steps(m2, max(20, args.steps // 2), 2047)
            extra["ctx2048_decode"] = {"workload": workload_name(2048), "value": 1e3 / ms2, "unit": "tokens/s", "ms_per_step": ms2,
                                       "roofline": roofline_of(ms2, 2048)}
            # the f16 KV option (parity budget in tests/test_gpu_kv_f16.py): the same step with half the KV bytes
            m3 = th.LlamaModel.synthetic(dev, V, E, NMULT, H, L, 2048, kv_f16=True)
            m3.fill_kv(2047)
            m3.set_token(1)
            for _ in range(5):
                m3.step_async(2047)
            ms3 = timed_steps(m3, max(20, args.steps // 2), 2047)
            m3.close()
            extra["ctx2048_decode_f16kv"] = {"workload": workload_name(2048) + ", f16 KV cache option (not the reference's f32 cache)",
                                             "value": 1e3 / ms3, "unit": "tokens/s", "ms_per_step": ms3,
                                             "roofline": roofline_of(ms3, 2048, kv_elem_bytes=2)}
            # configs[2]: a 128-token prompt in one batched pass (tcgen05 GEMM path) at n_past = 0, then one decode step
            prompt = (np.arange(128, dtype=np.int64) * 7919 % V).astype(np.int32).tolist()
            for _ in range(2):
                m2.eval(prompt, 0)
            reps = 5
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                tok, _ = m2.eval(prompt, 0)
            torch.cuda.synchronize()
            ms_pre = (time.perf_counter() - t0) * 1e3 / reps
            t0 = time.perf_counter()
            for _ in range(reps):
                m2.eval([tok], 128)
            ms_dec = (time.perf_counter() - t0) * 1e3 / reps
            flops = 2.0 * 128 * (L * (4 * E * E + 3 * E * F) + V * E)          # weight GEMMs of one 128-token pass
            extra["prefill128"] = {"workload": "LLaMA-7B f16, 128-token prompt batched prefill + 1 decode step, 1xB200 (tensor-core matmul path)",
                                   "prefill_ms": ms_pre, "prefill_tokens_per_s": 128e3 / ms_pre, "decode_after_prefill_ms": ms_dec,
                                   "weight_gemm_tflops": flops / (ms_pre * 1e-3) / 1e12,
                                   "weights_GBps": W_BYTES / (ms_pre * 1e-3) / 1e9,
                                   "timing": "host wall clock around th_eval_gpu (host tokens in, logits out), 5 repetitions after 2 warm-ups"}
            m2.close()

        except Exception as e:      # the headline measurement above stands; say what is missing instead of losing the line
            extra["error"] = f"{type(e).__name__}: {e}"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = oracle_steps(args.ctx, 5, 30.0)

    line = {"metric": METRIC, "value": tok_s, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(args.ctx, args.layers, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roof, "cpu_baseline": cpu,
            "parity_check": parity, "allreduce": allreduce, "extra": extra}
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
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
