"""Reads the decode kernel's in-kernel timeline (bench.py --phase-profile -> gpurun_out/timeline_*.npz) and
reports, per phase kind, where the time goes: barrier skew/latency, fence, prologue, first-tile wait,
tile streaming span, consumer ring-wait share, producer ring-wait share."""
import json
import sys

import numpy as np

E, F, V = 4096, 11008, 32000
NAMES = ["qkv", "att", "wo", "w13", "w2"]


def summarize(marks, prod, n_layer, ctx, verbose=False):
    # phase p (= grid barriers passed): 5l+k layer phases (k: qkv, att, wo, w13, w2); 5L logits; 5L+1 = end of kernel
    start, pro, first, last, arrive, prod_last, waitc, prod_first = (marks[:, :, i] for i in range(8))
    nph = 1 + 5 * n_layer
    out = {}

    def stats(ph_list, bytes_per_phase):
        rows = []
        for ph in ph_list:
            s, a = start[:, ph], arrive[:, ph]
            nxt = start[:, ph + 1]
            dur = nxt.max() - s.min()                      # phase wall time
            skew = a.max() - a.min()                       # arrival skew at the closing barrier
            lat = nxt.min() - a.max()                      # last arrival -> first pass
            fence = np.median(np.where(prod_first[:, ph] > 0, prod_first[:, ph] - s, 0))   # producer's first issue relative to phase start (negative: ahead)
            prol = np.median(np.where(pro[:, ph] > 0, pro[:, ph] - s, 0))
            ft = np.median(np.where(first[:, ph] > 0, first[:, ph] - np.maximum(pro[:, ph], s), 0))
            span = np.where(last[:, ph] > 0, last[:, ph] - first[:, ph], 0)
            rows.append((dur, skew, lat, fence, prol, ft, np.median(span), span.max(), np.median(waitc[:, ph])))
        r = np.array(rows, dtype=np.float64)
        m = r.mean(0)
        return {"us": m[0] / 1e3, "ideal_us_at_6.57TBps": bytes_per_phase / 6.5732e12 * 1e6, "arrive_skew_us": m[1] / 1e3,
                "barrier_latency_us": m[2] / 1e3, "producer_lead_us": -m[3] / 1e3, "prologue_us": m[4] / 1e3, "first_tile_wait_us": m[5] / 1e3,
                "tile_span_med_us": m[6] / 1e3, "tile_span_max_us": m[7] / 1e3, "consumer_ring_wait_kcyc": m[8] / 1e3}

    mb = {"qkv": 3 * E * E * 2, "att": 2 * ctx * E * 4, "wo": E * E * 2, "w13": 2 * E * F * 2, "w2": E * F * 2}
    for k, n in enumerate(NAMES):
        out[n] = stats([5 * l + k for l in range(n_layer)], mb[n])
        out[n]["total_us"] = out[n]["us"] * n_layer
        out[n]["GBps"] = mb[n] / (out[n]["us"] * 1e-6) / 1e9
    out["out"] = stats([5 * n_layer], V * E * 2)
    out["kernel_us"] = float(start[:, nph].max() - start[:, 0].min()) / 1e3
    out["producer_wait_frac_med"] = float(np.median(prod[:, 0] / np.maximum(prod[:, 1], 1)))
    out["producer_tiles_med"] = float(np.median(prod[:, 2]))
    return out


if __name__ == "__main__":
    z = np.load(sys.argv[1])
    s = summarize(z["marks"], z["prod"], int(z["n_layer"]), int(z["ctx"]))
    for k, v in s.items():
        if isinstance(v, dict):
            print(f"{k:6s} " + "  ".join(f"{kk}={vv:.2f}" for kk, vv in v.items()))
        else:
            print(f"{k}: {v:.3f}")
