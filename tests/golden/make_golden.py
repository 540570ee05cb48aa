"""Generates tests/golden/*.npz|json from the REAL reference host code.

Run in the build container (needs /root/reference): `python tests/golden/make_golden.py`.
It builds oracle/_ref/libthref_host.so (reference th.cpp / th-llama.cpp / th-llama-loader.cpp
compiled unmodified against oracle/webgpu_stub) and records what the reference itself computes:

  fp16_table.npz   ggml_compute_fp16_to_fp32 for all 65536 codes (th.cpp:312) and
                   ggml_compute_fp32_to_fp16 on 20000 seeded floats (th.cpp:335)
  greedy.json      llama_sample_top_p_top_k(temp=0) on seeded logit vectors incl. ties
                   (th-llama.cpp:826-838)
  loader_tiny.json load_llama_file on a tiny ggjt file written by the oracle's writer
                   (th-llama-loader.cpp:485): hparams + per tensor dtype / shape / sha256

The GPU box has no /root/reference; tests there read only these committed files.
"""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TINY_SEED = 0x7B5EED


def greedy_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for n in (2, 7, 512, 4000, 32000):
        for k in range(4):
            v = rng.standard_normal(n).astype(np.float32)
            if k == 1:   # exact tie at the maximum: lowest index must win
                i, j = sorted(rng.choice(n, 2, replace=False).tolist())
                v[i] = v[j] = np.float32(v.max() + 1.0)
            if k == 2:   # all equal
                v[:] = np.float32(0.25)
            if k == 3:   # maximum at the last position, with -inf entries
                v[rng.integers(0, n - 1)] = -np.inf
                v[n - 1] = np.float32(v[np.isfinite(v)].max() + 0.5)
            cases.append(v)
    return cases


def main():
    o.build()
    R = o.ref_lib()
    assert R is not None, "oracle/_ref not built (no /root/reference?)"

    table = np.array([R.ref_fp16_to_fp32(i) for i in range(65536)], dtype=np.float32)
    rng = np.random.default_rng(7)
    f = np.concatenate([
        rng.standard_normal(8000).astype(np.float32),
        (rng.standard_normal(4000) * 1e-6).astype(np.float32),
        (rng.standard_normal(4000) * 7e4).astype(np.float32),
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0, 65520.0, 6e-8, 5.96e-8, 2.98e-8],
                 dtype=np.float32),
        (rng.standard_normal(3990) * 0.02).astype(np.float32)])
    h = np.array([R.ref_fp32_to_fp16(float(x)) for x in f], dtype=np.uint16)
    np.savez_compressed(os.path.join(HERE, "fp16_table.npz"), f16_to_f32_bits=table.view(np.uint32),
                        f32_in_bits=f.view(np.uint32), f16_out=h)

    g = []
    for v in greedy_cases():
        vv = np.ascontiguousarray(v)
        g.append(int(R.ref_greedy(vv.ctypes.data_as(C.POINTER(C.c_float)), vv.size)))
    json.dump({"seed": 2024, "expected": g}, open(os.path.join(HERE, "greedy.json"), "w"))

    m = o.Model.synthetic(o.TINY, TINY_SEED)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "tiny.ggjt")
        m.write_ggjt(path)
        file_sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
        h_ = R.ref_load(path.encode())
        assert h_, "reference loader rejected the file"
        hp = (C.c_int32 * 9)()
        R.ref_hparams(h_, hp)
        tensors = {}
        names = o.tensor_names(o.TINY.n_layer) + ["output.weight-split1", "output.weight-split2",
                                                  "layers.0.key_cache"]
        for name in names:
            ty, shp, nb = C.c_int(), (C.c_int64 * 4)(), C.c_int64()
            p = R.ref_tensor(h_, name.encode(), C.byref(ty), shp, C.byref(nb))
            assert p, name
            raw = C.string_at(p, nb.value)
            tensors[name] = {"f16": ty.value, "shape_lbrc": list(shp), "nbytes": nb.value,
                             "sha256": hashlib.sha256(raw).hexdigest()}
        toks = (C.c_int32 * 64)()
        n = R.ref_tokenize(h_, b"<5><17>", 1, toks, 64)
        R.ref_free(h_)
    json.dump({"seed": TINY_SEED, "file_sha256": file_sha, "hparams": list(hp),
               "dispatches_during_load": int(R.ref_dispatch_count()),
               "tokenize_<5><17>": list(toks[:n]), "tensors": tensors},
              open(os.path.join(HERE, "loader_tiny.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
