#!/usr/bin/env python
"""tests/golden/make_graph_trace.py -- records the command stream the REFERENCE's own th_eval_gpu encodes
(/root/reference/th-llama.cpp:464-660: build_layer_cmdbuf x n_layer + build_final_compute_cmdbuf) for a single
token, by running the unmodified reference host code on the host-memory WebGPU stub (oracle/_ref, built by
oracle/Makefile from the sources where they lie), and commits it as tests/golden/graph_trace_tiny.json.

Only the ENCODING is observed: pipeline label, workgroup counts, bound buffers (name, offset, size), uniform words,
buffer-to-buffer copies, queue writes, submits.  No shader runs, so the arithmetic of the WGSL stays unpinned.
Needs /root/reference (this container); tests/test_graph_trace.py reads the fixture everywhere.
"""
import ctypes as C
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

TINY_SEED = 0x7B5EED
CASES = [(5, 3), (7, 0)]          # (token, n_past)


def record(R, handle, token, n_past):
    R.ref_trace_eval.restype = C.c_char_p
    R.ref_trace_eval.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int]
    toks = (C.c_int32 * 1)(token)
    d = json.loads(R.ref_trace_eval(handle, toks, 1, n_past).decode())
    names, anon = d["buffers"], {}

    def name(i):
        if str(i) in names:
            return names[str(i)]
        return anon.setdefault(i, "uniform#%d" % len(anon))     # the per-call dims uniforms (th-llama.hpp:195-204)

    out = []
    for c in d["commands"]:
        if c["kind"] == "dispatch":
            out.append(["dispatch", c["label"], c["wg"], [[name(b[1]), b[2], b[3]] for b in c["binds"]],
                        {k: v for k, v in c["uniforms"].items()}])
        elif c["kind"] == "copy":
            out.append(["copy", name(c["src"]), c["src_off"], name(c["dst"]), c["dst_off"], c["size"]])
        elif c["kind"] == "write":
            out.append(["write", name(c["dst"]), c["dst_off"], c["size"]])
        else:
            out.append(["submit"])
    return out


def generate():
    o.build()
    R = o.ref_lib()
    assert R is not None and hasattr(R, "ref_trace_eval"), "oracle/_ref not built (no /root/reference?)"
    m = o.Model.synthetic(o.TINY, TINY_SEED)
    cases = []
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "tiny.ggjt")
        m.write_ggjt(path)
        for token, n_past in CASES:
            h = R.ref_load(path.encode())          # a fresh model per case: the recorded stream must not depend on history
            assert h
            cases.append({"token": token, "n_past": n_past, "commands": record(R, h, token, n_past)})
            R.ref_free(h)
    t = o.TINY
    return {"reference": "th_eval_gpu, /root/reference/th-llama.cpp:464-660, n_tokens == 1, on oracle/webgpu_stub",
            "model": {"n_vocab": t.n_vocab, "n_embd": t.n_embd, "n_mult": t.n_mult, "n_head": t.n_head, "n_layer": t.n_layer,
                      "n_ctx": t.n_ctx, "seed": TINY_SEED},
            "cases": cases}


if __name__ == "__main__":
    doc = generate()
    with open(os.path.join(HERE, "graph_trace_tiny.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote graph_trace_tiny.json:", [len(c["commands"]) for c in doc["cases"]], "events")
