"""The graph, pinned by the reference itself.

tests/golden/graph_trace_tiny.json holds the command stream the reference's own th_eval_gpu
(/root/reference/th-llama.cpp:464-660 -> build_layer_cmdbuf :270-452 x n_layer, build_final_compute_cmdbuf :240-268)
encodes for one token, recorded by running the unmodified reference host code on the host-memory WebGPU stub
(tests/golden/make_graph_trace.py).  Here the oracle's restated graph is compared with it command by command: pipeline
label, operand buffers, uniform words, copy offsets and sizes.  What stays unpinned is the ARITHMETIC inside each
pipeline: no WGSL can run in this container, the stub's dispatches do nothing.
"""
import json
import os
import struct

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "graph_trace_tiny.json")

# words of the uniform block that carry meaning per pipeline (th-llama.hpp:181-204: LlamaNetworkUniforms {n_past,
# n_tokens}, LlamaTensorDimsUniforms {A_B, A_M, A_N, scale, B_B, B_M, B_N, offset})
UNIFORM_WORDS = {"RoPE": 2, "transpose": 3, "mat_mul": 7, "row_softmax": 3}
PER_LAYER, FINAL = 28, 6      # commands per layer / after the last layer (th-llama.cpp:592-643: 32 * 28 + 6 = 902 at 7B)


@pytest.fixture(scope="module")
def fixture():
    return json.load(open(FIXTURE))


def commands(case):
    """dispatches and copies (the contents of the command buffers), without queue writes and submits"""
    return [c for c in case["commands"] if c[0] in ("dispatch", "copy")]


def canonical(case):
    """the reference stream in the oracle trace's vocabulary"""
    out = []
    for c in commands(case):
        if c[0] == "copy":
            _, src, src_off, dst, dst_off, size = c
            assert src_off == 0
            out.append(["copy", src, dst, dst_off, size])
            continue
        _, label, wg, binds, uniforms = c
        operands = [b[0] for b in binds if not b[0].startswith("uniform#") and b[0] != "networkUniforms"]
        assert all(b[1] == 0 for b in binds), "the single-token path binds every buffer at offset 0"
        words = []
        if label in UNIFORM_WORDS:
            assert len(uniforms) <= 1
            for w in uniforms.values():
                words = w[:UNIFORM_WORDS[label]]
        out.append(["op", label, operands, words])
    return out


def test_reference_stream_has_the_documented_length(fixture):
    n_layer = fixture["model"]["n_layer"]
    for case in fixture["cases"]:
        assert len(commands(case)) == PER_LAYER * n_layer + FINAL
    assert PER_LAYER * 32 + FINAL == 902                      # SURVEY / th-llama.cpp:592-643 at 7B
    # one submit for the layers + final, one for the result copy; embedding row written to inp0 and inp6 (:577-585)
    case = fixture["cases"][0]
    assert sum(1 for c in case["commands"] if c[0] == "submit") == 2
    E = fixture["model"]["n_embd"]
    writes = [c for c in case["commands"] if c[0] == "write" and c[1] in ("inp0", "inp6")]
    assert [w[1:] for w in writes] == [["inp0", 0, 4 * E], ["inp6", 0, 4 * E]]


def test_every_layer_issues_the_same_28_commands(fixture):
    for case in fixture["cases"]:
        can = canonical(case)
        n_layer = fixture["model"]["n_layer"]
        layers = [can[PER_LAYER * l:PER_LAYER * (l + 1)] for l in range(n_layer)]
        strip = lambda cmds, l: json.loads(json.dumps(cmds).replace("layers.%d." % l, "layers.L."))
        for l in range(1, n_layer):
            assert strip(layers[l], l) == strip(layers[0], 0)


def test_oracle_graph_equals_the_reference_stream(fixture, oracle):
    m = fixture["model"]
    cfg = oracle.Config(n_vocab=m["n_vocab"], n_embd=m["n_embd"], n_mult=m["n_mult"], n_head=m["n_head"], n_layer=m["n_layer"],
                        n_ctx=m["n_ctx"])
    model = oracle.Model.synthetic(cfg, m["seed"])
    for case in fixture["cases"]:
        model.reset()
        if case["n_past"]:
            model.fill_kv_synthetic(case["n_past"])
        ours = oracle.eval_trace(model, case["token"], case["n_past"])
        ref = canonical(case)
        body, tail = ref[:PER_LAYER * m["n_layer"]], ref[PER_LAYER * m["n_layer"]:]
        assert ours[:len(body)] == body
        # after the last layer: rms_norm, row_element_multiply, then the logits.  The reference splits output.weight in two
        # halves, multiplies each into its own scratch row and adds them (th-llama.cpp:255-262; SURVEY F2/F3: the intended
        # result is the full dot product); the oracle and the product issue ONE matvec over the whole matrix.
        assert ours[len(body):len(body) + 2] == tail[:2]
        assert [c[1] for c in tail[2:5]] == ["vector_mat_mul_split", "vector_mat_mul_split", "vector_reduce"]
        assert tail[2][2][:2] == ["inp0", "output-split1"] and tail[3][2][:2] == ["inp0", "output-split2"]
        assert tail[4][2][0] == tail[2][2][2] == "out" and tail[4][2][1] == tail[3][2][2]      # out += second half
        assert ours[len(body) + 2:] == [["op", "output_matvec", ["inp0", "output", "out"], []]]
        assert tail[5] == ["copy", "out", "resultBuffer", 0, 4 * m["n_vocab"]]


def test_uniform_words_follow_the_shapes(fixture):
    """n_past enters the stream in exactly four places per layer: RoPE's uniform, the KV append offsets, the transposes of
    the caches and the attention matmuls / softmax (N = n_past + 1)."""
    m = fixture["model"]
    E, H = m["n_embd"], m["n_head"]
    D = E // H
    for case in fixture["cases"]:
        n_past, N = case["n_past"], case["n_past"] + 1
        layer0 = canonical(case)[:PER_LAYER]
        by_label = {}
        for c in layer0:
            by_label.setdefault(c[1] if c[0] == "op" else "copy", []).append(c)
        assert [c[3] for c in by_label["RoPE"]] == [[n_past, 1]] * 2
        assert [c[3:] for c in by_label["copy"][:2]] == [[4 * E * n_past, 4 * E]] * 2
        assert [c[3] for c in by_label["transpose"]] == [[N, H, D], [N, H, D], [1, H, D], []]
        scale = struct.unpack("<I", struct.pack("<f", 1.0 / D ** 0.5))[0]
        one = struct.unpack("<I", struct.pack("<f", 1.0))[0]
        assert [c[3] for c in by_label["mat_mul"]] == [[H, 1, D, scale, H, D, N], [H, 1, N, one, H, N, D]]
        assert by_label["row_softmax"][0][3] == [H, 1, N]


def test_fixture_is_what_the_reference_encodes_today():
    """Where /root/reference is present (this container) the stream is recorded again and must equal the committed file."""
    if not os.path.exists("/root/reference/th-llama.cpp"):
        pytest.skip("no /root/reference here")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_graph_trace", os.path.join(HERE, "golden", "make_graph_trace.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    live = json.loads(json.dumps(mod.generate()))
    assert live == json.load(open(FIXTURE))
