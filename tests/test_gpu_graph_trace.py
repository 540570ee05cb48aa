"""GPU: the product's op-graph path (EvalPath_OpGraph: th_eval_gpu -> build_layer_cmdbuf x n_layer ->
build_final_compute_cmdbuf behind the reference's cmdbuf_* surface) issues the command sequence the reference's own
th_eval_gpu encodes (tests/golden/graph_trace_tiny.json, recorded on the WebGPU stub), command for command.
Documented differences: the embedding row is selected on the device (the reference's kUseGpuEmbeddingSelection branch,
th-llama.cpp:552-575: one f16->f32 conversion + the copy to inp6) instead of two queue writes; the logits are ONE matvec
over output.weight instead of two half-matrix matvecs + a reduce; the result is read back by th_eval_gpu itself."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LABEL = {"rms_norm": "cmdbuf_rms_norm", "row_element_multiply": "cmdbuf_row_element_multiply",
                 "vector_mat_mul": "cmdbuf_vector_mat_mul_trans", "RoPE": "cmdbuf_RoPE", "transpose": "cmdbuf_transpose",
                 "mat_mul": "cmdbuf_mat_mul", "row_softmax": "cmdbuf_row_softmax", "addition": "cmdbuf_addition",
                 "silu": "cmdbuf_silu", "hadamard_in_place": "cmdbuf_element_mult_in_place"}


def test_opgraph_launch_sequence_equals_the_reference_stream(oracle):
    import token_hawk_b200 as th
    fx = json.load(open(os.path.join(HERE, "golden", "graph_trace_tiny.json")))
    m = fx["model"]
    dev = th.Device(0)
    g = th.LlamaModel.synthetic(dev, m["n_vocab"], m["n_embd"], m["n_mult"], m["n_head"], m["n_layer"], m["n_ctx"], m["seed"])
    g.set_eval_path(th.EVAL_OPGRAPH)
    case = fx["cases"][0]
    for i in range(case["n_past"]):
        g.eval([1 + i], i)
    labels = th.trace_commands(lambda: g.eval([case["token"]], case["n_past"]))
    ref = [c for c in case["commands"] if c[0] in ("dispatch", "copy")]
    n_body = 28 * m["n_layer"]
    expect = ["cmdbuf_f16_f32_conversion", "copy"]                                   # th-llama.cpp:552-575
    expect += ["copy" if c[0] == "copy" else PRODUCT_LABEL[c[1]] for c in ref[:n_body]]
    assert [c[1] for c in ref[n_body:n_body + 5]] == ["rms_norm", "row_element_multiply", "vector_mat_mul_split",
                                                      "vector_mat_mul_split", "vector_reduce"]
    expect += ["cmdbuf_rms_norm", "cmdbuf_row_element_multiply", "cmdbuf_vector_mat_mul_trans"]
    assert labels == expect
    # and the fused path issues none of them: one persistent kernel per token
    g.set_eval_path(th.EVAL_FUSED)
    assert th.trace_commands(lambda: g.eval([case["token"]], case["n_past"] + 1)) == []
    g.close()
    dev.close()
