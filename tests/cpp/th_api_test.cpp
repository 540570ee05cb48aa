// th_api_test.cpp -- exercises the th:: op surface from C++, the way the reference's own graph code uses it
// (th-llama.cpp:270-452): TensorBuffer allocation/upload, the two-phase ComputePipeline idiom, shape views by
// mutating .shape, validators returning an empty CommandBuffer, deferred CommandBuffers + queue_submit, and a
// whole-model evaluation through load-free synthetic weights.  Prints "PASS" / "FAIL: ..." lines; exit code 0/1.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "th/th-llama-loader.hpp"

using namespace th;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL: " __VA_ARGS__); printf("  [%s:%d]\n", __FILE__, __LINE__); ++g_fail; } } while (0)

static std::vector<float> download(thk_ctx* ctx, const TensorBuffer& t) {
    std::vector<float> v(t.get_size_bytes() / 4);
    thk_download(ctx, v.data(), t.gpu, 0, t.get_size_bytes());
    return v;
}

int main() {
    thk_ctx* ctx = nullptr;
    if (thk_init(0, &ctx) != THK_OK) { printf("FAIL: thk_init: %s\n", thk_last_error()); return 1; }
    EncoderTag* enc = reinterpret_cast<EncoderTag*>(1);

    // ---- matvec with f16 weights: y = W x (cmdbuf_vector_mat_mul_trans) ----
    const int R = 96, C = 512;
    std::vector<float> x(C), yref(R, 0.f);
    std::vector<ggml_fp16_t> W((size_t)R * C);
    for (int c = 0; c < C; ++c) x[c] = sinf(0.1f * c);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) {
            const float w = 0.01f * (float)((r * 31 + c * 7) % 41 - 20);
            W[(size_t)r * C + c] = ggml_compute_fp32_to_fp16(w);
            yref[r] += x[c] * ggml_compute_fp16_to_fp32(W[(size_t)r * C + c]);
        }
    TensorBuffer tx(x.data(), TensorShape{0, 0, 1, C}, TensorType_F32, false, ctx, ctx);
    TensorBuffer tW(W.data(), TensorShape{0, 0, R, C}, TensorType_F16, false, ctx, ctx);
    TensorBuffer ty(TensorShape{0, 0, 1, R}, TensorType_F32, ctx);
    CHECK(tx.is_valid() && tW.is_valid() && ty.is_valid(), "tensor allocation");

    ComputePipeline pipe{true};                                   // buildPipelineFlag: first call only builds
    CommandBuffer c0 = cmdbuf_vector_mat_mul_trans(ctx, nullptr, nullptr, &pipe, tx, tW, ty, 0);
    CHECK(!c0.is_valid() && pipe.is_valid(), "two-phase: first call builds the pipeline and records nothing");
    thk_memset(ctx, ty.gpu, 0, ty.get_size_bytes());
    CommandBuffer c1 = cmdbuf_vector_mat_mul_trans(ctx, enc, nullptr, &pipe, tx, tW, ty, 0);
    CHECK(c1.is_valid(), "matvec with encoder enqueues");
    auto y = download(ctx, ty);
    float err = 0;
    for (int r = 0; r < R; ++r) err = fmaxf(err, fabsf(y[r] - yref[r]));
    CHECK(err < 1e-4f, "matvec result (max err %g)", err);

    // cached shapes are validated (th.cpp:99-124): a differently shaped A must be rejected
    TensorBuffer tbad(TensorShape{0, 0, 1, C / 2}, TensorType_F32, ctx);
    CHECK(!cmdbuf_vector_mat_mul_trans(ctx, enc, nullptr, &pipe, tbad, tW, ty, 0).is_valid(), "validator rejects A.c != B.c");

    // deferred command buffer: no encoder / pass -> nothing runs until queue_submit
    thk_memset(ctx, ty.gpu, 0, ty.get_size_bytes());
    CommandBuffer cd = cmdbuf_vector_mat_mul_trans(ctx, nullptr, nullptr, nullptr, tx, tW, ty, 0);
    CHECK(cd.is_valid(), "deferred command buffer is valid");
    auto y0 = download(ctx, ty);
    CHECK(y0[0] == 0.f && y0[R - 1] == 0.f, "deferred command buffer did not execute early");
    CHECK(queue_submit(ctx, cd), "queue_submit");
    auto y1 = download(ctx, ty);
    err = 0;
    for (int r = 0; r < R; ++r) err = fmaxf(err, fabsf(y1[r] - yref[r]));
    CHECK(err < 1e-4f, "deferred matvec result (max err %g)", err);

    // ---- rms_norm + gain + shape views (reference mutates .shape, th-llama.cpp:317-319) ----
    std::vector<float> a(2 * C), g(C);
    for (int i = 0; i < 2 * C; ++i) a[i] = cosf(0.05f * i) * 3.f;
    for (int i = 0; i < C; ++i) g[i] = 1.f + 0.001f * i;
    TensorBuffer ta(a.data(), TensorShape{0, 0, 2, C}, TensorType_F32, false, ctx, ctx);
    TensorBuffer tg(g.data(), TensorShape{0, 0, 1, C}, TensorType_F32, false, ctx, ctx);
    CHECK(cmdbuf_rms_norm(ctx, enc, nullptr, nullptr, ta).is_valid(), "rms_norm");
    CHECK(cmdbuf_row_element_multiply(ctx, enc, nullptr, nullptr, ta, tg).is_valid(), "row_element_multiply");
    auto an = download(ctx, ta);
    for (int row = 0; row < 2; ++row) {
        double ss = 0;
        for (int i = 0; i < C; ++i) ss += (double)a[row * C + i] * a[row * C + i];
        const float inv = 1.0f / sqrtf((float)(ss / C) + 1e-6f);
        float e = 0;
        for (int i = 0; i < C; ++i) e = fmaxf(e, fabsf(an[row * C + i] - a[row * C + i] * inv * g[i]));
        CHECK(e < 1e-4f, "rms_norm*gain row %d (max err %g)", row, e);
    }
    ta.shape = TensorShape{0, 2, 8, C / 8};                       // view as [tokens=2][heads=8][dim]
    CHECK(ta.get_size_bytes() == (size_t)2 * C * 4, "view keeps the byte size");
    ta.reset_shape();
    CHECK(ta.shape == ta.originalShape, "reset_shape");
    TensorBuffer tg2(TensorShape{0, 0, 2, C}, TensorType_F32, ctx);
    CHECK(!cmdbuf_row_element_multiply(ctx, enc, nullptr, nullptr, ta, tg2).is_valid(), "validator: gain must have one row");
    CHECK(!cmdbuf_addition(ctx, enc, nullptr, nullptr, ta, tg, ta).is_valid(), "validator: addition shapes must match");

    // ---- whole model through the C++ entry points: synthetic weights, both evaluation paths ----
    auto m = create_synthetic_llama(ctx, ctx, /*vocab*/ 512, /*embd*/ 512, /*mult*/ 256, /*head*/ 8, /*layer*/ 2, /*ctx*/ 64, 0x7B5EEDull);
    CHECK(m && m->decoder && m->n_ff == 1536, "create_synthetic_llama");
    if (m) {
        std::vector<tk_llama_token> prompt = {1, 17, 400, 33};
        std::vector<float> logits_fused, logits_graph;
        tk_llama_token tf = -1, tg_ = -1;
        for (size_t i = 0; i < prompt.size(); ++i) tf = th_eval_gpu(ctx, ctx, m, &prompt[i], 1, (int)i);
        logits_fused = m->lastLogits;
        CHECK(m->gpuLaunches == 1, "fused path: one launch per token (got %lld)", (long long)m->gpuLaunches);
        m->evalPath = EvalPath_OpGraph;
        for (size_t i = 0; i < prompt.size(); ++i) tg_ = th_eval_gpu(ctx, ctx, m, &prompt[i], 1, (int)i);
        logits_graph = m->lastLogits;
        CHECK(m->gpuLaunches == 24 * 2 + 4, "op graph: 24 launches per layer + 4 (got %lld)", (long long)m->gpuLaunches);
        float md = 0, mx = 0;
        for (size_t i = 0; i < logits_fused.size(); ++i) { md = fmaxf(md, fabsf(logits_fused[i] - logits_graph[i])); mx = fmaxf(mx, fabsf(logits_graph[i])); }
        CHECK(tf == tg_ && tf >= 0, "greedy token equal on both paths (%d vs %d)", tf, tg_);
        CHECK(md / mx < 1e-4f, "fused vs op-graph logits (rel %g)", md / mx);
        tk_llama_token bad = 100000;
        CHECK(th_eval_gpu(ctx, ctx, m, &bad, 1, 0) == -1, "token outside the vocabulary is rejected");
        CHECK(th_eval_gpu(ctx, ctx, m, &prompt[0], 1, 64) == -1, "full context is rejected");
    }
    m.reset();
    thk_destroy(ctx);
    if (g_fail == 0) printf("PASS th_api_test\n");
    return g_fail ? 1 : 0;
}
