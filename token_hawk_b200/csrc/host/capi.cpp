// capi.cpp -- extern "C" window onto the host layer (th::LlamaModel, loader, th_eval_gpu), in the
// spirit of the reference's web front-end exports (capi_* in web/main.cpp:71-179).  This is what the
// Python harness (tests, bench.py) binds with ctypes; it adds no arithmetic of its own.
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "th/th-llama-loader.hpp"

using namespace th;

namespace th { extern int64_t g_launch_count; extern std::vector<std::string>* g_op_trace; extern bool g_default_kv_f16; }

struct CapiModel { std::shared_ptr<LlamaModel> m; thk_ctx* ctx; };
static thread_local std::string g_capi_err;

extern "C" {

const char* capi_last_error(void) { return g_capi_err.empty() ? thk_last_error() : g_capi_err.c_str(); }

void* capi_device_create(int ordinal) {
    thk_ctx* ctx = nullptr;
    g_capi_err.clear();
    if (thk_init(ordinal, &ctx) != THK_OK) return nullptr;
    return ctx;
}
void capi_device_destroy(void* ctx) { thk_destroy((thk_ctx*)ctx); }

void* capi_model_synthetic(void* ctx, int n_vocab, int n_embd, int n_mult, int n_head, int n_layer, int n_ctx, uint64_t seed) {
    g_capi_err.clear();
    auto m = create_synthetic_llama((thk_ctx*)ctx, (thk_ctx*)ctx, n_vocab, n_embd, n_mult, n_head, n_layer, n_ctx, seed);
    if (!m) { if (!*thk_last_error()) g_capi_err = "create_synthetic_llama failed"; return nullptr; }
    return new CapiModel{m, (thk_ctx*)ctx};
}
void* capi_model_synthetic_tp(void* ctx, int n_vocab, int n_embd, int n_mult, int n_head, int n_layer, int n_ctx, uint64_t seed, int tp_rank,
                              int tp_size) {
    g_capi_err.clear();
    auto m = create_synthetic_llama((thk_ctx*)ctx, (thk_ctx*)ctx, n_vocab, n_embd, n_mult, n_head, n_layer, n_ctx, seed, tp_rank, tp_size);
    if (!m) { if (!*thk_last_error()) g_capi_err = "create_synthetic_llama failed"; return nullptr; }
    return new CapiModel{m, (thk_ctx*)ctx};
}
void* capi_model_load_tp(void* ctx, const char* path, int n_ctx, int tp_rank, int tp_size) {
    g_capi_err.clear();
    auto m = load_llama_file((thk_ctx*)ctx, (thk_ctx*)ctx, path, n_ctx, tp_rank, tp_size);
    if (!m) { g_capi_err = std::string("load_llama_file failed: ") + path; return nullptr; }
    return new CapiModel{m, (thk_ctx*)ctx};
}
// tensor-parallel wiring: this rank's exchange region, and the peers' regions as mapped here
int capi_exchange_info(void* h, void** buf, uint64_t* bytes) {
    auto& m = ((CapiModel*)h)->m;
    size_t n = 0;
    if (!m->decoder) return -1;
    const int rc = thk_decoder_exchange_info(m->decoder, buf, &n, nullptr, nullptr);
    *bytes = n;
    return rc;
}
int capi_set_peers(void* h, void* const* bufs, int n) {
    auto& m = ((CapiModel*)h)->m;
    return m->decoder ? thk_decoder_set_peers(m->decoder, bufs, nullptr, n) : -1;
}
// th_eval_gpu in two halves (see th-llama.hpp)
int capi_eval_launch(void* h, const int32_t* tokens, int n_tokens, int n_past) {
    CapiModel* cm = (CapiModel*)h;
    g_capi_err.clear();
    return th_eval_gpu_launch(cm->ctx, cm->ctx, cm->m, tokens, n_tokens, n_past);
}
int capi_eval_finish(void* h, float* logits_out) {
    CapiModel* cm = (CapiModel*)h;
    const int tok = th_eval_gpu_finish(cm->ctx, cm->ctx, cm->m);
    if (tok >= 0 && logits_out) memcpy(logits_out, cm->m->lastLogits.data(), sizeof(float) * cm->m->lastLogits.size());
    return tok;
}
void* capi_model_load(void* ctx, const char* path, int n_ctx) {
    g_capi_err.clear();
    auto m = load_llama_file((thk_ctx*)ctx, (thk_ctx*)ctx, path, n_ctx);
    if (!m) { g_capi_err = std::string("load_llama_file failed: ") + path; return nullptr; }
    return new CapiModel{m, (thk_ctx*)ctx};
}
void capi_model_free(void* h) { delete (CapiModel*)h; }

// out9: n_vocab n_embd n_mult n_head n_layer n_ctx n_ff has_fused eval_path
int capi_model_dims(void* h, int32_t* out9) {
    auto& m = ((CapiModel*)h)->m;
    const int32_t v[9] = {m->n_vocab, m->n_embd, m->n_mult, m->n_head, m->n_layer, m->n_ctx, m->n_ff, m->decoder ? 1 : 0, (int)m->evalPath};
    memcpy(out9, v, sizeof v);
    return 0;
}
int capi_model_tp(void* h, int32_t* out2) {
    auto& m = ((CapiModel*)h)->m;
    out2[0] = m->tp_rank; out2[1] = m->tp_size;
    return 0;
}
int capi_set_eval_path(void* h, int path) {
    auto& m = ((CapiModel*)h)->m;
    if (path == EvalPath_Fused && !m->decoder) { g_capi_err = "fused decoder unavailable"; return -1; }
    if (path == EvalPath_OpGraph && m->tp_size > 1) { g_capi_err = "the op graph runs unsharded only"; return -1; }
    m->evalPath = (EvalPath)path;
    return 0;
}
void capi_reset(void* h) { ((CapiModel*)h)->m->n_past = 0; }
void capi_set_batch_prefill(void* h, int on) { ((CapiModel*)h)->m->batchPrefill = on != 0; }

// th_eval_gpu: tokens (host), returns sampled (greedy) token or -1; logits_out (host, n_vocab) optional
int capi_eval(void* h, const int32_t* tokens, int n_tokens, int n_past, float* logits_out) {
    CapiModel* cm = (CapiModel*)h;
    g_capi_err.clear();
    const int tok = th_eval_gpu(cm->ctx, cm->ctx, cm->m, tokens, n_tokens, n_past);
    if (tok >= 0 && logits_out) memcpy(logits_out, cm->m->lastLogits.data(), sizeof(float) * cm->m->lastLogits.size());
    return tok;
}
int64_t capi_last_launches(void* h) { return ((CapiModel*)h)->m->gpuLaunches; }

int capi_generate(void* h, const int32_t* prompt, int n_prompt, int n_new, int32_t* out) {
    CapiModel* cm = (CapiModel*)h;
    std::vector<tk_llama_token> p(prompt, prompt + n_prompt);
    auto r = generate_greedy(cm->ctx, cm->ctx, cm->m, p, n_new);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
}

// device-resident greedy loop: n_steps chained fused steps, no host round trip per token
int capi_generate_device(void* h, int first_token, int n_past, int n_steps, int32_t* out_host, float* last_logits_host) {
    CapiModel* cm = (CapiModel*)h;
    auto& m = cm->m;
    g_capi_err.clear();
    if (!m->decoder) { g_capi_err = "fused decoder unavailable"; return -1; }
    void* d_out = nullptr;
    if (thk_malloc(cm->ctx, sizeof(int32_t) * (size_t)n_steps, &d_out)) return -1;
    int rc = thk_upload(cm->ctx, m->d_token, 0, &first_token, sizeof(int32_t));
    if (!rc) rc = thk_decoder_generate(m->decoder, m->d_token, n_past, n_steps, (int32_t*)d_out, (float*)m->out.gpu);
    if (!rc) rc = thk_download(cm->ctx, out_host, d_out, 0, sizeof(int32_t) * (size_t)n_steps);
    if (!rc && last_logits_host) rc = thk_download(cm->ctx, last_logits_host, m->out.gpu, 0, sizeof(float) * (size_t)(m->n_vocab / m->tp_size));
    if (!rc) rc = thk_decoder_check(m->decoder);
    thk_free(cm->ctx, d_out);
    m->gpuLaunches = n_steps;
    return rc;
}

// enqueue-only variants for timing on the stream (bench.py): no host sync inside
int capi_step_async(void* h, int n_past) {   // token already in m->d_token; argmax lands in m->d_next
    CapiModel* cm = (CapiModel*)h;
    auto& m = cm->m;
    return thk_decoder_step(m->decoder, m->d_token, n_past, (float*)m->out.gpu, m->d_next, nullptr);
}
int capi_set_token(void* h, int token) { CapiModel* cm = (CapiModel*)h; return thk_upload(cm->ctx, cm->m->d_token, 0, &token, sizeof(int32_t)); }
int capi_sync(void* h) { return thk_sync(((CapiModel*)h)->ctx); }
int capi_check(void* h) { auto& m = ((CapiModel*)h)->m; return m->decoder ? thk_decoder_check(m->decoder) : 0; }
void* capi_stream(void* h) { return thk_stream(((CapiModel*)h)->ctx); }

int capi_profile(void* h, int enable, unsigned long long* out, int n) {
    auto& m = ((CapiModel*)h)->m;
    return m->decoder ? thk_decoder_profile(m->decoder, enable, out, n) : -1;
}
// models created after this call keep the fused path's KV cache in f16 (1) or f32 (0, the reference's)
void capi_set_default_kv_f16(int on) { th::g_default_kv_f16 = on != 0; }

// test hook: record the labels of the commands issued through the op surface (th.cpp: g_op_trace)
void capi_trace_begin(void) {
    static std::vector<std::string> trace;
    trace.clear();
    th::g_op_trace = &trace;
}
// -> number of commands; the labels, newline separated, into buf (truncated at cap)
int capi_trace_end(char* buf, int cap) {
    std::vector<std::string>* t = th::g_op_trace;
    th::g_op_trace = nullptr;
    if (!t) return -1;
    std::string all;
    for (const std::string& l : *t) { all += l; all += '\n'; }
    if (buf && cap > 0) { const size_t n = all.size() < (size_t)cap - 1 ? all.size() : (size_t)cap - 1; memcpy(buf, all.data(), n); buf[n] = 0; }
    return (int)t->size();
}

int capi_tune(void* h, const char* key, int value) {
    auto& m = ((CapiModel*)h)->m;
    return m->decoder ? thk_decoder_tune(m->decoder, key, value) : -1;
}
int capi_fill_kv(void* h, uint64_t seed, int n_positions) { return fill_kv_synthetic(((CapiModel*)h)->m, seed, n_positions) ? 0 : -1; }

// residual stream after the last evaluated token's layers (fused path: decoder scratch; op graph: inp[6])
int capi_hidden(void* h, float* out_host) {
    CapiModel* cm = (CapiModel*)h;
    auto& m = cm->m;
    const float* src = nullptr;
    if (m->evalPath == EvalPath_Fused) { if (thk_decoder_hidden(m->decoder, &src)) return -1; }
    else src = (const float*)m->inp[6].gpu;
    return thk_download(cm->ctx, out_host, src, 0, sizeof(float) * (size_t)m->n_embd);
}

static TensorBuffer* find_tensor(LlamaModel* m, const std::string& name) {
    if (name == "tok_embeddings.weight") return &m->tok_embeddings;
    if (name == "norm.weight") return &m->norm;
    if (name == "output.weight") return &m->outputMat;
    if (name.rfind("layers.", 0) == 0) {
        const size_t dot = name.find('.', 7);
        if (dot == std::string::npos) return nullptr;
        const int l = atoi(name.substr(7, dot - 7).c_str());
        if (l < 0 || l >= (int)m->layers.size()) return nullptr;
        const std::string s = name.substr(dot + 1);
        LlamaLayer& L = m->layers[l];
        if (s == "attention_norm.weight") return &L.attention_norm;
        if (s == "attention.wq.weight") return &L.wq;
        if (s == "attention.wk.weight") return &L.wk;
        if (s == "attention.wv.weight") return &L.wv;
        if (s == "attention.wo.weight") return &L.wo;
        if (s == "ffn_norm.weight") return &L.ffn_norm;
        if (s == "feed_forward.w1.weight") return &L.w1;
        if (s == "feed_forward.w2.weight") return &L.w2;
        if (s == "feed_forward.w3.weight") return &L.w3;
        if (s == "key_cache") return &L.key_cache;
        if (s == "value_cache") return &L.value_cache;
        if (s == "key_cache_hpd") return &L.key_cache_hpd;
        if (s == "value_cache_hpd") return &L.value_cache_hpd;
    }
    return nullptr;
}
// info3: is_f16, rows(or b*r), cols; returns byte size or -1
int64_t capi_tensor_info(void* h, const char* name, int64_t* info3) {
    TensorBuffer* t = find_tensor(((CapiModel*)h)->m.get(), name);
    if (!t || !t->gpu) return -1;
    const TensorShape s = t->originalShape;
    info3[0] = t->type == TensorType_F16;
    info3[1] = (s.b ? s.b : 1) * (s.r ? s.r : 1);
    info3[2] = s.c;
    return (int64_t)(info3[1] * info3[2] * (int64_t)get_TensorType_size(t->type));
}
int capi_tensor_download(void* h, const char* name, void* out_host, int64_t bytes) {
    CapiModel* cm = (CapiModel*)h;
    TensorBuffer* t = find_tensor(cm->m.get(), name);
    if (!t || !t->gpu) return -1;
    return thk_download(cm->ctx, out_host, t->gpu, 0, (size_t)bytes);
}

// ---- sampler / tokenizer / generation driver (th-llama.cpp:111-238, 814-1107); the first three need no device ----
int capi_sample(int n_vocab, const float* logits, const int32_t* last_n, int n_last, int top_k, float top_p, float temp,
                float repeat_penalty, uint32_t seed) {
    std::mt19937 rng(seed);
    std::vector<tk_llama_token> last(last_n, last_n + (last_n ? n_last : 0));
    return llama_sample_logits(rng, logits, n_vocab, last, top_k, top_p, temp, repeat_penalty);
}
void* capi_vocab_create(const char* const* tokens, const int32_t* lengths, const float* scores, int n) {
    LlamaVocab* v = new LlamaVocab;
    v->id_to_token.resize(n);
    for (int i = 0; i < n; ++i) {
        const std::string t(tokens[i], (size_t)lengths[i]);
        v->token_to_id[t] = i;                      // later duplicates win, like the reference loader (th-llama-loader.cpp:560)
        v->id_to_token[i].tok = t;
        v->id_to_token[i].score = scores[i];
    }
    return v;
}
void capi_vocab_free(void* v) { delete (LlamaVocab*)v; }
int capi_vocab_tokenize(void* v, const char* text, int text_len, int add_bos, int32_t* out, int cap) {
    const std::vector<tk_llama_token> t = tk_llama_tokenize(*(LlamaVocab*)v, std::string(text, (size_t)text_len), add_bos != 0);
    for (size_t i = 0; i < t.size() && (int)i < cap; ++i) out[i] = t[i];
    return (int)t.size();
}
int capi_tokenize(void* h, const char* text, int add_bos, int32_t* out, int cap) {
    const std::vector<tk_llama_token> t = tk_llama_tokenize(((CapiModel*)h)->m, std::string(text), add_bos != 0);
    for (size_t i = 0; i < t.size() && (int)i < cap; ++i) out[i] = t[i];
    return (int)t.size();
}
const char* capi_token_str(void* h, int token) { return tk_llama_token_to_str(((CapiModel*)h)->m, token); }
void capi_set_sampler(void* h, float temp, uint32_t seed) {
    auto& m = ((CapiModel*)h)->m;
    m->samplerTemp = temp;
    m->rng.seed(seed);
}
// do_inference: returns the number of bytes of generated text (copied, NUL-terminated, into out when it fits), < 0 on error
int capi_inference(void* h, const char* prompt, int max_new_tokens, char* out, int cap) {
    CapiModel* cm = (CapiModel*)h;
    g_capi_err.clear();
    bool failed = false;
    cm->m->onError = [&](std::string e) { g_capi_err = e; failed = true; };
    const std::string msg = do_inference(cm->ctx, cm->ctx, cm->m, prompt, max_new_tokens);
    cm->m->onError = nullptr;
    if (failed) return -1;
    if (out && cap > 0) {
        const size_t n = std::min(msg.size(), (size_t)cap - 1);
        memcpy(out, msg.data(), n);
        out[n] = 0;
    }
    return (int)msg.size();
}

int capi_vocab_size(void* h) { return (int)((CapiModel*)h)->m->vocab.id_to_token.size(); }

}  // extern "C"
