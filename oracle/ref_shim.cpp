// ref_shim.cpp -- extern "C" window onto the REAL reference host code (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libthref_host.so).  TEST
// INFRASTRUCTURE ONLY: used by tests/ and tests/golden/make_golden.py to pin the oracle.
// Nothing here is reference source; it only calls reference functions.
#include "th.hpp"
#include "th-llama.hpp"
#include "th-llama-loader.hpp"
#include <webgpu/webgpu.h>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace th {
// declared at th-llama.cpp:36-47 (external linkage, no header)
tk_llama_token llama_sample_top_p_top_k(std::shared_ptr<LlamaModel> m,
        const std::vector<tk_llama_token>& last_n_tokens, int top_k, float top_p, float temp,
        float repeat_penalty, std::vector<float>& logits);
std::vector<tk_llama_token> tk_llama_tokenize(std::shared_ptr<LlamaModel> m, const std::string& text, bool add_bos);
// declared at th-llama.cpp:28-33, defined :464
tk_llama_token th_eval_gpu(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, const tk_llama_token* tokens,
                           int n_tokens, int n_past);
}

struct RefModel { std::shared_ptr<th::LlamaModel> m; };

extern "C" {

float ref_fp16_to_fp32(uint16_t h) { return th::ggml_compute_fp16_to_fp32(h); }     // th.cpp:312
uint16_t ref_fp32_to_fp16(float f) { return th::ggml_compute_fp32_to_fp16(f); }     // th.cpp:335

// greedy branch (temp <= 0) of the reference sampler, th-llama.cpp:826-838
int ref_greedy(const float* logits, int n) {
    auto m = std::make_shared<th::LlamaModel>();
    m->n_vocab = n;
    std::vector<float> l(logits, logits + n);
    return th::llama_sample_top_p_top_k(m, {}, 40, 0.95f, 0.0f, 1.10f, l);
}

// full reference loader on a ggjt file (th-llama-loader.cpp:485), against the host-memory stub
void* ref_load(const char* path) {
    auto m = th::load_llama_file(thstub_device(), thstub_queue(), path);
    if (!m) return nullptr;
    return new RefModel{m};
}
void ref_free(void* h) { delete (RefModel*)h; }

int ref_hparams(void* h, int32_t* out9) {
    auto& m = ((RefModel*)h)->m;
    out9[0] = m->n_vocab; out9[1] = m->n_embd; out9[2] = m->n_mult; out9[3] = m->n_head;
    out9[4] = m->n_layer; out9[5] = m->n_rot; out9[6] = m->f16; out9[7] = m->n_ctx;
    out9[8] = (int32_t)m->layers.size();
    return 0;
}

static th::TensorBuffer* find(th::LlamaModel* m, const std::string& name) {
    if (name == "tok_embeddings.weight") return &m->tok_embeddings;
    if (name == "norm.weight") return &m->norm;
    if (name == "output.weight") return &m->outputMat;
    if (name == "output.weight-split1") return &m->outputMatSplit1;
    if (name == "output.weight-split2") return &m->outputMatSplit2;
    if (name.rfind("layers.", 0) == 0) {
        size_t dot = name.find('.', 7);
        int l = std::stoi(name.substr(7, dot - 7));
        if (l < 0 || l >= (int)m->layers.size()) return nullptr;
        std::string s = name.substr(dot + 1);
        th::LlamaLayer& L = m->layers[l];
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
    }
    return nullptr;
}

// -> pointer to the tensor bytes as the reference stored them (stub "GPU" buffer, or cpuBackup)
const void* ref_tensor(void* h, const char* name, int* type, int64_t* shape4, int64_t* nbytes) {
    th::TensorBuffer* t = find(((RefModel*)h)->m.get(), name);
    if (!t || !t->is_valid()) return nullptr;
    *type = (t->type == th::TensorType_F16) ? 1 : 0;
    // originalShape: layer 0's w1/w2/w3 .shape is left row/col-swapped by the batch dry-run in
    // build_pipelines_llama (th-llama.cpp:427-428,444) until the next reset_layer_tensors.
    shape4[0] = t->originalShape.l; shape4[1] = t->originalShape.b; shape4[2] = t->originalShape.r; shape4[3] = t->originalShape.c;
    *nbytes = (int64_t)t->get_size_bytes();
    if (t->gpu) return thstub_buffer_data(t->gpu);
    return t->cpuBackup.data();
}

int ref_tokenize(void* h, const char* text, int add_bos, int32_t* out, int cap) {
    auto toks = th::tk_llama_tokenize(((RefModel*)h)->m, text, add_bos != 0);
    int n = (int)toks.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = toks[i];
    return n;
}

// a model that only carries a vocabulary (no file): for pinning the tokenizer on arbitrary vocabularies
void* ref_vocab_model(const char* const* tokens, const int32_t* lengths, const float* scores, int n) {
    auto m = std::make_shared<th::LlamaModel>();
    m->n_vocab = n;
    m->vocab.id_to_token.resize(n);
    for (int i = 0; i < n; ++i) {
        std::string t(tokens[i], (size_t)lengths[i]);
        m->vocab.token_to_id[t] = i;
        m->vocab.id_to_token[i].tok = t;
        m->vocab.id_to_token[i].score = scores[i];
    }
    return new RefModel{m};
}

// the reference sampler (th-llama.cpp:814-907) with its mt19937 seeded explicitly
int ref_sample(int n_vocab, const float* logits, const int32_t* last_n, int n_last, int top_k, float top_p, float temp,
               float repeat_penalty, uint32_t seed) {
    auto m = std::make_shared<th::LlamaModel>();
    m->n_vocab = n_vocab;
    m->rng.seed(seed);
    std::vector<float> l(logits, logits + n_vocab);
    std::vector<th::tk_llama_token> last(last_n, last_n + (last_n ? n_last : 0));
    return th::llama_sample_top_p_top_k(m, last, top_k, top_p, temp, repeat_penalty, l);
}

uint64_t ref_dispatch_count(void) { return thstub_dispatch_count(); }

// Run the reference's own th_eval_gpu (th-llama.cpp:464-660) on the stub and return the command stream it encoded:
// a JSON document {"buffers": {id: name}, "commands": [...]}.  The logits are whatever the stub buffers hold (zeros:
// no shader runs) -- only the ENCODING is observed.  The returned pointer is valid until the next call.
const char* ref_trace_eval(void* h, const int32_t* tokens, int n_tokens, int n_past) {
    static std::string out;
    auto& m = ((RefModel*)h)->m;
    std::string names = "{";
    auto collect = [&]() {
    auto add = [&](const std::string& name, th::TensorBuffer& t) {
        if (!t.gpu) return;
        const std::string key = "\"" + std::to_string(thstub_buffer_id(t.gpu)) + "\": ";
        if (names.find(key) != std::string::npos) return;
        if (names.size() > 1) names += ", ";
        names += key + "\"" + name + "\"";
    };
    add("tok_embeddings", m->tok_embeddings); add("norm", m->norm); add("output", m->outputMat);
    add("output-split1", m->outputMatSplit1); add("output-split2", m->outputMatSplit2);
    add("out", m->out); add("outScratch", m->outScratch); add("resultBuffer", m->resultBuffer);
    for (int i = 0; i < th::LlamaModel::nInpBuffers; ++i) add("inp" + std::to_string(i), m->inp[i]);
    for (int i = 0; i < th::LlamaModel::nSplitScratch; ++i) add("splitScratch" + std::to_string(i), m->splitScratch[i]);
    add("ffWorking0", m->ffWorking[0]); add("ffWorking1", m->ffWorking[1]);
    add("working_key_cache", m->working_key_cache); add("working_val_cache", m->working_val_cache);
    if (m->networkUniforms) {
        const std::string key = "\"" + std::to_string(thstub_buffer_id(m->networkUniforms)) + "\": ";
        if (names.find(key) == std::string::npos) { if (names.size() > 1) names += ", "; names += key + "\"networkUniforms\""; }
    }
    for (size_t l = 0; l < m->layers.size(); ++l) {
        th::LlamaLayer& L = m->layers[l];
        const std::string p = "layers." + std::to_string(l) + ".";
        add(p + "attention_norm", L.attention_norm); add(p + "wq", L.wq); add(p + "wk", L.wk); add(p + "wv", L.wv);
        add(p + "wo", L.wo); add(p + "ffn_norm", L.ffn_norm); add(p + "w1", L.w1); add(p + "w2", L.w2); add(p + "w3", L.w3);
        add(p + "key_cache", L.key_cache); add(p + "value_cache", L.value_cache);
    }
    };
    collect();
    thstub_trace_begin();
    std::vector<th::tk_llama_token> toks(tokens, tokens + n_tokens);
    th::th_eval_gpu(thstub_device(), thstub_queue(), m, toks.data(), n_tokens, n_past);
    std::string lines = thstub_trace_end();
    collect();                      // th_eval_gpu re-creates m->out (th-llama.cpp): name the buffers that exist now as well
    names += "}";
    out = "{\"buffers\": " + names + ", \"commands\": [";
    size_t pos = 0;
    bool first = true;
    while (pos < lines.size()) {
        size_t nl = lines.find('\n', pos);
        if (nl == std::string::npos) nl = lines.size();
        if (nl > pos) { if (!first) out += ", "; out += lines.substr(pos, nl - pos); first = false; }
        pos = nl + 1;
    }
    out += "]}";
    return out.c_str();
}

}  // extern "C"
