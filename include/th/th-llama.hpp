// th-llama.hpp -- LLaMA model graph on the th:: op surface (mirrors kayvr/token-hawk th-llama.hpp).
// Same structs and entry points; two evaluation paths behind th_eval_gpu:
//   EvalPath_OpGraph : the reference's own graph, op for op (build_layer_cmdbuf, th-llama.cpp:270-452)
//   EvalPath_Fused   : one persistent sm_100a kernel per token (thk_decoder_step) -- the default
#pragma once

#include <array>
#include <functional>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "th/th.hpp"

namespace th {

static const bool kSplitFinalMultiply = false;   // no 256 MB buffer cap on CUDA (th-llama.hpp:20, th.cpp:3793)
static const bool kUseGpuEmbeddingSelection = true;   // embedding row gathered on the device (th-llama.hpp:21)

typedef int tk_llama_token;

enum EvalPath { EvalPath_Fused = 0, EvalPath_OpGraph = 1 };

struct LlamaLayer {            // th-llama.hpp:37-55
    int64_t index{};
    TensorBuffer attention_norm{};
    TensorBuffer wq{}, wk{}, wv{}, wo{};
    TensorBuffer ffn_norm{};
    TensorBuffer w1{}, w2{}, w3{};
    TensorBuffer key_cache{};      // [n_ctx][n_head][head_dim] f32 -- op-graph path (reference layout)
    TensorBuffer value_cache{};
    TensorBuffer key_cache_hpd{};  // [n_head][n_ctx][head_dim] f32 (f16 when LlamaModel::kvF16) -- fused path (no per-token transposes)
    TensorBuffer value_cache_hpd{};
};

struct LlamaLayerComputePipeline {   // th-llama.hpp:57-78
    ComputePipeline p01{true}, p02{true}, p03_mm{true}, p03_mm_reduce{true}, p04_rope{true}, p05_trans{true},
        p06_mm{true}, p07_softmax{true}, p08_mm{true}, p09_t{true}, p10_mm{true}, p10_mm_reduce{true}, p11_add{true},
        p12_rms{true}, p13_norm{true}, p14_mm{true}, p15_silu{true}, p16_hadamard{true}, p17_mm{true}, p18_add{true};
};
struct LlamaFinalComputePipeline {   // th-llama.hpp:80-85
    ComputePipeline p01{true}, p02{true}, p03{true}, p03_reduce{true};
};

struct LlamaVocab {                  // th-llama.hpp:87-98
    using id = int32_t;
    using token = std::string;
    struct token_score { token tok; float score; };
    std::unordered_map<token, id> token_to_id;
    std::vector<token_score> id_to_token;
};

struct LlamaModel {                  // th-llama.hpp:100-177
    std::mt19937 rng{};
    int32_t n_vocab = 32000, n_ctx = 512, n_embd = 4096, n_mult = 256, n_head = 32, n_layer = 32, n_rot = 64;
    int32_t n_batch = 128;           // tokens per batched-prompt pass (reference: 8, th-llama.cpp:14)
    int32_t f16 = 1;
    int32_t n_ff = 11008;            // derived at load (th-llama-loader.cpp:349); no 11008 assert here

    TensorBuffer tok_embeddings{};   // f16 on the device (the reference keeps an f32 CPU copy)
    std::vector<LlamaLayer> layers{};
    LlamaLayerComputePipeline ps{}, pb{};
    LlamaFinalComputePipeline pfs{}, pfb{};
    TensorBuffer norm{};
    TensorBuffer outputMat{};
    TensorBuffer out{};
    TensorBuffer outScratch{};
    static const int nInpBuffers = 7;
    TensorBuffer inp[nInpBuffers]{};
    TensorBuffer ffWorking[2]{};
    TensorBuffer working_key_cache{};
    TensorBuffer working_val_cache{};
    TensorBuffer resultBuffer{};     // cpuBackup only: pinned host staging for the logits
    WGPUBuffer networkUniforms{};
    std::array<WGPUBuffer, 5> dimsUniforms{};
    LlamaVocab vocab{};

    std::function<void(std::string, std::string)> onNewToken;
    std::function<void(std::string)> onInferenceComplete;
    std::function<void(std::string)> onError;

    bool loadFailed = false;
    std::unordered_map<std::string, TensorBuffer> loadedMapping;

    std::vector<tk_llama_token> embd_inp{}, embd{};
    int n_past = 0;
    int n_consumed = 0;
    tk_llama_token lastGeneratedToken{};
    std::string generatedMessage;
    std::vector<tk_llama_token> last_n_tokens{};

    // ---- CUDA engine state (no reference analogue) ----
    WGPUDevice device{};
    EvalPath evalPath = EvalPath_Fused;
    // The two evaluation paths keep the KV cache in different layouts ([pos][head][dim] as in th-llama-loader.cpp:335 for
    // the op graph, [head][pos][dim] for the fused kernel).  Number of leading positions valid in each; th_eval_gpu
    // copies the missing rows across before it evaluates, so the paths can be mixed freely within one context.
    int32_t kvValidPhd = 0, kvValidHpd = 0;
    // f16 KV for the fused path (no reference analogue -- the reference's cache is f32, th-llama-loader.cpp:335):
    // the [head][pos][dim] cache holds f16, K (after RoPE) and V rounded to nearest-even when appended.  The op graph's
    // [pos][head][dim] cache stays f32; rows crossing between the layouts are rounded / widened (thk_kv_*_hpd_f16).
    bool kvF16 = false;
    float samplerTemp = 0.0f;        // reference hard-codes 0.8 (th-llama.cpp:721); greedy is the parity mode
    thk_decoder* decoder = nullptr;
    int32_t* d_token = nullptr;      // device: token id in, greedy id out
    int32_t* d_next = nullptr;
    float* pinnedLogits = nullptr;
    std::vector<float> lastLogits;
    int64_t gpuLaunches = 0;         // kernels launched by the last th_eval_gpu call
    bool batchPrefill = true;        // n_tokens > 1: one batched pass (tensor-core matmuls) instead of n single-token steps
    int32_t tp_rank = 0, tp_size = 1; // tensor parallel: this model holds rank tp_rank's shard (fused path only)

    ~LlamaModel();
};

static const size_t kLlamaUniformsSize = 32;
typedef thk_network_uniforms LlamaNetworkUniforms;    // th-llama.hpp:181-192
typedef thk_dims_uniforms LlamaTensorDimsUniforms;    // th-llama.hpp:195-204
static_assert(sizeof(LlamaNetworkUniforms) == kLlamaUniformsSize, "uniform block must be 32 bytes");
static_assert(sizeof(LlamaTensorDimsUniforms) == kLlamaUniformsSize, "uniform block must be 32 bytes");

void reset_layer_tensors(LlamaLayer& layer);
void reset_working_memory_tensors(LlamaModel& m);
void build_pipelines_llama(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m);
void build_layer_cmdbuf(WGPUDevice device, WGPUCommandEncoder encoder, std::shared_ptr<LlamaModel> m, LlamaLayer& l,
                        LlamaLayerComputePipeline& p, int n_tokens, int n_past);
void build_final_compute_cmdbuf(WGPUDevice device, WGPUCommandEncoder encoder, std::shared_ptr<LlamaModel> m,
                                LlamaFinalComputePipeline& p, int n_tokens);
// th-llama.cpp:464-660.  Returns the sampled token (greedy), or -1 on error.  n_tokens > 1 is
// evaluated as n_tokens consecutive single-token steps (logits of the last one are kept).
tk_llama_token th_eval_gpu(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, const tk_llama_token* tokens,
                           int n_tokens, int n_past);
// The two halves of th_eval_gpu (enqueue / wait + read back + sample), exposed so one host thread can drive
// several tensor-parallel ranks: launch all of them, then finish all of them.
int th_eval_gpu_launch(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, const tk_llama_token* tokens, int n_tokens,
                       int n_past);
tk_llama_token th_eval_gpu_finish(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m);
// th-llama.cpp:814-907: greedy for temp <= 0, else temperature / repetition penalty / top-k / top-p sampling with m->rng
tk_llama_token llama_sample_top_p_top_k(std::shared_ptr<LlamaModel> m, const std::vector<tk_llama_token>& last_n_tokens, int top_k,
                                        float top_p, float temp, float repeat_penalty, std::vector<float>& logits);
// n_new greedy tokens after feeding `prompt` one token at a time (do_inference's sync loop,
// th-llama.cpp:199-238, without tokenizer / printing)
std::vector<tk_llama_token> generate_greedy(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m,
                                            const std::vector<tk_llama_token>& prompt, int n_new);

// the sampler on a bare logits vector (what llama_sample_top_p_top_k runs after picking the last n_vocab logits)
tk_llama_token llama_sample_logits(std::mt19937& rng, const float* logits, int n_logits, const std::vector<tk_llama_token>& last_n_tokens,
                                   int top_k, float top_p, float temp, float repeat_penalty);
// tokenizer (th-llama.cpp:909-1107): SentencePiece-style merges driven by the vocabulary scores, byte fallback (id = byte + 3)
std::vector<tk_llama_token> tk_llama_tokenize(const LlamaVocab& vocab, const std::string& text, bool add_bos);
std::vector<tk_llama_token> tk_llama_tokenize(std::shared_ptr<LlamaModel> m, const std::string& text, bool add_bos);
int tk_llama_tokenize(std::shared_ptr<LlamaModel> m, const char* text, tk_llama_token* tokens, int n_max_tokens, bool add_bos);
const char* tk_llama_token_to_str(std::shared_ptr<LlamaModel> m, tk_llama_token token);
tk_llama_token tk_llama_token_bos();
tk_llama_token tk_llama_token_eos();
// do_inference (th-llama.cpp:111-168), synchronous: tokenise, feed the prompt, generate until EOS / max_new_tokens /
// full context; streams pieces through m->onNewToken and returns the generated text
std::string do_inference(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m, std::string prompt, int max_new_tokens);

}  // namespace th
