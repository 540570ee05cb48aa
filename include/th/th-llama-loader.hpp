// th-llama-loader.hpp -- GGML `ggjt` v1 loader (mirrors kayvr/token-hawk th-llama-loader.hpp:8-16).
#pragma once
#include "th/th-llama.hpp"

namespace th {

// th-llama-loader.cpp:485-635.  n_ctx is not stored in the file; the reference hard-codes 512.
std::shared_ptr<LlamaModel> load_llama_file(WGPUDevice device, WGPUQueue queue, const std::string& filename, int32_t n_ctx = 512,
                                            int32_t tp_rank = 0, int32_t tp_size = 1);
// th-llama-loader.cpp:330-435: working buffers, KV caches, uniforms, pipelines, fused decoder
void post_load_init_model(WGPUDevice device, WGPUQueue queue, std::shared_ptr<LlamaModel> m);
bool load_header(LlamaModel* m, void* data, int64_t dataSize, int64_t vocabSize);
bool load_weights(LlamaModel* m, WGPUDevice device, WGPUQueue queue, void* data, int64_t dataSize, int64_t numElementsInFile,
                  int64_t originalFileOffset);
// Synthetic model with the oracle's counter PRNG; weights are generated on the device.
std::shared_ptr<LlamaModel> create_synthetic_llama(WGPUDevice device, WGPUQueue queue, int32_t n_vocab, int32_t n_embd, int32_t n_mult,
                                                   int32_t n_head, int32_t n_layer, int32_t n_ctx, uint64_t seed, int32_t tp_rank = 0,
                                                   int32_t tp_size = 1);
// Fill both KV cache layouts for positions [0, n_positions) with the oracle's synthetic values.
bool fill_kv_synthetic(std::shared_ptr<LlamaModel> m, uint64_t seed, int n_positions);

}  // namespace th
