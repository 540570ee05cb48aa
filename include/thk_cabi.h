/*
 * thk_cabi.h -- C ABI of libthk_sm100a.so: the hand-written sm_100a CUDA kernels that replace
 * TokenHawk's WebGPU/WGSL backend for the single-token LLaMA decode path.
 *
 * Boundary rules
 *   - extern "C", plain pointers and sizes only; every data pointer is a DEVICE pointer unless
 *     the parameter name says host.  No torch / STL types.
 *   - every call is stream-ordered on the context's stream and returns immediately; only
 *     thk_sync / thk_download block.
 *   - return value: 0 = ok, negative = THK_E_*; thk_last_error() gives the message of the last
 *     failing call on the calling thread.  Nothing falls back to the CPU: without a CUDA device
 *     thk_init fails with THK_E_CUDA.
 *   - kernels never allocate; scratch is owned by the context / decoder objects.
 *
 * Each entry cites the reference interface it replaces (file:line under kayvr/token-hawk).
 * The reference-side binding is shown in INTEGRATION.md.
 */
#ifndef THK_CABI_H
#define THK_CABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THK_OK            0
#define THK_E_INVALID    -1   /* bad shape / dtype / alignment / null pointer (validators th.cpp:541-608 etc.) */
#define THK_E_CUDA       -2   /* CUDA runtime error (message has cudaGetErrorString) */
#define THK_E_UNSUPPORTED -3  /* valid request the kernels do not implement */
#define THK_E_NCCL       -4
#define THK_E_TIMEOUT    -5   /* in-kernel watchdog fired (grid barrier or mbarrier wait) */

typedef struct thk_ctx thk_ctx;

/* ---- device / queue: replaces WGPUDevice + WGPUQueue creation (cli/main.cpp:93-128) ---- */
int  thk_init(int device_ordinal, thk_ctx** out);
/* same, but issue all work on an existing cudaStream_t (e.g. torch's current stream) */
int  thk_init_on_stream(int device_ordinal, void* cuda_stream, thk_ctx** out);
int  thk_destroy(thk_ctx* ctx);
const char* thk_last_error(void);
const char* thk_version(void);
int  thk_device_info(thk_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
void* thk_stream(thk_ctx* ctx);

/* ---- buffers: wgpuDeviceCreateBuffer / wgpuBufferRelease (th.cpp:210-221, th.hpp:110-115) ---- */
int  thk_malloc(thk_ctx* ctx, size_t bytes, void** dptr);
int  thk_free(thk_ctx* ctx, void* dptr);
int  thk_memset(thk_ctx* ctx, void* dptr, int value, size_t bytes);
/* wgpuQueueWriteBuffer (th.cpp:205-208, th-llama.cpp:484-584) */
int  thk_upload(thk_ctx* ctx, void* dst, size_t dst_offset, const void* host_src, size_t bytes);
/* wgpuBufferMapAsync + GetConstMappedRange + Unmap (th-llama.cpp:666-706); blocks until done */
int  thk_download(thk_ctx* ctx, void* host_dst, const void* src, size_t src_offset, size_t bytes);
/* wgpuCommandEncoderCopyBufferToBuffer (th-llama.cpp:337-338,412,450,647) */
int  thk_copy(thk_ctx* ctx, void* dst, size_t dst_offset, const void* src, size_t src_offset, size_t bytes);
/* wgpuQueueSubmit + wait (th-llama.cpp:640, 700-706) */
int  thk_sync(thk_ctx* ctx);
/* pinned host memory for the logits read-back buffer (resultBuffer, th-llama-loader.cpp:363) */
int  thk_host_alloc(thk_ctx* ctx, size_t bytes, void** hptr);
int  thk_host_free(thk_ctx* ctx, void* hptr);

/* ---- mapping device memory across ranks (tensor parallel wiring; no reference analogue) ----
 * thk_ipc_export/import wrap cudaIpcGetMemHandle/OpenMemHandle (64-byte handle) for one-process-per-GPU
 * launches; thk_enable_peer_access is the same-process equivalent. */
int  thk_ipc_export(thk_ctx* ctx, void* dptr, unsigned char* handle64);
int  thk_ipc_import(thk_ctx* ctx, const unsigned char* handle64, void** dptr);
int  thk_ipc_close(thk_ctx* ctx, void* dptr);
int  thk_enable_peer_access(thk_ctx* ctx, thk_ctx* peer);

/* ---- the reference's 32-byte uniform blocks (th-llama.hpp:181-217), device resident ---- */
typedef struct { uint32_t n_past, n_tokens; float pad2, pad3; uint32_t pad4[4]; } thk_network_uniforms;
typedef struct { uint32_t A_B, A_M, A_N; float scale; uint32_t B_B, B_M, B_N; float offset; } thk_dims_uniforms;

/* ---- op-level kernels: one per cmdbuf_* op of th.hpp:302-452 ----
 * Shapes are passed by value for validation and grid sizing exactly as the reference derives its
 * dispatch counts from TensorBuffer::shape at encode time; where the reference shader reads its
 * dimensions from a uniform buffer, `uniforms` (device pointer, may be NULL) is read by the kernel
 * instead of the by-value copy, so a recorded command stream can be replayed with new n_past. */

/* cmdbuf_vector_mat_mul_trans, th.cpp:3046-3139 (shader 2839-2892): c[b][r] = sum_k a[b][k]*B[b][r][k].
 * a_offset_bytes mirrors aOffset.  B is f16 (b_is_f16=1) or f32.  Needs C % 8 == 0 for f16. */
int thk_vector_mat_mul_trans(thk_ctx* ctx, const float* a, size_t a_offset_bytes, const void* B, float* c,
                             int64_t R, int64_t C, int64_t batch, int b_is_f16);
/* cmdbuf_vector_multi_mat_mul_split_trans + cmdbuf_vector_reduce, th.cpp:3795-3912, 4042-4127:
 * logits over nsplit column-split buffers B[i] of shape [R, C/nsplit].  Computes the full dot
 * product (the reference's reduce drops part of it). */
int thk_vector_multi_mat_mul_split_trans(thk_ctx* ctx, const float* a, size_t a_offset_bytes,
                                         const void* const* B_splits_host_array, int nsplit, float* c,
                                         float* scratch, int64_t R, int64_t C_total, int b_is_f16);
/* cmdbuf_vector_reduce, th.cpp:4042-4127: a += b over n (full coverage) */
int thk_vector_reduce(thk_ctx* ctx, float* a, const float* b, int64_t n);
/* cmdbuf_rms_norm, th.cpp:1229-1296 (shader 1153-1200), eps 1e-6, in place */
int thk_rms_norm(thk_ctx* ctx, float* x, int64_t rows, int64_t N);
/* cmdbuf_row_element_multiply, th.cpp:1368-1449 (shader 1298-1315) */
int thk_row_element_multiply(thk_ctx* ctx, float* x, const float* gain, int64_t rows, int64_t N);
/* cmdbuf_RoPE, th.cpp:1536-1616 (shader 1457-1492); x viewed [n_tokens][n_head][head_dim];
 * n_past is read from uniforms->n_past (device) */
int thk_rope(thk_ctx* ctx, float* x, int64_t n_tokens, int64_t n_head, int64_t head_dim,
             const thk_network_uniforms* uniforms);
/* cmdbuf_transpose, th.cpp:1036-1151 (shader 876-912): zy [B,M,N]->[M,B,N]; else [B,M,N]->[B,N,M].
 * With uniforms: B=A_B, M=A_M, N=A_N are read on the device. */
int thk_transpose(thk_ctx* ctx, const float* in, float* out, int64_t B, int64_t M, int64_t N, int zy,
                  const thk_dims_uniforms* uniforms);
/* cmdbuf_mat_mul, th.cpp:750-861 (shader 396-539): batched C = A*B or A*B^T, optional scale.
 * With uniforms: M=A_M, N=B_N, K=A_N, scale are read on the device and the scale is applied. */
int thk_mat_mul(thk_ctx* ctx, const float* A, const void* B, float* C, int64_t batch, int64_t M, int64_t K,
                int64_t N, int transposeB, int b_is_f16, const thk_dims_uniforms* uniforms);
/* cmdbuf_row_softmax, th.cpp:2034-2119 (shader 1865-1961); with uniforms: M=A_M, N=A_N */
int thk_row_softmax(thk_ctx* ctx, float* a, int64_t batch, int64_t M, int64_t N,
                    const thk_dims_uniforms* uniforms);
/* cmdbuf_masked_softmax, th.cpp:1774-1863: causal softmax, row i sees columns <= n_past + i
 * (n_past = N - M; the reference shader ignores n_past and is dormant) */
int thk_masked_softmax(thk_ctx* ctx, float* a, int64_t batch, int64_t M, int64_t N,
                       const thk_dims_uniforms* uniforms);
/* cmdbuf_addition th.cpp:2235-2335; cmdbuf_silu 2754-2837; cmdbuf_element_mult_in_place 2589-2677 */
int thk_addition(thk_ctx* ctx, const float* a, const float* b, float* c, int64_t n);
int thk_silu(thk_ctx* ctx, float* a, int64_t n);
int thk_element_mult_in_place(thk_ctx* ctx, float* a, const float* b, int64_t n);
/* cmdbuf_f16_f32_conversion, th.cpp:4264-4351: one f16 row -> f32 (offsets in bytes) */
int thk_f16_f32_conversion(thk_ctx* ctx, float* out, size_t out_offset_bytes, const uint16_t* in,
                           size_t in_offset_bytes, int64_t n);

/* KV rows [pos0, pos0+npos) from the op graph's [pos][head][dim] cache into the fused decoder's
 * [head][n_ctx][dim] cache (hand-over from batched prefill to single-token decode) */
int thk_kv_to_hpd(thk_ctx* ctx, const float* src_phd, float* dst_hpd, int64_t pos0, int64_t npos, int64_t n_ctx, int64_t H, int64_t D);
/* ... and back, so that the op graph / a batched pass can continue a context the fused decoder extended (one
 * authoritative cache per position: th_eval_gpu tracks which layout holds which rows) */
int thk_kv_from_hpd(thk_ctx* ctx, const float* src_hpd, float* dst_phd, int64_t pos0, int64_t npos, int64_t n_ctx, int64_t H, int64_t D);
/* the same two copies for an f16 fused-layout cache (thk_llama_dims.kv_f16): f32 rows are rounded to nearest-even on the way in,
 * widened exactly on the way out */
int thk_kv_to_hpd_f16(thk_ctx* ctx, const float* src_phd, uint16_t* dst_hpd, int64_t pos0, int64_t npos, int64_t n_ctx, int64_t H, int64_t D);
int thk_kv_from_hpd_f16(thk_ctx* ctx, const uint16_t* src_hpd, float* dst_phd, int64_t pos0, int64_t npos, int64_t n_ctx, int64_t H, int64_t D);
/* dst[i] = f16(src[i]), round to nearest even (no reference analogue: the reference only widens, th.cpp:4130-4200) */
int thk_f32_to_f16(thk_ctx* ctx, const float* src, uint16_t* dst, int64_t n);

/* ---- synthetic tensors (no reference analogue): same counter PRNG as the oracle ---- */
int thk_fill_f16(thk_ctx* ctx, uint16_t* dst, uint64_t seed, uint64_t tensor_id, int64_t rows, int64_t cols,
                 int64_t row0, int64_t col0, int64_t full_cols);
int thk_fill_gain(thk_ctx* ctx, float* dst, uint64_t seed, uint64_t tensor_id, int64_t n);
/* KV value for logical index ((pos*H + head)*D + d) of the FULL model, written into the fused
 * decoder's [head_local][pos][d] layout for heads [head0, head0+n_head_local) */
int thk_fill_kv(thk_ctx* ctx, float* dst, uint64_t seed, uint64_t tensor_id, int64_t n_pos, int64_t n_ctx,
                int64_t n_head_total, int64_t head0, int64_t n_head_local, int64_t head_dim);

/* ---- fused decode engine: th_eval_gpu for n_tokens == 1 as ONE persistent kernel ----
 * Replaces build_layer_cmdbuf x n_layer + build_final_compute_cmdbuf (th-llama.cpp:270-452,
 * 240-268, 592-640): ~900 GPU commands per token become one launch. */
typedef struct {
    int32_t n_vocab, n_embd, n_head, n_layer, n_ff, n_ctx;
    int32_t tp_rank, tp_size;   /* tensor parallel: this device owns heads/ff rows/vocab rows of rank */
    int32_t kv_f16;             /* 0: f32 KV cache (the reference's, th-llama-loader.cpp:335-339); 1: the fused layout holds f16,
                                 * K (after RoPE) and V rounded to nearest-even when appended -- half the KV bytes per token */
} thk_llama_dims;

typedef struct {                /* LlamaLayer, th-llama.hpp:37-55 (device pointers, local shards) */
    const float*    attention_norm;   /* [n_embd] f32 */
    const uint16_t* wq;               /* [n_embd/tp, n_embd] f16 rows of the local heads */
    const uint16_t* wk;
    const uint16_t* wv;
    const uint16_t* wo;               /* [n_embd, n_embd/tp] f16 (columns of the local heads) */
    const float*    ffn_norm;
    const uint16_t* w1;               /* [n_ff/tp, n_embd] */
    const uint16_t* w2;               /* [n_embd, n_ff/tp] */
    const uint16_t* w3;               /* [n_ff/tp, n_embd] */
    float*          key_cache;        /* [n_head/tp][n_ctx][head_dim] f32, or f16 when dims.kv_f16 (fused-path layout) */
    float*          value_cache;
} thk_llama_layer;

typedef struct thk_decoder thk_decoder;

/* layers: host array of n_layer structs holding device pointers.  tok_embeddings [n_vocab,n_embd]
 * f16 (replicated), norm [n_embd] f32, output [n_vocab/tp, n_embd] f16. */
int thk_decoder_create(thk_ctx* ctx, const thk_llama_dims* dims, const thk_llama_layer* layers,
                       const uint16_t* tok_embeddings, const float* norm, const uint16_t* output,
                       thk_decoder** out);
int thk_decoder_destroy(thk_decoder* dec);
/* One decode step.  token: device int32 (token id to embed).  n_past by value (also mirrored into
 * the device-resident uniforms like th-llama.cpp:479-485).  logits: device f32 [n_vocab/tp] or NULL.
 * next_token: device int32 receiving the greedy argmax over the local vocab rows (global id), and
 * next_logit its value (for the cross-rank argmax); either may be NULL. */
int thk_decoder_step(thk_decoder* dec, const int32_t* token, int32_t n_past, float* logits,
                     int32_t* next_token, float* next_logit);
/* n_steps chained greedy steps without host round trips: step i embeds the previous argmax.
 * tokens_out: device int32 [n_steps].  Single-GPU only (tp_size == 1). */
int thk_decoder_generate(thk_decoder* dec, const int32_t* first_token, int32_t n_past, int32_t n_steps,
                         int32_t* tokens_out, float* last_logits);
/* residual stream after the last step's layers (debug/parity): device f32 [n_embd] */
int thk_decoder_hidden(thk_decoder* dec, const float** hidden);
/* number of kernel launches issued by the last step / generate call (bench's gpu_launches) */
int thk_decoder_last_launches(thk_decoder* dec);
/* blocks on the stream and reports an in-kernel abort (watchdog on a barrier wait -> THK_E_TIMEOUT,
 * token id out of range -> THK_E_INVALID); the reference's validators assert instead (th-llama.cpp:606) */
int thk_decoder_check(thk_decoder* dec);
/* in-kernel timeline (no reference analogue; the reference only has wall-clock stats, th.cpp:45-87).  Only in the
 * profiling build of this library (-DTHK_PROFILE, token_hawk_b200/lib_prof); the production build returns
 * THK_E_UNSUPPORTED when asked to enable it.  Layout of host_out (u64 %globaltimer ns of the last launch):
 * [cta][phase < 256][8 marks] | [cta][4 producer stats] | [cta][64 tile-retired] | [cta][64 tile-issued] (prof_phase). */
int thk_decoder_profile(thk_decoder* dec, int enable, unsigned long long* host_out, int n);
/* tuning knobs of the persistent kernel (no reference analogue; scripts/tune.py sweeps them on the GPU):
 *   "l2_ahead_kb"  KB of a CTA's upcoming rows the producer warp asks L2 for at a phase boundary (0 = off)
 *   "prof_phase"   phase index whose per-tile issue / retire times thk_decoder_profile records
 *   "dataflow"     1: Wo -> W13 -> W2 -> next QKV synchronise through epoch-stamped vectors instead of grid barriers (tp 1)
 *   "poll_single"  1: a stale read of such a vector spins on the one stale element before re-reading everything
 * Takes effect at the next step.  Unknown key -> THK_E_INVALID. */
int thk_decoder_tune(thk_decoder* dec, const char* key, int value);
/* The static tile schedule of CTA `cta` (of `grid`) in matvec phase `phase` (0 QKV, 1 Wo, 2 W1/W3, 3 W2, 4 output), computed
 * on the host by the code the kernel runs: out = {C, KT, CT, paired, n_groups, then (segment, first row, rows) per row group}.
 * No reference analogue (the reference dispatches one workgroup per row, th.cpp:3046-3139); used by the CPU tests. */
int thk_decoder_plan(const thk_llama_dims* dims, int phase, int grid, int cta, int32_t* out, int cap);
/* tensor-parallel wiring: peer pointers obtained by the host via CUDA IPC (or same-process P2P).
 * peer_bufs[r] / peer_flags[r] are device-visible addresses of rank r's exchange buffer / flag
 * array as returned by thk_decoder_exchange_info on that rank. */
int thk_decoder_exchange_info(thk_decoder* dec, void** buf, size_t* buf_bytes, void** flags, size_t* flag_bytes);
int thk_decoder_set_peers(thk_decoder* dec, void* const* peer_bufs, void* const* peer_flags, int n);

/* ---- batched-prompt path (tensor cores): Y[M,N] = X[M,K] * W[N,K]^T, W f16, X/Y f32 ----
 * replaces cmdbuf_mat_mul with an f16 B operand on the n_tokens > 1 path (th-llama.cpp:308-310,
 * 404, 429-430, 444). tcgen05 + TMEM + TMA; X is split into f16 hi + lo terms so the result keeps
 * f32-activation accuracy.  N % 32 == 0 and K % 64 == 0; N % 128 == 0 (every LLaMA matrix) takes 128 x 128 tiles with
 * split-K over ~one CTA per SM, the K slices added in slice order (deterministic).  Launches on the context's stream. */
int thk_gemm_f16_tc(thk_ctx* ctx, const float* X, const uint16_t* W, float* Y, int64_t M, int64_t N, int64_t K);
/* sizes the context's workspace (hi/lo panels of X up to [max_M, max_K] + split-K partial tiles) once (post_load_init_model
 * allocates all working buffers at load time, th-llama-loader.cpp:330-435), so that thk_gemm_f16_tc itself never allocates */
int thk_gemm_reserve(thk_ctx* ctx, int64_t max_M, int64_t max_K);
/* blocks on the stream; THK_E_TIMEOUT if a GEMM launch hit its in-kernel watchdog */
int thk_gemm_check(thk_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* THK_CABI_H */
