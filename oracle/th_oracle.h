/*
 * th_oracle.h -- CPU oracle for the TokenHawk single-token LLaMA decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call it.  The CUDA engine never routes through this code.
 *
 * What it is: a plain-C restatement of the arithmetic the reference executes in
 * its WGSL shaders (kayvr/token-hawk, th.cpp) in the order its graph issues them
 * (th-llama.cpp:240-452).  Every function cites the reference lines it follows.
 *
 * Pinning status: the reference ships no tests, golden vectors or fixtures, and
 * its arithmetic runs only under Dawn/WebGPU, which cannot be built here
 * (SURVEY.md 8c).  The shader arithmetic is therefore PARITY UNPINNED by the
 * reference.  What IS pinned against real reference code compiled here
 * (oracle/_ref, see oracle/Makefile): the fp16<->fp32 helpers (all 65536 codes),
 * the greedy sampler branch, and the ggjt loader's tensor bytes/shapes.  The
 * shader restatement is additionally cross-checked by an independent PyTorch
 * fp32 LLaMA forward (tests/test_oracle_vs_torch.py).
 */
#ifndef TH_ORACLE_H
#define TH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar helpers (th.cpp:291-359) ---- */
float    tho_fp16_to_fp32(uint16_t h);
uint16_t tho_fp32_to_fp16(float f);

/* ---- op restatements: same names as the reference's cmdbuf_* ops ---- */
/* th.cpp:2839-2892  y[b][r] = sum_c x[b][c] * W[b][r][c]; shader summation order when C%256==0 */
void tho_vector_mat_mul_trans_f16(const float* x, const uint16_t* W, float* y,
                                  int64_t R, int64_t C, int64_t batch);
void tho_vector_mat_mul_trans_f32(const float* x, const float* W, float* y,
                                  int64_t R, int64_t C, int64_t batch);
/* th.cpp:1153-1200 */
void tho_rms_norm(float* x, int64_t rows, int64_t N);
/* th.cpp:1298-1315 */
void tho_row_element_multiply(float* x, const float* g, int64_t rows, int64_t N);
/* th.cpp:1457-1492; x viewed [n_tokens][n_head][head_dim] */
void tho_rope(float* x, int64_t n_tokens, int64_t n_head, int64_t head_dim, uint32_t n_past);
/* th.cpp:876-912; zy: [B,M,N]->[M,B,N]; yx: [B,M,N]->[B,N,M] */
void tho_transpose(const float* a, float* c, int64_t B, int64_t M, int64_t N, int zy);
/* th.cpp:420-539; C[b][m][n] = scale * sum_k A[b][m][k] * (transposeB ? B[b][n][k] : B[b][k][n]) */
void tho_mat_mul(const float* A, const void* Bm, float* Cm, int64_t batch, int64_t M, int64_t K,
                 int64_t N, int transposeB, int do_scale, float scale, int b_is_f16);
/* th.cpp:1885-1961 */
void tho_row_softmax(float* a, int64_t batch, int64_t M, int64_t N);
/* intended function of th.cpp:1619-1700 (causal mask with n_past) */
void tho_causal_softmax(float* a, int64_t batch, int64_t M, int64_t N, int64_t n_past);
/* th.cpp:2121-2149, 2680-2709, 2498-2526 */
void tho_addition(const float* a, const float* b, float* c, int64_t n);
void tho_silu(float* a, int64_t n);
void tho_element_mult_in_place(float* a, const float* b, int64_t n);
/* th.cpp:3914-3945; refbug!=0 reproduces the truncated coverage of th.cpp:3990-3996 */
void tho_vector_reduce(float* a, const float* b, int64_t n, int numSplits, int refbug);
/* th.cpp:4129-4165 */
void tho_f16_f32_conversion(float* out, const uint16_t* in, int64_t n);
/* th-llama.cpp:826-838; lowest index wins ties */
int32_t tho_greedy(const float* logits, int32_t n);

/* ---- synthetic weights: counter-based PRNG shared with the CUDA fill kernel ---- */
uint64_t tho_hash(uint64_t seed, uint64_t tensor_id, uint64_t idx);
/* f16 matrix element (row r, col c of the FULL [rows,cols] tensor) */
uint16_t tho_synth_f16(uint64_t seed, uint64_t tensor_id, uint64_t idx);
float    tho_synth_gain(uint64_t seed, uint64_t tensor_id, uint64_t idx);
void     tho_fill_f16(uint16_t* dst, uint64_t seed, uint64_t tensor_id, int64_t rows, int64_t cols,
                      int64_t row0, int64_t col0, int64_t full_cols);
void     tho_fill_gain(float* dst, uint64_t seed, uint64_t tensor_id, int64_t n);
void     tho_fill_kv(float* dst, uint64_t seed, uint64_t tensor_id, int64_t n); /* U(-1,1) f32 */

/* ---- model ---- */
typedef struct {
    int32_t n_vocab, n_embd, n_mult, n_head, n_layer, n_rot, f16;
    int32_t n_ctx;   /* not in the file; reference hard-codes 512 (th-llama.hpp:105) */
} tho_hparams;

typedef struct tho_model tho_model;

enum { THO_F32 = 0, THO_F16 = 1 };   /* ggml ftype numbering (th-llama-loader.cpp:18-19) */

int32_t    tho_n_ff(const tho_hparams* hp);                 /* th-llama-loader.cpp:349 */
tho_model* tho_model_create(const tho_hparams* hp);
void       tho_model_free(tho_model* m);
const tho_hparams* tho_model_hparams(const tho_model* m);
/* names as in the ggjt file (th-llama-loader.cpp:355-427); data is copied */
int        tho_model_set_tensor(tho_model* m, const char* name, const void* data, int ftype,
                                int64_t rows, int64_t cols);
const void* tho_model_get_tensor(const tho_model* m, const char* name, int* ftype,
                                 int64_t* rows, int64_t* cols);
int        tho_model_ready(const tho_model* m);             /* 1 when every tensor is present */
void       tho_model_fill_synthetic(tho_model* m, uint64_t seed);
void       tho_model_fill_kv_synthetic(tho_model* m, uint64_t seed, int n_positions);
void       tho_model_reset(tho_model* m);                   /* clears the KV cache */
/* caches as the reference stores them: [n_ctx][n_head][head_dim] f32 per layer */
float*     tho_model_key_cache(tho_model* m, int layer);
float*     tho_model_value_cache(tho_model* m, int layer);

/*
 * One evaluation (th_eval_gpu, th-llama.cpp:464-660): n_tokens==1 follows
 * build_layer_cmdbuf / build_final_compute_cmdbuf op for op.  n_tokens>1 is defined as
 * n_tokens sequential single-token evaluations (the reference's own batch path is dormant
 * and its mask ignores n_past); logits_out receives the LAST token's logits.
 * hidden_out (optional, n_layer+1 rows of n_embd) receives the residual stream after each
 * layer and, in the last row, the final normed vector.
 */
/* Product option, NOT the reference's (its cache is f32, th-llama-loader.cpp:335): round K (after RoPE) and V to f16, nearest
 * even, when they are appended to the cache, and tho_model_fill_kv_synthetic's values likewise.  Oracle twin of
 * thk_llama_dims.kv_f16. */
void tho_model_set_kv_f16(tho_model* m, int on);

/* Op trace of tho_eval (test infrastructure, tests/test_graph_trace.py): one line per command, reference labels and buffer
 * names; see th_oracle.c.  Not thread safe. */
void tho_trace_begin(void);
const char* tho_trace_end(void);

int tho_eval(tho_model* m, const int32_t* tokens, int n_tokens, int n_past,
             float* logits_out, float* hidden_out);

/* ---- ggjt v1 file I/O (th-llama-loader.cpp:485-635; layout in SURVEY appendix B) ---- */
tho_model* tho_load_ggjt(const char* path, int32_t n_ctx);
int        tho_write_ggjt(const tho_model* m, const char* path);

int  tho_num_threads(void);
void tho_set_num_threads(int n);   /* omp_set_num_threads: overrides OMP_NUM_THREADS (torchrun sets it to 1) */
void tho_set_strict_order(int on);  /* 1 (default): shader summation order; 0: plain sequential */

#ifdef __cplusplus
}
#endif
#endif
