/*
 * th_oracle.c -- CPU restatement of the TokenHawk decode path.  TEST INFRASTRUCTURE ONLY
 * (see th_oracle.h for the rules and the pinning status: shader arithmetic is PARITY
 * UNPINNED by the reference; helpers/sampler/loader are pinned against oracle/_ref).
 *
 * All arithmetic is f32 with separate multiply and add (build with -ffp-contract=off), f16
 * weights are decoded to f32 first, exactly as the WGSL does.  Summation follows the
 * shaders' order (per-thread contiguous chunk, then a pairwise tree over the 256-thread
 * workgroup) whenever the shader's own divisibility requirements hold, so that the only
 * difference to a GPU implementation is the GPU's own reduction order and libm.
 */
#include "th_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define THO_WG 256 /* workgroup size used by the 1-D reduction shaders (th.cpp:1154, 2994) */

static int g_strict_order = 1;
void tho_set_strict_order(int on) { g_strict_order = on; }

int tho_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* bench.py's CPU arms set this explicitly: torchrun exports OMP_NUM_THREADS=1 to every rank */
void tho_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * fp16 <-> fp32: the Maratyszcza bit trick the reference copied from ggml (th.cpp:291-359);
 * the WGSL twin (th.cpp:363-394) is the same computation.
 * ---------------------------------------------------------------------------------------- */
static inline float bits_to_f32(uint32_t w) { float f; memcpy(&f, &w, 4); return f; }
static inline uint32_t f32_to_bits(float f) { uint32_t w; memcpy(&w, &f, 4); return w; }

float tho_fp16_to_fp32(uint16_t h) {
    const uint32_t w = (uint32_t)h << 16;
    const uint32_t sign = w & 0x80000000u;
    const uint32_t two_w = w + w;
    const uint32_t exp_offset = 0xE0u << 23;
    const float exp_scale = bits_to_f32(0x7800000u); /* 2^-112 */
    const float normalized = bits_to_f32((two_w >> 4) + exp_offset) * exp_scale;
    const uint32_t magic_mask = 126u << 23;
    const float denormalized = bits_to_f32((two_w >> 17) | magic_mask) - 0.5f;
    const uint32_t cutoff = 1u << 27;
    const uint32_t r = sign | (two_w < cutoff ? f32_to_bits(denormalized) : f32_to_bits(normalized));
    return bits_to_f32(r);
}

uint16_t tho_fp32_to_fp16(float f) {
    const float scale_to_inf = bits_to_f32(0x77800000u);
    const float scale_to_zero = bits_to_f32(0x08800000u);
    float base = (fabsf(f) * scale_to_inf) * scale_to_zero;
    const uint32_t w = f32_to_bits(f);
    const uint32_t shl1_w = w + w;
    const uint32_t sign = w & 0x80000000u;
    uint32_t bias = shl1_w & 0xFF000000u;
    if (bias < 0x71000000u) bias = 0x71000000u;
    base = bits_to_f32((bias >> 1) + 0x07800000u) + base;
    const uint32_t bits = f32_to_bits(base);
    const uint32_t exp_bits = (bits >> 13) & 0x00007C00u;
    const uint32_t mantissa_bits = bits & 0x00000FFFu;
    const uint32_t nonsign = exp_bits + mantissa_bits;
    return (uint16_t)((sign >> 16) | (shl1_w > 0xFF000000u ? 0x7E00u : nonsign));
}

/* pairwise tree over a 256-entry workgroup array: for stride = 128..1: s[i] += s[i+stride] */
static inline float wg_tree_sum(float* s) {
    for (int stride = THO_WG / 2; stride > 0; stride >>= 1)
        for (int i = 0; i < stride; ++i) s[i] = s[i] + s[i + stride];
    return s[0];
}
static inline float wg_tree_max(float* s) {
    for (int stride = THO_WG / 2; stride > 0; stride >>= 1)
        for (int i = 0; i < stride; ++i) s[i] = s[i] > s[i + stride] ? s[i] : s[i + stride];
    return s[0];
}

/* ------------------------------------------------------------------------------------------
 * cmdbuf_vector_mat_mul_trans (th.cpp:2839-2892, host 2951-3139).
 * One workgroup per output row; thread t sums kTileSize = C/256 contiguous products
 * starting at t*kTileSize, then the tree.  The reference rejects C < 256 or C % 256 != 0
 * (th.cpp:2996-3006); for such C (only reachable through our own generalisation) the sum
 * is a plain left-to-right loop.
 * ---------------------------------------------------------------------------------------- */
static float dot_f16_row(const float* x, const uint16_t* w, int64_t C) {
    if (g_strict_order && C >= THO_WG && C % THO_WG == 0) {
        float part[THO_WG];
        const int64_t tile = C / THO_WG;
        for (int t = 0; t < THO_WG; ++t) {
            float sum = 0.0f;
            const float* xa = x + t * tile;
            const uint16_t* wb = w + t * tile;
            for (int64_t i = 0; i < tile; ++i) sum = sum + xa[i] * tho_fp16_to_fp32(wb[i]);
            part[t] = sum;
        }
        return wg_tree_sum(part);
    }
    float sum = 0.0f;
    for (int64_t i = 0; i < C; ++i) sum = sum + x[i] * tho_fp16_to_fp32(w[i]);
    return sum;
}
static float dot_f32_row(const float* x, const float* w, int64_t C) {
    if (g_strict_order && C >= THO_WG && C % THO_WG == 0) {
        float part[THO_WG];
        const int64_t tile = C / THO_WG;
        for (int t = 0; t < THO_WG; ++t) {
            float sum = 0.0f;
            for (int64_t i = 0; i < tile; ++i) sum = sum + x[t * tile + i] * w[t * tile + i];
            part[t] = sum;
        }
        return wg_tree_sum(part);
    }
    float sum = 0.0f;
    for (int64_t i = 0; i < C; ++i) sum = sum + x[i] * w[i];
    return sum;
}

void tho_vector_mat_mul_trans_f16(const float* x, const uint16_t* W, float* y, int64_t R, int64_t C,
                                  int64_t batch) {
    if (batch <= 0) batch = 1;
    for (int64_t b = 0; b < batch; ++b) {
        const float* xb = x + b * C;               /* matStrideA = C * wid.z       */
        const uint16_t* Wb = W + b * R * C;        /* matStrideB = R * C * wid.z   */
        float* yb = y + b * R;                     /* matStrideO = R * wid.z       */
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < R; ++r) yb[r] = dot_f16_row(xb, Wb + r * C, C);
    }
}
void tho_vector_mat_mul_trans_f32(const float* x, const float* W, float* y, int64_t R, int64_t C,
                                  int64_t batch) {
    if (batch <= 0) batch = 1;
    for (int64_t b = 0; b < batch; ++b) {
        const float* xb = x + b * C;
        const float* Wb = W + b * R * C;
        float* yb = y + b * R;
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < R; ++r) yb[r] = dot_f32_row(xb, Wb + r * C, C);
    }
}

/* cmdbuf_rms_norm (th.cpp:1153-1200): per row, 256 threads x N/256 contiguous squares, tree,
 * inv = 1/sqrt(sum/N + 1e-6), x *= inv. */
void tho_rms_norm(float* x, int64_t rows, int64_t N) {
    const float kEpsilon = 1e-6f;
    for (int64_t r = 0; r < rows; ++r) {
        float* row = x + r * N;
        float total;
        if (g_strict_order && N % THO_WG == 0) {
            float part[THO_WG];
            const int64_t cols = N / THO_WG;
            for (int t = 0; t < THO_WG; ++t) {
                float sum = 0.0f;
                for (int64_t i = 0; i < cols; ++i) sum = sum + row[t * cols + i] * row[t * cols + i];
                part[t] = sum;
            }
            total = wg_tree_sum(part);
        } else {
            total = 0.0f;
            for (int64_t i = 0; i < N; ++i) total = total + row[i] * row[i];
        }
        const float inv = 1.0f / sqrtf(total / (float)N + kEpsilon);
        for (int64_t i = 0; i < N; ++i) row[i] = row[i] * inv;
    }
}

/* cmdbuf_row_element_multiply (th.cpp:1298-1315) */
void tho_row_element_multiply(float* x, const float* g, int64_t rows, int64_t N) {
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < N; ++c) x[r * N + c] = x[r * N + c] * g[c];
}

/* cmdbuf_RoPE (th.cpp:1457-1492, constants 1494-1534): A viewed [z=token][y=head][x=dim];
 * for even x: theta = pow(10000, -x/kDims), p = n_past + z, rotate (x, x+1). */
void tho_rope(float* x, int64_t n_tokens, int64_t n_head, int64_t head_dim, uint32_t n_past) {
    for (int64_t z = 0; z < n_tokens; ++z) {
        const float p = (float)(n_past + (uint32_t)z);
        for (int64_t y = 0; y < n_head; ++y) {
            float* v = x + z * n_head * head_dim + y * head_dim;
            for (int64_t i = 0; i < head_dim; i += 2) {
                const float theta = powf(10000.0f, (-(float)i) / (float)head_dim);
                const float c = cosf(p * theta);
                const float s = sinf(p * theta);
                const float x0 = v[i], x1 = v[i + 1];
                v[i] = x0 * c - x1 * s;
                v[i + 1] = x0 * s + x1 * c;
            }
        }
    }
}

/* cmdbuf_transpose (th.cpp:876-912) */
void tho_transpose(const float* a, float* c, int64_t B, int64_t M, int64_t N, int zy) {
    if (B <= 0) B = 1;
    for (int64_t z = 0; z < B; ++z)
        for (int64_t y = 0; y < M; ++y)
            for (int64_t x = 0; x < N; ++x) {
                if (zy) c[y * (B * N) + z * N + x] = a[z * M * N + y * N + x];
                else    c[z * M * N + x * M + y] = a[z * M * N + y * N + x];
            }
}

/* cmdbuf_mat_mul (th.cpp:420-539; 8x8 workgroup, 1x1 tile => K consumed in chunks of 8:
 * inner `sum` over the chunk, then suma += sum; optional scale after the full sum). */
void tho_mat_mul(const float* A, const void* Bm, float* Cm, int64_t batch, int64_t M, int64_t K,
                 int64_t N, int transposeB, int do_scale, float scale, int b_is_f16) {
    if (batch <= 0) batch = 1;
    const int64_t KC = 8; /* kSharedMemDimX = kWorkgroupX * kTileSizeX (th.cpp:611-615, 402) */
    for (int64_t z = 0; z < batch; ++z) {
        const float* a = A + z * M * K;
        float* c = Cm + z * M * N;
#pragma omp parallel for schedule(static) collapse(2)
        for (int64_t m = 0; m < M; ++m) {
            for (int64_t n = 0; n < N; ++n) {
                float suma = 0.0f;
                for (int64_t k0 = 0; k0 < K; k0 += KC) {
                    float sum = 0.0f;
                    for (int64_t k = k0; k < k0 + KC; ++k) {
                        float av = 0.0f, bv = 0.0f; /* out-of-bounds tiles are zero-filled */
                        if (k < K) {
                            av = a[m * K + k];
                            const int64_t bi = z * K * N + (transposeB ? n * K + k : k * N + n);
                            bv = b_is_f16 ? tho_fp16_to_fp32(((const uint16_t*)Bm)[bi])
                                          : ((const float*)Bm)[bi];
                        }
                        sum = sum + av * bv;
                    }
                    suma = suma + sum;
                }
                if (do_scale) suma = suma * scale;
                c[m * N + n] = suma;
            }
        }
    }
}

/* cmdbuf_row_softmax (th.cpp:1885-1961): 256 threads, ceil(N/256) contiguous columns each,
 * tree max (-1e14 sentinel), exp, tree sum, divide. */
void tho_row_softmax(float* a, int64_t batch, int64_t M, int64_t N) {
    if (batch <= 0) batch = 1;
    const float kNegativeInf = -1e14f;
    int64_t cpt = N / THO_WG;
    if (N % THO_WG != 0) cpt += 1;
    for (int64_t z = 0; z < batch; ++z)
        for (int64_t y = 0; y < M; ++y) {
            float* row = a + z * M * N + y * N;
            float sh[THO_WG];
            for (int t = 0; t < THO_WG; ++t) {
                float mx = kNegativeInf;
                for (int64_t i = 0; i < cpt; ++i) {
                    const int64_t xx = t * cpt + i;
                    if (xx < N) mx = mx > row[xx] ? mx : row[xx];
                }
                sh[t] = mx;
            }
            const float row_max = wg_tree_max(sh);
            for (int t = 0; t < THO_WG; ++t) {
                float s = 0.0f;
                for (int64_t i = 0; i < cpt; ++i) {
                    const int64_t xx = t * cpt + i;
                    if (xx < N) {
                        const float e = expf(row[xx] - row_max);
                        row[xx] = e;
                        s = s + e;
                    }
                }
                sh[t] = s;
            }
            const float row_sum = wg_tree_sum(sh);
            for (int64_t xx = 0; xx < N; ++xx) row[xx] = row[xx] / row_sum;
        }
}

/* Intended function of cmdbuf_masked_softmax (th.cpp:1619-1700): row i of M query rows may
 * see columns j <= n_past + i.  (The reference shader ignores n_past and only handles N==8;
 * it is never executed -- SURVEY C9.)  Masked entries become 0. */
void tho_causal_softmax(float* a, int64_t batch, int64_t M, int64_t N, int64_t n_past) {
    if (batch <= 0) batch = 1;
    for (int64_t z = 0; z < batch; ++z)
        for (int64_t y = 0; y < M; ++y) {
            float* row = a + z * M * N + y * N;
            int64_t lim = n_past + y + 1;
            if (lim > N) lim = N;
            float mx = -1e14f;
            for (int64_t j = 0; j < lim; ++j) mx = mx > row[j] ? mx : row[j];
            float s = 0.0f;
            for (int64_t j = 0; j < lim; ++j) { row[j] = expf(row[j] - mx); s = s + row[j]; }
            for (int64_t j = 0; j < lim; ++j) row[j] = row[j] / s;
            for (int64_t j = lim; j < N; ++j) row[j] = 0.0f;
        }
}

void tho_addition(const float* a, const float* b, float* c, int64_t n) {
    for (int64_t i = 0; i < n; ++i) c[i] = a[i] + b[i];
}
void tho_silu(float* a, int64_t n) {
    for (int64_t i = 0; i < n; ++i) { const float v = a[i]; a[i] = v / (1.0f + expf(-v)); }
}
void tho_element_mult_in_place(float* a, const float* b, int64_t n) {
    for (int64_t i = 0; i < n; ++i) a[i] = a[i] * b[i];
}

/* cmdbuf_vector_reduce (th.cpp:3914-3945, host 3985-4127).  Intended: a += b over n.
 * refbug: kTileSize = (n/numSplits)/256 truncates, so each of numSplits workgroups covers
 * only 256*kTileSize of its n/numSplits elements. */
void tho_vector_reduce(float* a, const float* b, int64_t n, int numSplits, int refbug) {
    if (!refbug) { for (int64_t i = 0; i < n; ++i) a[i] = a[i] + b[i]; return; }
    int64_t split = n / numSplits;
    int64_t tile = split / THO_WG;
    if (tile == 0) tile = 1;
    for (int64_t wg = 0; wg < numSplits; ++wg)
        for (int64_t t = 0; t < THO_WG; ++t)
            for (int64_t i = 0; i < tile; ++i) {
                const int64_t xi = t * tile + wg * split + i;
                if (xi < n) a[xi] = a[xi] + b[xi];
            }
}

void tho_f16_f32_conversion(float* out, const uint16_t* in, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = tho_fp16_to_fp32(in[i]);
}

/* greedy branch of llama_sample_top_p_top_k (th-llama.cpp:826-838) */
int32_t tho_greedy(const float* logits, int32_t n) {
    float max_logit = logits[0];
    int32_t max_id = 0;
    for (int32_t i = 1; i < n; ++i)
        if (logits[i] > max_logit) { max_logit = logits[i]; max_id = i; }
    return max_id;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic weights.  Counter-based so CPU and GPU generate identical tensors without a file.
 * Matrix/embedding element: k in [-1024,1023] scaled by 2^-15 (exactly representable in f16,
 * std ~0.018).  Gains: 1 + k*2^-12, k in [-512,511].
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
uint64_t tho_hash(uint64_t seed, uint64_t tensor_id, uint64_t idx) {
    return splitmix64(splitmix64(seed ^ (tensor_id * 0xD1B54A32D192ED03ull)) + idx);
}
uint16_t tho_synth_f16(uint64_t seed, uint64_t tensor_id, uint64_t idx) {
    const int32_t k = (int32_t)(tho_hash(seed, tensor_id, idx) >> 53) - 1024;
    return tho_fp32_to_fp16((float)k * (1.0f / 32768.0f));
}
float tho_synth_gain(uint64_t seed, uint64_t tensor_id, uint64_t idx) {
    const int32_t k = (int32_t)(tho_hash(seed, tensor_id, idx) >> 54) - 512;
    return 1.0f + (float)k * (1.0f / 4096.0f);
}
void tho_fill_f16(uint16_t* dst, uint64_t seed, uint64_t tensor_id, int64_t rows, int64_t cols,
                  int64_t row0, int64_t col0, int64_t full_cols) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c)
            dst[r * cols + c] =
                tho_synth_f16(seed, tensor_id, (uint64_t)((row0 + r) * full_cols + col0 + c));
}
void tho_fill_gain(float* dst, uint64_t seed, uint64_t tensor_id, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = tho_synth_gain(seed, tensor_id, (uint64_t)i);
}
void tho_fill_kv(float* dst, uint64_t seed, uint64_t tensor_id, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const int32_t k = (int32_t)(tho_hash(seed, tensor_id, (uint64_t)i) >> 48) - 32768;
        dst[i] = (float)k * (1.0f / 32768.0f);
    }
}

/* ------------------------------------------------------------------------------------------
 * Model container (LlamaModel / LlamaLayer, th-llama.hpp:37-177).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    void* data; int ftype; int64_t rows, cols;
} tho_tensor;

enum { T_ATTN_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_PER_LAYER };

struct tho_model {
    tho_hparams hp;
    int32_t n_ff, head_dim;
    tho_tensor tok_embeddings, norm, output;
    tho_tensor* layers; /* [n_layer][T_PER_LAYER] */
    float** key_cache;  /* [n_layer] -> [n_ctx][n_head][head_dim] */
    float** value_cache;
    /* working buffers (post_load_init_model, th-llama-loader.cpp:330-363) */
    float *inp[7], *ff[2], *work_k, *work_v, *out;
    int kv_f16;         /* NOT the reference's: K (after RoPE) and V are rounded to f16 when appended (product option) */
};

int32_t tho_n_ff(const tho_hparams* hp) {
    return ((2 * (4 * hp->n_embd) / 3 + hp->n_mult - 1) / hp->n_mult) * hp->n_mult;
}

static const char* k_layer_suffix[T_PER_LAYER] = {
    "attention_norm.weight", "attention.wq.weight", "attention.wk.weight", "attention.wv.weight",
    "attention.wo.weight",   "ffn_norm.weight",     "feed_forward.w1.weight",
    "feed_forward.w2.weight", "feed_forward.w3.weight"};

tho_model* tho_model_create(const tho_hparams* hp) {
    tho_model* m = (tho_model*)calloc(1, sizeof(tho_model));
    m->hp = *hp;
    if (m->hp.n_ctx <= 0) m->hp.n_ctx = 512;
    m->n_ff = tho_n_ff(hp);
    m->head_dim = hp->n_embd / hp->n_head;
    m->layers = (tho_tensor*)calloc((size_t)hp->n_layer * T_PER_LAYER, sizeof(tho_tensor));
    m->key_cache = (float**)calloc(hp->n_layer, sizeof(float*));
    m->value_cache = (float**)calloc(hp->n_layer, sizeof(float*));
    const size_t kv = (size_t)m->hp.n_ctx * hp->n_embd;
    for (int l = 0; l < hp->n_layer; ++l) {
        m->key_cache[l] = (float*)calloc(kv, sizeof(float));
        m->value_cache[l] = (float*)calloc(kv, sizeof(float));
    }
    for (int i = 0; i < 7; ++i) m->inp[i] = (float*)calloc((size_t)hp->n_embd > (size_t)m->hp.n_ctx * hp->n_head ? (size_t)hp->n_embd : (size_t)m->hp.n_ctx * hp->n_head, sizeof(float));
    for (int i = 0; i < 2; ++i) m->ff[i] = (float*)calloc(m->n_ff, sizeof(float));
    m->work_k = (float*)calloc(kv, sizeof(float));
    m->work_v = (float*)calloc(kv, sizeof(float));
    m->out = (float*)calloc(hp->n_vocab, sizeof(float));
    return m;
}

static void free_tensor(tho_tensor* t) { free(t->data); t->data = NULL; }

void tho_model_free(tho_model* m) {
    if (!m) return;
    free_tensor(&m->tok_embeddings); free_tensor(&m->norm); free_tensor(&m->output);
    for (int i = 0; i < m->hp.n_layer * T_PER_LAYER; ++i) free_tensor(&m->layers[i]);
    for (int l = 0; l < m->hp.n_layer; ++l) { free(m->key_cache[l]); free(m->value_cache[l]); }
    for (int i = 0; i < 7; ++i) free(m->inp[i]);
    free(m->ff[0]); free(m->ff[1]); free(m->work_k); free(m->work_v); free(m->out);
    free(m->layers); free(m->key_cache); free(m->value_cache); free(m);
}

const tho_hparams* tho_model_hparams(const tho_model* m) { return &m->hp; }

static tho_tensor* find_tensor(tho_model* m, const char* name) {
    if (!strcmp(name, "tok_embeddings.weight")) return &m->tok_embeddings;
    if (!strcmp(name, "norm.weight")) return &m->norm;
    if (!strcmp(name, "output.weight")) return &m->output;
    int l = -1, off = 0;
    if (sscanf(name, "layers.%d.%n", &l, &off) >= 1 && off > 0 && l >= 0 && l < m->hp.n_layer)
        for (int k = 0; k < T_PER_LAYER; ++k)
            if (!strcmp(name + off, k_layer_suffix[k])) return &m->layers[l * T_PER_LAYER + k];
    return NULL;
}

int tho_model_set_tensor(tho_model* m, const char* name, const void* data, int ftype,
                         int64_t rows, int64_t cols) {
    tho_tensor* t = find_tensor(m, name);
    if (!t) return -1;
    if (ftype != THO_F32 && ftype != THO_F16) return -2; /* quantised: rejected (loader :595-602) */
    const size_t bytes = (size_t)rows * cols * (ftype == THO_F16 ? 2 : 4);
    free(t->data);
    t->data = malloc(bytes);
    if (data) memcpy(t->data, data, bytes);
    t->ftype = ftype; t->rows = rows; t->cols = cols;
    return 0;
}

const void* tho_model_get_tensor(const tho_model* m, const char* name, int* ftype, int64_t* rows,
                                 int64_t* cols) {
    tho_tensor* t = find_tensor((tho_model*)m, name);
    if (!t || !t->data) return NULL;
    if (ftype) *ftype = t->ftype;
    if (rows) *rows = t->rows;
    if (cols) *cols = t->cols;
    return t->data;
}

int tho_model_ready(const tho_model* m) {
    if (!m->tok_embeddings.data || !m->norm.data || !m->output.data) return 0;
    for (int i = 0; i < m->hp.n_layer * T_PER_LAYER; ++i) if (!m->layers[i].data) return 0;
    return 1;
}

float* tho_model_key_cache(tho_model* m, int layer) { return m->key_cache[layer]; }
float* tho_model_value_cache(tho_model* m, int layer) { return m->value_cache[layer]; }

void tho_model_set_kv_f16(tho_model* m, int on) { m->kv_f16 = on != 0; }

void tho_model_reset(tho_model* m) {
    const size_t kv = (size_t)m->hp.n_ctx * m->hp.n_embd * sizeof(float);
    for (int l = 0; l < m->hp.n_layer; ++l) { memset(m->key_cache[l], 0, kv); memset(m->value_cache[l], 0, kv); }
}

/* tensor ids: 0 tok_embeddings, 1 norm, 2 output, 3 + 9*l + k for layer tensors */
void tho_model_fill_synthetic(tho_model* m, uint64_t seed) {
    const int64_t E = m->hp.n_embd, V = m->hp.n_vocab, F = m->n_ff;
    tho_model_set_tensor(m, "tok_embeddings.weight", NULL, THO_F16, V, E);
    tho_fill_f16((uint16_t*)m->tok_embeddings.data, seed, 0, V, E, 0, 0, E);
    tho_model_set_tensor(m, "norm.weight", NULL, THO_F32, 1, E);
    tho_fill_gain((float*)m->norm.data, seed, 1, E);
    tho_model_set_tensor(m, "output.weight", NULL, THO_F16, V, E);
    tho_fill_f16((uint16_t*)m->output.data, seed, 2, V, E, 0, 0, E);
    char name[128];
    for (int l = 0; l < m->hp.n_layer; ++l)
        for (int k = 0; k < T_PER_LAYER; ++k) {
            snprintf(name, sizeof name, "layers.%d.%s", l, k_layer_suffix[k]);
            const uint64_t id = 3 + 9ull * l + k;
            if (k == T_ATTN_NORM || k == T_FFN_NORM) {
                tho_model_set_tensor(m, name, NULL, THO_F32, 1, E);
                tho_fill_gain((float*)find_tensor(m, name)->data, seed, id, E);
            } else {
                int64_t R = E, C = E;
                if (k == T_W1 || k == T_W3) { R = F; C = E; }
                if (k == T_W2) { R = E; C = F; }
                tho_model_set_tensor(m, name, NULL, THO_F16, R, C);
                tho_fill_f16((uint16_t*)find_tensor(m, name)->data, seed, id, R, C, 0, 0, C);
            }
        }
}

/* throughput-only runs may fill the cache synthetically instead of running n steps
 * (SURVEY 8d).  Logical element index = ((pos*H + head)*D + d) in the reference layout; tensor
 * ids 1000+2*l (K) and 1001+2*l (V). */
void tho_model_fill_kv_synthetic(tho_model* m, uint64_t seed, int n_positions) {
    const int64_t n = (int64_t)n_positions * m->hp.n_embd;
    for (int l = 0; l < m->hp.n_layer; ++l) {
        tho_fill_kv(m->key_cache[l], seed, 1000 + 2ull * l, n);
        tho_fill_kv(m->value_cache[l], seed, 1001 + 2ull * l, n);
        if (m->kv_f16) {
            for (int64_t i = 0; i < n; ++i) {
                m->key_cache[l][i] = tho_fp16_to_fp32(tho_fp32_to_fp16(m->key_cache[l][i]));
                m->value_cache[l][i] = tho_fp16_to_fp32(tho_fp32_to_fp16(m->value_cache[l][i]));
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * One decode step, op for op as th_eval_gpu / build_layer_cmdbuf / build_final_compute_cmdbuf
 * issue them for n_tokens == 1 (th-llama.cpp:464-660, 270-452, 240-268).
 * Buffer roles: inp0 = x (normed in place), inp6 = residual copy, inp1..3 = Q/K/V,
 * inp4 = Q^T, inp5 = scores.
 * ---------------------------------------------------------------------------------------- */
/* Op trace (test infrastructure for tests/test_graph_trace.py): while a trace is open, eval_one logs every op it issues --
 * the reference pipeline label, the reference names of the operand buffers and the words the reference puts into the op's
 * uniform block -- so that the restated graph can be compared, command by command, with the stream the reference's own
 * th_eval_gpu encodes on the WebGPU stub (tests/golden/graph_trace_tiny.json).  One line per command:
 *   op|<label>|<operand>,<operand>,...|<uniform word>,...        copy|<src>|<dst>|<dst byte offset>|<bytes>          */
static char* g_trace = NULL;
static size_t g_trace_len = 0, g_trace_cap = 0;
static int g_trace_on = 0;
void tho_trace_begin(void) { g_trace_len = 0; if (g_trace) g_trace[0] = 0; g_trace_on = 1; }
const char* tho_trace_end(void) { g_trace_on = 0; return g_trace ? g_trace : ""; }
static void TR(const char* fmt, ...) {
    if (!g_trace_on) return;
    char line[256];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(line, sizeof line, fmt, ap);
    va_end(ap);
    if (n <= 0) return;
    if (g_trace_len + (size_t)n + 2 > g_trace_cap) {
        g_trace_cap = (g_trace_cap + (size_t)n + 2) * 2;
        g_trace = (char*)realloc(g_trace, g_trace_cap);
    }
    memcpy(g_trace + g_trace_len, line, (size_t)n);
    g_trace_len += (size_t)n;
    g_trace[g_trace_len++] = '\n';
    g_trace[g_trace_len] = 0;
}
static uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static void matvec_t(const tho_tensor* W, const float* x, float* y) {
    if (W->ftype == THO_F16) tho_vector_mat_mul_trans_f16(x, (const uint16_t*)W->data, y, W->rows, W->cols, 1);
    else tho_vector_mat_mul_trans_f32(x, (const float*)W->data, y, W->rows, W->cols, 1);
}

static int eval_one(tho_model* m, int32_t token, int n_past, float* logits_out, float* hidden_out) {
    const int64_t E = m->hp.n_embd, H = m->hp.n_head, D = m->head_dim, F = m->n_ff;
    const int64_t N = n_past + 1;
    if (token < 0 || token >= m->hp.n_vocab) return -1;
    if (n_past < 0 || N > m->hp.n_ctx) return -2;
    float *inp0 = m->inp[0], *q = m->inp[1], *k = m->inp[2], *v = m->inp[3], *qT = m->inp[4],
          *scores = m->inp[5], *inp6 = m->inp[6];

    /* embedding row -> inp0 and inp6 (th-llama.cpp:577-585; table is f16 -> f32,
     * th-llama-loader.cpp:185-196) */
    if (m->tok_embeddings.ftype == THO_F16)
        tho_f16_f32_conversion(inp0, (const uint16_t*)m->tok_embeddings.data + (int64_t)token * E, E);
    else
        memcpy(inp0, (const float*)m->tok_embeddings.data + (int64_t)token * E, E * sizeof(float));
    memcpy(inp6, inp0, E * sizeof(float));

    const long long Eb = (long long)(E * (int64_t)sizeof(float));
    const float att_scale = 1.0f / sqrtf((float)D);
    for (int l = 0; l < m->hp.n_layer; ++l) {
        const tho_tensor* T = &m->layers[l * T_PER_LAYER];
        float* kc = m->key_cache[l];
        float* vc = m->value_cache[l];
        tho_rms_norm(inp0, 1, E);                                              /* :299 */
        TR("op|rms_norm|inp0|");
        tho_row_element_multiply(inp0, (const float*)T[T_ATTN_NORM].data, 1, E); /* :300 */
        TR("op|row_element_multiply|inp0,layers.%d.attention_norm|", l);
        matvec_t(&T[T_WQ], inp0, q);                                           /* :304 */
        TR("op|vector_mat_mul|inp0,layers.%d.wq,inp1|", l);
        matvec_t(&T[T_WK], inp0, k);                                           /* :305 */
        TR("op|vector_mat_mul|inp0,layers.%d.wk,inp2|", l);
        matvec_t(&T[T_WV], inp0, v);                                           /* :306 */
        TR("op|vector_mat_mul|inp0,layers.%d.wv,inp3|", l);
        tho_rope(q, 1, H, D, (uint32_t)n_past);                                /* :321 */
        TR("op|RoPE|inp1|%d,1", n_past);
        tho_rope(k, 1, H, D, (uint32_t)n_past);                                /* :322 */
        TR("op|RoPE|inp2|%d,1", n_past);
        if (m->kv_f16) {   /* product option: the cache holds f16 (round to nearest even); the f32 path below is the reference's */
            for (int64_t i = 0; i < E; ++i) { k[i] = tho_fp16_to_fp32(tho_fp32_to_fp16(k[i])); v[i] = tho_fp16_to_fp32(tho_fp32_to_fp16(v[i])); }
        }
        memcpy(kc + (int64_t)n_past * E, k, E * sizeof(float));                /* :337 */
        TR("copy|inp2|layers.%d.key_cache|%lld|%lld", l, (long long)n_past * Eb, Eb);
        memcpy(vc + (int64_t)n_past * E, v, E * sizeof(float));                /* :338 */
        TR("copy|inp3|layers.%d.value_cache|%lld|%lld", l, (long long)n_past * Eb, Eb);
        tho_transpose(kc, m->work_k, N, H, D, 1);                              /* :353 [N,H,D]->[H,N,D] */
        TR("op|transpose|layers.%d.key_cache,working_key_cache|%lld,%lld,%lld", l, (long long)N, (long long)H, (long long)D);
        tho_transpose(vc, m->work_v, N, H, D, 1);                              /* :354 */
        TR("op|transpose|layers.%d.value_cache,working_val_cache|%lld,%lld,%lld", l, (long long)N, (long long)H, (long long)D);
        tho_transpose(q, qT, 1, H, D, 1);                                      /* :355 */
        TR("op|transpose|inp1,inp4|1,%lld,%lld", (long long)H, (long long)D);
        tho_mat_mul(qT, m->work_k, scores, H, 1, D, N, 1, 1, att_scale, 0);    /* :365 */
        TR("op|mat_mul|inp4,working_key_cache,inp5|%lld,1,%lld,%u,%lld,%lld,%lld", (long long)H, (long long)D, f32_bits(att_scale),
           (long long)H, (long long)D, (long long)N);
        tho_row_softmax(scores, H, 1, N);                                      /* :373 */
        TR("op|row_softmax|inp5|%lld,1,%lld", (long long)H, (long long)N);
        tho_mat_mul(scores, m->work_v, k, H, 1, N, D, 0, 1, 1.0f, 0);          /* :380 -> keyBuf */
        TR("op|mat_mul|inp5,working_val_cache,inp2|%lld,1,%lld,%u,%lld,%lld,%lld", (long long)H, (long long)N, f32_bits(1.0f),
           (long long)H, (long long)N, (long long)D);
        tho_transpose(k, v, H, 1, D, 1);                                       /* :397 -> valueBuf */
        TR("op|transpose|inp2,inp3|");
        matvec_t(&T[T_WO], v, q);                                              /* :402 inp1 = Wo*ctx */
        TR("op|vector_mat_mul|inp3,layers.%d.wo,inp1|", l);
        tho_addition(q, inp6, k, E);                                           /* :409 inp2 = inp1+inp6 */
        TR("op|addition|inp1,inp6,inp2|");
        memcpy(v, k, E * sizeof(float));                                       /* :412 inp3 = inp2 */
        TR("copy|inp2|inp3|0|%lld", Eb);
        tho_rms_norm(k, 1, E);                                                 /* :415 */
        TR("op|rms_norm|inp2|");
        tho_row_element_multiply(k, (const float*)T[T_FFN_NORM].data, 1, E);   /* :416 */
        TR("op|row_element_multiply|inp2,layers.%d.ffn_norm|", l);
        matvec_t(&T[T_W1], k, m->ff[0]);                                       /* :423 */
        TR("op|vector_mat_mul|inp2,layers.%d.w1,ffWorking0|", l);
        matvec_t(&T[T_W3], k, m->ff[1]);                                       /* :424 */
        TR("op|vector_mat_mul|inp2,layers.%d.w3,ffWorking1|", l);
        tho_silu(m->ff[0], F);                                                 /* :436 */
        TR("op|silu|ffWorking0|");
        tho_element_mult_in_place(m->ff[0], m->ff[1], F);                      /* :438 */
        TR("op|hadamard_in_place|ffWorking0,ffWorking1|");
        matvec_t(&T[T_W2], m->ff[0], k);                                       /* :441 */
        TR("op|vector_mat_mul|ffWorking0,layers.%d.w2,inp2|", l);
        tho_addition(v, k, inp0, E);                                           /* :447 */
        TR("op|addition|inp3,inp2,inp0|");
        memcpy(inp6, inp0, E * sizeof(float));                                 /* :450 */
        TR("copy|inp0|inp6|0|%lld", Eb);
        if (hidden_out) memcpy(hidden_out + (int64_t)l * E, inp0, E * sizeof(float));
    }
    tho_rms_norm(inp0, 1, E);                                                  /* :252 */
    TR("op|rms_norm|inp0|");
    tho_row_element_multiply(inp0, (const float*)m->norm.data, 1, E);          /* :253 */
    TR("op|row_element_multiply|inp0,norm|");
    if (hidden_out) memcpy(hidden_out + (int64_t)m->hp.n_layer * E, inp0, E * sizeof(float));
    /* :255-262 split matvec + reduce; intended result = full dot product */
    matvec_t(&m->output, inp0, m->out);
    TR("op|output_matvec|inp0,output,out|");          /* = the reference's two vector_mat_mul_split + vector_reduce */
    if (logits_out) memcpy(logits_out, m->out, (size_t)m->hp.n_vocab * sizeof(float));
    return 0;
}

int tho_eval(tho_model* m, const int32_t* tokens, int n_tokens, int n_past, float* logits_out,
             float* hidden_out) {
    if (!tho_model_ready(m)) return -3;
    for (int i = 0; i < n_tokens; ++i) {
        const int last = (i == n_tokens - 1);
        int rc = eval_one(m, tokens[i], n_past + i, last ? logits_out : NULL, last ? hidden_out : NULL);
        if (rc) return rc;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * ggjt v1 (th-llama-loader.cpp:485-635, load_weights :121-265).
 * ---------------------------------------------------------------------------------------- */
#define GGJT_MAGIC 0x67676a74u
#define GGML_MAGIC_UNVERSIONED 0x67676d6cu

tho_model* tho_load_ggjt(const char* path, int32_t n_ctx) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    uint32_t magic = 0, version = 0;
    if (fread(&magic, 4, 1, f) != 1 || magic != GGJT_MAGIC) { fclose(f); return NULL; }
    if (fread(&version, 4, 1, f) != 1 || version != 1) { fclose(f); return NULL; }
    tho_hparams hp; memset(&hp, 0, sizeof hp);
    int32_t h7[7];
    if (fread(h7, 4, 7, f) != 7) { fclose(f); return NULL; }
    hp.n_vocab = h7[0]; hp.n_embd = h7[1]; hp.n_mult = h7[2]; hp.n_head = h7[3];
    hp.n_layer = h7[4]; hp.n_rot = h7[5]; hp.f16 = h7[6]; hp.n_ctx = n_ctx;
    for (int i = 0; i < hp.n_vocab; ++i) { /* vocab: u32 len, bytes, f32 score */
        uint32_t len; float score;
        if (fread(&len, 4, 1, f) != 1) { fclose(f); return NULL; }
        fseek(f, (long)len, SEEK_CUR);
        if (fread(&score, 4, 1, f) != 1) { fclose(f); return NULL; }
    }
    tho_model* m = tho_model_create(&hp);
    for (;;) {
        int32_t n_dims, name_len, ftype;
        if (fread(&n_dims, 4, 1, f) != 1) break; /* EOF */
        if (fread(&name_len, 4, 1, f) != 1 || fread(&ftype, 4, 1, f) != 1) break;
        int32_t ne[2] = {1, 1};
        if (n_dims < 1 || n_dims > 2 || name_len < 0 || name_len > 256) { tho_model_free(m); fclose(f); return NULL; }
        for (int i = 0; i < n_dims; ++i) if (fread(&ne[i], 4, 1, f) != 1) { tho_model_free(m); fclose(f); return NULL; }
        char name[260]; memset(name, 0, sizeof name);
        if (fread(name, 1, (size_t)name_len, f) != (size_t)name_len) { tho_model_free(m); fclose(f); return NULL; }
        long pos = ftell(f);
        pos = (pos + 31) & ~31L;
        fseek(f, pos, SEEK_SET);
        if (ftype != THO_F32 && ftype != THO_F16) { tho_model_free(m); fclose(f); return NULL; }
        const int64_t cols = ne[0], rows = (n_dims == 2) ? ne[1] : 1; /* ne[0] = columns */
        const size_t bytes = (size_t)rows * cols * (ftype == THO_F16 ? 2 : 4);
        void* buf = malloc(bytes);
        if (fread(buf, 1, bytes, f) != bytes) { free(buf); tho_model_free(m); fclose(f); return NULL; }
        tho_model_set_tensor(m, name, buf, ftype, rows, cols);
        free(buf);
    }
    fclose(f);
    return m;
}

int tho_write_ggjt(const tho_model* m, const char* path) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    const uint32_t magic = GGJT_MAGIC, version = 1;
    fwrite(&magic, 4, 1, f); fwrite(&version, 4, 1, f);
    const int32_t h7[7] = {m->hp.n_vocab, m->hp.n_embd, m->hp.n_mult, m->hp.n_head, m->hp.n_layer, m->hp.n_rot, m->hp.f16};
    fwrite(h7, 4, 7, f);
    for (int i = 0; i < m->hp.n_vocab; ++i) { /* synthetic vocab: "<i>" with score -i */
        char tok[32];
        const uint32_t len = (uint32_t)snprintf(tok, sizeof tok, "<%d>", i);
        const float score = -(float)i;
        fwrite(&len, 4, 1, f); fwrite(tok, 1, len, f); fwrite(&score, 4, 1, f);
    }
    const int n_named = 3 + m->hp.n_layer * T_PER_LAYER;
    for (int t = 0; t < n_named; ++t) {
        char name[128]; const tho_tensor* T;
        if (t == 0) { strcpy(name, "tok_embeddings.weight"); T = &m->tok_embeddings; }
        else if (t == 1) { strcpy(name, "norm.weight"); T = &m->norm; }
        else if (t == 2) { strcpy(name, "output.weight"); T = &m->output; }
        else { int l = (t - 3) / T_PER_LAYER, k = (t - 3) % T_PER_LAYER;
               snprintf(name, sizeof name, "layers.%d.%s", l, k_layer_suffix[k]); T = &m->layers[l * T_PER_LAYER + k]; }
        if (!T->data) { fclose(f); return -2; }
        const int32_t n_dims = (T->rows == 1 && T->ftype == THO_F32) ? 1 : 2;
        const int32_t name_len = (int32_t)strlen(name), ftype = T->ftype;
        fwrite(&n_dims, 4, 1, f); fwrite(&name_len, 4, 1, f); fwrite(&ftype, 4, 1, f);
        const int32_t ne0 = (int32_t)T->cols, ne1 = (int32_t)T->rows;
        fwrite(&ne0, 4, 1, f); if (n_dims == 2) fwrite(&ne1, 4, 1, f);
        fwrite(name, 1, (size_t)name_len, f);
        long pos = ftell(f); const long pad = ((pos + 31) & ~31L) - pos;
        static const char zeros[32] = {0};
        fwrite(zeros, 1, (size_t)pad, f);
        fwrite(T->data, T->ftype == THO_F16 ? 2 : 4, (size_t)(T->rows * T->cols), f);
    }
    fclose(f);
    return 0;
}
