// ops.cu -- op-level sm_100a kernels, one per cmdbuf_* op of the reference (th.hpp:302-452).
// These back the reference-shaped graph (build_layer_cmdbuf) one op per launch; the fused
// persistent decoder (decoder.cu) is the fast path.  All arithmetic is f32; f16 weights are
// converted exactly with cvt (same values as the reference's bit-trick decode, th.cpp:363-394).
#include "common.cuh"

// -------------------------------------------------------------------------------------------
// cmdbuf_vector_mat_mul_trans (th.cpp:2839-2892): c[b][r] = sum_k a[b][k] * B[b][r][k]
// One warp per output row, 128-bit streaming loads of B, 4 independent loads in flight per lane.
// -------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256) matvec_kernel(const float* __restrict__ a, const void* __restrict__ Bv,
                                                     float* __restrict__ c, int64_t R, int64_t C) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t b = blockIdx.y;
    const float* x = a + b * C;
    float* y = c + b * R;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < R; r += warps_per_grid) {
        float acc0 = 0.f, acc1 = 0.f;
        if (F16) {
            const uint16_t* w = (const uint16_t*)Bv + (b * R + r) * C;
            const int64_t nvec = C >> 3;  // 8 halves per uint4
            int64_t j = lane;
            for (; j + 96 < nvec; j += 128) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = ld_stream_v4(w + (j + 32 * u) * 8);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 x0 = __ldg((const float4*)(x + (j + 32 * u) * 8));
                    const float4 x1 = __ldg((const float4*)(x + (j + 32 * u) * 8 + 4));
                    float2 f;
                    f = h2_to_f2(v[u].x); acc0 = fmaf(x0.x, f.x, acc0); acc1 = fmaf(x0.y, f.y, acc1);
                    f = h2_to_f2(v[u].y); acc0 = fmaf(x0.z, f.x, acc0); acc1 = fmaf(x0.w, f.y, acc1);
                    f = h2_to_f2(v[u].z); acc0 = fmaf(x1.x, f.x, acc0); acc1 = fmaf(x1.y, f.y, acc1);
                    f = h2_to_f2(v[u].w); acc0 = fmaf(x1.z, f.x, acc0); acc1 = fmaf(x1.w, f.y, acc1);
                }
            }
            for (; j < nvec; j += 32) {
                const uint4 v = ld_stream_v4(w + j * 8);
                const float4 x0 = __ldg((const float4*)(x + j * 8));
                const float4 x1 = __ldg((const float4*)(x + j * 8 + 4));
                float2 f;
                f = h2_to_f2(v.x); acc0 = fmaf(x0.x, f.x, acc0); acc1 = fmaf(x0.y, f.y, acc1);
                f = h2_to_f2(v.y); acc0 = fmaf(x0.z, f.x, acc0); acc1 = fmaf(x0.w, f.y, acc1);
                f = h2_to_f2(v.z); acc0 = fmaf(x1.x, f.x, acc0); acc1 = fmaf(x1.y, f.y, acc1);
                f = h2_to_f2(v.w); acc0 = fmaf(x1.z, f.x, acc0); acc1 = fmaf(x1.w, f.y, acc1);
            }
        } else {
            const float* w = (const float*)Bv + (b * R + r) * C;
            const int64_t nvec = C >> 2;
            for (int64_t j = lane; j < nvec; j += 32) {
                const uint4 v = ld_stream_v4(w + j * 4);
                const float4 x0 = __ldg((const float4*)(x + j * 4));
                acc0 = fmaf(x0.x, __uint_as_float(v.x), acc0); acc1 = fmaf(x0.y, __uint_as_float(v.y), acc1);
                acc0 = fmaf(x0.z, __uint_as_float(v.z), acc0); acc1 = fmaf(x0.w, __uint_as_float(v.w), acc1);
            }
            for (int64_t k = (nvec << 2) + lane; k < C; k += 32) acc0 = fmaf(x[k], w[k], acc0);
        }
        const float s = warp_sum(acc0 + acc1);
        if (lane == 0) y[r] = s;
    }
}

extern "C" int thk_vector_mat_mul_trans(thk_ctx* ctx, const float* a, size_t a_offset_bytes, const void* B,
                                        float* c, int64_t R, int64_t C, int64_t batch, int b_is_f16) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && a && B && c, "thk_vector_mat_mul_trans: null argument");
    THK_CHECK_ARG(R > 0 && C > 0, "thk_vector_mat_mul_trans: R=%lld C=%lld", (long long)R, (long long)C);
    THK_CHECK_ARG(a_offset_bytes % 16 == 0, "thk_vector_mat_mul_trans: aOffset %zu not 16-byte aligned", a_offset_bytes);
    THK_CHECK_ARG(!b_is_f16 || C % 8 == 0, "thk_vector_mat_mul_trans: f16 rows need C %% 8 == 0 (C=%lld)", (long long)C);
    THK_CHECK_ARG(b_is_f16 || C % 4 == 0, "thk_vector_mat_mul_trans: f32 rows need C %% 4 == 0 (C=%lld)", (long long)C);
    if (batch <= 0) batch = 1;
    const float* x = (const float*)((const char*)a + a_offset_bytes);
    int64_t blocks = (R + 7) / 8;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)batch);
    if (b_is_f16) matvec_kernel<true><<<grid, 256, 0, ctx->stream>>>(x, B, c, R, C);
    else matvec_kernel<false><<<grid, 256, 0, ctx->stream>>>(x, B, c, R, C);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

__global__ void add_inplace_kernel(float* a, const float* b, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        a[i] = a[i] + b[i];
}
static inline unsigned ew_blocks(const thk_ctx* ctx, int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

extern "C" int thk_vector_reduce(thk_ctx* ctx, float* a, const float* b, int64_t n) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && a && b && n > 0, "thk_vector_reduce: bad argument");
    add_inplace_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(a, b, n);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

extern "C" int thk_vector_multi_mat_mul_split_trans(thk_ctx* ctx, const float* a, size_t a_offset_bytes,
                                                    const void* const* B_splits, int nsplit, float* c,
                                                    float* scratch, int64_t R, int64_t C_total, int b_is_f16) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && a && B_splits && c, "thk_vector_multi_mat_mul_split_trans: null argument");
    THK_CHECK_ARG(nsplit >= 1 && C_total % nsplit == 0, "split count %d does not divide C=%lld", nsplit, (long long)C_total);
    THK_CHECK_ARG(nsplit == 1 || scratch, "thk_vector_multi_mat_mul_split_trans: scratch needed for %d splits", nsplit);
    const int64_t Cs = C_total / nsplit;
    for (int s = 0; s < nsplit; ++s) {
        float* dst = (s == 0) ? c : scratch;
        int rc = thk_vector_mat_mul_trans(ctx, a, a_offset_bytes + (size_t)s * Cs * sizeof(float), B_splits[s], dst, R, Cs, 1, b_is_f16);
        if (rc) return rc;
        if (s > 0) { rc = thk_vector_reduce(ctx, c, scratch, R); if (rc) return rc; }
    }
    return THK_OK;
}

// -------------------------------------------------------------------------------------------
// cmdbuf_rms_norm (th.cpp:1153-1200): x <- x / sqrt(mean(x^2) + 1e-6), one block per row
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rms_norm_kernel(float* __restrict__ x, int64_t N) {
    __shared__ float red[8];
    __shared__ float inv_s;
    float* row = x + (int64_t)blockIdx.x * N;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < N; i += 256) { const float v = row[i]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        inv_s = 1.0f / sqrtf(t / (float)N + 1e-6f);
    }
    __syncthreads();
    const float inv = inv_s;
    for (int64_t i = threadIdx.x; i < N; i += 256) row[i] = row[i] * inv;
}
extern "C" int thk_rms_norm(thk_ctx* ctx, float* x, int64_t rows, int64_t N) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && x && rows > 0 && N > 0, "thk_rms_norm: bad argument");
    rms_norm_kernel<<<(unsigned)rows, 256, 0, ctx->stream>>>(x, N);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

__global__ void row_mul_kernel(float* x, const float* g, int64_t total, int64_t N) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = x[i] * g[i % N];
}
extern "C" int thk_row_element_multiply(thk_ctx* ctx, float* x, const float* gain, int64_t rows, int64_t N) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && x && gain && rows > 0 && N > 0, "thk_row_element_multiply: bad argument");
    row_mul_kernel<<<ew_blocks(ctx, rows * N), 256, 0, ctx->stream>>>(x, gain, rows * N, N);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

// -------------------------------------------------------------------------------------------
// cmdbuf_RoPE (th.cpp:1457-1492): adjacent pairs, theta = 10000^(-x/head_dim), p = n_past + token
// -------------------------------------------------------------------------------------------
__global__ void rope_kernel(float* x, int64_t n_tokens, int64_t n_head, int64_t head_dim,
                            const thk_network_uniforms* __restrict__ u) {
    const int64_t half = head_dim >> 1;
    const int64_t total = n_tokens * n_head * half;
    const uint32_t n_past = u->n_past;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pair = i % half;
        const int64_t z = i / (half * n_head);
        const float p = (float)(n_past + (uint32_t)z);
        const float theta = powf(10000.0f, (-(float)(2 * pair)) / (float)head_dim);
        float s, c;
        sincosf(p * theta, &s, &c);
        float* v = x + 2 * i;
        const float x0 = v[0], x1 = v[1];
        v[0] = x0 * c - x1 * s;
        v[1] = x0 * s + x1 * c;
    }
}
extern "C" int thk_rope(thk_ctx* ctx, float* x, int64_t n_tokens, int64_t n_head, int64_t head_dim,
                        const thk_network_uniforms* uniforms) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && x && uniforms, "thk_rope: null argument");
    THK_CHECK_ARG(n_tokens > 0 && n_head > 0 && head_dim > 0 && head_dim % 2 == 0, "thk_rope: bad shape");
    rope_kernel<<<ew_blocks(ctx, n_tokens * n_head * head_dim / 2), 256, 0, ctx->stream>>>(x, n_tokens, n_head, head_dim, uniforms);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

// -------------------------------------------------------------------------------------------
// cmdbuf_transpose (th.cpp:876-912)
// -------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t B, int64_t M,
                                 int64_t N, int zy, const thk_dims_uniforms* __restrict__ u) {
    if (u) { B = u->A_B; M = u->A_M; N = u->A_N; }
    const int64_t total = B * M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t xx = i % N, y = (i / N) % M, z = i / (N * M);
        if (zy) out[y * (B * N) + z * N + xx] = in[i];
        else out[z * M * N + xx * M + y] = in[i];
    }
}
extern "C" int thk_transpose(thk_ctx* ctx, const float* in, float* out, int64_t B, int64_t M, int64_t N, int zy,
                             const thk_dims_uniforms* uniforms) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && in && out, "thk_transpose: null argument");
    THK_CHECK_ARG(in != out, "thk_transpose: in-place transpose is not supported");
    if (B <= 0) B = 1;
    THK_CHECK_ARG(M > 0 && N > 0, "thk_transpose: bad shape");
    transpose_kernel<<<ew_blocks(ctx, B * M * N), 256, 0, ctx->stream>>>(in, out, B, M, N, zy, uniforms);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

// -------------------------------------------------------------------------------------------
// cmdbuf_mat_mul (th.cpp:396-539): batched C = A*B or A*B^T (+ scale).  transposeB: one warp per
// output element (contiguous dot); else one thread per output column.
// -------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ float load_b(const void* B, int64_t i) {
    if (F16) return __half2float(__ushort_as_half(((const uint16_t*)B)[i]));
    return ((const float*)B)[i];
}
template <bool F16>
__global__ void __launch_bounds__(256) mat_mul_tb_kernel(const float* __restrict__ A, const void* __restrict__ Bm,
                                                         float* __restrict__ Cm, int64_t batch, int64_t M, int64_t K,
                                                         int64_t N, int do_scale, const thk_dims_uniforms* __restrict__ u) {
    float scale = 1.0f;
    if (u) { M = u->A_M; N = u->B_N; K = u->A_N; scale = u->scale; do_scale = 1; }
    const int lane = threadIdx.x & 31;
    const int64_t total = batch * M * N;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); o < total; o += wstride) {
        const int64_t n = o % N, m = (o / N) % M, z = o / (N * M);
        const float* a = A + (z * M + m) * K;
        const int64_t boff = (z * N + n) * K;
        float acc = 0.f;
        for (int64_t k = lane; k < K; k += 32) acc = fmaf(a[k], load_b<F16>(Bm, boff + k), acc);
        acc = warp_sum(acc);
        if (lane == 0) Cm[o] = do_scale ? acc * scale : acc;
    }
}
template <bool F16>
__global__ void __launch_bounds__(128) mat_mul_nn_kernel(const float* __restrict__ A, const void* __restrict__ Bm,
                                                         float* __restrict__ Cm, int64_t batch, int64_t M, int64_t K,
                                                         int64_t N, int do_scale, const thk_dims_uniforms* __restrict__ u) {
    float scale = 1.0f;
    if (u) { M = u->A_M; N = u->B_N; K = u->A_N; scale = u->scale; do_scale = 1; }
    const int64_t total = batch * M * N;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = o % N, m = (o / N) % M, z = o / (N * M);
        const float* a = A + (z * M + m) * K;
        float acc = 0.f;
        for (int64_t k = 0; k < K; ++k) acc = fmaf(a[k], load_b<F16>(Bm, (z * K + k) * N + n), acc);
        Cm[o] = do_scale ? acc * scale : acc;
    }
}
// f32 x f32 with more than a few rows of A (the attention matmuls of a batched prompt pass: [H][T][D] x [H][N][D]^T and
// [H][T][N] x [H][N][D]): 64 x 64 output tile per CTA, K in steps of 16 through shared memory, 4 x 4 outputs per thread.
// (The per-output-warp kernels above took 142 + 41 us per layer at T = 128 -- a quarter of the 128-token prefill,
// profiles/v7_prefill_launches.md; they remain for M == 1 and for f16 B.)
template <bool TB>
__global__ void __launch_bounds__(256) mat_mul_tiled_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Cm,
                                                            int64_t batch, int64_t M, int64_t K, int64_t N,
                                                            const thk_dims_uniforms* __restrict__ u) {
    float scale = 1.0f;
    bool do_scale = false;
    if (u) { M = u->A_M; N = u->B_N; K = u->A_N; scale = u->scale; do_scale = true; }
    __shared__ float As[16][65], Bs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t tiles_m = (M + 63) / 64, tiles_n = (N + 63) / 64;
    for (int64_t t = blockIdx.x; t < batch * tiles_m * tiles_n; t += gridDim.x) {
        const int64_t z = t / (tiles_m * tiles_n), rem = t % (tiles_m * tiles_n);
        const int64_t m0 = (rem / tiles_n) * 64, n0 = (rem % tiles_n) * 64;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int64_t k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = (int)threadIdx.x + i * 256;
                {
                    const int r = idx >> 4, kk = idx & 15;
                    const int64_t m = m0 + r, k = k0 + kk;
                    As[kk][r] = (m < M && k < K) ? A[(z * M + m) * K + k] : 0.f;
                }
                if (TB) {
                    const int c = idx >> 4, kk = idx & 15;
                    const int64_t n = n0 + c, k = k0 + kk;
                    Bs[kk][c] = (n < N && k < K) ? Bm[(z * N + n) * K + k] : 0.f;
                } else {
                    const int kk = idx >> 6, c = idx & 63;
                    const int64_t n = n0 + c, k = k0 + kk;
                    Bs[kk][c] = (n < N && k < K) ? Bm[(z * K + k) * N + n] : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty + 16 * i]; b[i] = Bs[kk][tx + 16 * i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t m = m0 + ty + 16 * i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t n = n0 + tx + 16 * j;
                if (n < N) Cm[(z * M + m) * N + n] = do_scale ? acc[i][j] * scale : acc[i][j];
            }
        }
    }
}
extern "C" int thk_mat_mul(thk_ctx* ctx, const float* A, const void* B, float* C, int64_t batch, int64_t M, int64_t K,
                           int64_t N, int transposeB, int b_is_f16, const thk_dims_uniforms* uniforms) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && A && B && C, "thk_mat_mul: null argument");
    if (batch <= 0) batch = 1;
    THK_CHECK_ARG(M > 0 && K > 0 && N > 0, "thk_mat_mul: one of the dimensions is zero");
    const int64_t total = batch * M * N;
    // NOTE: with uniforms the kernels take M, N, K from the device; the host values only size the grid
    if (!b_is_f16 && M >= 8) {
        int64_t blocks = batch * ((M + 63) / 64) * ((N + 63) / 64); const int64_t cap = (int64_t)ctx->sm_count * 8; if (blocks > cap) blocks = cap;
        if (transposeB) mat_mul_tiled_kernel<true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(A, (const float*)B, C, batch, M, K, N, uniforms);
        else mat_mul_tiled_kernel<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(A, (const float*)B, C, batch, M, K, N, uniforms);
    } else if (transposeB) {
        int64_t blocks = (total + 7) / 8; const int64_t cap = (int64_t)ctx->sm_count * 8; if (blocks > cap) blocks = cap;
        if (b_is_f16) mat_mul_tb_kernel<true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(A, B, C, batch, M, K, N, 0, uniforms);
        else mat_mul_tb_kernel<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(A, B, C, batch, M, K, N, 0, uniforms);
    } else {
        int64_t blocks = (total + 127) / 128; const int64_t cap = (int64_t)ctx->sm_count * 16; if (blocks > cap) blocks = cap;
        if (b_is_f16) mat_mul_nn_kernel<true><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A, B, C, batch, M, K, N, 0, uniforms);
        else mat_mul_nn_kernel<false><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A, B, C, batch, M, K, N, 0, uniforms);
    }
    THK_LAUNCH_CHECK();
    return THK_OK;
}

// -------------------------------------------------------------------------------------------
// cmdbuf_row_softmax (th.cpp:1865-1961) and causal variant (intended cmdbuf_masked_softmax)
// one block per (row, batch)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
    return t;
}
__global__ void __launch_bounds__(256) softmax_kernel(float* a, int64_t M, int64_t N, int causal,
                                                      const thk_dims_uniforms* __restrict__ u) {
    __shared__ float red[8];
    if (u) { M = u->A_M; N = u->A_N; }
    const int64_t y = blockIdx.x, z = blockIdx.y;
    if (y >= M) return;
    float* row = a + (z * M + y) * N;
    int64_t lim = N;
    if (causal) { lim = (N - M) + y + 1; if (lim > N) lim = N; }
    float mx = -1e14f;
    for (int64_t i = threadIdx.x; i < lim; i += 256) mx = fmaxf(mx, row[i]);
    mx = block_reduce(mx, red, true);
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < lim; i += 256) { const float e = expf(row[i] - mx); row[i] = e; s += e; }
    s = block_reduce(s, red, false);
    for (int64_t i = threadIdx.x; i < lim; i += 256) row[i] = row[i] / s;
    for (int64_t i = lim + threadIdx.x; i < N; i += 256) row[i] = 0.f;
}
static int softmax_launch(thk_ctx* ctx, float* a, int64_t batch, int64_t M, int64_t N, int causal,
                          const thk_dims_uniforms* uniforms, const char* who) {
    THK_CHECK_ARG(ctx && a, "%s: null argument", who);
    if (batch <= 0) batch = 1;
    THK_CHECK_ARG(M > 0 && N > 0, "%s: bad shape", who);
    softmax_kernel<<<dim3((unsigned)M, (unsigned)batch), 256, 0, ctx->stream>>>(a, M, N, causal, uniforms);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
extern "C" int thk_row_softmax(thk_ctx* ctx, float* a, int64_t batch, int64_t M, int64_t N, const thk_dims_uniforms* u) {
    THK_ENTER(ctx);
    return softmax_launch(ctx, a, batch, M, N, 0, u, "thk_row_softmax");
}
extern "C" int thk_masked_softmax(thk_ctx* ctx, float* a, int64_t batch, int64_t M, int64_t N, const thk_dims_uniforms* u) {
    THK_ENTER(ctx);
    return softmax_launch(ctx, a, batch, M, N, 1, u, "thk_masked_softmax");
}

// -------------------------------------------------------------------------------------------
// elementwise: cmdbuf_addition (2121-2149), cmdbuf_silu (2680-2709), cmdbuf_element_mult_in_place
// (2498-2526), cmdbuf_f16_f32_conversion (4129-4165)
// -------------------------------------------------------------------------------------------
__global__ void add_kernel(const float* a, const float* b, float* c, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c[i] = a[i] + b[i];
}
__global__ void silu_kernel(float* a, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = a[i];
        a[i] = v / (1.0f + expf(-v));
    }
}
__global__ void mul_kernel(float* a, const float* b, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = a[i] * b[i];
}
__global__ void f16_f32_kernel(float* out, const uint16_t* in, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __half2float(__ushort_as_half(in[i]));
}
extern "C" int thk_addition(thk_ctx* ctx, const float* a, const float* b, float* c, int64_t n) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && a && b && c && n > 0, "thk_addition: bad argument");
    add_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(a, b, c, n);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
extern "C" int thk_silu(thk_ctx* ctx, float* a, int64_t n) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && a && n > 0, "thk_silu: bad argument");
    silu_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(a, n);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
extern "C" int thk_element_mult_in_place(thk_ctx* ctx, float* a, const float* b, int64_t n) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && a && b && n > 0, "thk_element_mult_in_place: bad argument");
    mul_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(a, b, n);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
extern "C" int thk_f16_f32_conversion(thk_ctx* ctx, float* out, size_t out_off, const uint16_t* in, size_t in_off, int64_t n) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && out && in && n > 0, "thk_f16_f32_conversion: bad argument");
    THK_CHECK_ARG(out_off % 4 == 0 && in_off % 2 == 0, "thk_f16_f32_conversion: misaligned offsets");
    f16_f32_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>((float*)((char*)out + out_off), (const uint16_t*)((const char*)in + in_off), n);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

// -------------------------------------------------------------------------------------------
// synthetic tensors (oracle twins: tho_fill_f16 / tho_fill_gain / tho_fill_kv)
// -------------------------------------------------------------------------------------------
__global__ void fill_f16_kernel(uint16_t* dst, uint64_t seed, uint64_t tid, int64_t rows, int64_t cols, int64_t row0,
                                int64_t col0, int64_t full_cols) {
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i % cols;
        const int32_t k = (int32_t)(thk_hash(seed, tid, (uint64_t)((row0 + r) * full_cols + col0 + c)) >> 53) - 1024;
        dst[i] = __half_as_ushort(__float2half_rn((float)k * (1.0f / 32768.0f)));
    }
}
__global__ void fill_gain_kernel(float* dst, uint64_t seed, uint64_t tid, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t k = (int32_t)(thk_hash(seed, tid, (uint64_t)i) >> 54) - 512;
        dst[i] = 1.0f + (float)k * (1.0f / 4096.0f);
    }
}
__global__ void fill_kv_kernel(float* dst, uint64_t seed, uint64_t tid, int64_t n_pos, int64_t n_ctx, int64_t H,
                               int64_t head0, int64_t Hl, int64_t D) {
    const int64_t total = Hl * n_pos * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = i % D, pos = (i / D) % n_pos, hl = i / (D * n_pos);
        const uint64_t logical = (uint64_t)((pos * H + head0 + hl) * D + d);
        const int32_t k = (int32_t)(thk_hash(seed, tid, logical) >> 48) - 32768;
        dst[(hl * n_ctx + pos) * D + d] = (float)k * (1.0f / 32768.0f);
    }
}
extern "C" int thk_fill_f16(thk_ctx* ctx, uint16_t* dst, uint64_t seed, uint64_t tid, int64_t rows, int64_t cols,
                            int64_t row0, int64_t col0, int64_t full_cols) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dst && rows > 0 && cols > 0, "thk_fill_f16: bad argument");
    fill_f16_kernel<<<ew_blocks(ctx, rows * cols), 256, 0, ctx->stream>>>(dst, seed, tid, rows, cols, row0, col0, full_cols);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
extern "C" int thk_fill_gain(thk_ctx* ctx, float* dst, uint64_t seed, uint64_t tid, int64_t n) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dst && n > 0, "thk_fill_gain: bad argument");
    fill_gain_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(dst, seed, tid, n);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
extern "C" int thk_fill_kv(thk_ctx* ctx, float* dst, uint64_t seed, uint64_t tid, int64_t n_pos, int64_t n_ctx,
                           int64_t H, int64_t head0, int64_t Hl, int64_t D) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dst && n_pos > 0 && n_pos <= n_ctx, "thk_fill_kv: bad argument");
    fill_kv_kernel<<<ew_blocks(ctx, Hl * n_pos * D), 256, 0, ctx->stream>>>(dst, seed, tid, n_pos, n_ctx, H, head0, Hl, D);
    THK_LAUNCH_CHECK();
    return THK_OK;
}

// -------------------------------------------------------------------------------------------
// KV rows [pos0, pos0+npos) from the reference layout [pos][head][dim] into the fused decoder's
// [head][n_ctx][dim] layout (after a batched prefill through the op graph; no reference analogue --
// the reference re-transposes the whole cache every token instead, th-llama.cpp:353-354)
// -------------------------------------------------------------------------------------------
__global__ void kv_to_hpd_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t pos0, int64_t npos, int64_t n_ctx,
                                 int64_t H, int64_t D) {
    const int64_t total = npos * H * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = i % D, h = (i / D) % H, pz = i / (D * H);
        dst[(h * n_ctx + pos0 + pz) * D + d] = src[((pos0 + pz) * H + h) * D + d];
    }
}
extern "C" int thk_kv_to_hpd(thk_ctx* ctx, const float* src_phd, float* dst_hpd, int64_t pos0, int64_t npos, int64_t n_ctx, int64_t H,
                             int64_t D) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && src_phd && dst_hpd, "thk_kv_to_hpd: null argument");
    THK_CHECK_ARG(pos0 >= 0 && npos > 0 && pos0 + npos <= n_ctx && H > 0 && D > 0, "thk_kv_to_hpd: bad range");
    kv_to_hpd_kernel<<<ew_blocks(ctx, npos * H * D), 256, 0, ctx->stream>>>(src_phd, dst_hpd, pos0, npos, n_ctx, H, D);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
// ... into an f16 fused-layout cache (thk_llama_dims.kv_f16): rounded to nearest even, like the fused kernel's own append
__global__ void kv_to_hpd_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t pos0, int64_t npos, int64_t n_ctx,
                                     int64_t H, int64_t D) {
    const int64_t total = npos * H * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = i % D, h = (i / D) % H, pz = i / (D * H);
        dst[(h * n_ctx + pos0 + pz) * D + d] = __float2half_rn(src[((pos0 + pz) * H + h) * D + d]);
    }
}
extern "C" int thk_kv_to_hpd_f16(thk_ctx* ctx, const float* src_phd, uint16_t* dst_hpd, int64_t pos0, int64_t npos, int64_t n_ctx, int64_t H,
                                 int64_t D) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && src_phd && dst_hpd, "thk_kv_to_hpd_f16: null argument");
    THK_CHECK_ARG(pos0 >= 0 && npos > 0 && pos0 + npos <= n_ctx && H > 0 && D > 0, "thk_kv_to_hpd_f16: bad range");
    kv_to_hpd_f16_kernel<<<ew_blocks(ctx, npos * H * D), 256, 0, ctx->stream>>>(src_phd, (__half*)dst_hpd, pos0, npos, n_ctx, H, D);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
__global__ void kv_from_hpd_f16_kernel(const __half* __restrict__ src, float* __restrict__ dst, int64_t pos0, int64_t npos, int64_t n_ctx,
                                       int64_t H, int64_t D) {
    const int64_t total = npos * H * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = i % D, h = (i / D) % H, pz = i / (D * H);
        dst[((pos0 + pz) * H + h) * D + d] = __half2float(src[(h * n_ctx + pos0 + pz) * D + d]);
    }
}
extern "C" int thk_kv_from_hpd_f16(thk_ctx* ctx, const uint16_t* src_hpd, float* dst_phd, int64_t pos0, int64_t npos, int64_t n_ctx,
                                   int64_t H, int64_t D) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && src_hpd && dst_phd, "thk_kv_from_hpd_f16: null argument");
    THK_CHECK_ARG(pos0 >= 0 && npos > 0 && pos0 + npos <= n_ctx && H > 0 && D > 0, "thk_kv_from_hpd_f16: bad range");
    kv_from_hpd_f16_kernel<<<ew_blocks(ctx, npos * H * D), 256, 0, ctx->stream>>>((const __half*)src_hpd, dst_phd, pos0, npos, n_ctx, H, D);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __float2half_rn(src[i]);
}
extern "C" int thk_f32_to_f16(thk_ctx* ctx, const float* src, uint16_t* dst, int64_t n) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && src && dst && n > 0, "thk_f32_to_f16: bad argument");
    f32_to_f16_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(src, (__half*)dst, n);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
// the inverse hand-over: rows [pos0, pos0+npos) of the fused decoder's [head][n_ctx][dim] cache back into the op graph's
// [pos][head][dim] cache, so that a batched pass or the op graph can continue a context the fused kernel extended
__global__ void kv_from_hpd_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t pos0, int64_t npos, int64_t n_ctx,
                                   int64_t H, int64_t D) {
    const int64_t total = npos * H * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = i % D, h = (i / D) % H, pz = i / (D * H);
        dst[((pos0 + pz) * H + h) * D + d] = src[(h * n_ctx + pos0 + pz) * D + d];
    }
}
extern "C" int thk_kv_from_hpd(thk_ctx* ctx, const float* src_hpd, float* dst_phd, int64_t pos0, int64_t npos, int64_t n_ctx, int64_t H,
                               int64_t D) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && src_hpd && dst_phd, "thk_kv_from_hpd: null argument");
    THK_CHECK_ARG(pos0 >= 0 && npos > 0 && pos0 + npos <= n_ctx && H > 0 && D > 0, "thk_kv_from_hpd: bad range");
    kv_from_hpd_kernel<<<ew_blocks(ctx, npos * H * D), 256, 0, ctx->stream>>>(src_hpd, dst_phd, pos0, npos, n_ctx, H, D);
    THK_LAUNCH_CHECK();
    return THK_OK;
}
