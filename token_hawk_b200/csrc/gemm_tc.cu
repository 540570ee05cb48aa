// gemm_tc.cu -- batched-prompt GEMM on the 5th-gen tensor cores:  Y[M,N] = X[M,K] * W[N,K]^T
//
// Replaces cmdbuf_mat_mul with an f16 B operand on the n_tokens > 1 path (th-llama.cpp:308-310, 404,
// 429-430, 444; shader th.cpp:396-539, an 8x8-thread scalar GEMM).  W is the ggjt f16 weight matrix,
// row-major [N,K]: already "K-major", exactly what tcgen05 wants for operand B.  X is f32 (the
// reference keeps activations in f32), so it is split on the fly into two f16 terms,
//     X = hi + lo,   hi = f16(X),  lo = f16(X - f32(hi)),
// and both products accumulate into the same f32 TMEM accumulator.  f16 x f16 products are exact in
// f32, so the result carries ~22 bits of X instead of 11 and stays inside the 1e-3 logit budget.
//
// Kernel shape: one CTA per 128 x BN output tile and K slice.  BN = 128 whenever N allows it (every LLaMA matrix), with
// split-K so that ~one CTA per SM is busy (N = 4096: 32 column tiles x 4 K slices); the K slices write partial tiles to a
// per-context workspace and a second kernel adds them in slice order (deterministic).  Why: every CTA streams the whole
// hi + lo activation panel of its K range from L2, so with 128 x 32 tiles (round 1) a K = 4096 GEMM moved 268 MB of
// activations L2->SM against 33 MB of weights and took 53 us on average (profiles/v7_prefill_launches.md).
//   warp 0 lane 0 : TMA producer   cp.async.bulk.tensor.2d (128B swizzle) -> smem ring (4 x 48 KB, or 5 x 36 KB at BN = 32)
//   warp 1 lane 0 : MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M=128 N=BN K=16, D in TMEM
//   warp 2        : TMEM alloc / dealloc (BN columns)
//   warps 0..3    : epilogue       tcgen05.ld 32x32b.x32 per 32 columns -> registers -> global f32
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64;                    // BK f16 = 128 bytes = one swizzle-128B row
constexpr int kABytes = BM * BK * 2;                // 16 KB
constexpr int kMaxSplits = 8;
template <int BN> struct Tile {
    static constexpr int kBBytes = BN * BK * 2;                 // 4 KB (BN 32) / 16 KB (BN 128)
    static constexpr int kStageBytes = 2 * kABytes + kBBytes;   // hi + lo + W
    static constexpr int kStages = BN == 128 ? 4 : 5;
    static constexpr int kTmemCols = BN;                        // power of two >= 32
    static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024;
};
constexpr unsigned long long kTimeoutNs = 2000000000ull;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void bar_init(uint32_t b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint32_t b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bar_try(uint32_t b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool bar_wait(uint32_t b, uint32_t parity, unsigned* status) {
    if (bar_try(b, parity)) return true;
    const unsigned long long t0 = gtime();
    unsigned it = 0;
    while (!bar_try(b, parity)) {
        if ((++it & 255u) == 0u) {
            if (*(volatile unsigned*)status != 0u) return false;
            if (gtime() - t0 > kTimeoutNs) { atomicCAS(status, 0u, 0x500u | (b & 0xffu)); return false; }
        }
    }
    return true;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);         // start address
    d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: A=f16, B=f16, D=f32, both K-major, N at [17,23) as N>>3, M at [24,29) as M>>4
__device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct GemmParams {
    CUtensorMap map_hi, map_lo, map_w;
    float* Y;            // output, or (splits > 1) the partial-tile workspace [splits][M][N]
    int M, N, K, splits;
    unsigned* status;
};

template <int BN>
__global__ void __launch_bounds__(128, 1) gemm_f16_tc_kernel(const __grid_constant__ GemmParams p) {
    constexpr int kStages = Tile<BN>::kStages, kStageBytes = Tile<BN>::kStageBytes, kTmemCols = Tile<BN>::kTmemCols;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[kStages], empty_bar[kStages], accum_bar;
    __shared__ uint32_t tmem_base_holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int nkb_all = p.K / BK;
    const int kb0 = (int)(((long long)nkb_all * blockIdx.z) / p.splits), kb1 = (int)(((long long)nkb_all * (blockIdx.z + 1)) / p.splits);
    const int nkb = kb1 - kb0;                        // K blocks of this CTA's slice (>= 1: the host caps splits at K / BK)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { bar_init(s32(&full_bar[s]), 1); bar_init(s32(&empty_bar[s]), 1); }
        bar_init(s32(&accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_holder)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_holder;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer ----
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_w) : "memory");
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            const uint32_t par = ((kb / kStages) & 1) ^ 1;
            if (!bar_wait(s32(&empty_bar[s]), par, p.status)) break;
            const uint32_t fb = s32(&full_bar[s]);
            unsigned char* st = smem + (size_t)s * kStageBytes;
            bar_expect_tx(fb, kStageBytes);
            tma_load_2d(s32(st), &p.map_hi, (kb0 + kb) * BK, m0, fb);
            tma_load_2d(s32(st + kABytes), &p.map_lo, (kb0 + kb) * BK, m0, fb);
            tma_load_2d(s32(st + 2 * kABytes), &p.map_w, (kb0 + kb) * BK, n0, fb);
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer: one thread drives the tensor core for the whole CTA ----
        const uint32_t idesc = umma_idesc_f16(BM, BN);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            if (!bar_wait(s32(&full_bar[s]), (kb / kStages) & 1, p.status)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = s32(smem + (size_t)s * kStageBytes), a_lo = a_hi + kABytes, b = a_hi + 2 * kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {          // UMMA_K = 16 f16 = 32 bytes along the swizzled row
                const uint64_t db = umma_desc_sw128(b + k * 32);
                umma_f16(tmem_d, umma_desc_sw128(a_hi + k * 32), db, idesc, (kb | k) != 0);
                umma_f16(tmem_d, umma_desc_sw128(a_lo + k * 32), db, idesc, 1u);
            }
            umma_commit(s32(&empty_bar[s]));             // frees the smem stage once these MMAs retire
        }
        umma_commit(s32(&accum_bar));                    // accumulator complete
    }
    __syncwarp();

    // ---- epilogue: all four warps, warp w reads TMEM lanes 32w..32w+31 (= output rows) ----
    const bool ok = bar_wait(s32(&accum_bar), 0, p.status);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        const int m = m0 + warp * 32 + lane;
        float* out = p.Y + (size_t)blockIdx.z * p.M * p.N;            // splits > 1: this slice's partial tile
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < p.M) {
                float4* dst = (float4*)(out + (size_t)m * p.N + n0 + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
}

// X (f32) -> hi, lo (f16): X = hi + lo to ~22 bits
__global__ void split_hi_lo_kernel(const float* __restrict__ X, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = X[i];
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}

// Y = sum of the K slices' partial tiles, in slice order (deterministic); n4 = M * N / 4
__global__ void splitk_reduce_kernel(const float4* __restrict__ ws, float4* __restrict__ Y, int64_t n4, int splits) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = ws[i];
        for (int z = 1; z < splits; ++z) {
            const float4 b = ws[(int64_t)z * n4 + i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        Y[i] = a;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(EncodeTiledFn fn, CUtensorMap* map, const void* base, int64_t rows, int64_t K, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { thk_set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld", (int)r, (long long)rows, (long long)K); return THK_E_CUDA; }
    return THK_OK;
}

}  // namespace

static EncodeTiledFn lookup_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    return (EncodeTiledFn)fn;
}

// The hi/lo split workspace and the status word belong to the context (one stream): thk_gemm_reserve sizes them once at
// model load, so that thk_gemm_f16_tc itself never allocates ("kernels never allocate", thk_cabi.h); a larger request
// than reserved still works (drain this context's stream, grow) but is not the intended path.
static int gemm_workspace(thk_ctx* ctx, size_t need) {
    if (ctx->gemm_ws_bytes < need) {
        THK_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->gemm_ws);
        ctx->gemm_ws = nullptr; ctx->gemm_ws_bytes = 0;
        THK_CUDA(cudaMalloc(&ctx->gemm_ws, need));
        ctx->gemm_ws_bytes = need;
    }
    if (!ctx->gemm_status) { THK_CUDA(cudaMalloc(&ctx->gemm_status, 16)); THK_CUDA(cudaMemsetAsync(ctx->gemm_status, 0, 16, ctx->stream)); }
    if (!ctx->gemm_attr_set) {
        THK_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<32>::kSmem));
        THK_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<128>::kSmem));
        ctx->gemm_attr_set = true;
    }
    return THK_OK;
}

extern "C" int thk_gemm_reserve(thk_ctx* ctx, int64_t max_M, int64_t max_K) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && max_M > 0 && max_K > 0, "thk_gemm_reserve: bad argument");
    // hi + lo panels, and the split-K partial tiles: splits * (N / 128) <= sm_count, so splits * N <= 128 * sm_count
    return gemm_workspace(ctx, (size_t)max_M * max_K * 2 * 2 + (size_t)max_M * 128 * (size_t)ctx->sm_count * sizeof(float));
}

extern "C" int thk_gemm_f16_tc(thk_ctx* ctx, const float* X, const uint16_t* W, float* Y, int64_t M, int64_t N, int64_t K) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && X && W && Y, "thk_gemm_f16_tc: null argument");
    THK_CHECK_ARG(M > 0 && N > 0 && K > 0, "thk_gemm_f16_tc: bad shape");
    THK_CHECK_ARG(N % 32 == 0, "thk_gemm_f16_tc: N must be a multiple of %d (N=%lld)", 32, (long long)N);
    THK_CHECK_ARG(K % BK == 0, "thk_gemm_f16_tc: K must be a multiple of %d (K=%lld)", BK, (long long)K);
    static const EncodeTiledFn encode = lookup_encode();        // (thread-safe: C++11 static initialisation)
    if (!encode) { thk_set_error("cuTensorMapEncodeTiled not available in this driver"); return THK_E_UNSUPPORTED; }
    const int bn = (N % 128 == 0) ? 128 : 32;
    const int64_t tiles = (N / bn) * ((M + BM - 1) / BM);
    int splits = 1;
    if (bn == 128) {                                  // fill the SMs: K slices of at least 4 K blocks each
        splits = (int)((int64_t)ctx->sm_count / tiles);
        if (splits > kMaxSplits) splits = kMaxSplits;
        if (splits > (int)(K / BK / 4)) splits = (int)(K / BK / 4);
        if (splits < 1) splits = 1;
    }
    const size_t panel_bytes = (size_t)M * K * 2 * 2;
    const size_t ws_bytes = splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
    int rc = gemm_workspace(ctx, panel_bytes + ws_bytes);
    if (rc) return rc;
    __half* hi = (__half*)ctx->gemm_ws;
    __half* lo = hi + (size_t)M * K;
    float* ws = (float*)((unsigned char*)ctx->gemm_ws + panel_bytes);     // (panel_bytes is a multiple of 256: K % 64 == 0)
    {
        int64_t blocks = ((int64_t)M * K + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
        split_hi_lo_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(X, hi, lo, (int64_t)M * K);
        THK_LAUNCH_CHECK();
    }
    GemmParams p{};
    rc = make_map(encode, &p.map_hi, hi, M, K, BM);
    if (!rc) rc = make_map(encode, &p.map_lo, lo, M, K, BM);
    if (!rc) rc = make_map(encode, &p.map_w, W, N, K, bn);
    if (rc) return rc;
    p.Y = splits > 1 ? ws : Y; p.M = (int)M; p.N = (int)N; p.K = (int)K; p.splits = splits; p.status = ctx->gemm_status;
    dim3 grid((unsigned)(N / bn), (unsigned)((M + BM - 1) / BM), (unsigned)splits);
    if (bn == 128) gemm_f16_tc_kernel<128><<<grid, 128, Tile<128>::kSmem, ctx->stream>>>(p);
    else gemm_f16_tc_kernel<32><<<grid, 128, Tile<32>::kSmem, ctx->stream>>>(p);
    THK_LAUNCH_CHECK();
    if (splits > 1) {
        const int64_t n4 = (int64_t)M * N / 4;
        int64_t blocks = (n4 + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
        splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float4*)ws, (float4*)Y, n4, splits);
        THK_LAUNCH_CHECK();
    }
    return THK_OK;
}

// reports a watchdog abort of the GEMM kernel (bounded mbarrier waits); blocks on the stream
extern "C" int thk_gemm_check(thk_ctx* ctx) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx, "thk_gemm_check: bad ctx");
    if (!ctx->gemm_status) return THK_OK;
    unsigned st = 0;
    THK_CUDA(cudaMemcpyAsync(&st, ctx->gemm_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    THK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (st) { thk_set_error("gemm_f16_tc_kernel aborted: code 0x%x (mbarrier wait timed out)", st); cudaMemsetAsync(ctx->gemm_status, 0, 4, ctx->stream); return THK_E_TIMEOUT; }
    return THK_OK;
}
