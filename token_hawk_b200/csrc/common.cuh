// common.cuh -- shared host/device helpers for libthk_sm100a.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "thk_cabi.h"

struct thk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    // tensor-core GEMM (gemm_tc.cu): hi/lo split workspace and watchdog status word of THIS context -- two contexts on one
    // device (or two models on different streams) never share operands in flight
    void* gemm_ws = nullptr;
    size_t gemm_ws_bytes = 0;
    unsigned* gemm_status = nullptr;
    bool gemm_attr_set = false;       // cudaFuncSetAttribute is per device; a context belongs to one device
};

void thk_set_error(const char* fmt, ...);

// every entry point makes its context's device current (several contexts may live in one process)
#define THK_ENTER(ctx) do { if (ctx) cudaSetDevice((ctx)->device); } while (0)

#define THK_CHECK_ARG(cond, ...)                                    \
    do {                                                            \
        if (!(cond)) { thk_set_error(__VA_ARGS__); return THK_E_INVALID; } \
    } while (0)

#define THK_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            thk_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return THK_E_CUDA;                                                              \
        }                                                                                   \
    } while (0)

#define THK_LAUNCH_CHECK()                                                                   \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            thk_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return THK_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

// ---- device helpers ----
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {  // 128-bit streaming load, no L1 allocate
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t u) {   // exact f16x2 -> f32x2 (== th.cpp:312-333)
    __half2 h = *reinterpret_cast<__half2*>(&u);
    return __half22float2(h);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// counter PRNG shared with oracle/th_oracle.c (tho_hash)
__host__ __device__ __forceinline__ uint64_t thk_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t thk_hash(uint64_t seed, uint64_t tid, uint64_t idx) {
    return thk_splitmix64(thk_splitmix64(seed ^ (tid * 0xD1B54A32D192ED03ull)) + idx);
}
