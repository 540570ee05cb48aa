// microbench_cvt.cu -- standalone sm_100a measurement for the next optimisation step of decoder.cu (not part of the product).
//
// ncu on the v4 kernel shows the tile body FMA-pipe-bound: per warp and 32 KB tile 64 HADD2.F32 (exact f16 -> f32) + 32
// FFMA2, ~13 warp-cycles each.  Before redesigning the body this measures, per SM and with the kernel's occupancy (8 math
// warps, 2 per sub-partition), the issue cost of the candidates:
//   0  HADD2.F32 convert + FFMA2          (what the kernel does)
//   1  HADD2.F32 convert + scalar FFMA
//   2  convert only (HADD2.F32)
//   3  FFMA2 only
//   4  integer convert (shift/mask into an f32 scaled by 2^-112, activation pre-scaled by 2^112) + FFMA2: moves the
//      conversion from the FMA pipe to the ALU pipe; exact for finite f16 including subnormals
// Output: cycles per 32 KB tile per SM and the implied us/tile at the measured clock.
// (Round 1 ran a first version whose plain shared loads were hoisted out of the loop; it only confirmed the FFMA2 rate:
//  64 FFMA2 per sub-partition and tile in 130 cycles = 2 cycles per warp instruction.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_cvt scripts/microbench_cvt.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float2 h2f2(uint32_t u) { __half2 h = *reinterpret_cast<__half2*>(&u); return __half22float2(h); }
// f16x2 -> two f32 bit patterns equal to value * 2^-112 (exact; subnormals map to f32 subnormals)
__device__ __forceinline__ float2 h2f2_int(uint32_t u) {
    const uint32_t lo = ((u << 13) & 0x0fffe000u) | ((u << 16) & 0x80000000u);
    const uint32_t hi = ((u >> 3) & 0x0fffe000u) | (u & 0x80000000u);
    return make_float2(__uint_as_float(lo), __uint_as_float(hi));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) body(const uint4* __restrict__ w, float* out, int iters, long long* cycles) {
    __shared__ uint4 tile[8 * 256];                       // 8 rows x 256 lanes-worth of weights (32 KB)
    for (int i = threadIdx.x; i < 8 * 256; i += 256) tile[i] = w[i];
    __syncthreads();
    unsigned long long acc[8];
    float accs[8][2];
    for (int r = 0; r < 8; ++r) { acc[r] = 0ull; accs[r][0] = accs[r][1] = 0.f; }
    const unsigned long long xp[4] = {pack2(1.f, 2.f), pack2(3.f, 4.f), pack2(5.f, 6.f), pack2(7.f, 8.f)};
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint4 v[8];
        // volatile shared loads by address, as in the kernel: the compiler must not hoist them (and the conversions that
        // depend on them) out of the loop -- a first version with plain loads measured nothing but the FFMA2 chain
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile) + (((uint32_t)threadIdx.x + (uint32_t)(it & 7)) & 255u) * 16u;
#pragma unroll
        for (int r = 0; r < 8; ++r)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[r].x), "=r"(v[r].y), "=r"(v[r].z), "=r"(v[r].w) : "r"(base + r * 4096u) : "memory");
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint32_t u[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (MODE == 0) { const float2 f = h2f2(u[q]); acc[r] = ffma2(xp[q], pack2(f.x, f.y), acc[r]); }
                if (MODE == 1) { const float2 f = h2f2(u[q]); accs[r][0] = fmaf(1.5f, f.x, accs[r][0]); accs[r][1] = fmaf(2.5f, f.y, accs[r][1]); }
                if (MODE == 2) { const float2 f = h2f2(u[q]); sink ^= __float_as_uint(f.x) ^ __float_as_uint(f.y); }
                if (MODE == 3) { acc[r] = ffma2(xp[q], (unsigned long long)u[q] | 0x3f80000000000000ull, acc[r]); }
                if (MODE == 4) { const float2 f = h2f2_int(u[q]); acc[r] = ffma2(xp[q], pack2(f.x, f.y), acc[r]); }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int r = 0; r < 8; ++r) s += __uint_as_float((uint32_t)acc[r]) + __uint_as_float((uint32_t)(acc[r] >> 32)) + accs[r][0] + accs[r][1];
    out[blockIdx.x * 256 + threadIdx.x] = s + __uint_as_float(sink);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static int run(const char* name, const uint4* w, float* out, long long* cyc, int grid, int khz) {
    const int iters = 2000;
    body<MODE><<<grid, 256>>>(w, out, iters, cyc);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per_tile = (double)mx / iters;           // one "tile" = 8 warps x (8 rows x 256 cols) = 32 KB of weights per SM
    printf("%-44s %8.1f cycles per 32 KB tile per SM  -> %6.2f us/tile at %d MHz (HBM delivers one per ~0.62 us)\n", name, per_tile,
           per_tile / (khz / 1e3), khz / 1000);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    uint4* w; float* out; long long* cyc;
    CK(cudaMalloc(&w, 8 * 256 * sizeof(uint4)));
    CK(cudaMemset(w, 0x3c, 8 * 256 * sizeof(uint4)));     // finite f16 values
    CK(cudaMalloc(&out, grid * 256 * sizeof(float)));
    CK(cudaMalloc(&cyc, 256 * sizeof(long long)));
    printf("device %s, %d SMs\n", prop.name, grid);
    if (run<0>("HADD2.F32 convert + FFMA2 (kernel today)", w, out, cyc, grid, khz)) return 1;
    if (run<1>("HADD2.F32 convert + scalar FFMA", w, out, cyc, grid, khz)) return 1;
    if (run<2>("convert only", w, out, cyc, grid, khz)) return 1;
    if (run<3>("FFMA2 only", w, out, cyc, grid, khz)) return 1;
    if (run<4>("integer convert (ALU pipe) + FFMA2", w, out, cyc, grid, khz)) return 1;
    return 0;
}
