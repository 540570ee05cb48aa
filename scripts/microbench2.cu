// microbench2.cu -- consumer-side and barrier experiments for the persistent decode kernel (not product code).
//   * consumer speed: how fast can the math warps drain 32 KB tiles from shared memory when the producer
//     does not copy at all (pure consumer bound), for 8 math warps (all on every tile) and 16 math warps
//     (two groups of 8 taking alternate tiles), and the same with real copies (HBM bound).
//   * grid barrier variants incl. a flag array without atomics.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench2 scripts/microbench2.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kSlot = 32 * 1024;
constexpr int kMaxSlots = 8;

struct Cfg {
    const unsigned char* base;
    unsigned long long bytes_total;
    int nslots, piece, copy;       // copy 0: producer only signals (consumer-bound run)
    int nphase, reps, tiles[8];
    int barrier;                   // 0 none, 1 counter (decoder.cu), 4 flag array
    int prologue;                  // 0 none, 1 vector + gain + reduce, 3 same with the gain loaded before the barrier
    int group_sync;                // 1: every 2 tiles reduce the accumulators and hand them to the epilogue warp (named barriers)
    float* vec; float* gain;
    unsigned* bar; unsigned* flags;
    unsigned long long* out;
    unsigned* status;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_relaxed_v4(const unsigned* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }

__device__ __forceinline__ void wait_bar(const Cfg& c, uint32_t bar, uint32_t par) {
    if (mbar_try_wait(bar, par)) return;
    unsigned it = 0;
    while (!mbar_try_wait(bar, par)) { if (++it > 20000000u) { atomicExch(c.status, 1u); __trap(); } }
}
// all lanes poll; lanes cover the 148 flags with uint4 loads (37 lanes x 16 B)... generic: n flags
__device__ __forceinline__ void wait_flags(const Cfg& c, const unsigned* flags, int n, unsigned epoch, int lane) {
    unsigned it = 0;
    while (true) {
        bool ok = true;
        for (int i = lane * 4; i < n; i += 128) {
            const uint4 v = ld_relaxed_v4(flags + i);
            ok = ok && (int)(v.x - epoch) >= 0 && (i + 1 >= n || (int)(v.y - epoch) >= 0) && (i + 2 >= n || (int)(v.z - epoch) >= 0) && (i + 3 >= n || (int)(v.w - epoch) >= 0);
        }
        if (__all_sync(0xffffffffu, ok)) break;
        if (++it > 20000000u) { atomicExch(c.status, 2u); __trap(); }
    }
    asm volatile("fence.acquire.gpu;" ::: "memory");
}

__device__ __forceinline__ float2 h2f2(uint32_t u) { __half2 h = *reinterpret_cast<__half2*>(&u); return __half22float2(h); }
__device__ __forceinline__ void fma8(const uint4& w, const float4& x0, const float4& x1, float& a0, float& a1) {
    float2 f;
    f = h2f2(w.x); a0 = fmaf(x0.x, f.x, a0); a1 = fmaf(x0.y, f.y, a1);
    f = h2f2(w.y); a0 = fmaf(x0.z, f.x, a0); a1 = fmaf(x0.w, f.y, a1);
    f = h2f2(w.z); a0 = fmaf(x1.x, f.x, a0); a1 = fmaf(x1.y, f.y, a1);
    f = h2f2(w.w); a0 = fmaf(x1.z, f.x, a0); a1 = fmaf(x1.w, f.y, a1);
}

template <int MW>
struct SmemT {
    unsigned long long full[kMaxSlots], empty[kMaxSlots];
    float part[MW];
    float red[2][MW][8];
    float xs[4096];
};

template <int MW>
__global__ void __launch_bounds__(64 + MW * 32, 1) k_stream(const __grid_constant__ Cfg c) {
    constexpr int MT = MW * 32, NT = 64 + MT, G = MW / 8;
    extern __shared__ __align__(1024) unsigned char smem[];
    SmemT<MW>* sm = (SmemT<MW>*)(smem + (size_t)c.nslots * kSlot);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < c.nslots; ++i) { mbar_init(smem_u32(&sm->full[i]), 1); mbar_init(smem_u32(&sm->empty[i]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 4096; i += NT) sm->xs[i] = 1.0f;
    __syncthreads();
    const unsigned long long t_start = gtimer();
    const uint32_t slot_base = smem_u32(smem), full_base = smem_u32(&sm->full[0]), empty_base = smem_u32(&sm->empty[0]);
    const unsigned nb = gridDim.x, b = blockIdx.x;
    const unsigned long long per_cta = (c.bytes_total / nb) & ~(unsigned long long)(kSlot - 1);

    if (tid < 32) {
        uint32_t sl = 0, par = 0;
        unsigned long long off = 0;
        for (int rep = 0; rep < c.reps; ++rep)
            for (int ph = 0; ph < c.nphase; ++ph)
                for (int t = 0; t < c.tiles[ph]; ++t) {
                    const uint32_t eb = empty_base + sl * 8, fb = full_base + sl * 8;
                    wait_bar(c, eb, par ^ 1u);
                    if (c.copy) {
                        if (lane == 0) mbar_expect_tx(fb, kSlot);
                        __syncwarp();
                        const unsigned char* src = c.base + b * per_cta + (off % per_cta);
                        const uint32_t dst = slot_base + sl * kSlot;
                        for (uint32_t o = (uint32_t)lane * c.piece; o < (uint32_t)kSlot; o += 32u * c.piece) bulk_g2s(dst + o, src + o, (uint32_t)c.piece, fb);
                    } else {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(fb);
                    }
                    off += kSlot;
                    if (++sl == (uint32_t)c.nslots) { sl = 0; par ^= 1u; }
                }
    } else if (tid < 32 + MT) {
        const int ct = tid - 32, w = ct >> 5, grp = w >> 3, cw = w & 7;
        uint32_t sl = 0, par = 0;
        unsigned nbar = 0, tcount = 0, gq = 0;
        float acc[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
        const uint32_t xs_a = smem_u32(&sm->xs[0]) + (uint32_t)(cw * 256 + lane * 4) * 4;
        for (int rep = 0; rep < c.reps; ++rep)
            for (int ph = 0; ph < c.nphase; ++ph) {
                for (int t = 0; t < c.tiles[ph]; ++t, ++tcount) {
                    if ((int)(tcount % G) == grp) {
                        wait_bar(c, full_base + sl * 8, par);
                        const uint4 xr0 = lds128(xs_a), xr1 = lds128(xs_a + 512);
                        const float4 x0 = *(const float4*)&xr0, x1 = *(const float4*)&xr1;
                        const uint32_t wa = slot_base + sl * kSlot + (uint32_t)(cw * 256 + lane * 8) * 2;
                        uint4 wv[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) wv[r] = lds128(wa + (uint32_t)r * 4096);
#pragma unroll
                        for (int r = 0; r < 8; ++r) fma8(wv[r], x0, x1, acc[r][0], acc[r][1]);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty_base + sl * 8);
                    }
                    if (++sl == (uint32_t)c.nslots) { sl = 0; par ^= 1u; }
                    if (c.group_sync && (t & 1) == 1) {
                        // row group finished (2 K tiles): transposing reduction, hand the 8 row sums to the epilogue warp
                        float v[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) { v[r] = acc[r][0] + acc[r][1]; acc[r][0] = 0.f; acc[r][1] = 0.f; }
                        int off = 16;
#pragma unroll
                        for (int n = 8; n > 1; n >>= 1) {
                            const int half = n >> 1;
                            const bool upper = (lane & off) != 0;
#pragma unroll
                            for (int i = 0; i < half; ++i) {
                                const float send = upper ? v[i] : v[i + half];
                                const float keep = upper ? v[i + half] : v[i];
                                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                            off >>= 1;
                        }
                        float rs = v[0];
                        rs += __shfl_xor_sync(0xffffffffu, rs, 2);
                        rs += __shfl_xor_sync(0xffffffffu, rs, 1);
                        const int buf = gq & 1;
                        if (gq >= 2) bar_sync(5 + buf, MT + 32);
                        if ((lane & 3) == 0) sm->red[buf][w][lane >> 2] = rs;
                        bar_arrive(3 + buf, MT + 32);
                        ++gq;
                    }
                }
                if (c.group_sync) { for (unsigned k = gq >= 2 ? gq - 2 : 0; k < gq; ++k) bar_sync(5 + (k & 1), MT + 32); gq = 0; }
                // -------- phase transition --------
                ++nbar;
                constexpr int KH = 1024 / MT > 0 ? 1024 / MT : 1;     // float4s per thread
                float4 g4[KH];
                if (c.prologue == 3) {
#pragma unroll
                    for (int k = 0; k < KH; ++k) if (ct + k * MT < 1024) g4[k] = __ldg((const float4*)c.gain + ct + k * MT);
                }
                if (c.barrier == 1) {
                    bar_sync(1, MT + 32);
                    if (ct == 0) {
                        red_release_add(c.bar, 1u);
                        const unsigned target = nbar * nb; unsigned it = 0;
                        while (ld_acquire_u32(c.bar) < target) { if (++it > 20000000u) { atomicExch(c.status, 3u); __trap(); } }
                    }
                    bar_sync(1, MT + 32);
                } else if (c.barrier == 4) {
                    bar_sync(1, MT + 32);                          // epilogue warp's stores are done
                    if (w == 0) {
                        if (lane == 0) st_release_u32(c.flags + b, nbar);
                        wait_flags(c, c.flags, (int)nb, nbar, lane);
                    }
                    bar_sync(1, MT + 32);
                }
                if (c.prologue) {
                    float4 v[KH];
                    float ss = 0.f;
#pragma unroll
                    for (int k = 0; k < KH; ++k) if (ct + k * MT < 1024) v[k] = __ldcg((const float4*)(c.vec) + ct + k * MT);
                    if (c.prologue == 1) {
#pragma unroll
                        for (int k = 0; k < KH; ++k) if (ct + k * MT < 1024) g4[k] = __ldg((const float4*)c.gain + ct + k * MT);
                    }
#pragma unroll
                    for (int k = 0; k < KH; ++k) if (ct + k * MT < 1024) ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                    if (lane == 0) sm->part[w] = ss;
                    bar_sync(2, MT);
                    float tot = 0.f;
#pragma unroll
                    for (int i = 0; i < MW; ++i) tot += sm->part[i];
                    const float inv = rsqrtf(tot / 4096.f + 1e-6f);
#pragma unroll
                    for (int k = 0; k < KH; ++k) if (ct + k * MT < 1024) {
                        float4 o = v[k]; const float4 gg = g4[k];
                        o.x = o.x * inv * gg.x; o.y = o.y * inv * gg.y; o.z = o.z * inv * gg.z; o.w = o.w * inv * gg.w;
                        *((float4*)sm->xs + ct + k * MT) = o;
                    }
                    bar_sync(2, MT);
                }
            }
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += acc[r][0] + acc[r][1];
        if (s == 123.456f) c.out[b * 4 + 3] = 1;
    } else {
        // epilogue warp
        unsigned nbar = 0;
        for (int rep = 0; rep < c.reps; ++rep)
            for (int ph = 0; ph < c.nphase; ++ph) {
                ++nbar;
                if (c.group_sync) {
                    unsigned gq = 0;
                    for (int t = 1; t < c.tiles[ph]; t += 2, ++gq) {
                        const int buf = gq & 1;
                        bar_sync(3 + buf, MT + 32);
                        float y = 0.f;
                        if (lane < 8) { for (int i = 0; i < MW; ++i) y += sm->red[buf][i][lane]; c.vec[(b * 8 + lane) & 4095] = 1.0f + y * 1e-30f; }
                        __syncwarp();
                        bar_arrive(5 + buf, MT + 32);
                    }
                } else if (c.barrier) {
                    for (int i = b * 32 + lane; i < 4096; i += nb * 32) c.vec[i] = 1.0f + (float)(nbar & 3);
                }
                if (c.barrier) { bar_sync(1, MT + 32); bar_sync(1, MT + 32); }
            }
    }
    if (tid == 32) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        c.out[b * 4 + 0] = t_start; c.out[b * 4 + 1] = gtimer(); c.out[b * 4 + 2] = smid;
    }
}

__global__ void __launch_bounds__(320, 1) barrier_kernel(unsigned* ctr, unsigned* flags, int n, int variant) {
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned nb = gridDim.x, b = blockIdx.x;
    Cfg c{}; c.status = ctr + 200;
    for (int i = 1; i <= n; ++i) {
        __syncthreads();
        if (variant == 1) {
            if (tid == 0) { red_release_add(ctr, 1u); unsigned it = 0; while (ld_acquire_u32(ctr) < (unsigned)i * nb) { if (++it > 20000000u) __trap(); } }
        } else if (variant == 4) {          // flag array, one warp polls with 128-bit loads
            if (tid < 32) { if (lane == 0) st_release_u32(flags + b, (unsigned)i); wait_flags(c, flags, (int)nb, (unsigned)i, lane); }
        } else if (variant == 5) {          // counter, but arrivals spread over 4 counters in different lines; poller sums them
            if (tid == 0) {
                red_release_add(ctr + (b & 3) * 32, 1u);
                unsigned it = 0;
                while (true) {
                    const unsigned s = ld_relaxed_u32(ctr) + ld_relaxed_u32(ctr + 32) + ld_relaxed_u32(ctr + 64) + ld_relaxed_u32(ctr + 96);
                    if (s >= (unsigned)i * nb) break;
                    if (++it > 20000000u) __trap();
                }
                asm volatile("fence.acquire.gpu;" ::: "memory");
            }
        } else if (variant == 6) {          // counter with relaxed red after an explicit release fence, relaxed polling
            if (tid == 0) {
                asm volatile("fence.release.gpu;" ::: "memory");
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
                unsigned it = 0; while (ld_relaxed_u32(ctr) < (unsigned)i * nb) { if (++it > 20000000u) __trap(); }
                asm volatile("fence.acquire.gpu;" ::: "memory");
            }
        }
        __syncthreads();
    }
}

__global__ void fill_kernel(uint4* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount;
    const unsigned long long bytes = 12ull << 30;
    unsigned char* base; CK(cudaMalloc(&base, bytes));
    fill_kernel<<<grid * 8, 256>>>((uint4*)base, bytes / 16); CK(cudaDeviceSynchronize());
    float *vec, *gain; CK(cudaMalloc(&vec, 4096 * 4)); CK(cudaMemset(vec, 0, 4096 * 4)); CK(cudaMalloc(&gain, 4096 * 4)); CK(cudaMemset(gain, 0, 4096 * 4));
    unsigned* ctl; CK(cudaMalloc(&ctl, 8192 * 4));
    unsigned long long* out; CK(cudaMalloc(&out, grid * 4 * 8));
    std::vector<unsigned long long> h(grid * 4);

    for (int variant : {1, 4, 5, 6}) {
        CK(cudaMemset(ctl, 0, 8192 * 4));
        unsigned* flags = ctl + 1024; int n = 2000;
        void* args[] = {&ctl, &flags, &n, &variant};
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(grid), dim3(320), args, 0, 0));
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("barrier variant %d: %.3f us per barrier\n", variant, ms * 1e3 / n);
    }

    auto run = [&](const char* name, int mw, int copy, int barrier, int prologue, int group_sync, int nslots, bool phases) {
        Cfg c{};
        c.base = base; c.bytes_total = bytes; c.nslots = nslots; c.piece = 4096; c.copy = copy;
        const double mb[5] = {100.66e6, 16.8e6, 33.55e6, 180.35e6, 90.18e6};
        unsigned long long tot = 0;
        if (phases) { c.nphase = 5; c.reps = 32; for (int i = 0; i < 5; ++i) { c.tiles[i] = ((int)(mb[i] / grid / kSlot + 0.5) + 1) & ~1; tot += c.tiles[i]; } tot *= 32; }
        else { c.nphase = 1; c.reps = 1; c.tiles[0] = (int)(13.2e9 / grid / kSlot) & ~1; tot = c.tiles[0]; }
        c.barrier = barrier; c.prologue = prologue; c.group_sync = group_sync;
        c.vec = vec; c.gain = gain; c.bar = ctl; c.flags = ctl + 1024; c.out = out; c.status = ctl + 512;
        const size_t smem = (size_t)nslots * kSlot + (mw == 8 ? sizeof(SmemT<8>) : sizeof(SmemT<16>)) + 1024;
        void* fn = mw == 8 ? (void*)k_stream<8> : (void*)k_stream<16>;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        double best = 1e30;
        for (int it = 0; it < 4; ++it) {
            CK(cudaMemset(ctl, 0, 8192 * 4));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            void* args[] = {&c};
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(64 + mw * 32), args, smem, 0));
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 && ms < best) best = ms;
            CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
        }
        const double gb = (double)tot * kSlot * grid / 1e9;
        printf("%-58s %8.3f ms  %8.1f GB/s  (%.2f GB)\n", name, best, gb / best * 1e3, gb);
        fflush(stdout);
    };
    //   name                                                   mw copy bar pro gsync slots phases
    run("consumer only, 8 warps",                                 8, 0, 0, 0, 0, 5, false);
    run("consumer only, 16 warps (2 groups, alternate tiles)",   16, 0, 0, 0, 0, 5, false);
    run("consumer only, 8 warps, group handoff",                  8, 0, 0, 0, 1, 5, false);
    run("consumer only, 16 warps, group handoff",                16, 0, 0, 0, 1, 5, false);
    run("stream, 8 warps",                                        8, 1, 0, 0, 0, 5, false);
    run("stream, 16 warps",                                      16, 1, 0, 0, 0, 5, false);
    run("stream, 16 warps, 4 slots",                             16, 1, 0, 0, 0, 4, false);
    run("stream, 8 warps, group handoff",                         8, 1, 0, 0, 1, 5, false);
    run("stream, 16 warps, group handoff",                       16, 1, 0, 0, 1, 5, false);
    run("phases, 8w, barrier 1, prologue 1, handoff",             8, 1, 1, 1, 1, 5, true);
    run("phases, 16w, barrier 1, prologue 1, handoff",           16, 1, 1, 1, 1, 5, true);
    run("phases, 16w, barrier 1, prologue 3 (gain early)",       16, 1, 1, 3, 1, 5, true);
    run("phases, 16w, barrier 4 (flags), prologue 3",            16, 1, 4, 3, 1, 5, true);
    run("phases, 8w, barrier 4 (flags), prologue 3",              8, 1, 4, 3, 1, 5, true);
    run("phases, 16w, barrier 4, prologue 3, 4 slots",           16, 1, 4, 3, 1, 4, true);
    run("phases, 16w, barrier 4, prologue 3, 6 slots",           16, 1, 4, 3, 1, 6, true);
    run("phases, 16w, no barrier, prologue 3",                   16, 1, 0, 3, 1, 5, true);
    run("phases consumer only, 16w, barrier 4, prologue 3",      16, 0, 4, 3, 1, 5, true);
    run("phases consumer only, 8w, barrier 1, prologue 1",        8, 0, 1, 1, 1, 5, true);
    return 0;
}
