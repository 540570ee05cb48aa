// microbench.cu -- standalone sm_100a measurements that bound the persistent decode kernel's design.
// Not part of the product: a synthetic skeleton of decoder.cu (producer warp + cp.async.bulk ring +
// math warps + phase transitions) whose features can be switched on and off one by one, so that one
// GPU run shows what each of them costs under the real HBM load:
//   * pure-read HBM peak through the ring (no phases)                      -> the real roofline
//   * + grid barrier per phase (variants)    * + prologue (16 KB vector from L2, reduce, stage)
//   * static vs dynamic (atomic-claimed) tile ownership                    -> arrival skew
//   * L2 look-ahead prefetch, ring geometry, math on/off
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench scripts/microbench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kMathWarps = 8;
constexpr int kMathThreads = kMathWarps * 32;
constexpr int kThreads = 32 + kMathThreads + 32;   // producer, math, epilogue (same shape as decoder.cu)
constexpr int kMaxSlots = 12;

struct Cfg {
    const unsigned char* base;     // weights
    unsigned long long bytes_total;
    int slot_bytes, nslots, piece;
    int nphase;                    // phases in the pattern (repeated `reps` times)
    int tiles[8];                  // tiles per CTA per phase (static) ; dynamic: total units = tiles*grid
    int reps;
    int barrier;                   // 0 none, 1 decoder.cu barrier, 2 split arrive/wait (epilogue warp arrives, math warps poll)
    int prologue;                  // 0 none, 1 load 16 KB vector + reduce + stage, 2 no reduce (deferred norm)
    int math;                      // 0 consumers only release slots, 1 generic-LD fma, 2 ld.shared fma
    int dynamic;                   // 0 static contiguous ranges, 1 atomic-claimed units of `unit_tiles` tiles
    int unit_tiles;
    int l2_ahead_kb;               // 0 off; else producer prefetches this much beyond the ring when it finds the ring full
    float* vec;                    // 4096-float vector rewritten every phase (epilogue) and read in the prologue
    unsigned* bar;                 // barrier counter
    unsigned* claim;               // per (rep, phase) claim counters
    unsigned long long* out;       // per CTA: t_start, t_end, smid, checksum bits
    unsigned long long timeout_ns;
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
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }

struct Smem {
    unsigned long long full[kMaxSlots], empty[kMaxSlots];
    int meta[kMaxSlots];          // 0 data tile, 1 end of phase, 2 end of stream
    float part[kMathWarps];
    float xs[4096];
};

__device__ __forceinline__ bool wait_bar(const Cfg& c, uint32_t bar, uint32_t par) {
    if (mbar_try_wait(bar, par)) return true;
    const unsigned long long t0 = gtimer();
    unsigned it = 0;
    while (!mbar_try_wait(bar, par)) {
        if ((++it & 255u) == 0u) {
            if (*(volatile unsigned*)c.status) return false;
            if (gtimer() - t0 > c.timeout_ns) { atomicExch(c.status, 1u); __trap(); }
        }
    }
    return true;
}

__device__ __forceinline__ float2 h2f2(uint32_t u) { __half2 h = *reinterpret_cast<__half2*>(&u); return __half22float2(h); }
__device__ __forceinline__ void fma8(const uint4& w, const float4& x0, const float4& x1, float& a0, float& a1) {
    float2 f;
    f = h2f2(w.x); a0 = fmaf(x0.x, f.x, a0); a1 = fmaf(x0.y, f.y, a1);
    f = h2f2(w.y); a0 = fmaf(x0.z, f.x, a0); a1 = fmaf(x0.w, f.y, a1);
    f = h2f2(w.z); a0 = fmaf(x1.x, f.x, a0); a1 = fmaf(x1.y, f.y, a1);
    f = h2f2(w.w); a0 = fmaf(x1.z, f.x, a0); a1 = fmaf(x1.w, f.y, a1);
}

__global__ void __launch_bounds__(kThreads, 1) stream_kernel(const __grid_constant__ Cfg c) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* slots = smem;
    Smem* sm = (Smem*)(smem + (size_t)c.nslots * c.slot_bytes);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < c.nslots; ++i) { mbar_init(smem_u32(&sm->full[i]), 1); mbar_init(smem_u32(&sm->empty[i]), kMathWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 4096; i += kThreads) sm->xs[i] = 1.0f;
    __syncthreads();
    const unsigned long long t_start = gtimer();
    const uint32_t slot_base = smem_u32(slots), full_base = smem_u32(&sm->full[0]), empty_base = smem_u32(&sm->empty[0]);
    const unsigned nb = gridDim.x, b = blockIdx.x;
    const unsigned long long per_cta = (c.bytes_total / nb) & ~(unsigned long long)(c.slot_bytes - 1);

    if (tid < 32) {
        // ---------------- producer ----------------
        uint32_t sl = 0, par = 0;
        unsigned long long off_static = 0;                 // static: running offset inside this CTA's region
        bool ok = true;
        unsigned long long pf_issued = 0;                  // static mode L2 look-ahead cursor (absolute offset in region)
        auto push = [&](int meta, const unsigned char* src) {
            const uint32_t eb = empty_base + sl * 8, fb = full_base + sl * 8;
            if (c.l2_ahead_kb && !c.dynamic && meta == 0 && !mbar_try_wait(eb, par ^ 1u)) {
                // ring full: ask L2 for more of this CTA's stream beyond the ring
                const unsigned long long lo = off_static + (unsigned long long)c.nslots * c.slot_bytes;
                if (pf_issued < lo) pf_issued = lo;
                if (pf_issued < lo + (unsigned long long)c.l2_ahead_kb * 1024ull && pf_issued + 32768 <= per_cta) {
                    if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c.base + b * per_cta + pf_issued), "r"(32768u) : "memory");
                    pf_issued += 32768;
                }
            }
            if (ok) ok = wait_bar(c, eb, par ^ 1u);
            if (!ok) return;
            if (lane == 0) sm->meta[sl] = meta;
            if (meta == 0) {
                if (lane == 0) mbar_expect_tx(fb, (uint32_t)c.slot_bytes);
                __syncwarp();
                const uint32_t dst = slot_base + sl * c.slot_bytes;
                for (uint32_t o = (uint32_t)lane * c.piece; o < (uint32_t)c.slot_bytes; o += 32u * c.piece)
                    bulk_g2s(dst + o, src + o, (uint32_t)c.piece, fb);
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(fb);            // sentinel: completes the phase of the full barrier without data
            }
            if (++sl == (uint32_t)c.nslots) { sl = 0; par ^= 1u; }
        };
        for (int rep = 0; rep < c.reps && ok; ++rep) {
            for (int ph = 0; ph < c.nphase && ok; ++ph) {
                if (!c.dynamic) {
                    for (int t = 0; t < c.tiles[ph] && ok; ++t) {
                        const unsigned long long o = off_static % per_cta;
                        push(0, c.base + b * per_cta + o);
                        off_static += c.slot_bytes;
                    }
                } else {
                    // units of unit_tiles tiles claimed from a per-phase counter; the phase's units tile the whole buffer region
                    const unsigned units = (unsigned)((long long)c.tiles[ph] * nb / c.unit_tiles);
                    unsigned* ctr = c.claim + (rep * c.nphase + ph);
                    unsigned long long phase_base = ((unsigned long long)(rep * c.nphase + ph) * 0x9E3779B1ull * 65536ull) % (c.bytes_total / 2);
                    phase_base &= ~(unsigned long long)(c.slot_bytes - 1);
                    unsigned u = 0;
                    if (lane == 0) u = atomicAdd(ctr, 1u);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    while (u < units && ok) {
                        unsigned un = 0;
                        if (lane == 0) un = atomicAdd(ctr, 1u);                 // claim ahead: latency hidden behind this unit's loads
                        for (int t = 0; t < c.unit_tiles && ok; ++t) {
                            unsigned long long o = phase_base + ((unsigned long long)u * c.unit_tiles + t) * c.slot_bytes;
                            o %= (c.bytes_total - c.slot_bytes);
                            o &= ~(unsigned long long)1023;
                            push(0, c.base + o);
                        }
                        u = __shfl_sync(0xffffffffu, un, 0);
                    }
                }
                push(1, nullptr);
            }
        }
        push(2, nullptr);
    } else if (tid < 32 + kMathThreads) {
        // ---------------- math warps ----------------
        const int ct = tid - 32, cw = ct >> 5;
        uint32_t sl = 0, par = 0;
        unsigned nbar = 0;
        float acc[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
        bool ok = true;
        const uint32_t xs_a = smem_u32(&sm->xs[0]);
        while (true) {
            if (ok) ok = wait_bar(c, full_base + sl * 8, par);
            if (!ok) break;
            const int meta = sm->meta[sl];
            if (meta == 0 && c.math) {
                // 8 rows x 2048 cols tile; warp cw owns 256 columns; lane owns 8 consecutive columns (as tile_fma RPW=8, CPW=1)
                const int ncols = c.slot_bytes / 16;      // columns per row when the slot holds 8 rows of f16
                for (int cb = cw * 256; cb < ncols; cb += kMathWarps * 256) {
                    const int col = cb + lane * 8;
                    if (c.math == 1) {
                        const float4 x0 = *(const float4*)(sm->xs + ((cb & 4095 & ~255) + lane * 4));
                        const float4 x1 = *(const float4*)(sm->xs + ((cb & 4095 & ~255) + 128 + lane * 4));
                        const unsigned char* wp = slots + (size_t)sl * c.slot_bytes + (size_t)col * 2;
                        uint4 w[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) w[r] = *(const uint4*)(wp + (size_t)r * ncols * 2);
#pragma unroll
                        for (int r = 0; r < 8; ++r) fma8(w[r], x0, x1, acc[r][0], acc[r][1]);
                    } else {
                        const uint32_t xa = xs_a + (uint32_t)(((cb & 4095 & ~255) + lane * 4) * 4);
                        const uint4 xr0 = lds128(xa), xr1 = lds128(xa + 512);
                        const float4 x0 = *(const float4*)&xr0, x1 = *(const float4*)&xr1;
                        const uint32_t wa = slot_base + sl * c.slot_bytes + (uint32_t)col * 2;
                        uint4 w[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) w[r] = lds128(wa + (uint32_t)r * ncols * 2);
#pragma unroll
                        for (int r = 0; r < 8; ++r) fma8(w[r], x0, x1, acc[r][0], acc[r][1]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_base + sl * 8);
            if (++sl == (uint32_t)c.nslots) { sl = 0; par ^= 1u; }
            if (meta == 2) break;
            if (meta == 1) {
                // -------- phase transition --------
                ++nbar;
                if (c.barrier == 1) {
                    bar_sync(1, kMathThreads + 32);
                    if (ct == 0) {
                        red_release_add(c.bar, 1u);
                        const unsigned target = nbar * nb;
                        unsigned it = 0; unsigned long long t0 = 0;
                        while (ld_acquire_u32(c.bar) < target) {
                            if ((++it & 63u) == 0u) { if (!t0) t0 = gtimer(); if (*(volatile unsigned*)c.status) break; if (gtimer() - t0 > c.timeout_ns) { atomicExch(c.status, 2u); __trap(); } }
                        }
                    }
                    bar_sync(1, kMathThreads + 32);
                } else if (c.barrier == 2) {
                    // the epilogue warp arrives for the CTA; every math warp polls for itself (no CTA-wide bar.sync)
                    bar_sync(2, kMathThreads + 32);            // hand the phase's partials to the epilogue warp (it does the global writes)
                    const unsigned target = nbar * nb;
                    if (lane == 0) {
                        unsigned it = 0; unsigned long long t0 = 0;
                        while (ld_acquire_u32(c.bar) < target) {
                            if ((++it & 63u) == 0u) { if (!t0) t0 = gtimer(); if (*(volatile unsigned*)c.status) break; if (gtimer() - t0 > c.timeout_ns) { atomicExch(c.status, 2u); __trap(); } }
                        }
                    }
                    __syncwarp();
                }
                if (c.prologue) {
                    float4 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = __ldcg((const float4*)(c.vec) + ct + k * kMathThreads);
                    if (c.prologue == 1) {
                        float ss = 0.f;
#pragma unroll
                        for (int k = 0; k < 4; ++k) ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                        if (lane == 0) sm->part[cw] = ss;
                        bar_sync(3, kMathThreads);
                        float tot = 0.f;
#pragma unroll
                        for (int w = 0; w < kMathWarps; ++w) tot += sm->part[w];
                        const float inv = rsqrtf(tot / 4096.f + 1e-6f);
#pragma unroll
                        for (int k = 0; k < 4; ++k) { v[k].x *= inv; v[k].y *= inv; v[k].z *= inv; v[k].w *= inv; }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) *((float4*)sm->xs + ct + k * kMathThreads) = v[k];
                    bar_sync(3, kMathThreads);
                }
            }
        }
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += acc[r][0] + acc[r][1];
        if (s == 123.456f) c.out[b * 4 + 3] = 1;          // keep the math alive
    } else {
        // ---------------- epilogue warp ----------------
        if (c.barrier) {
            unsigned nbar = 0;
            for (int rep = 0; rep < c.reps; ++rep)
                for (int ph = 0; ph < c.nphase; ++ph) {
                    ++nbar;
                    if (c.barrier == 1) {
                        // writes this CTA's share of the vector, then joins the CTA-wide barrier
                        for (int i = b * 32 + lane; i < 4096; i += nb * 32) c.vec[i] = 1.0f + (float)(nbar & 3);
                        bar_sync(1, kMathThreads + 32);
                        bar_sync(1, kMathThreads + 32);
                    } else {
                        bar_sync(2, kMathThreads + 32);
                        for (int i = b * 32 + lane; i < 4096; i += nb * 32) c.vec[i] = 1.0f + (float)(nbar & 3);
                        __syncwarp();
                        if (lane == 0) red_release_add(c.bar, 1u);
                    }
                    if (*(volatile unsigned*)c.status) return;
                }
        }
    }
    if (tid == 32) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        c.out[b * 4 + 0] = t_start; c.out[b * 4 + 1] = gtimer(); c.out[b * 4 + 2] = smid;
    }
}

// pure barrier latency (no loads): nbar barriers back to back, variant as above (1 = decoder.cu style)
__global__ void __launch_bounds__(kThreads, 1) barrier_kernel(unsigned* ctr, int n, int variant, unsigned long long* out) {
    const int tid = threadIdx.x;
    const unsigned nb = gridDim.x;
    const unsigned long long t0 = gtimer();
    for (int i = 1; i <= n; ++i) {
        if (variant == 1) {
            __syncthreads();
            if (tid == 0) { red_release_add(ctr, 1u); { unsigned it = 0; while (ld_acquire_u32(ctr) < (unsigned)i * nb) { if (++it > 50000000u) __trap(); } } }
            __syncthreads();
        } else if (variant == 2) {       // relaxed polling + one acquire fence at the end
            __syncthreads();
            if (tid == 0) { red_release_add(ctr, 1u); { unsigned it = 0; while (ld_relaxed_u32(ctr) < (unsigned)i * nb) { if (++it > 50000000u) __trap(); } } asm volatile("fence.acquire.gpu;" ::: "memory"); }
            __syncthreads();
        } else {                         // every warp polls; one thread arrives
            __syncthreads();
            if (tid == 0) red_release_add(ctr, 1u);
            if ((tid & 31) == 0) { unsigned it = 0; while (ld_acquire_u32(ctr) < (unsigned)i * nb) { if (++it > 50000000u) __trap(); } }
            __syncwarp();
        }
    }
    if (tid == 0) { out[blockIdx.x * 2] = t0; out[blockIdx.x * 2 + 1] = gtimer(); }
}

__global__ void timer_res_kernel(unsigned long long* out) {
    unsigned long long prev = gtimer(), mind = ~0ull, maxd = 0; int changes = 0;
    const long long c0 = clock64();
    for (int i = 0; i < 200000; ++i) { const unsigned long long t = gtimer(); if (t != prev) { const unsigned long long d = t - prev; if (d < mind) mind = d; if (d > maxd) maxd = d; prev = t; ++changes; } }
    out[0] = mind; out[1] = maxd; out[2] = changes; out[3] = (unsigned long long)(clock64() - c0);
}

__global__ void fill_kernel(uint4* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
}

struct Result { double ms, gbs; double cta_min, cta_med, cta_max; };

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int grid = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, grid, prop.clockRate);
    const unsigned long long bytes = 12ull << 30;
    unsigned char* base; CK(cudaMalloc(&base, bytes));
    fill_kernel<<<grid * 8, 256>>>((uint4*)base, bytes / 16); CK(cudaDeviceSynchronize());
    float* vec; CK(cudaMalloc(&vec, 4096 * 4)); CK(cudaMemset(vec, 0, 4096 * 4));
    unsigned* ctl; CK(cudaMalloc(&ctl, 4096 * 4));
    unsigned long long* out; CK(cudaMalloc(&out, grid * 4 * 8));
    std::vector<unsigned long long> h(grid * 4);

    {   // globaltimer resolution
        timer_res_kernel<<<1, 1>>>(out); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, 32, cudaMemcpyDeviceToHost));
        printf("globaltimer: min step %llu ns, max step %llu ns, %llu changes over %llu cycles\n", h[0], h[1], h[2], h[3]);
    }
    for (int variant = 1; variant <= 3; ++variant) {   // pure barrier latency
        CK(cudaMemset(ctl, 0, 4096 * 4));
        void* args[] = {&ctl, nullptr, &variant, &out};
        int n = 2000; args[1] = &n;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(grid), dim3(kThreads), args, 0, 0));
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("barrier variant %d: %.3f us per barrier (%d barriers, %d CTAs x %d threads)\n", variant, ms * 1e3 / n, n, grid, kThreads);
    }

    // LLaMA-7B phase pattern in 32 KB tiles per CTA: QKV 100.7 MB, att 16.8 MB, Wo 33.5 MB, W13 180 MB, W2 90 MB
    auto run = [&](const char* name, int barrier, int prologue, int math, int dynamic, int unit_tiles, int l2kb, int slot_kb, int nslots, int piece, bool phases) {
        Cfg c{};
        c.base = base; c.bytes_total = bytes; c.slot_bytes = slot_kb * 1024; c.nslots = nslots; c.piece = piece;
        const double mb[5] = {100.66e6, 16.8e6, 33.55e6, 180.35e6, 90.18e6};
        unsigned long long tot_tiles = 0;
        if (phases) {
            c.nphase = 5; c.reps = 32;
            for (int i = 0; i < 5; ++i) { c.tiles[i] = (int)(mb[i] / grid / c.slot_bytes + 0.5); if (dynamic) c.tiles[i] = ((c.tiles[i] + unit_tiles - 1) / unit_tiles) * unit_tiles; tot_tiles += c.tiles[i]; }
            tot_tiles *= 32;
        } else {
            c.nphase = 1; c.reps = 1; c.tiles[0] = (int)(13.2e9 / grid / c.slot_bytes); tot_tiles = c.tiles[0];
        }
        c.barrier = barrier; c.prologue = prologue; c.math = math; c.dynamic = dynamic; c.unit_tiles = unit_tiles; c.l2_ahead_kb = l2kb;
        c.vec = vec; c.bar = ctl; c.claim = ctl + 16; c.out = out; c.timeout_ns = 2000000000ull; c.status = ctl + 8;
        const size_t smem = (size_t)nslots * c.slot_bytes + sizeof(Smem) + 1024;
        if (smem > 227 * 1024) { printf("%-44s skipped (smem %zu)\n", name, smem); return; }
        CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        double best = 1e30; std::vector<double> spans;
        for (int it = 0; it < 4; ++it) {
            CK(cudaMemset(ctl, 0, 4096 * 4));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            void* args[] = {&c};
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((void*)stream_kernel, dim3(grid), dim3(kThreads), args, smem, 0));
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            unsigned st[2]; CK(cudaMemcpy(st, ctl + 8, 4, cudaMemcpyDeviceToHost));
            if (st[0]) { printf("%-44s ABORTED (status %u)\n", name, st[0]); return; }
            if (it > 0 && ms < best) best = ms;
            CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
        }
        CK(cudaMemcpy(h.data(), out, grid * 32, cudaMemcpyDeviceToHost));
        for (int i = 0; i < grid; ++i) spans.push_back((double)(h[i * 4 + 1] - h[i * 4]) * 1e-6);
        std::sort(spans.begin(), spans.end());
        const double gb = (double)tot_tiles * c.slot_bytes * grid / 1e9;
        printf("%-44s %8.3f ms  %7.1f GB/s  (%.2f GB; CTA span min/med/max %.3f/%.3f/%.3f ms)\n", name, best, gb / best * 1e3, gb, spans.front(), spans[grid / 2], spans.back());
        fflush(stdout);
    };
    //   name                                     bar pro math dyn unit l2kb slotKB nslots piece phases
    run("stream, no math, 5x32K, piece 4K",          0, 0, 0, 0, 1, 0, 32, 5, 4096, false);
    run("stream, no math, 5x32K, piece 32K",         0, 0, 0, 0, 1, 0, 32, 5, 32768, false);
    run("stream, no math, 5x32K, piece 1K",          0, 0, 0, 0, 1, 0, 32, 5, 1024, false);
    run("stream, no math, 6x32K, piece 4K",          0, 0, 0, 0, 1, 0, 32, 6, 4096, false);
    run("stream, no math, 3x32K, piece 4K",          0, 0, 0, 0, 1, 0, 32, 3, 4096, false);
    run("stream, no math, 2x32K, piece 4K",          0, 0, 0, 0, 1, 0, 32, 2, 4096, false);
    run("stream, no math, 10x16K, piece 4K",         0, 0, 0, 0, 1, 0, 16, 10, 4096, false);
    run("stream, math generic LD, 5x32K",            0, 0, 1, 0, 1, 0, 32, 5, 4096, false);
    run("stream, math ld.shared, 5x32K",             0, 0, 2, 0, 1, 0, 32, 5, 4096, false);
    run("stream, math ld.shared, 6x32K",             0, 0, 2, 0, 1, 0, 32, 6, 4096, false);
    run("phases only (sentinels), math lds",         0, 0, 2, 0, 1, 0, 32, 5, 4096, true);
    run("phases + barrier v1, math lds",             1, 0, 2, 0, 1, 0, 32, 5, 4096, true);
    run("phases + barrier v2 (split), math lds",     2, 0, 2, 0, 1, 0, 32, 5, 4096, true);
    run("phases + barrier v1 + prologue, lds",       1, 1, 2, 0, 1, 0, 32, 5, 4096, true);
    run("phases + barrier v1 + prologue, generic",   1, 1, 1, 0, 1, 0, 32, 5, 4096, true);
    run("phases + barrier v2 + prologue, lds",       2, 1, 2, 0, 1, 0, 32, 5, 4096, true);
    run("phases + barrier v2 + prologue(noreduce)",  2, 2, 2, 0, 1, 0, 32, 5, 4096, true);
    run("  same, 6x32K ring",                        2, 2, 2, 0, 1, 0, 32, 6, 4096, true);
    run("  same, L2 ahead 128K",                     2, 2, 2, 0, 1, 128, 32, 5, 4096, true);
    run("  same, L2 ahead 256K",                     2, 2, 2, 0, 1, 256, 32, 5, 4096, true);
    run("  same, L2 ahead 512K",                     2, 2, 2, 0, 1, 512, 32, 5, 4096, true);
    run("v1 + prologue, L2 ahead 256K",              1, 1, 2, 0, 1, 256, 32, 5, 4096, true);
    run("dynamic unit 1, v2 + prologue(noreduce)",   2, 2, 2, 1, 1, 0, 32, 5, 4096, true);
    run("dynamic unit 2, v2 + prologue(noreduce)",   2, 2, 2, 1, 2, 0, 32, 5, 4096, true);
    run("dynamic unit 2, v1 + prologue",             1, 1, 2, 1, 2, 0, 32, 5, 4096, true);
    run("dynamic unit 2, no barrier",                0, 0, 2, 1, 2, 0, 32, 5, 4096, true);
    run("dynamic unit 4, v2 + prologue(noreduce)",   2, 2, 2, 1, 4, 0, 32, 5, 4096, true);
    run("v1 + prologue, no math",                    1, 1, 0, 0, 1, 0, 32, 5, 4096, true);
    return 0;
}
