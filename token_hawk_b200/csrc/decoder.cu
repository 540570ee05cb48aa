// decoder.cu -- th_eval_gpu for n_tokens == 1 as ONE persistent sm_100a kernel.
//
// Replaces the ~900 WebGPU commands the reference encodes per token (build_layer_cmdbuf x n_layer
// + build_final_compute_cmdbuf, th-llama.cpp:270-452, 240-268, 592-640) with a single launch of
// one CTA per SM.  Design (DESIGN.md has the full write-up):
//
//   * warp 0 of every CTA is a PRODUCER: it walks the CTA's static list of weight / KV tiles for
//     the whole token (all layers, all phases) and streams them HBM -> shared memory with
//     cp.async.bulk (UBLKCP) into a ring of 32 KB slots guarded by full/empty mbarriers.  It never
//     waits for activations, so HBM stays busy across phase boundaries and grid barriers.
//   * warps 1..8 are CONSUMERS: they wait for tiles, do the f16->f32 FMA work out of shared memory
//     (f32 activations, f32 accumulate: the reference's arithmetic), run the fused epilogues
//     (RMSNorm prologue, RoPE + KV append, SiLU*mul, residual adds, argmax) and take part in the
//     grid barriers that separate the data-dependent phases.
//   * phases per layer: QKV | attention (split-KV, online softmax) | Wo | W1,W3 | W2 ; then logits.
//
// Arithmetic follows oracle/th_oracle.c (the restatement of the WGSL); only summation order
// differs.  No tensor cores: at M=1 the work is 1 FLOP/byte and HBM-bound.
#include <cooperative_groups.h>

#include "common.cuh"

namespace {

constexpr int kSlotBytes = 32 * 1024;
constexpr int kNumSlots = 5;
constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = 32 + kConsumerThreads;
constexpr int kMaxTilePos = 128;     // attention: positions per tile cap (score staging size)
constexpr int kMaxHeadDim = 128;
constexpr int kMaxGroupRows = 64;    // RT <= 64

struct MatCfg { int WC, RPW, CPW; };   // col-warps, rows per warp, 256-col chunks per warp per tile

struct DecParams {
    int n_vocab, n_embd, n_head, n_layer, n_ff, n_ctx, head_dim;
    int Eh, Fh, Hl, Vl;                 // local (tensor-parallel shard) sizes
    int tp_rank, tp_size;
    const thk_llama_layer* layers;      // device array [n_layer]
    const uint16_t* emb;
    const float* norm;
    const uint16_t* out_w;
    float *x, *h1, *q, *o, *ff, *part;
    unsigned* head_ctr;
    float* amax_val;
    int* amax_idx;
    unsigned *bar_ctr, *bar_next;
    unsigned* status;                   // [0] abort code, [1..3] diagnostics
    const int* token;
    int n_past;
    float* logits;
    int* next_token;
    float* next_logit;
    MatCfg cfg_qkv, cfg_wo, cfg_w13, cfg_w2, cfg_out;
    int att_tpos;                       // positions per attention tile
    int att_max_split;
    unsigned long long timeout_ns;
};

struct Seg { const uint16_t* W; int rows; };

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

__device__ __forceinline__ bool aborted(const DecParams& p) { return ld_volatile_u32(p.status) != 0; }
__device__ __noinline__ void raise_abort(const DecParams& p, unsigned code, unsigned a, unsigned b) {
    if (atomicCAS(p.status, 0u, code) == 0u) { p.status[1] = a; p.status[2] = b; p.status[3] = blockIdx.x; }
}

// bounded mbarrier wait: returns false when the watchdog fired or another CTA aborted
__device__ __forceinline__ bool mbar_wait(const DecParams& p, uint32_t bar, uint32_t parity, unsigned tag) {
    if (mbar_try_wait(bar, parity)) return true;
    const unsigned long long t0 = gtimer();
    unsigned it = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++it & 255u) == 0u) {
            if (aborted(p)) return false;
            if (gtimer() - t0 > p.timeout_ns) { raise_abort(p, 0x100u | tag, bar, parity); return false; }
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// shared memory carve-up
// ------------------------------------------------------------------------------------------
struct SmemMisc {
    unsigned long long full[kNumSlots];
    unsigned long long empty[kNumSlots];
    alignas(16) float red[2][8][kMaxGroupRows];     // [buffer][col-warp][row in group]
    alignas(16) float sc[2][kMaxTilePos];           // attention scores, double buffered
    alignas(16) float redo[kConsumerWarps][kMaxHeadDim];
    float redl[kConsumerWarps];
    float norm_part[kConsumerWarps];
    float bval[kMaxGroupRows];
    int bidx[kMaxGroupRows];
    int flag;
};

// conflict-free activation layout: 256-col chunk c, lane l owns cols 8l..8l+7; its first float4 sits
// at (c*64 + l)*16 B and its second at (c*64 + 32 + l)*16 B.
__device__ __forceinline__ int xs_index(int col) {
    const int within = col & 255;
    return (col & ~255) + ((within & 4) << 5) + ((within >> 3) << 2) + (within & 3);
}

// ------------------------------------------------------------------------------------------
// tile schedule of one matvec phase for this CTA (shared by producer and consumers)
// ------------------------------------------------------------------------------------------
struct MatSched {
    int RT, CT, KT, C;
    int g0, g1;
    int gs[3];        // row groups per segment
    bool paired;
};
__device__ __forceinline__ MatSched make_sched(const Seg* seg, int nseg, int C, const MatCfg& cfg, bool paired) {
    MatSched s;
    s.C = C;
    s.RT = (kConsumerWarps / cfg.WC) * cfg.RPW;
    s.CT = cfg.WC * 256 * cfg.CPW;
    s.KT = (C + s.CT - 1) / s.CT;
    s.paired = paired;
    int G = 0;
    for (int i = 0; i < 3; ++i) {
        s.gs[i] = (i < nseg) ? (seg[i].rows + s.RT - 1) / s.RT : 0;
        if (!paired || i == 0) G += s.gs[i];
    }
    s.g0 = (int)(((long long)G * blockIdx.x) / gridDim.x);
    s.g1 = (int)(((long long)G * (blockIdx.x + 1)) / gridDim.x);
    return s;
}
__device__ __forceinline__ void locate_group(const MatSched& s, int g, int& segi, int& lg) {
    segi = 0; lg = g;
    if (!s.paired) {
        while (segi < 2 && lg >= s.gs[segi]) { lg -= s.gs[segi]; ++segi; }
    }
}

// ------------------------------------------------------------------------------------------
// PRODUCER
// ------------------------------------------------------------------------------------------
struct Ring {
    uint32_t slot_base, full_base, empty_base;
    uint32_t tc;   // tiles issued / consumed so far
    __device__ __forceinline__ uint32_t slot() const { return tc % kNumSlots; }
    __device__ __forceinline__ uint32_t round() const { return tc / kNumSlots; }
    __device__ __forceinline__ uint32_t slot_addr() const { return slot_base + slot() * kSlotBytes; }
    __device__ __forceinline__ uint32_t full_bar() const { return full_base + slot() * 8; }
    __device__ __forceinline__ uint32_t empty_bar() const { return empty_base + slot() * 8; }
};

__device__ __forceinline__ bool produce_mat_phase(const DecParams& p, Ring& ring, const Seg* seg, int nseg, int C,
                                                  const MatCfg& cfg, bool paired) {
    const int lane = threadIdx.x & 31;
    const MatSched s = make_sched(seg, nseg, C, cfg, paired);
    for (int g = s.g0; g < s.g1; ++g) {
        int segi, lg;
        locate_group(s, g, segi, lg);
        const int nsub = paired ? 2 : 1;
        for (int sub = 0; sub < nsub; ++sub) {
            const Seg& sg = seg[paired ? sub : segi];
            const int row0 = lg * s.RT;
            const int nrows = min(s.RT, sg.rows - row0);
            for (int kt = 0; kt < s.KT; ++kt) {
                const int col0 = kt * s.CT;
                const int ncols = min(s.CT, C - col0);
                if (!mbar_wait(p, ring.empty_bar(), (ring.round() & 1u) ^ 1u, 1)) return false;
                const uint32_t dst = ring.slot_addr(), fb = ring.full_bar();
                if (lane == 0) mbar_expect_tx(fb, (uint32_t)nrows * ncols * 2u);
                __syncwarp();
                const uint16_t* src = sg.W + (size_t)row0 * C + col0;
                if (ncols == C) {
                    if (lane == 0) bulk_g2s(dst, src, (uint32_t)nrows * ncols * 2u, fb);
                } else {
                    for (int r = lane; r < nrows; r += 32)
                        bulk_g2s(dst + (uint32_t)r * ncols * 2u, src + (size_t)r * C, (uint32_t)ncols * 2u, fb);
                }
                ++ring.tc;
            }
        }
    }
    return true;
}

// attention work split (shared by producer and consumers)
struct AttSched { int S, N; };
__device__ __forceinline__ AttSched make_att(const DecParams& p) {
    AttSched a;
    a.N = p.n_past + 1;
    int S = (int)gridDim.x / p.Hl;
    if (S < 1) S = 1;
    const int by_len = (a.N + 15) / 16;          // at least ~16 positions per split
    if (S > by_len) S = by_len;
    if (S > p.att_max_split) S = p.att_max_split;
    if (S > a.N) S = a.N;
    if (S < 1) S = 1;
    a.S = S;
    return a;
}

__device__ __forceinline__ bool produce_att_phase(const DecParams& p, Ring& ring, const thk_llama_layer& L) {
    const int lane = threadIdx.x & 31;
    const AttSched a = make_att(p);
    const int D = p.head_dim;
    for (int u = blockIdx.x; u < p.Hl * a.S; u += gridDim.x) {
        const int h = u / a.S, sp = u % a.S;
        const int pa = (int)(((long long)a.N * sp) / a.S);
        int pb = (int)(((long long)a.N * (sp + 1)) / a.S);
        if (pb > p.n_past) pb = p.n_past;            // position n_past is read from global by the consumers
        for (int pos = pa; pos < pb; pos += p.att_tpos) {
            const int np = min(p.att_tpos, pb - pos);
            const uint32_t bytes = (uint32_t)np * D * 4u;
            for (int kv = 0; kv < 2; ++kv) {
                if (!mbar_wait(p, ring.empty_bar(), (ring.round() & 1u) ^ 1u, 2)) return false;
                const float* base = (kv == 0 ? L.key_cache : L.value_cache) + ((size_t)h * p.n_ctx + pos) * D;
                if (lane == 0) { mbar_expect_tx(ring.full_bar(), bytes); bulk_g2s(ring.slot_addr(), base, bytes, ring.full_bar()); }
                __syncwarp();
                ++ring.tc;
            }
        }
    }
    return true;
}

__device__ void producer_main(const DecParams& p, Ring ring) {
    for (int l = 0; l < p.n_layer; ++l) {
        const thk_llama_layer L = p.layers[l];
        { Seg s[3] = {{L.wq, p.Eh}, {L.wk, p.Eh}, {L.wv, p.Eh}};
          if (!produce_mat_phase(p, ring, s, 3, p.n_embd, p.cfg_qkv, false)) return; }
        if (!produce_att_phase(p, ring, L)) return;
        { Seg s[3] = {{L.wo, p.n_embd}, {nullptr, 0}, {nullptr, 0}};
          if (!produce_mat_phase(p, ring, s, 1, p.Eh, p.cfg_wo, false)) return; }
        { Seg s[3] = {{L.w1, p.Fh}, {L.w3, p.Fh}, {nullptr, 0}};
          if (!produce_mat_phase(p, ring, s, 2, p.n_embd, p.cfg_w13, true)) return; }
        { Seg s[3] = {{L.w2, p.n_embd}, {nullptr, 0}, {nullptr, 0}};
          if (!produce_mat_phase(p, ring, s, 1, p.Fh, p.cfg_w2, false)) return; }
    }
    Seg s[3] = {{p.out_w, p.Vl}, {nullptr, 0}, {nullptr, 0}};
    produce_mat_phase(p, ring, s, 1, p.n_embd, p.cfg_out, false);
}

// ------------------------------------------------------------------------------------------
// CONSUMERS
// ------------------------------------------------------------------------------------------
struct Cons {
    const DecParams& p;
    Ring ring;
    unsigned char* slots;
    float* xs;
    SmemMisc* sm;
    int ct, cw, lane;        // consumer thread id 0..255, consumer warp 0..7, lane
    unsigned nbar;           // grid barriers passed
    bool ok;
    int redbuf;
};

__device__ __forceinline__ void grid_barrier(Cons& c) {
    // all consumer threads have issued their global writes for this phase
    __threadfence();
    consumer_bar();
    if (c.ct == 0) {
        ++c.nbar;
        red_release_add(c.p.bar_ctr, 1u);
        const unsigned target = c.nbar * gridDim.x;
        if (ld_acquire_u32(c.p.bar_ctr) < target) {
            const unsigned long long t0 = gtimer();
            unsigned it = 0;
            while (ld_acquire_u32(c.p.bar_ctr) < target) {
                if ((++it & 63u) == 0u) {
                    if (aborted(c.p)) break;
                    if (gtimer() - t0 > c.p.timeout_ns) { raise_abort(c.p, 0x200u, c.nbar, target); break; }
                }
            }
        }
    } else {
        ++c.nbar;
    }
    consumer_bar();
}

// xs <- rmsnorm(src) * gain   (cmdbuf_rms_norm + cmdbuf_row_element_multiply, th.cpp:1153-1200,1298-1315)
__device__ __forceinline__ void prologue_norm(Cons& c, const float* src, const float* gain, int n) {
    float ss = 0.f;
    for (int i = c.ct * 4; i < n; i += kConsumerThreads * 4) {
        const float4 v = __ldcg((const float4*)(src + i));
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    ss = warp_sum(ss);
    if (c.lane == 0) c.sm->norm_part[c.cw] = ss;
    consumer_bar();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) tot += c.sm->norm_part[w];
    const float inv = 1.0f / sqrtf(tot / (float)n + 1e-6f);
    for (int i = c.ct * 4; i < n; i += kConsumerThreads * 4) {
        const float4 v = __ldcg((const float4*)(src + i));
        const float4 g = __ldg((const float4*)(gain + i));
        float4 o;
        o.x = (v.x * inv) * g.x; o.y = (v.y * inv) * g.y; o.z = (v.z * inv) * g.z; o.w = (v.w * inv) * g.w;
        *(float4*)(c.xs + xs_index(i)) = o;
    }
    consumer_bar();
}
__device__ __forceinline__ void prologue_copy(Cons& c, const float* src, int n) {
    for (int i = c.ct * 4; i < n; i += kConsumerThreads * 4)
        *(float4*)(c.xs + xs_index(i)) = __ldcg((const float4*)(src + i));
    consumer_bar();
}

// one tile: acc[r] += sum over this warp's columns of W[row][col] * x[col]
template <int RPW, int CPW>
__device__ __forceinline__ void tile_fma(const unsigned char* slot, const float* xs, int WC, int wr, int wc, int lane,
                                         int col0, int nrows, int ncols, float (&acc)[8]) {
#pragma unroll
    for (int cp = 0; cp < CPW; ++cp) {
        const int col = ((cp * WC + wc) << 8) + (lane << 3);
        if (col < ncols) {
            const float* xp = xs + (col0 + ((cp * WC + wc) << 8)) + (lane << 2);
            const float4 x0 = *(const float4*)xp;
            const float4 x1 = *(const float4*)(xp + 128);
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int row = wr * RPW + r;
                if (row < nrows) {
                    const uint4 w = *(const uint4*)(slot + ((size_t)row * ncols + col) * 2);
                    float2 f;
                    float a = acc[r];
                    f = h2_to_f2(w.x); a = fmaf(x0.x, f.x, a); a = fmaf(x0.y, f.y, a);
                    f = h2_to_f2(w.y); a = fmaf(x0.z, f.x, a); a = fmaf(x0.w, f.y, a);
                    f = h2_to_f2(w.z); a = fmaf(x1.x, f.x, a); a = fmaf(x1.y, f.y, a);
                    f = h2_to_f2(w.w); a = fmaf(x1.z, f.x, a); a = fmaf(x1.w, f.y, a);
                    acc[r] = a;
                }
            }
        }
    }
}

// transposing warp reduction of RPW accumulators: afterwards lane (row << (5-log2 RPW)) holds row's sum
template <int RPW>
__device__ __forceinline__ float reduce_rows(float (&v)[8], int lane) {
    int off = 16;
#pragma unroll
    for (int n = RPW; n > 1; n >>= 1) {
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
    float r = v[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

enum EpiKind { EPI_QKV, EPI_WO, EPI_W13, EPI_W2, EPI_OUT };

struct EpiState {
    float gate;       // W1 result held between the paired sub-groups
    float best;       // running argmax
    int best_idx;
};

__device__ __forceinline__ void rope_pair(float& a, float& b, int i_even, int head_dim, int n_past) {
    // cmdbuf_RoPE, th.cpp:1457-1492
    const float theta = powf(10000.0f, (-(float)i_even) / (float)head_dim);
    float s, c;
    sincosf((float)n_past * theta, &s, &c);
    const float x0 = a, x1 = b;
    a = x0 * c - x1 * s;
    b = x0 * s + x1 * c;
}

template <int RPW, int CPW>
__device__ __forceinline__ void consume_mat_phase_t(Cons& c, const Seg* seg, int nseg, int C, const MatCfg& cfg, bool paired,
                                                    EpiKind kind, const thk_llama_layer* L, EpiState& es) {
    const DecParams& p = c.p;
    const MatSched s = make_sched(seg, nseg, C, cfg, paired);
    const int WC = cfg.WC;
    const int wr = c.cw / WC, wc = c.cw % WC;
    constexpr int kLog = (RPW == 8) ? 3 : (RPW == 4) ? 2 : 1;
    for (int g = s.g0; g < s.g1; ++g) {
        int segi, lg;
        locate_group(s, g, segi, lg);
        const int nsub = paired ? 2 : 1;
        for (int sub = 0; sub < nsub; ++sub) {
            const int si = paired ? sub : segi;
            const int row0 = lg * s.RT;
            const int nrows = min(s.RT, seg[si].rows - row0);
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = 0.f;
            for (int kt = 0; kt < s.KT; ++kt) {
                const int col0 = kt * s.CT;
                const int ncols = min(s.CT, C - col0);
                if (c.ok) c.ok = mbar_wait(p, c.ring.full_bar(), c.ring.round() & 1u, 3);
                if (c.ok) tile_fma<RPW, CPW>(c.slots + (size_t)c.ring.slot() * kSlotBytes, c.xs, WC, wr, wc, c.lane, col0, nrows, ncols, acc);
                __syncwarp();
                if (c.lane == 0) mbar_arrive(c.ring.empty_bar());
                ++c.ring.tc;
            }
            const float rsum = reduce_rows<RPW>(acc, c.lane);
            const int rb = c.redbuf;
            if ((c.lane & ((32 >> kLog) - 1)) == 0) c.sm->red[rb][wc][wr * RPW + (c.lane >> (5 - kLog))] = rsum;
            consumer_bar();
            c.redbuf ^= 1;
            // ---- epilogue: thread t owns row row0 + t of this group ----
            const int t = c.ct;
            if (t < nrows) {
                float y = 0.f;
                for (int w = 0; w < WC; ++w) y += c.sm->red[rb][w][t];
                const int r = row0 + t;
                switch (kind) {
                case EPI_QKV: {
                    const int D = p.head_dim;
                    if (si == 2) {                                   // V: append (th-llama.cpp:338)
                        L->value_cache[((size_t)(r / D) * p.n_ctx + p.n_past) * D + (r % D)] = y;
                    } else if ((t & 1) == 0) {                       // Q / K: rotate the pair (t, t+1)
                        float y1 = 0.f;
                        for (int w = 0; w < WC; ++w) y1 += c.sm->red[rb][w][t + 1];
                        rope_pair(y, y1, r % D, D, p.n_past);
                        if (si == 0) { p.q[r] = y; p.q[r + 1] = y1; }
                        else {
                            float* kc = L->key_cache + ((size_t)(r / D) * p.n_ctx + p.n_past) * D + (r % D);
                            kc[0] = y; kc[1] = y1;                  // th-llama.cpp:337
                        }
                    }
                } break;
                case EPI_WO: p.h1[r] = __ldcg(p.x + r) + y; break;                    // th-llama.cpp:409
                case EPI_W13:
                    if (sub == 0) es.gate = y;
                    else { const float gv = es.gate; p.ff[r] = (gv / (1.0f + expf(-gv))) * y; }   // :436,:438
                    break;
                case EPI_W2: p.x[r] = __ldcg(p.h1 + r) + y; break;                    // th-llama.cpp:447
                case EPI_OUT: {
                    if (p.logits) p.logits[r] = y;
                    const int gid = p.tp_rank * p.Vl + r;
                    if (es.best_idx < 0 || y > es.best) { es.best = y; es.best_idx = gid; }
                } break;
                }
            }
        }
    }
}

__device__ __forceinline__ void consume_mat_phase(Cons& c, const Seg* seg, int nseg, int C, const MatCfg& cfg, bool paired,
                                                  EpiKind kind, const thk_llama_layer* L, EpiState& es) {
    if (cfg.RPW == 8) consume_mat_phase_t<8, 1>(c, seg, nseg, C, cfg, paired, kind, L, es);
    else if (cfg.RPW == 4) consume_mat_phase_t<4, 2>(c, seg, nseg, C, cfg, paired, kind, L, es);
    else consume_mat_phase_t<2, 4>(c, seg, nseg, C, cfg, paired, kind, L, es);
}

// ------------------------------------------------------------------------------------------
// attention: single query, split-KV, online softmax
// (cmdbuf_mat_mul QK^T * 1/sqrt(D), cmdbuf_row_softmax, cmdbuf_mat_mul P*V; th-llama.cpp:365-380)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    float s = a.x * b.x;
    s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    return s;
}

__device__ void consume_att_phase(Cons& c, const thk_llama_layer& L) {
    const DecParams& p = c.p;
    const AttSched a = make_att(p);
    const int D = p.head_dim, nvec = D >> 2;
    const float scale = 1.0f / sqrtf((float)D);
    const bool act = c.lane < nvec;
    for (int u = blockIdx.x; u < p.Hl * a.S; u += gridDim.x) {
        const int h = u / a.S, sp = u % a.S;
        const int pa = (int)(((long long)a.N * sp) / a.S);
        const int pb_full = (int)(((long long)a.N * (sp + 1)) / a.S);
        const int pb = min(pb_full, p.n_past);
        const bool has_new = (pb_full == a.N);
        float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) q4 = __ldcg((const float4*)(p.q + h * D) + c.lane);
        float m = -INFINITY, lsum = 0.f;
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int buf = 0;
        for (int pos = pa; pos < pb; pos += p.att_tpos) {
            const int np = min(p.att_tpos, pb - pos);
            // K tile -> scores
            if (c.ok) c.ok = mbar_wait(p, c.ring.full_bar(), c.ring.round() & 1u, 4);
            if (c.ok) {
                const float* kt = (const float*)(c.slots + (size_t)c.ring.slot() * kSlotBytes);
                for (int j = c.cw; j < np; j += kConsumerWarps) {
                    float sdot = 0.f;
                    if (act) sdot = dot4(q4, *(const float4*)(kt + j * D + c.lane * 4));
                    sdot = warp_sum(sdot) * scale;
                    if (c.lane == 0) c.sm->sc[buf][j] = sdot;
                }
            }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.ring.empty_bar());
            ++c.ring.tc;
            consumer_bar();
            float bmax = -INFINITY;
            for (int j = c.lane; j < np; j += 32) bmax = fmaxf(bmax, c.sm->sc[buf][j]);
            bmax = warp_max(bmax);
            const float m_new = fmaxf(m, bmax);
            const float corr = expf(m - m_new);
            lsum *= corr; o4.x *= corr; o4.y *= corr; o4.z *= corr; o4.w *= corr;
            m = m_new;
            // V tile -> weighted sum
            if (c.ok) c.ok = mbar_wait(p, c.ring.full_bar(), c.ring.round() & 1u, 5);
            if (c.ok) {
                const float* vt = (const float*)(c.slots + (size_t)c.ring.slot() * kSlotBytes);
                for (int j = c.cw; j < np; j += kConsumerWarps) {
                    const float pj = expf(c.sm->sc[buf][j] - m);
                    lsum += pj;
                    if (act) {
                        const float4 v4 = *(const float4*)(vt + j * D + c.lane * 4);
                        o4.x = fmaf(pj, v4.x, o4.x); o4.y = fmaf(pj, v4.y, o4.y);
                        o4.z = fmaf(pj, v4.z, o4.z); o4.w = fmaf(pj, v4.w, o4.w);
                    }
                }
            }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.ring.empty_bar());
            ++c.ring.tc;
            buf ^= 1;
        }
        if (has_new) {   // the token's own K/V row, written by the QKV epilogue of this launch
            const size_t off = ((size_t)h * p.n_ctx + p.n_past) * D;
            float sdot = 0.f;
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (act) {
                sdot = dot4(q4, __ldcg((const float4*)(L.key_cache + off) + c.lane));
                v4 = __ldcg((const float4*)(L.value_cache + off) + c.lane);
            }
            sdot = warp_sum(sdot) * scale;
            const float m_new = fmaxf(m, sdot);
            const float corr = expf(m - m_new);
            lsum *= corr; o4.x *= corr; o4.y *= corr; o4.z *= corr; o4.w *= corr;
            m = m_new;
            if (c.cw == 0) {
                const float pj = expf(sdot - m);
                lsum += pj;
                o4.x = fmaf(pj, v4.x, o4.x); o4.y = fmaf(pj, v4.y, o4.y);
                o4.z = fmaf(pj, v4.z, o4.z); o4.w = fmaf(pj, v4.w, o4.w);
            }
        }
        // combine the 8 warps' partial sums (same running max in every warp)
        if (act) *(float4*)(&c.sm->redo[c.cw][c.lane * 4]) = o4;
        if (c.lane == 0) c.sm->redl[c.cw] = lsum;
        consumer_bar();
        float* part = p.part + (size_t)(h * p.att_max_split + sp) * (D + 2);
        if (c.ct < D) {
            float od = 0.f;
#pragma unroll
            for (int w = 0; w < kConsumerWarps; ++w) od += c.sm->redo[w][c.ct];
            part[2 + c.ct] = od;
        }
        if (c.ct == 0) {
            float lt = 0.f;
#pragma unroll
            for (int w = 0; w < kConsumerWarps; ++w) lt += c.sm->redl[w];
            part[0] = m; part[1] = lt;
        }
        // last split to finish combines the head (deterministic: fixed split order)
        __threadfence();
        consumer_bar();
        if (c.ct == 0) {
            const unsigned prev = atomicAdd(p.head_ctr + h, 1u);
            c.sm->flag = (prev == (unsigned)(a.S - 1));
            if (c.sm->flag) p.head_ctr[h] = 0u;
        }
        consumer_bar();
        if (c.sm->flag) {
            __threadfence();
            if (c.ct < D) {
                const float* ph = p.part + (size_t)h * p.att_max_split * (D + 2);
                float M = -INFINITY;
                for (int s2 = 0; s2 < a.S; ++s2) M = fmaxf(M, __ldcg(ph + (size_t)s2 * (D + 2)));
                float Lt = 0.f, od = 0.f;
                for (int s2 = 0; s2 < a.S; ++s2) {
                    const float* ps = ph + (size_t)s2 * (D + 2);
                    const float w = expf(__ldcg(ps) - M);
                    Lt = fmaf(__ldcg(ps + 1), w, Lt);
                    od = fmaf(__ldcg(ps + 2 + c.ct), w, od);
                }
                p.o[h * D + c.ct] = od / Lt;
            }
        }
        consumer_bar();   // sm->flag / redo reuse
    }
}

__device__ void consumer_main(const DecParams& p, Ring ring, unsigned char* slots, float* xs, SmemMisc* sm) {
    Cons c{p, ring, slots, xs, sm, (int)threadIdx.x - 32, ((int)threadIdx.x - 32) >> 5, (int)threadIdx.x & 31, 0u, true, 0};
    EpiState es{0.f, 0.f, -1};

    // phase E: x <- f32(tok_embeddings[token]) (th-llama.cpp:577-585; device-side like :552-575)
    {
        int tok = *p.token;
        if (tok < 0 || tok >= p.n_vocab) { if (c.ct == 0) raise_abort(p, 0x300u, (unsigned)tok, 0); tok = 0; }
        const uint16_t* row = p.emb + (size_t)tok * p.n_embd;
        for (int i = (blockIdx.x * kConsumerThreads + c.ct); i < p.n_embd; i += gridDim.x * kConsumerThreads)
            p.x[i] = __half2float(__ushort_as_half(row[i]));
        if (blockIdx.x == 0 && c.ct == 0) *p.bar_next = 0u;   // arm the next launch's barrier counter
    }
    grid_barrier(c);

    for (int l = 0; l < p.n_layer; ++l) {
        const thk_llama_layer* Lp = p.layers + l;
        const thk_llama_layer L = *Lp;
        prologue_norm(c, p.x, L.attention_norm, p.n_embd);
        { Seg s[3] = {{L.wq, p.Eh}, {L.wk, p.Eh}, {L.wv, p.Eh}};
          consume_mat_phase(c, s, 3, p.n_embd, p.cfg_qkv, false, EPI_QKV, Lp, es); }
        grid_barrier(c);
        consume_att_phase(c, L);
        grid_barrier(c);
        prologue_copy(c, p.o, p.Eh);
        { Seg s[3] = {{L.wo, p.n_embd}, {nullptr, 0}, {nullptr, 0}};
          consume_mat_phase(c, s, 1, p.Eh, p.cfg_wo, false, EPI_WO, Lp, es); }
        grid_barrier(c);
        prologue_norm(c, p.h1, L.ffn_norm, p.n_embd);
        { Seg s[3] = {{L.w1, p.Fh}, {L.w3, p.Fh}, {nullptr, 0}};
          consume_mat_phase(c, s, 2, p.n_embd, p.cfg_w13, true, EPI_W13, Lp, es); }
        grid_barrier(c);
        prologue_copy(c, p.ff, p.Fh);
        { Seg s[3] = {{L.w2, p.n_embd}, {nullptr, 0}, {nullptr, 0}};
          consume_mat_phase(c, s, 1, p.Fh, p.cfg_w2, false, EPI_W2, Lp, es); }
        grid_barrier(c);
    }
    // final: rmsnorm * norm, logits, greedy argmax (th-llama.cpp:240-268, 826-838)
    prologue_norm(c, p.x, p.norm, p.n_embd);
    { Seg s[3] = {{p.out_w, p.Vl}, {nullptr, 0}, {nullptr, 0}};
      consume_mat_phase(c, s, 1, p.n_embd, p.cfg_out, false, EPI_OUT, nullptr, es); }
    if (p.next_token || p.next_logit) {
        // CTA-level argmax over the epilogue threads (thread t saw rows row0+t in ascending groups)
        if (c.ct < kMaxGroupRows) { sm->bval[c.ct] = es.best; sm->bidx[c.ct] = es.best_idx; }
        consumer_bar();
        if (c.ct == 0) {
            float bv = 0.f; int bi = -1;
            for (int t = 0; t < kMaxGroupRows; ++t) {
                const int idx = sm->bidx[t];
                if (idx < 0) continue;
                const float v = sm->bval[t];
                if (bi < 0 || v > bv || (v == bv && idx < bi)) { bv = v; bi = idx; }
            }
            p.amax_val[blockIdx.x] = bv; p.amax_idx[blockIdx.x] = bi;
        }
        grid_barrier(c);
        if (blockIdx.x == 0 && c.ct == 0) {
            float bv = 0.f; int bi = -1;
            for (unsigned b = 0; b < gridDim.x; ++b) {
                const int idx = __ldcg(p.amax_idx + b);
                if (idx < 0) continue;
                const float v = __ldcg(p.amax_val + b);
                if (bi < 0 || v > bv || (v == bv && idx < bi)) { bv = v; bi = idx; }
            }
            if (p.next_token) *p.next_token = bi < 0 ? 0 : bi;
            if (p.next_logit) *p.next_logit = bv;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) decode_kernel(const __grid_constant__ DecParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* slots = smem;
    SmemMisc* sm = (SmemMisc*)(smem + kNumSlots * kSlotBytes);
    float* xs = (float*)(smem + kNumSlots * kSlotBytes + ((sizeof(SmemMisc) + 127) & ~127));
    if (threadIdx.x == 0) {
        for (int i = 0; i < kNumSlots; ++i) {
            mbar_init(smem_u32(&sm->full[i]), 1);
            mbar_init(smem_u32(&sm->empty[i]), kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Ring ring{smem_u32(slots), smem_u32(&sm->full[0]), smem_u32(&sm->empty[0]), 0u};
    if (threadIdx.x < 32) producer_main(p, ring);
    else consumer_main(p, ring, slots, xs, sm);
}

size_t decode_smem_bytes(int max_vec) {
    return (size_t)kNumSlots * kSlotBytes + ((sizeof(SmemMisc) + 127) & ~127) + (size_t)((max_vec + 255) & ~255) * sizeof(float);
}

MatCfg choose_cfg(int rows_total, int C, bool paired_rows_counted_once, int n_cta) {
    (void)paired_rows_counted_once;
    const int chunks = (C + 255) / 256;
    int WC = 1;
    while (WC * 2 <= chunks && WC < 8) WC *= 2;
    const int WR = kConsumerWarps / WC;
    int cap = 1;
    while (cap * WC < chunks && cap < 4) cap *= 2;     // chunks per warp needed to span C, pow2, <= 4
    const int opts[3][2] = {{8, 1}, {4, 2}, {2, 4}};
    MatCfg best{WC, 8, 1};
    for (int i = 0; i < 3; ++i) {
        if (opts[i][1] > cap) break;
        best = MatCfg{WC, opts[i][0], opts[i][1]};
        const int G = (rows_total + WR * opts[i][0] - 1) / (WR * opts[i][0]);
        if (G >= 16 * n_cta) break;
    }
    return best;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct thk_decoder {
    thk_ctx* ctx = nullptr;
    DecParams p{};
    thk_llama_layer* d_layers = nullptr;
    float* scratch = nullptr;
    unsigned* ctrl = nullptr;       // [64 barrier counters][n_head counters][4 status]
    int* d_tok = nullptr;           // chained-token scratch for generate
    unsigned launch_seq = 0;
    int grid = 0;
    size_t smem = 0;
    int last_launches = 0;
};

static int check_status(thk_decoder* d) {
    unsigned st[4];
    THK_CUDA(cudaMemcpyAsync(st, d->p.status, sizeof st, cudaMemcpyDeviceToHost, d->ctx->stream));
    THK_CUDA(cudaStreamSynchronize(d->ctx->stream));
    if (st[0] != 0) {
        thk_set_error("decode kernel aborted: code 0x%x a=%u b=%u cta=%u (0x1xx mbarrier wait, 0x200 grid barrier, 0x300 bad token)",
                      st[0], st[1], st[2], st[3]);
        cudaMemsetAsync(d->p.status, 0, sizeof st, d->ctx->stream);
        return st[0] == 0x300u ? THK_E_INVALID : THK_E_TIMEOUT;
    }
    return THK_OK;
}

extern "C" int thk_decoder_create(thk_ctx* ctx, const thk_llama_dims* dims, const thk_llama_layer* layers,
                                  const uint16_t* tok_embeddings, const float* norm, const uint16_t* output,
                                  thk_decoder** out) {
    THK_CHECK_ARG(ctx && dims && layers && tok_embeddings && norm && output && out, "thk_decoder_create: null argument");
    *out = nullptr;
    const int tp = dims->tp_size > 0 ? dims->tp_size : 1;
    THK_CHECK_ARG(tp == 1, "thk_decoder_create: tensor parallel sizes > 1 go through thk_decoder_create + set_peers (not built yet)");
    THK_CHECK_ARG(dims->n_embd > 0 && dims->n_head > 0 && dims->n_embd % dims->n_head == 0, "bad n_embd/n_head");
    const int D = dims->n_embd / dims->n_head;
    THK_CHECK_ARG(D % 4 == 0 && D <= kMaxHeadDim && D % 2 == 0, "head_dim %d unsupported (need multiple of 4, <= %d)", D, kMaxHeadDim);
    THK_CHECK_ARG(dims->n_embd % 8 == 0 && dims->n_ff % 8 == 0, "n_embd and n_ff must be multiples of 8 (16-byte f16 rows)");
    THK_CHECK_ARG(dims->n_head % tp == 0 && dims->n_ff % tp == 0 && dims->n_vocab % tp == 0, "tp_size must divide n_head, n_ff, n_vocab");
    THK_CHECK_ARG(dims->n_ctx > 0 && dims->n_layer > 0 && dims->n_vocab > 0, "bad dims");
    THK_CUDA(cudaSetDevice(ctx->device));

    thk_decoder* d = new thk_decoder;
    d->ctx = ctx;
    DecParams& p = d->p;
    p.n_vocab = dims->n_vocab; p.n_embd = dims->n_embd; p.n_head = dims->n_head; p.n_layer = dims->n_layer;
    p.n_ff = dims->n_ff; p.n_ctx = dims->n_ctx; p.head_dim = D;
    p.tp_rank = dims->tp_rank; p.tp_size = tp;
    p.Eh = dims->n_embd / tp; p.Fh = dims->n_ff / tp; p.Hl = dims->n_head / tp; p.Vl = dims->n_vocab / tp;
    p.emb = tok_embeddings; p.norm = norm; p.out_w = output;
    d->grid = ctx->sm_count;
    p.cfg_qkv = choose_cfg(3 * p.Eh, p.n_embd, false, d->grid);
    p.cfg_wo = choose_cfg(p.n_embd, p.Eh, false, d->grid);
    p.cfg_w13 = choose_cfg(p.Fh, p.n_embd, true, d->grid);
    p.cfg_w2 = choose_cfg(p.n_embd, p.Fh, false, d->grid);
    p.cfg_out = choose_cfg(p.Vl, p.n_embd, false, d->grid);
    p.att_tpos = kSlotBytes / (D * 4);
    if (p.att_tpos > kMaxTilePos) p.att_tpos = kMaxTilePos;
    p.att_max_split = d->grid / p.Hl > 0 ? d->grid / p.Hl : 1;
    p.timeout_ns = 4000000000ull;
    const int max_vec = p.n_embd > p.Fh ? p.n_embd : p.Fh;
    d->smem = decode_smem_bytes(max_vec);
    if (d->smem > 227 * 1024) {
        thk_set_error("thk_decoder_create: needs %zu bytes of shared memory (> 227 KB); n_ff/tp=%d too large", d->smem, p.Fh);
        delete d;
        return THK_E_UNSUPPORTED;
    }
    THK_CUDA(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));

    THK_CUDA(cudaMalloc(&d->d_layers, sizeof(thk_llama_layer) * p.n_layer));
    THK_CUDA(cudaMemcpy(d->d_layers, layers, sizeof(thk_llama_layer) * p.n_layer, cudaMemcpyHostToDevice));
    p.layers = d->d_layers;
    // scratch: x, h1 [E]; q, o [Eh]; ff [Fh]; part [Hl*S*(D+2)]; amax [grid] x2
    const size_t part_n = (size_t)p.Hl * p.att_max_split * (D + 2);
    const size_t nfl = (size_t)2 * p.n_embd + 2 * p.Eh + p.Fh + part_n + 2 * d->grid + 64;
    THK_CUDA(cudaMalloc(&d->scratch, nfl * sizeof(float)));
    THK_CUDA(cudaMemset(d->scratch, 0, nfl * sizeof(float)));
    float* f = d->scratch;
    p.x = f; f += p.n_embd; p.h1 = f; f += p.n_embd; p.q = f; f += p.Eh; p.o = f; f += p.Eh;
    p.ff = f; f += (p.Fh + 3) & ~3; p.part = f; f += part_n; p.amax_val = f; f += d->grid; p.amax_idx = (int*)f;
    const size_t nctrl = 64 + p.Hl + 4;
    THK_CUDA(cudaMalloc(&d->ctrl, nctrl * sizeof(unsigned)));
    THK_CUDA(cudaMemset(d->ctrl, 0, nctrl * sizeof(unsigned)));
    p.head_ctr = d->ctrl + 64;
    p.status = d->ctrl + 64 + p.Hl;
    THK_CUDA(cudaMalloc(&d->d_tok, sizeof(int) * 2));
    *out = d;
    return THK_OK;
}

extern "C" int thk_decoder_destroy(thk_decoder* d) {
    if (!d) return THK_OK;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    cudaFree(d->d_layers); cudaFree(d->scratch); cudaFree(d->ctrl); cudaFree(d->d_tok);
    delete d;
    return THK_OK;
}

static int launch_step(thk_decoder* d, const int32_t* token, int32_t n_past, float* logits, int32_t* next_token, float* next_logit) {
    DecParams p = d->p;
    p.token = token; p.n_past = n_past; p.logits = logits; p.next_token = next_token; p.next_logit = next_logit;
    p.bar_ctr = d->ctrl + (d->launch_seq % 64);
    p.bar_next = d->ctrl + ((d->launch_seq + 1) % 64);
    ++d->launch_seq;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)d->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = d->smem;
    cfg.stream = d->ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // co-residency of all CTAs is required by the grid barriers
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    THK_CUDA(cudaLaunchKernelEx(&cfg, decode_kernel, p));
    return THK_OK;
}

extern "C" int thk_decoder_step(thk_decoder* d, const int32_t* token, int32_t n_past, float* logits,
                                int32_t* next_token, float* next_logit) {
    THK_CHECK_ARG(d && token, "thk_decoder_step: null argument");
    THK_CHECK_ARG(n_past >= 0 && n_past < d->p.n_ctx, "thk_decoder_step: n_past %d outside context [0,%d)", n_past, d->p.n_ctx);
    int rc = launch_step(d, token, n_past, logits, next_token, next_logit);
    d->last_launches = 1;
    return rc;
}

extern "C" int thk_decoder_generate(thk_decoder* d, const int32_t* first_token, int32_t n_past, int32_t n_steps,
                                    int32_t* tokens_out, float* last_logits) {
    THK_CHECK_ARG(d && first_token && tokens_out && n_steps > 0, "thk_decoder_generate: bad argument");
    THK_CHECK_ARG(d->p.tp_size == 1, "thk_decoder_generate: single-GPU only");
    THK_CHECK_ARG(n_past >= 0 && n_past + n_steps <= d->p.n_ctx, "thk_decoder_generate: steps exceed context");
    for (int i = 0; i < n_steps; ++i) {
        const int32_t* tok = (i == 0) ? first_token : tokens_out + (i - 1);
        int rc = launch_step(d, tok, n_past + i, (i == n_steps - 1) ? last_logits : nullptr, tokens_out + i, nullptr);
        if (rc) return rc;
    }
    d->last_launches = n_steps;
    return THK_OK;
}

extern "C" int thk_decoder_hidden(thk_decoder* d, const float** hidden) {
    THK_CHECK_ARG(d && hidden, "thk_decoder_hidden: null argument");
    *hidden = d->p.x;
    return THK_OK;
}
extern "C" int thk_decoder_last_launches(thk_decoder* d) { return d ? d->last_launches : 0; }

// surfaces in-kernel aborts (watchdog / bad token); blocks on the stream
extern "C" int thk_decoder_check(thk_decoder* d) {
    THK_CHECK_ARG(d, "thk_decoder_check: null argument");
    return check_status(d);
}

extern "C" int thk_decoder_exchange_info(thk_decoder*, void**, size_t*, void**, size_t*) {
    thk_set_error("tensor-parallel exchange not built yet");
    return THK_E_UNSUPPORTED;
}
extern "C" int thk_decoder_set_peers(thk_decoder*, void* const*, void* const*, int) {
    thk_set_error("tensor-parallel exchange not built yet");
    return THK_E_UNSUPPORTED;
}
