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
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kSlotBytes = 32 * 1024;
constexpr int kNumSlots = 5;
#ifndef THK_MATH_WARPS
#define THK_MATH_WARPS 8
#endif
constexpr int kMathWarps = THK_MATH_WARPS;            // 8 or 16
constexpr int kRC = 64 / kMathWarps;                  // (rows per warp) x (256-col chunks per warp) per 32 KB tile
constexpr int kMathThreads = kMathWarps * 32;
constexpr int kMathBase = 32;                         // warp 0 = producer, warps 1..8 = math, warp 9 = epilogue
constexpr int kThreads = kMathBase + kMathThreads + 32;
constexpr int kConsumerWarps = kMathWarps;
enum NamedBarrier { BAR_ALL = 1, BAR_MATH = 2, BAR_A0 = 3, BAR_A1 = 4, BAR_B0 = 5, BAR_B1 = 6 };
constexpr int kMaxTilePos = 128;     // attention: positions per tile cap (score staging size)
constexpr int kMaxHeadDim = 128;
constexpr int kMaxGroupRows = 64;    // RT <= 64

struct MatCfg { int WC, RPW, CPW; };   // col-warps, rows per warp, 256-col chunks per warp per tile

// Host-computed tile schedule of one matvec phase (no integer divisions on the device's critical path)
struct PhaseDesc {
    int C;            // columns (input dimension)
    int WC, RPW, CPW; // col-warps, rows per warp, 256-col chunks per warp per tile
    int RT, CT, KT;   // rows per group, columns per K tile, K tiles per group
    int nseg, paired; // segments (weight matrices) and whether seg0/seg1 rows are processed pairwise
    int rows[3];      // rows per segment
    int gs[3];        // row groups per segment
    int G;            // row groups in the phase (paired: groups of seg0)
};
enum PhaseKind { PH_QKV = 0, PH_WO = 1, PH_W13 = 2, PH_W2 = 3, PH_OUT = 4 };

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
    PhaseDesc ph[5];
    int att_tpos;                       // positions per attention tile
    int att_max_split;
    unsigned long long timeout_ns;
    int l2_ahead;                       // producer prefetches ahead into L2 at phase start (THK_L2_AHEAD=0 disables)
    // tensor parallel exchange (tp_size > 1): every rank owns one region laid out as
    //   float xb[2][tp][n_embd] | float amax_val[tp] | int amax_idx[tp] | unsigned flags[tp] | unsigned flags2[tp]
    // and writes its partial vectors / argmax candidates straight into every peer's region over NVLink.
    unsigned char* xchg[8];             // region base per rank (peer-mapped pointers; [tp_rank] is local)
    unsigned epoch_base;                // flags are monotonic: exchange k of this launch uses epoch_base + k + 1
    unsigned long long* prof;           // optional timeline: [cta][phase<256][8] u64 (see prof_mark), then [cta][4] producer stats
};


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
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// exchange region accessors
__device__ __forceinline__ float* xb_ptr(const DecParams& p, int rank, int which, int src) {
    return (float*)p.xchg[rank] + ((size_t)which * p.tp_size + src) * p.n_embd;
}
__device__ __forceinline__ size_t xchg_tail(const DecParams& p) { return (size_t)2 * p.tp_size * p.n_embd * sizeof(float); }
__device__ __forceinline__ float* xamax_val(const DecParams& p, int rank) { return (float*)(p.xchg[rank] + xchg_tail(p)); }
__device__ __forceinline__ int* xamax_idx(const DecParams& p, int rank) { return (int*)(p.xchg[rank] + xchg_tail(p)) + p.tp_size; }
__device__ __forceinline__ unsigned* xflags(const DecParams& p, int rank, int set) {
    return (unsigned*)(p.xchg[rank] + xchg_tail(p)) + (2 + set) * p.tp_size;
}
__device__ __forceinline__ void st_release_sys(unsigned* ptr, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* ptr) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}

constexpr int kProfPhases = 256;
enum ProfSlot { PROF_START = 0, PROF_PROLOGUE = 1, PROF_FIRST_TILE = 2, PROF_LAST_TILE = 3, PROF_ARRIVE = 4, PROF_FENCED = 5, PROF_WAIT_FULL = 6 };

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
    float2 rope[kMaxHeadDim / 2];       // (cos, sin) of n_past * theta_i for this token
    int flag;
};

// conflict-free activation layout: 256-col chunk c, lane l owns cols 8l..8l+7; its first float4 sits
// at (c*64 + l)*16 B and its second at (c*64 + 32 + l)*16 B.
__device__ __forceinline__ int xs_index(int col) {
    const int within = col & 255;
    return (col & ~255) + ((within & 4) << 5) + ((within >> 3) << 2) + (within & 3);
}

// ------------------------------------------------------------------------------------------
// this CTA's share of a phase: contiguous row-group range [g0, g1)
// ------------------------------------------------------------------------------------------
// Rows, not row groups, are divided between the CTAs: CTA b owns the (even-aligned) row range
// [R*b/n, R*(b+1)/n) of the concatenated segments and walks it in groups of up to RT rows; its last
// group may be short.  With R/n = 83 (QKV) or 27.7 (Wo, W2) rows per CTA this balances the phase to
// ~1-4% where whole 8-row groups left up to 33% (4 groups vs 3) -- the arrival skew at the barrier.
struct RowIt {
    int r, r_end;              // current / end row in the concatenated row space (paired: rows of segment 0)
    int si, row0, nrows;       // segment, first row within it, rows in this group
    __device__ __forceinline__ void init(const PhaseDesc& d) {
        const unsigned total = (unsigned)(d.paired ? d.rows[0] : d.rows[0] + d.rows[1] + d.rows[2]);
        const unsigned n = gridDim.x, b = blockIdx.x;
        r = (int)(((total * b) / n) & ~1u);
        r_end = (b + 1 == n) ? (int)total : (int)(((total * (b + 1)) / n) & ~1u);
        place(d);
    }
    __device__ __forceinline__ bool valid() const { return r < r_end; }
    __device__ __forceinline__ void place(const PhaseDesc& d) {
        if (r >= r_end) return;
        int off = 0;
        si = 0;
        if (!d.paired) { while (si < 2 && r >= off + d.rows[si]) { off += d.rows[si]; ++si; } }
        row0 = r - off;
        nrows = min(min(d.RT, d.rows[si] - row0), r_end - r);
    }
    __device__ __forceinline__ void next(const PhaseDesc& d) { r += nrows; place(d); }
};

// ------------------------------------------------------------------------------------------
// PRODUCER
// ------------------------------------------------------------------------------------------
struct Ring {
    uint32_t slot_base, full_base, empty_base;
    uint32_t tc;          // tiles issued / consumed so far
    long long wait_cyc;   // profiling: cycles spent waiting on the ring
    uint32_t sl, par;     // current slot and its phase parity (kept incrementally: no modulo per tile)
    __device__ __forceinline__ Ring(uint32_t sb, uint32_t fb, uint32_t eb, uint32_t tc_, long long w)
        : slot_base(sb), full_base(fb), empty_base(eb), tc(tc_), wait_cyc(w), sl(tc_ % kNumSlots), par((tc_ / kNumSlots) & 1u) {}
    __device__ __forceinline__ void advance() { ++tc; if (++sl == kNumSlots) { sl = 0; par ^= 1u; } }
    __device__ __forceinline__ uint32_t slot() const { return sl; }
    __device__ __forceinline__ uint32_t full_parity() const { return par; }
    __device__ __forceinline__ uint32_t empty_parity() const { return par ^ 1u; }
    __device__ __forceinline__ uint32_t slot_addr() const { return slot_base + sl * kSlotBytes; }
    __device__ __forceinline__ uint32_t full_bar() const { return full_base + sl * 8; }
    __device__ __forceinline__ uint32_t empty_bar() const { return empty_base + sl * 8; }
};

// While the consumers sit in a grid barrier + prologue (~5 us) the ring (160 KB) fills up and the
// producer would go idle -- and with it HBM.  At the start of a phase the producer therefore also
// asks L2 for the next kL2Ahead bytes of this CTA's rows beyond what the ring can hold; after the
// stall those tiles stream from L2 faster than HBM could deliver them.
constexpr uint32_t kL2Ahead = 256u * 1024u;
__device__ __forceinline__ void l2_prefetch_block(const void* base, uint32_t bytes, uint32_t skip, uint32_t amount, int lane) {
    if (bytes <= skip) return;
    const uint32_t end = min(bytes, skip + amount);
    for (uint32_t off = skip + (uint32_t)lane * 16384u; off < end; off += 32u * 16384u)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)base + off), "r"(min(16384u, end - off)) : "memory");
}

__device__ __forceinline__ bool produce_mat_phase(const DecParams& p, Ring& ring, const PhaseDesc& dref, const uint16_t* w0p,
                                               const uint16_t* w1p, const uint16_t* w2p) {
    const int lane = threadIdx.x & 31;
    const PhaseDesc d = dref;
    const bool prof = p.prof != nullptr;
    const int C = d.C;
    RowIt it;
    if (p.l2_ahead == 1) {
        it.init(d);
        if (it.valid()) {
            const uint32_t ring_bytes = kNumSlots * kSlotBytes;
            if (d.paired) {                        // rows [row0, ..) of both matrices, consumed alternately
                const uint32_t bytes = (uint32_t)(it.r_end - it.r) * (uint32_t)C * 2u;
                l2_prefetch_block(w0p + (size_t)it.row0 * C, bytes, ring_bytes / 2, kL2Ahead / 2, lane);
                l2_prefetch_block(w1p + (size_t)it.row0 * C, bytes, ring_bytes / 2, kL2Ahead / 2, lane);
            } else {                               // the CTA's rows are one contiguous block (per segment)
                const uint16_t* w = it.si == 0 ? w0p : it.si == 1 ? w1p : w2p;
                const uint32_t rows_here = (uint32_t)min(it.r_end - it.r, d.rows[it.si] - it.row0);
                l2_prefetch_block(w + (size_t)it.row0 * C, rows_here * (uint32_t)C * 2u, ring_bytes, kL2Ahead, lane);
            }
        }
    }
    // mode 2: whenever the producer finds the ring full (steady-state back-pressure, and above all the
    // grid-barrier + prologue stall at a phase boundary) it asks L2 for one more tile-sized chunk of this
    // CTA's rows beyond the ring, up to kL2Ahead ahead -- HBM keeps working while the consumers cannot.
    const bool pf_on = p.l2_ahead == 2;
    const uint32_t ring_bytes = kNumSlots * kSlotBytes;
    int pf_seg = -1;
    uint32_t pf_off = 0, blk_bytes = 0;
    int blk_row0 = 0;
    for (it.init(d); it.valid(); it.next(d)) {
        const int nsub = d.paired ? 2 : 1;
        if (pf_on && it.si != pf_seg) {          // (re)anchor the look-ahead on this segment's block of rows
            pf_seg = it.si;
            blk_row0 = it.row0;
            blk_bytes = (uint32_t)min(it.r_end - it.r, d.rows[it.si] - it.row0) * (uint32_t)C * 2u;
            pf_off = 0;
        }
        for (int sub = 0; sub < nsub; ++sub) {
            const int si = d.paired ? sub : it.si;
            const int row0 = it.row0;
            const int nrows = it.nrows;
            for (int kt = 0; kt < d.KT; ++kt) {
                const int col0 = kt * d.CT;
                const int ncols = min(d.CT, C - col0);
                const long long t0 = prof ? clock64() : 0;
                if (pf_on && !mbar_try_wait(ring.empty_bar(), ring.empty_parity())) {
                    const uint32_t issue_off = (uint32_t)(row0 - blk_row0) * (uint32_t)C * 2u;
                    const uint32_t lo = issue_off + (d.paired ? ring_bytes / 2 : ring_bytes);
                    if (pf_off < lo) pf_off = lo;
                    if (pf_off < blk_bytes && pf_off < lo + kL2Ahead) {
                        const uint32_t n = min(32768u, blk_bytes - pf_off);
                        if (lane == 0) {
                            const uint16_t* b0 = (d.paired ? w0p : (si == 0 ? w0p : si == 1 ? w1p : w2p)) + (size_t)blk_row0 * C;
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)b0 + pf_off), "r"(n) : "memory");
                            if (d.paired) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)(w1p + (size_t)blk_row0 * C) + pf_off), "r"(n) : "memory");
                        }
                        pf_off += n;
                    }
                }
                if (!mbar_wait(p, ring.empty_bar(), ring.empty_parity(), 1)) return false;
                if (prof) ring.wait_cyc += clock64() - t0;
                const uint32_t dst = ring.slot_addr(), fb = ring.full_bar();
                if (lane == 0) mbar_expect_tx(fb, (uint32_t)nrows * ncols * 2u);
                __syncwarp();
                const uint16_t* src = (si == 0 ? w0p : si == 1 ? w1p : w2p) + (size_t)row0 * C + col0;
                if (ncols == C) {
                    // rows are contiguous: split the tile into <= 4 KB pieces issued by different lanes
                    const uint32_t total = (uint32_t)nrows * ncols * 2u;
                    for (uint32_t off = (uint32_t)lane * 4096u; off < total; off += 32u * 4096u)
                        bulk_g2s(dst + off, (const char*)src + off, min(4096u, total - off), fb);
                } else {
                    for (int r = lane; r < nrows; r += 32)
                        bulk_g2s(dst + (uint32_t)r * ncols * 2u, src + (size_t)r * C, (uint32_t)ncols * 2u, fb);
                }
                ring.advance();
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
                const long long w0 = p.prof ? clock64() : 0;
                if (!mbar_wait(p, ring.empty_bar(), ring.empty_parity(), 2)) return false;
                if (p.prof) ring.wait_cyc += clock64() - w0;
                const float* base = (kv == 0 ? L.key_cache : L.value_cache) + ((size_t)h * p.n_ctx + pos) * D;
                if (lane == 0) { mbar_expect_tx(ring.full_bar(), bytes); bulk_g2s(ring.slot_addr(), base, bytes, ring.full_bar()); }
                __syncwarp();
                ring.advance();
            }
        }
    }
    return true;
}

__device__ void producer_main(const DecParams& p, Ring ring) {
    const long long t0 = clock64();
    const int nsteps = 5 * p.n_layer + 1;
    int l = 0, k = 0;                                     // same step order as the consumers (StepKind)
    bool ok = true;
    for (int i = 0; i < nsteps && ok; ++i) {
        const thk_llama_layer* L = p.layers + (l < p.n_layer ? l : p.n_layer - 1);
        if (k == 1) {
            ok = produce_att_phase(p, ring, *L);
        } else {
            const int ph = k == 0 ? PH_QKV : k == 2 ? PH_WO : k == 3 ? PH_W13 : k == 4 ? PH_W2 : PH_OUT;
            const uint16_t* w0 = k == 0 ? L->wq : k == 2 ? L->wo : k == 3 ? L->w1 : k == 4 ? L->w2 : p.out_w;
            const uint16_t* w1 = k == 0 ? L->wk : k == 3 ? L->w3 : nullptr;
            const uint16_t* w2 = k == 0 ? L->wv : nullptr;
            ok = produce_mat_phase(p, ring, p.ph[ph], w0, w1, w2);
        }
        if (k == 4) { ++l; k = (l < p.n_layer) ? 0 : 5; } else ++k;
    }
    if (p.prof && (threadIdx.x & 31) == 0) {
        unsigned long long* stt = p.prof + (size_t)gridDim.x * kProfPhases * 8 + (size_t)blockIdx.x * 4;
        stt[0] = (unsigned long long)ring.wait_cyc; stt[1] = (unsigned long long)(clock64() - t0); stt[2] = ring.tc;
    }
}

// ------------------------------------------------------------------------------------------
// CONSUMERS: warps 1..8 do the math, warp 9 runs the epilogues
// ------------------------------------------------------------------------------------------
struct Cons {
    const DecParams& p;
    Ring ring;
    unsigned char* slots;
    float* xs;
    SmemMisc* sm;
    int ct, cw, lane;        // math thread id 0..255 (epilogue warp: 256..287), math warp 0..7, lane
    unsigned nbar;           // grid barriers passed
    bool ok;
};

__device__ __forceinline__ void prof_mark(const Cons& c, int slot) {
    if (c.p.prof && c.ct == 0 && c.nbar < (unsigned)kProfPhases)
        c.p.prof[((size_t)blockIdx.x * kProfPhases + c.nbar) * 8 + slot] = gtimer();
}

// Grid-wide barrier between data-dependent phases.  Every math and epilogue thread has issued its
// global writes; bar.sync orders them before thread 0's gpu-scope release (cumulativity), and the
// acquire + second bar.sync make the other CTAs' writes visible to all threads here (read with .cg).
__device__ __forceinline__ void grid_barrier(Cons& c, int xset = -1, unsigned epoch = 0u) {
    prof_mark(c, PROF_ARRIVE);
    bar_sync(BAR_ALL, kMathThreads + 32);
    ++c.nbar;
    if (c.ct == 0) {
        unsigned* const ctr = c.p.bar_ctr;              // read the parameter block once, not per poll
        red_release_add(ctr, 1u);
        const unsigned target = c.nbar * gridDim.x;
        unsigned long long t0 = 0;
        unsigned it = 0;
        while (ld_acquire_u32(ctr) < target) {
            if ((++it & 63u) == 0u) {
                if (t0 == 0) t0 = gtimer();
                if (aborted(c.p)) break;
                if (gtimer() - t0 > c.p.timeout_ns) { raise_abort(c.p, 0x200u, c.nbar, target); break; }
            }
        }
        if (xset >= 0) {
            // Cross-GPU step of the one-shot all-reduce: every local CTA has pushed its partial rows into all
            // peers (system-scope fenced before arriving here); rank-level flag tells the peers "my part is in".
            const DecParams& p = c.p;
            if (blockIdx.x == 0) {
                __threadfence_system();
                for (int d = 0; d < p.tp_size; ++d) st_release_sys(xflags(p, d, xset) + p.tp_rank, epoch);
            }
            t0 = 0; it = 0;
            for (int src = 0; src < p.tp_size; ++src) {
                const unsigned* f = xflags(p, p.tp_rank, xset) + src;
                while ((int)(ld_acquire_sys(f) - epoch) < 0) {
                    if ((++it & 63u) == 0u) {
                        if (t0 == 0) t0 = gtimer();
                        if (aborted(p)) break;
                        if (gtimer() - t0 > p.timeout_ns) { raise_abort(p, 0x400u, (unsigned)src, epoch); break; }
                    }
                }
            }
        }
    }
    bar_sync(BAR_ALL, kMathThreads + 32);
    prof_mark(c, PROF_START);
}

__device__ __forceinline__ void prefetch_l2(const void* base, int bytes, int tid, int nthreads) {
    for (int off = tid * 128; off < bytes; off += nthreads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)base + off));
}

// xs <- rmsnorm(v) * gain with v = src (+ the tp partial vectors of exchange `which`, in rank order)
// (cmdbuf_rms_norm + cmdbuf_row_element_multiply, th.cpp:1153-1200,1298-1315; the residual adds of
// th-llama.cpp:409/447 move here under tensor parallelism).  When `out` is given, v is also written
// back as the new residual stream, each float4 by exactly one CTA.
__device__ __forceinline__ float4 load_summed(const DecParams& p, const float* src, int which, int i) {
    float4 v = __ldcg((const float4*)(src + i));
    if (which >= 0) {
        for (int r = 0; r < p.tp_size; ++r) {
            const float4 a = __ldcg((const float4*)(xb_ptr(p, p.tp_rank, which, r) + i));
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
    }
    return v;
}
__device__ __forceinline__ void prologue_norm(Cons& c, const float* src, const float* gain, int n, int which = -1, float* out = nullptr) {
    constexpr int kHold = 4;                         // float4s kept in registers per thread (n <= 4096)
    float4 v[kHold], g[kHold];
    float ss = 0.f;
    const bool held = n <= kHold * kMathThreads * 4;
    if (held) {
#pragma unroll
        for (int k = 0; k < kHold; ++k) {
            const int i = (c.ct + k * kMathThreads) * 4;
            if (i < n) { v[k] = load_summed(c.p, src, which, i); g[k] = __ldg((const float4*)(gain + i)); }
        }
#pragma unroll
        for (int k = 0; k < kHold; ++k) {
            const int i = (c.ct + k * kMathThreads) * 4;
            if (i < n) { ss = fmaf(v[k].x, v[k].x, ss); ss = fmaf(v[k].y, v[k].y, ss); ss = fmaf(v[k].z, v[k].z, ss); ss = fmaf(v[k].w, v[k].w, ss); }
        }
    } else {
        for (int i = c.ct * 4; i < n; i += kMathThreads * 4) {
            const float4 t = load_summed(c.p, src, which, i);
            ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
        }
    }
    ss = warp_sum(ss);
    if (c.lane == 0) c.sm->norm_part[c.cw] = ss;
    bar_sync(BAR_MATH, kMathThreads);
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kMathWarps; ++w) tot += c.sm->norm_part[w];
    const float inv = 1.0f / sqrtf(tot / (float)n + 1e-6f);
    if (held) {
#pragma unroll
        for (int k = 0; k < kHold; ++k) {
            const int i = (c.ct + k * kMathThreads) * 4;
            if (i < n) {
                float4 o;
                o.x = (v[k].x * inv) * g[k].x; o.y = (v[k].y * inv) * g[k].y; o.z = (v[k].z * inv) * g[k].z; o.w = (v[k].w * inv) * g[k].w;
                *(float4*)(c.xs + xs_index(i)) = o;
                if (out && (unsigned)(i >> 2) % gridDim.x == blockIdx.x) *(float4*)(out + i) = v[k];
            }
        }
    } else {
        for (int i = c.ct * 4; i < n; i += kMathThreads * 4) {
            const float4 t = load_summed(c.p, src, which, i);
            const float4 gg = __ldg((const float4*)(gain + i));
            float4 o;
            o.x = (t.x * inv) * gg.x; o.y = (t.y * inv) * gg.y; o.z = (t.z * inv) * gg.z; o.w = (t.w * inv) * gg.w;
            *(float4*)(c.xs + xs_index(i)) = o;
            if (out && (unsigned)(i >> 2) % gridDim.x == blockIdx.x) *(float4*)(out + i) = t;
        }
    }
    bar_sync(BAR_MATH, kMathThreads);
    prof_mark(c, PROF_PROLOGUE);
}
__device__ __forceinline__ void prologue_copy(Cons& c, const float* src, int n) {
    for (int i = c.ct * 4; i < n; i += kMathThreads * 4)
        *(float4*)(c.xs + xs_index(i)) = __ldcg((const float4*)(src + i));
    bar_sync(BAR_MATH, kMathThreads);
    prof_mark(c, PROF_PROLOGUE);
}

// 8 weights (one uint4 of f16) times 8 activations into two independent accumulators
__device__ __forceinline__ void fma8(const uint4& w, const float4& x0, const float4& x1, float& a0, float& a1) {
    float2 f;
    f = h2_to_f2(w.x); a0 = fmaf(x0.x, f.x, a0); a1 = fmaf(x0.y, f.y, a1);
    f = h2_to_f2(w.y); a0 = fmaf(x0.z, f.x, a0); a1 = fmaf(x0.w, f.y, a1);
    f = h2_to_f2(w.z); a0 = fmaf(x1.x, f.x, a0); a1 = fmaf(x1.y, f.y, a1);
    f = h2_to_f2(w.w); a0 = fmaf(x1.z, f.x, a0); a1 = fmaf(x1.w, f.y, a1);
}

// One tile: acc[r][0..1] += sum over this warp's columns of W[row][col] * x[col].  Per 256-column
// chunk all shared-memory loads are issued before the FMAs; branches are warp-uniform except in the
// last partial chunk of a shard.
template <int RPW, int CPW>
__device__ __forceinline__ void tile_fma(const unsigned char* slot, const float* xs, int WC, int wr, int wc, int lane,
                                         int col0, int nrows, int ncols, float (&acc)[kRC][2]) {
    const int row_base = wr * RPW;
    const bool full_rows = row_base + RPW <= nrows;
#pragma unroll
    for (int cp = 0; cp < CPW; ++cp) {
        const int cbase = (cp * WC + wc) << 8;                 // first column of this warp's chunk in the tile
        if (cbase >= ncols) break;                             // warp-uniform
        const int col = cbase + (lane << 3);
        const float* xp = xs + col0 + cbase + (lane << 2);
        const unsigned char* wp = slot + ((size_t)row_base * ncols + col) * 2;
        if (cbase + 256 <= ncols && full_rows) {               // warp-uniform fast path
            const float4 x0 = *(const float4*)xp;
            const float4 x1 = *(const float4*)(xp + 128);
            uint4 w[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r) w[r] = *(const uint4*)(wp + (size_t)r * ncols * 2);
#pragma unroll
            for (int r = 0; r < RPW; ++r) fma8(w[r], x0, x1, acc[r][0], acc[r][1]);
        } else if (col < ncols) {
            const float4 x0 = *(const float4*)xp;
            const float4 x1 = *(const float4*)(xp + 128);
#pragma unroll
            for (int r = 0; r < RPW; ++r)
                if (row_base + r < nrows) fma8(*(const uint4*)(wp + (size_t)r * ncols * 2), x0, x1, acc[r][0], acc[r][1]);
        }
    }
}

// transposing warp reduction of RPW accumulators: afterwards lane (row << (5-log2 RPW)) holds row's sum
template <int RPW>
__device__ __forceinline__ float reduce_rows(float (&v)[kRC], int lane) {
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

// ---- math warps: stream tiles, leave per-warp partial row sums in sm->red[buf] for the epilogue warp ----
template <int RPW, int CPW>
__device__ __forceinline__ void math_mat_phase_t(Cons& c, const PhaseDesc& dref) {
    const DecParams& p = c.p;
    const PhaseDesc d = dref;                      // one read of the parameter block; the loops below use registers
    const bool prof = p.prof != nullptr;
    const int WC = d.WC, C = d.C;
    const int wr = c.cw / WC, wc = c.cw % WC;
    constexpr int kLog = (RPW == 8) ? 3 : (RPW == 4) ? 2 : (RPW == 2) ? 1 : 0;
    unsigned gq = 0;                                            // (group, sub) pairs handed to the epilogue warp
    bool first = true;
    RowIt it;
    for (it.init(d); it.valid(); it.next(d)) {
        const int nsub = d.paired ? 2 : 1;
        for (int sub = 0; sub < nsub; ++sub) {
            const int nrows = it.nrows;
            float acc[kRC][2];
#pragma unroll
            for (int r = 0; r < kRC; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
            for (int kt = 0; kt < d.KT; ++kt) {
                const int col0 = kt * d.CT;
                const int ncols = min(d.CT, C - col0);
                const long long w0 = prof ? clock64() : 0;
                if (c.ok) c.ok = mbar_wait(p, c.ring.full_bar(), c.ring.full_parity(), 3);
                if (prof) c.ring.wait_cyc += clock64() - w0;
                if (first) { if (prof) prof_mark(c, PROF_FIRST_TILE); first = false; }
                if (c.ok) tile_fma<RPW, CPW>(c.slots + (size_t)c.ring.slot() * kSlotBytes, c.xs, WC, wr, wc, c.lane, col0, nrows, ncols, acc);
                __syncwarp();
                if (c.lane == 0) mbar_arrive(c.ring.empty_bar());
                c.ring.advance();
            }
            float v[kRC];
#pragma unroll
            for (int r = 0; r < kRC; ++r) v[r] = acc[r][0] + acc[r][1];
            const float rsum = reduce_rows<RPW>(v, c.lane);
            const int buf = gq & 1;
            if (gq >= 2) bar_sync(BAR_B0 + buf, kMathThreads + 32);      // epilogue warp is done with this buffer
            if ((c.lane & ((32 >> kLog) - 1)) == 0) c.sm->red[buf][wc][wr * RPW + (c.lane >> (5 - kLog))] = rsum;
            bar_arrive(BAR_A0 + buf, kMathThreads + 32);                 // partials ready (non-blocking)
            ++gq;
        }
    }
    for (unsigned k = gq >= 2 ? gq - 2 : 0; k < gq; ++k) bar_sync(BAR_B0 + (k & 1), kMathThreads + 32);   // drain
    prof_mark(c, PROF_LAST_TILE);
    if (p.prof && c.ct == 0 && c.nbar < (unsigned)kProfPhases) {
        p.prof[((size_t)blockIdx.x * kProfPhases + c.nbar) * 8 + PROF_WAIT_FULL] = (unsigned long long)c.ring.wait_cyc;
        c.ring.wait_cyc = 0;
    }
}

// ---- epilogue warp ----
enum EpiKind { EPI_QKV, EPI_WO, EPI_W13, EPI_W2, EPI_OUT };
struct EpiState { float gate[2]; float best; int best_idx; };

__device__ __forceinline__ void epi_mat_phase(Cons& c, const PhaseDesc& dref, EpiKind kind, const thk_llama_layer* L, EpiState& es) {
    const DecParams& p = c.p;
    const PhaseDesc d = dref;
    const int WC = d.WC, D = p.head_dim, tp_size = p.tp_size, tp_rank = p.tp_rank, n_ctx = p.n_ctx, n_past = p.n_past, Vl = p.Vl;
    float* const px = p.x; float* const ph1 = p.h1; float* const pq = p.q; float* const pff = p.ff; float* const plogits = p.logits;
    float* const kcache = L ? L->key_cache : nullptr; float* const vcache = L ? L->value_cache : nullptr;
    unsigned gq = 0;
    RowIt it;
    for (it.init(d); it.valid(); it.next(d)) {
        const int nsub = d.paired ? 2 : 1;
        for (int sub = 0; sub < nsub; ++sub) {
            const int si = d.paired ? sub : it.si;
            const int row0 = it.row0;
            const int nrows = it.nrows;
            float resid[2] = {0.f, 0.f};
            if ((kind == EPI_WO || kind == EPI_W2) && tp_size == 1) {   // residual operand: load before waiting for the sums
                const float* rs = (kind == EPI_WO) ? px : ph1;
#pragma unroll
                for (int k = 0; k < 2; ++k) { const int t = c.lane + 32 * k; if (t < nrows) resid[k] = __ldcg(rs + row0 + t); }
            }
            const int buf = gq & 1;
            bar_sync(BAR_A0 + buf, kMathThreads + 32);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int t = c.lane + 32 * k;
                if (t >= nrows) continue;
                float y = 0.f;
                for (int w = 0; w < WC; ++w) y += c.sm->red[buf][w][t];
                const int r = row0 + t;
                switch (kind) {
                case EPI_QKV:
                    if (si == 2) {                                   // V: append (th-llama.cpp:338)
                        vcache[((size_t)(r / D) * n_ctx + n_past) * D + (r % D)] = y;
                    } else if ((t & 1) == 0) {                       // Q / K: rotate the pair (t, t+1), th.cpp:1457-1492
                        float y1 = 0.f;
                        for (int w = 0; w < WC; ++w) y1 += c.sm->red[buf][w][t + 1];
                        const float2 cs = c.sm->rope[(r % D) >> 1];
                        const float a = y * cs.x - y1 * cs.y, b = y * cs.y + y1 * cs.x;
                        if (si == 0) { pq[r] = a; pq[r + 1] = b; }
                        else {
                            float* kc = kcache + ((size_t)(r / D) * n_ctx + n_past) * D + (r % D);
                            kc[0] = a; kc[1] = b;                    // th-llama.cpp:337
                        }
                    }
                    break;
                case EPI_WO:                                                           // th-llama.cpp:409
                    if (tp_size == 1) ph1[r] = resid[k] + y;
                    else for (int dst = 0; dst < tp_size; ++dst) xb_ptr(p, dst, 0, tp_rank)[r] = y;   // partial -> every rank
                    break;
                case EPI_W13:
                    if (sub == 0) es.gate[k] = y;
                    else { const float gv = es.gate[k]; pff[r] = (gv / (1.0f + expf(-gv))) * y; }   // :436,:438
                    break;
                case EPI_W2:                                                           // th-llama.cpp:447
                    if (tp_size == 1) px[r] = resid[k] + y;
                    else for (int dst = 0; dst < tp_size; ++dst) xb_ptr(p, dst, 1, tp_rank)[r] = y;
                    break;
                case EPI_OUT: {
                    if (plogits) plogits[r] = y;
                    const int gid = tp_rank * Vl + r;
                    if (es.best_idx < 0 || y > es.best) { es.best = y; es.best_idx = gid; }
                } break;
                }
            }
            __syncwarp();
            bar_arrive(BAR_B0 + buf, kMathThreads + 32);
            ++gq;
        }
    }
}

// ------------------------------------------------------------------------------------------
// attention: single query, split-KV, online softmax (math warps only)
// (cmdbuf_mat_mul QK^T * 1/sqrt(D), cmdbuf_row_softmax, cmdbuf_mat_mul P*V; th-llama.cpp:365-380)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    float s = a.x * b.x;
    s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    return s;
}

__device__ __forceinline__ void math_att_phase(Cons& c, const thk_llama_layer& L) {
    const DecParams& p = c.p;
    const AttSched a = make_att(p);
    const int D = p.head_dim, nvec = D >> 2;
    const float scale = 1.0f / sqrtf((float)D);
    const bool act = c.lane < nvec;
    // K scoring: 4 threads per position, each owns a quarter of the head dimension (needs D % 16 == 0)
    const bool quad = (D % 16 == 0) && D <= 128;
    const int kq = c.ct & 3, kp = c.ct >> 2, nv = D >> 4;       // quarter, position slot 0..63, float4s per quarter
    int rot = 0;                                                // bank-conflict-free visiting order of the quarter's float4s
    if (nv == 8) rot = (kq + 4 * (kp & 1)) & 7;
    else if (nv == 4) rot = ((kq >> 1) + 2 * (kp & 1)) & 3;
    else if (nv == 2) rot = kp & 1;
    for (int u = blockIdx.x; u < p.Hl * a.S; u += gridDim.x) {
        const int h = u / a.S, sp = u % a.S;
        const int pa = (int)(((long long)a.N * sp) / a.S);
        const int pb_full = (int)(((long long)a.N * (sp + 1)) / a.S);
        const int pb = min(pb_full, p.n_past);
        const bool has_new = (pb_full == a.N);
        float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) q4 = __ldcg((const float4*)(p.q + h * D) + c.lane);
        float4 qr[8];
        if (quad) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nv) qr[j] = __ldcg((const float4*)(p.q + h * D + kq * (D >> 2)) + ((j + rot) & (nv - 1)));
        }
        float m = -INFINITY, lsum = 0.f;
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int buf = 0;
        for (int pos = pa; pos < pb; pos += p.att_tpos) {
            const int np = min(p.att_tpos, pb - pos);
            // K tile -> scores
            if (c.ok) c.ok = mbar_wait(p, c.ring.full_bar(), c.ring.full_parity(), 4);
            if (c.ok) {
                const float* kt = (const float*)(c.slots + (size_t)c.ring.slot() * kSlotBytes);
                if (quad) {
                    for (int j0 = 0; j0 < np; j0 += kMathThreads / 4) {
                        const int j = j0 + kp;
                        float sdot = 0.f;
                        if (j < np) {
                            const float4* kr = (const float4*)(kt + (size_t)j * D + kq * (D >> 2));
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (i < nv) sdot += dot4(qr[i], kr[(i + rot) & (nv - 1)]);
                        }
                        sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
                        sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
                        if (kq == 0 && j < np) c.sm->sc[buf][j] = sdot * scale;
                    }
                } else {
                    for (int j = c.cw; j < np; j += kMathWarps) {
                        float sdot = 0.f;
                        if (act) sdot = dot4(q4, *(const float4*)(kt + j * D + c.lane * 4));
                        sdot = warp_sum(sdot) * scale;
                        if (c.lane == 0) c.sm->sc[buf][j] = sdot;
                    }
                }
            }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(c.ring.empty_bar());
            c.ring.advance();
            bar_sync(BAR_MATH, kMathThreads);
            float bmax = -INFINITY;
            for (int j = c.lane; j < np; j += 32) bmax = fmaxf(bmax, c.sm->sc[buf][j]);
            bmax = warp_max(bmax);
            const float m_new = fmaxf(m, bmax);
            const float corr = expf(m - m_new);
            lsum *= corr; o4.x *= corr; o4.y *= corr; o4.z *= corr; o4.w *= corr;
            m = m_new;
            bar_sync(BAR_MATH, kMathThreads);                  // everyone has read the raw scores
            for (int j = c.ct; j < np; j += kMathThreads) c.sm->sc[buf][j] = expf(c.sm->sc[buf][j] - m);   // one exp per position
            bar_sync(BAR_MATH, kMathThreads);
            // V tile -> weighted sum
            if (c.ok) c.ok = mbar_wait(p, c.ring.full_bar(), c.ring.full_parity(), 5);
            if (c.ok) {
                const float* vt = (const float*)(c.slots + (size_t)c.ring.slot() * kSlotBytes);
                int j = c.cw;
                for (; j + 3 * kMathWarps < np; j += 4 * kMathWarps) {     // 4 positions in flight
                    float pj[4]; float4 v4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        pj[u] = c.sm->sc[buf][j + u * kMathWarps];
                        v4[u] = act ? *(const float4*)(vt + (j + u * kMathWarps) * D + c.lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        lsum += pj[u];
                        o4.x = fmaf(pj[u], v4[u].x, o4.x); o4.y = fmaf(pj[u], v4[u].y, o4.y);
                        o4.z = fmaf(pj[u], v4[u].z, o4.z); o4.w = fmaf(pj[u], v4[u].w, o4.w);
                    }
                }
                for (; j < np; j += kMathWarps) {
                    const float pj = c.sm->sc[buf][j];
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
            c.ring.advance();
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
        bar_sync(BAR_MATH, kMathThreads);
        float* part = p.part + (size_t)(h * p.att_max_split + sp) * (D + 2);
        if (c.ct < D) {
            float od = 0.f;
#pragma unroll
            for (int w = 0; w < kMathWarps; ++w) od += c.sm->redo[w][c.ct];
            part[2 + c.ct] = od;
        }
        if (c.ct == 0) {
            float lt = 0.f;
#pragma unroll
            for (int w = 0; w < kMathWarps; ++w) lt += c.sm->redl[w];
            part[0] = m; part[1] = lt;
        }
        // last split to finish combines the head (deterministic: fixed split order)
        bar_sync(BAR_MATH, kMathThreads);
        if (c.ct == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(p.head_ctr + h, 1u);
            c.sm->flag = (prev == (unsigned)(a.S - 1));
            if (c.sm->flag) { p.head_ctr[h] = 0u; __threadfence(); }
        }
        bar_sync(BAR_MATH, kMathThreads);
        if (c.sm->flag) {
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
        bar_sync(BAR_MATH, kMathThreads);   // sm->flag / redo reuse
    }
}

// ---- out-of-line phase bodies ----
// One out-of-line copy of each phase routine keeps the kernel's instruction footprint small (the
// fully inlined kernel was 400 KB of SASS and missed the instruction cache at every phase change).
// State crosses the call boundary BY VALUE (registers), never through a reference to a stack object.
enum StepKind { K_QKV = 0, K_ATT = 1, K_WO = 2, K_W13 = 3, K_W2 = 4, K_OUT = 5 };
__device__ __forceinline__ int phase_of(int k) { return k == K_QKV ? PH_QKV : k == K_WO ? PH_WO : k == K_W13 ? PH_W13 : k == K_W2 ? PH_W2 : PH_OUT; }

struct Shared { unsigned char* slots; float* xs; SmemMisc* sm; };
struct MState { uint32_t tc; int ok; unsigned nbar; long long wait_cyc; };
struct EpiRet { MState st; EpiState es; };

__device__ __forceinline__ Cons cons_from(const DecParams& p, const Shared& S, const MState& st, bool epi) {
    Ring ring(smem_u32(S.slots), smem_u32(&S.sm->full[0]), smem_u32(&S.sm->empty[0]), st.tc, st.wait_cyc);
    const int lane = (int)threadIdx.x & 31;
    if (epi) return Cons{p, ring, S.slots, S.xs, S.sm, kMathThreads + lane, kMathWarps, lane, st.nbar, st.ok != 0};
    return Cons{p, ring, S.slots, S.xs, S.sm, (int)threadIdx.x - kMathBase, ((int)threadIdx.x - kMathBase) >> 5, lane, st.nbar, st.ok != 0};
}
__device__ __forceinline__ MState state_of(const Cons& c) { return MState{c.ring.tc, c.ok ? 1 : 0, c.nbar, c.ring.wait_cyc}; }

__device__ __noinline__ MState nl_grid_barrier(const DecParams& p, Shared S, MState st, int epi, int xset = -1, unsigned epoch = 0u) {
    Cons c = cons_from(p, S, st, epi != 0);
    if (epi && xset >= 0) __threadfence_system();      // this warp's pushes to the peers are visible system-wide
    grid_barrier(c, xset, epoch);
    return state_of(c);
}
__device__ __noinline__ MState nl_prologue_norm(const DecParams& p, Shared S, MState st, const float* src, const float* gain, int n,
                                                int which = -1, float* out = nullptr) {
    Cons c = cons_from(p, S, st, false);
    prologue_norm(c, src, gain, n, which, out);
    return state_of(c);
}
__device__ __noinline__ MState nl_prologue_copy(const DecParams& p, Shared S, MState st, const float* src, int n) {
    Cons c = cons_from(p, S, st, false);
    prologue_copy(c, src, n);
    return state_of(c);
}
template <int RPW, int CPW>
__device__ __noinline__ MState nl_math_mat(const DecParams& p, Shared S, MState st, int ph) {
    Cons c = cons_from(p, S, st, false);
    math_mat_phase_t<RPW, CPW>(c, p.ph[ph]);
    return state_of(c);
}
__device__ __forceinline__ MState math_mat(const DecParams& p, Shared S, MState st, int ph) {
    const int cpw = p.ph[ph].CPW;
    if (cpw == 1) return nl_math_mat<kRC, 1>(p, S, st, ph);
    if (cpw == 2) return nl_math_mat<kRC / 2, 2>(p, S, st, ph);
    return nl_math_mat<kRC / 4, 4>(p, S, st, ph);
}
__device__ __noinline__ MState nl_math_att(const DecParams& p, Shared S, MState st, int layer) {
    Cons c = cons_from(p, S, st, false);
    math_att_phase(c, p.layers[layer]);
    return state_of(c);
}
__device__ __noinline__ EpiRet nl_epi_mat(const DecParams& p, Shared S, MState st, EpiState es, int ph, int kind, int layer) {
    Cons c = cons_from(p, S, st, true);
    epi_mat_phase(c, p.ph[ph], (EpiKind)kind, layer >= 0 ? p.layers + layer : nullptr, es);
    return EpiRet{state_of(c), es};
}

__device__ void math_main(const DecParams& p, Shared S) {
    MState st{0u, 1, 0u, 0};
    const int tid = (int)threadIdx.x - kMathBase;
    {
        Cons c = cons_from(p, S, st, false);
        prof_mark(c, PROF_START);
        // phase E: x <- f32(tok_embeddings[token]) (th-llama.cpp:577-585; device-side like :552-575)
        int tok = *p.token;
        if (tok < 0 || tok >= p.n_vocab) { if (c.ct == 0) raise_abort(p, 0x300u, (unsigned)tok, 0); tok = 0; }
        const uint16_t* row = p.emb + (size_t)tok * p.n_embd;
        for (int i = (blockIdx.x * kMathThreads + c.ct); i < p.n_embd; i += gridDim.x * kMathThreads)
            p.x[i] = __half2float(__ushort_as_half(row[i]));
        if (blockIdx.x == 0 && c.ct == 0) *p.bar_next = 0u;   // arm the next launch's barrier counter
        prefetch_l2(p.layers[0].attention_norm, p.n_embd * 4, c.ct, kMathThreads);
    }
    st = nl_grid_barrier(p, S, st, 0);
    const int nsteps = 5 * p.n_layer + 1;
    const bool tp = p.tp_size > 1;
    unsigned xk = 0;                                       // cross-GPU exchanges done in this launch
    int l = 0, k = K_QKV;
    for (int i = 0; i < nsteps; ++i) {
        const thk_llama_layer* L = p.layers + (l < p.n_layer ? l : p.n_layer - 1);
        if (k == K_QKV || k == K_W13 || k == K_OUT) {
            const float* gain = (k == K_QKV) ? L->attention_norm : (k == K_W13) ? L->ffn_norm : p.norm;
            if (!tp) {
                st = nl_prologue_norm(p, S, st, (k == K_W13) ? p.h1 : p.x, gain, p.n_embd);
            } else if (k == K_W13) {
                st = nl_prologue_norm(p, S, st, p.x, gain, p.n_embd, 0, p.h1);         // h1 = x + sum of Wo partials
            } else if (i == 0) {
                st = nl_prologue_norm(p, S, st, p.x, gain, p.n_embd);                    // first layer: x is the embedding
            } else {
                st = nl_prologue_norm(p, S, st, p.h1, gain, p.n_embd, 1, p.x);         // x = h1 + sum of W2 partials
            }
        } else if (k == K_WO || k == K_W2) {
            st = nl_prologue_copy(p, S, st, k == K_WO ? p.o : p.ff, k == K_WO ? p.Eh : p.Fh);
        }
        if (k == K_ATT) st = nl_math_att(p, S, st, l);
        else st = math_mat(p, S, st, phase_of(k));
        if (k == K_WO) prefetch_l2(L->ffn_norm, p.n_embd * 4, tid, kMathThreads);
        if (k == K_W2) prefetch_l2(l + 1 < p.n_layer ? p.layers[l + 1].attention_norm : p.norm, p.n_embd * 4, tid, kMathThreads);
        if (tp && (k == K_WO || k == K_W2)) { ++xk; st = nl_grid_barrier(p, S, st, 0, 0, p.epoch_base + xk); }
        else if (k != K_OUT || p.next_token || p.next_logit) st = nl_grid_barrier(p, S, st, 0);
        if (k == K_W2) { ++l; k = (l < p.n_layer) ? K_QKV : K_OUT; } else ++k;
    }
}

__device__ void epi_main(const DecParams& p, Shared S) {
    MState st{0u, 1, 0u, 0};
    EpiState es{{0.f, 0.f}, 0.f, -1};
    const int lane = (int)threadIdx.x & 31;
    // RoPE table for this token: cos/sin(n_past * 10000^(-2i/D)) (th.cpp:1478-1482), once per CTA
    for (int i = lane; i < (p.head_dim >> 1); i += 32) {
        const float theta = powf(10000.0f, (-(float)(2 * i)) / (float)p.head_dim);
        float sn, cs;
        sincosf((float)p.n_past * theta, &sn, &cs);
        S.sm->rope[i] = make_float2(cs, sn);
    }
    __syncwarp();
    st = nl_grid_barrier(p, S, st, 1);
    const int nsteps = 5 * p.n_layer + 1;
    unsigned xk = 0;
    int l = 0, k = K_QKV;
    for (int i = 0; i < nsteps; ++i) {
        if (k != K_ATT) {
            const int kind = k == K_QKV ? EPI_QKV : k == K_WO ? EPI_WO : k == K_W13 ? EPI_W13 : k == K_W2 ? EPI_W2 : EPI_OUT;
            const EpiRet r = nl_epi_mat(p, S, st, es, phase_of(k), kind, k == K_OUT ? -1 : l);
            st = r.st; es = r.es;
        }
        if (k == K_OUT) break;
        if (p.tp_size > 1 && (k == K_WO || k == K_W2)) { ++xk; st = nl_grid_barrier(p, S, st, 1, 0, p.epoch_base + xk); }
        else st = nl_grid_barrier(p, S, st, 1);
        if (k == K_W2) { ++l; k = (l < p.n_layer) ? K_QKV : K_OUT; } else ++k;
    }
    if (p.next_token || p.next_logit) {
        // greedy argmax (th-llama.cpp:826-838): lowest index wins ties; warp -> CTA -> grid
        float bv = es.best; int bi = es.best_idx;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { p.amax_val[blockIdx.x] = bv; p.amax_idx[blockIdx.x] = bi; }
        st = nl_grid_barrier(p, S, st, 1);
        if (blockIdx.x == 0 && lane == 0) {
            float gv = 0.f; int gi = -1;
            for (unsigned b = 0; b < gridDim.x; ++b) {
                const int idx = __ldcg(p.amax_idx + b);
                if (idx < 0) continue;
                const float v = __ldcg(p.amax_val + b);
                if (gi < 0 || v > gv || (v == gv && idx < gi)) { gv = v; gi = idx; }
            }
            if (p.tp_size > 1) {
                // cross-rank argmax: push this rank's candidate to every rank, wait for all, lowest id wins ties
                const unsigned epoch = p.epoch_base + 2u * (unsigned)p.n_layer + 1u;
                for (int d = 0; d < p.tp_size; ++d) { xamax_val(p, d)[p.tp_rank] = gv; xamax_idx(p, d)[p.tp_rank] = gi; }
                __threadfence_system();
                for (int d = 0; d < p.tp_size; ++d) st_release_sys(xflags(p, d, 1) + p.tp_rank, epoch);
                unsigned long long t0 = 0; unsigned it = 0;
                gv = 0.f; gi = -1;
                for (int src = 0; src < p.tp_size; ++src) {
                    const unsigned* f = xflags(p, p.tp_rank, 1) + src;
                    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
                        if ((++it & 63u) == 0u) {
                            if (t0 == 0) t0 = gtimer();
                            if (aborted(p)) break;
                            if (gtimer() - t0 > p.timeout_ns) { raise_abort(p, 0x401u, (unsigned)src, epoch); break; }
                        }
                    }
                    const int idx = __ldcv(xamax_idx(p, p.tp_rank) + src);
                    const float v = __ldcv(xamax_val(p, p.tp_rank) + src);
                    if (idx >= 0 && (gi < 0 || v > gv || (v == gv && idx < gi))) { gv = v; gi = idx; }
                }
            }
            if (p.next_token) *p.next_token = gi < 0 ? 0 : gi;
            if (p.next_logit) *p.next_logit = gv;
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
            mbar_init(smem_u32(&sm->empty[i]), kMathWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Ring ring(smem_u32(slots), smem_u32(&sm->full[0]), smem_u32(&sm->empty[0]), 0u, 0);
    const Shared S{slots, xs, sm};
    if (threadIdx.x < 32) producer_main(p, ring);
    else if (threadIdx.x < kMathBase + kMathThreads) math_main(p, S);
    else epi_main(p, S);
}

size_t decode_smem_bytes(int max_vec) {
    return (size_t)kNumSlots * kSlotBytes + ((sizeof(SmemMisc) + 127) & ~127) + (size_t)((max_vec + 255) & ~255) * sizeof(float);
}

PhaseDesc make_phase(int nseg, const int* rows, int C, bool paired, int n_cta) {
    PhaseDesc d{};
    d.C = C; d.nseg = nseg; d.paired = paired ? 1 : 0;
    for (int i = 0; i < 3; ++i) d.rows[i] = i < nseg ? rows[i] : 0;
    const int chunks = (C + 255) / 256;
    int WC = 1;
    while (WC * 2 <= chunks && WC < 8) WC *= 2;
    const int WR = kMathWarps / WC;
    int cap = 1;
    while (cap * WC < chunks && cap < 4) cap *= 2;     // chunks per warp needed to span C, pow2, <= 4
    int rows_total = 0;
    for (int i = 0; i < nseg; ++i) if (!paired || i == 0) rows_total += rows[i];
    const int opts[3][2] = {{kRC, 1}, {kRC / 2, 2}, {kRC / 4, 4}};
    int RPW = kRC, CPW = 1;
    for (int i = 0; i < 3; ++i) {
        if (opts[i][1] > cap) break;
        RPW = opts[i][0]; CPW = opts[i][1];
        const int G = (rows_total + WR * RPW - 1) / (WR * RPW);
        if (G >= 2 * n_cta) break;      // prefer the widest row group: it streams at the HBM rate
    }
    d.WC = WC; d.RPW = RPW; d.CPW = CPW;
    d.RT = WR * RPW;
    const int per_tile = WC * CPW;
    d.KT = (chunks + per_tile - 1) / per_tile;
    d.CT = ((chunks + d.KT - 1) / d.KT) * 256;          // K split evenly over the tiles
    d.G = 0;
    for (int i = 0; i < 3; ++i) {
        d.gs[i] = i < nseg ? (rows[i] + d.RT - 1) / d.RT : 0;
        if (!paired || i == 0) d.G += d.gs[i];
    }
    return d;
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
    unsigned char* xchg = nullptr;      // tensor-parallel exchange region (local)
    size_t xchg_bytes = 0;
    unsigned epoch = 0;
    bool peers_set = false;
    unsigned long long* d_prof = nullptr;
};

static int check_status(thk_decoder* d) {
    THK_ENTER(d->ctx);
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
    THK_CHECK_ARG(tp <= 8 && dims->tp_rank >= 0 && dims->tp_rank < tp, "thk_decoder_create: tp_size %d / tp_rank %d unsupported (1..8)", tp, dims->tp_rank);
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
    { const int r[3] = {p.Eh, p.Eh, p.Eh}; p.ph[PH_QKV] = make_phase(3, r, p.n_embd, false, d->grid); }
    { const int r[3] = {p.n_embd, 0, 0}; p.ph[PH_WO] = make_phase(1, r, p.Eh, false, d->grid); }
    { const int r[3] = {p.Fh, p.Fh, 0}; p.ph[PH_W13] = make_phase(2, r, p.n_embd, true, d->grid); }
    { const int r[3] = {p.n_embd, 0, 0}; p.ph[PH_W2] = make_phase(1, r, p.Fh, false, d->grid); }
    { const int r[3] = {p.Vl, 0, 0}; p.ph[PH_OUT] = make_phase(1, r, p.n_embd, false, d->grid); }
    if (getenv("THK_DEBUG"))
        for (int i = 0; i < 5; ++i)
            fprintf(stderr, "phase %d: C=%d WC=%d RPW=%d CPW=%d RT=%d CT=%d KT=%d G=%d\n", i, p.ph[i].C, p.ph[i].WC, p.ph[i].RPW, p.ph[i].CPW,
                    p.ph[i].RT, p.ph[i].CT, p.ph[i].KT, p.ph[i].G);
    p.att_tpos = kSlotBytes / (D * 4);
    if (p.att_tpos > kMaxTilePos) p.att_tpos = kMaxTilePos;
    p.att_max_split = d->grid / p.Hl > 0 ? d->grid / p.Hl : 1;
    p.timeout_ns = 4000000000ull;
    p.l2_ahead = getenv("THK_L2_AHEAD") ? atoi(getenv("THK_L2_AHEAD")) : 0;   // measured: hurts (294 vs 322 tok/s), see DESIGN.md
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
    if (tp > 1) {
        d->xchg_bytes = ((size_t)2 * tp * p.n_embd * sizeof(float) + (size_t)4 * tp * sizeof(unsigned) + 255) & ~(size_t)255;
        THK_CUDA(cudaMalloc(&d->xchg, d->xchg_bytes));
        THK_CUDA(cudaMemset(d->xchg, 0, d->xchg_bytes));
        p.xchg[p.tp_rank] = d->xchg;
    }
    THK_CUDA(cudaDeviceSynchronize());
    *out = d;
    return THK_OK;
}

extern "C" int thk_decoder_destroy(thk_decoder* d) {
    if (!d) return THK_OK;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    cudaFree(d->d_layers); cudaFree(d->scratch); cudaFree(d->ctrl); cudaFree(d->d_tok); cudaFree(d->d_prof); cudaFree(d->xchg);
    delete d;
    return THK_OK;
}

static int launch_step(thk_decoder* d, const int32_t* token, int32_t n_past, float* logits, int32_t* next_token, float* next_logit) {
    THK_ENTER(d->ctx);
    DecParams p = d->p;
    p.token = token; p.n_past = n_past; p.logits = logits; p.next_token = next_token; p.next_logit = next_logit;
    if (p.tp_size > 1) {
        if (!d->peers_set) { thk_set_error("thk_decoder_step: tensor-parallel decoder has no peers (call thk_decoder_set_peers)"); return THK_E_INVALID; }
        p.epoch_base = d->epoch;
        d->epoch += 2u * (unsigned)p.n_layer + 2u;
    }
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
    THK_CHECK_ARG(n_past >= 0 && n_past + n_steps <= d->p.n_ctx, "thk_decoder_generate: steps exceed context");
    for (int i = 0; i < n_steps; ++i) {
        const int32_t* tok = (i == 0) ? first_token : tokens_out + (i - 1);
        int rc = launch_step(d, tok, n_past + i, (i == n_steps - 1) ? last_logits : nullptr, tokens_out + i, nullptr);
        if (rc) return rc;
    }
    d->last_launches = n_steps;
    return THK_OK;
}

// timeline profile: per CTA, per phase (= grid barriers passed so far), 8 u64 slots of %globaltimer ns
// (ProfSlot), followed by per-CTA producer stats [empty-wait cycles, total cycles, tiles, 0]
extern "C" int thk_decoder_profile(thk_decoder* d, int enable, unsigned long long* host_out, int n) {
    THK_CHECK_ARG(d, "thk_decoder_profile: null argument");
    THK_ENTER(d->ctx);
    if (enable && !d->d_prof) {
        const size_t nprof = (size_t)d->grid * kProfPhases * 8 + (size_t)d->grid * 4;
        THK_CUDA(cudaMalloc(&d->d_prof, nprof * sizeof(unsigned long long)));
        THK_CUDA(cudaMemset(d->d_prof, 0, nprof * sizeof(unsigned long long)));
    }
    d->p.prof = enable ? d->d_prof : nullptr;
    if (host_out && n > 0 && d->d_prof) {
        THK_CUDA(cudaStreamSynchronize(d->ctx->stream));
        const size_t nprof = (size_t)d->grid * kProfPhases * 8 + (size_t)d->grid * 4;
        THK_CUDA(cudaMemcpy(host_out, d->d_prof, sizeof(unsigned long long) * ((size_t)n > nprof ? nprof : (size_t)n), cudaMemcpyDeviceToHost));
    }
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

extern "C" int thk_decoder_exchange_info(thk_decoder* d, void** buf, size_t* buf_bytes, void** flags, size_t* flag_bytes) {
    THK_CHECK_ARG(d && buf && buf_bytes, "thk_decoder_exchange_info: null argument");
    THK_CHECK_ARG(d->p.tp_size > 1, "thk_decoder_exchange_info: decoder is not tensor parallel");
    *buf = d->xchg; *buf_bytes = d->xchg_bytes;
    if (flags) *flags = d->xchg + (size_t)2 * d->p.tp_size * d->p.n_embd * sizeof(float) + (size_t)2 * d->p.tp_size * sizeof(unsigned);
    if (flag_bytes) *flag_bytes = (size_t)2 * d->p.tp_size * sizeof(unsigned);
    return THK_OK;
}
// peer_bufs[r]: rank r's exchange region as mapped into THIS process/device (cudaIpcOpenMemHandle or same-process
// peer access); the entry for the local rank is ignored.  peer_flags is unused (flags live inside the region).
extern "C" int thk_decoder_set_peers(thk_decoder* d, void* const* peer_bufs, void* const*, int n) {
    THK_CHECK_ARG(d && peer_bufs, "thk_decoder_set_peers: null argument");
    THK_CHECK_ARG(n == d->p.tp_size && n > 1, "thk_decoder_set_peers: expected %d peers, got %d", d->p.tp_size, n);
    for (int r = 0; r < n; ++r) {
        if (r == d->p.tp_rank) continue;
        THK_CHECK_ARG(peer_bufs[r] != nullptr, "thk_decoder_set_peers: peer %d is null", r);
        d->p.xchg[r] = (unsigned char*)peer_bufs[r];
    }
    d->peers_set = true;
    return THK_OK;
}
