// decoder.cu -- th_eval_gpu for n_tokens == 1 as ONE persistent sm_100a kernel.
//
// Replaces the ~900 WebGPU commands the reference encodes per token (build_layer_cmdbuf x n_layer
// + build_final_compute_cmdbuf, th-llama.cpp:270-452, 240-268, 592-640) with a single launch of
// one CTA per SM.  Design (DESIGN.md has the full write-up and the measurements behind it):
//
//   * 12 warps in three warpgroups.  Warps 0-7 = MATH (216 registers each, setmaxnreg.inc); warp 8 = PRODUCER, warp 9 =
//     EPILOGUE, warp 10 = exchange REDUCER (tensor parallel only), warp 11 idle (72 registers each, setmaxnreg.dec).
//     Nothing is out of line: one __noinline__ call caps every thread at the ABI's register budget (measured).
//   * the PRODUCER walks the CTA's static list of weight / KV tiles for the whole token (all layers, all phases) and
//     streams them HBM -> shared memory with cp.async.bulk (UBLKCP) into a ring of four 32 KB slots guarded by full/empty
//     mbarriers, L2 evict-first.  It never waits for activations, so HBM stays busy across phase boundaries; when a phase
//     boundary stalls the ring it asks L2 for the CTA's next rows (UBLKPF).
//   * a tile is 8 rows x <= 2048 columns of f16.  Math warp c owns the 8 rows of the 256-column chunk c: eight 128-bit
//     shared loads by shared-window address, exact f16 -> f32 conversion (HADD2.F32), packed FFMA2 against the activation
//     values of its 8 columns -- held in REGISTERS for the whole phase when the phase has <= 2 K tiles (every
//     4096-column matrix), else re-read from shared memory.  The row sums of a finished row group are reduced with a
//     transposing shuffle tree and handed to the epilogue warp through an 8-deep ring of records (mbarriers, ONE arrival
//     per warp after __syncwarp: lanes diverge at per-lane try_wait predicates, and several lanes arriving on one
//     mbarrier in the same instruction are not counted per lane).
//   * the EPILOGUE warp adds the eight warp sums per row in a fixed order and runs the fused epilogue (RMS scale, RoPE +
//     KV append, residual add, SiLU*mul, logits + argmax); it owns the grid barrier (one red.release + relaxed polling).
//   * phases per layer: QKV | attention (split-KV, warp-private online softmax) | Wo | W1,W3 | W2; then logits.
//     RMSNorm*gain is the prologue of the consuming phase; the split-KV combine is the prologue of Wo.  A grid barrier
//     separates QKV | attention | Wo; Wo -> W1,W3 -> W2 -> next QKV synchronise through epoch-stamped copies of the
//     vectors themselves (one 64-bit store per element, the consumer polls what it needs).
//   * tensor parallel (decode_kernel<true>): the Wo / W2 epilogues push their partial rows as epoch-stamped values into
//     every rank's exchange region over NVLink (one hop); on every rank the REDUCER warp of CTA b owns 1/grid of the
//     residual stream in registers, adds the tp partials in rank order (the read is the wait) and publishes the reduced
//     epoch-stamped vector locally -- so every math warp polls ONE 32 KB vector exactly as on a single GPU.
//
// Arithmetic follows oracle/th_oracle.c (the restatement of the WGSL); only summation order
// differs.  No tensor cores: at M=1 the work is 1 FLOP/byte and HBM-bound.  What bounds the kernel today (the latency
// chain of the 161 phase hand-offs, not the math) and everything that was tried: profiles/r2_timeline_and_experiments.md.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kSlotBytes = 32 * 1024;
#ifdef THK_SLOTS
constexpr int kNumSlots = THK_SLOTS;
#else
constexpr int kNumSlots = 4;                          // measured (r2, v7): 4 slots 2.633, 5 slots 2.638 ms/token -- equal; 4 leaves 40 KB of L1
#endif
constexpr int kChunks = 8;                            // 256-column chunks per tile
constexpr int kMathWarps = 8;                         // one per chunk
constexpr int kMathThreads = kMathWarps * 32;
constexpr int kSvcWarps = 4;                          // producer, epilogue, reducer (tensor parallel), one idle warp: 12 warps = 384 threads
constexpr int kMathBase = 0;                          // math warps 0-7; the service warps take the highest warp ids
constexpr int kSvcBase = kMathThreads;
constexpr int kThreads = kMathThreads + kSvcWarps * 32;   // 384 threads = 3 warpgroups: two of math warps, one of service warps
// A 12-warp CTA launches with 168 registers per thread; the service warpgroup hands registers to the two math warpgroups
// (setmaxnreg: 72 + 2 * 216 <= 512 per sub-partition lane).  Any ABI call in the kernel makes ptxas ignore the per-branch
// budgets, so the kernel has none (the timeline profiler is only compiled into lib_prof).
constexpr int kServiceRegs = 72, kMathRegs = 216;
constexpr int kRows = 8;                              // rows per row group (= per tile)
constexpr int kMaxTilePos = 128;                      // attention: positions per tile cap
constexpr int kMaxHeadDim = 128;
#ifndef THK_ATT_PER_ROUND
#define THK_ATT_PER_ROUND 4
#endif
constexpr int kAttPerRound = THK_ATT_PER_ROUND;       // attention: positions a warp scores per round (8 would halve the rounds of a 64-position tile but spills in the math warps: ptxas, not measured)
constexpr int kMaxSplit = 8;                          // attention: KV splits per head cap
#ifndef THK_COPY_CHUNKS
#define THK_COPY_CHUNKS 3
#endif
constexpr int kDumpBufs = 8;                          // row-group hand-off ring between the math warps and the epilogue warp
constexpr int kMaxOwn = 4;                            // tensor parallel: residual elements per reducer lane (n_embd <= 128 * grid)
enum NamedBarrier { BAR_ALL = 1, BAR_MATH = 2, BAR_PRE = 3 };

// Host-computed tile schedule of one matvec phase
struct PhaseDesc {
    int C;            // columns (input dimension)
    int KT, CT;       // K tiles per row group, columns per K tile (multiple of 256, <= 2048)
    int nseg, paired; // segments (weight matrices) and whether seg0/seg1 rows are processed pairwise
    int rows[3];      // rows per segment
};
enum PhaseKind { PH_QKV = 0, PH_WO = 1, PH_W13 = 2, PH_W2 = 3, PH_OUT = 4 };
enum StepKind { K_QKV = 0, K_ATT = 1, K_WO = 2, K_W13 = 3, K_W2 = 4, K_OUT = 5 };

struct DecParams {
    int n_vocab, n_embd, n_head, n_layer, n_ff, n_ctx, head_dim;
    int Eh, Fh, Hl, Vl;                 // local (tensor-parallel shard) sizes
    int tp_rank, tp_size;
    const thk_llama_layer* layers;      // device array [n_layer]
    const uint16_t* emb;
    const float* norm;
    const uint16_t* out_w;
    float *x, *q, *part;                // x: residual stream after the last layer (thk_decoder_hidden)
    // Epoch-stamped vectors: element = (epoch << 32) | f32 bits, written with ONE 64-bit store, so a reader that sees the
    // expected epoch sees the value.  xf: residual stream entering a layer (epoch flag_epoch + l), h1f: after attention
    // (flag_epoch + l + 1), fff: FFN hidden (flag_epoch + l + 1).
    unsigned long long *h1f, *fff, *xf;
    // att_flagged: q and the token's own K / V row also travel as epoch-stamped values, and the attention phase polls the
    // 3 x head_dim values of its head instead of waiting for the grid barrier behind the QKV phase
    unsigned long long *qf, *knf, *vnf;
    int att_flagged;
    unsigned flag_epoch;
    int nosync;                         // DEBUG (wrong results): skip every cross-CTA wait, to time the pipeline without synchronisation
    int poll_single;                    // flagged reads: after a stale read spin on the one stale element before re-reading all
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
    int kv_f16;                         // KV cache element type of the fused layout: 0 = f32 (reference), 1 = f16 (rounded at append)
    int att_max_split;
    unsigned long long timeout_ns;
    // tensor parallel exchange (tp_size > 1): every rank owns one region laid out as
    //   u64 xbf[2][tp][n_embd] (epoch-stamped partial vectors of the Wo / W2 exchange, one slot per source rank)
    //   | float amax_val[tp] | int amax_idx[tp] | unsigned flags[tp] | unsigned flags2[tp]
    // and writes its partial vectors / argmax candidates straight into every peer's region over NVLink.
    unsigned char* xchg[8];             // region base per rank (peer-mapped pointers; [tp_rank] is local)
    unsigned epoch_base;                // argmax exchange flags are monotonic across launches
    unsigned l2_ahead;                  // bytes of this CTA's rows the producer asks L2 for when a phase boundary stalls the ring (0: off)
    int prof_phase;                     // timeline: phase whose per-tile consume / issue times are recorded (kProfTiles each)
    unsigned long long* prof;           // optional timeline: [cta][phase<256][8] u64 (see ProfSlot), then [cta][4] producer stats
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
// weights and KV are read once per token: evict-first in L2, so the stream does not push the small hot
// data (activation vectors, gains, split results, barrier counters, code) out of the 126 MB L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
// shared-memory access by 32-bit shared-window address (no generic-pointer conversion on the hot path)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32f(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts128f(uint32_t a, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_flagged(unsigned long long* p, float v, unsigned epoch) {
    const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void st_flagged_sys(unsigned long long* p, float v, unsigned epoch) {     // peer memory over NVLink
    const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_flagged1_sys(const unsigned long long* p) {
    unsigned long long a;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ unsigned long long ld_flagged1(const unsigned long long* p) {
    unsigned long long a;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ void ld_flagged2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void st_release_sys(unsigned* ptr, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* ptr) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}

// exchange region accessors
__device__ __forceinline__ size_t xchg_tail(const DecParams& p) { return (size_t)2 * p.tp_size * p.n_embd * sizeof(unsigned long long); }
// flagged partial vector `which` (0: Wo, 1: W2) of source rank `src` inside rank `rank`'s exchange region
__device__ __forceinline__ unsigned long long* xbf_ptr(const DecParams& p, int rank, int which, int src) {
    return (unsigned long long*)p.xchg[rank] + ((size_t)which * p.tp_size + src) * p.n_embd;
}
__device__ __forceinline__ float* xamax_val(const DecParams& p, int rank) { return (float*)(p.xchg[rank] + xchg_tail(p)); }
__device__ __forceinline__ int* xamax_idx(const DecParams& p, int rank) { return (int*)(p.xchg[rank] + xchg_tail(p)) + p.tp_size; }
__device__ __forceinline__ unsigned* xflags(const DecParams& p, int rank, int set) {
    return (unsigned*)(p.xchg[rank] + xchg_tail(p)) + (2 + set) * p.tp_size;
}

// The in-kernel timeline is compiled in only with -DTHK_PROFILE (token_hawk_b200/lib_prof): its marks are out-of-line
// calls (an inlined %globaltimer read is hoisted above barriers by ptxas; a call is a scheduling fence).
#ifdef THK_PROFILE
#define PROF(p) ((p).prof)
#else
#define PROF(p) ((unsigned long long*)nullptr)
#endif
constexpr int kProfPhases = 256;
constexpr int kProfTiles = 64;   // after the producer stats: [cta][kProfTiles] tile-retired times, then [cta][kProfTiles] tile-issued times
enum ProfSlot { PROF_START = 0, PROF_PROLOGUE = 1, PROF_FIRST_TILE = 2, PROF_LAST_TILE = 3, PROF_ARRIVE = 4, PROF_PROD_LAST = 5, PROF_WAIT_FULL = 6, PROF_PROD_FIRST = 7 };
__device__ __noinline__ void prof_tile(unsigned long long* prof, int which, int tile) {        // which: 0 retired, 1 issued
    prof[(size_t)gridDim.x * (kProfPhases * 8 + 4 + which * kProfTiles) + (size_t)blockIdx.x * kProfTiles + tile] = gtimer();
}
__device__ __noinline__ void prof_mark(unsigned long long* prof, unsigned phase, int slot) {   // out of line: ~25 call sites
    if (phase < (unsigned)kProfPhases) prof[((size_t)blockIdx.x * kProfPhases + phase) * 8 + slot] = gtimer();
}
// math warps: `pm` is non-null only in the one thread that records
__device__ __forceinline__ void mark(unsigned long long* pm, const DecParams& p, unsigned phase, int slot) {
    if (pm != nullptr) prof_mark(PROF(p), phase, slot);
}

__device__ __forceinline__ void raise_abort(const DecParams& p, unsigned code, unsigned a, unsigned b) {
    if (atomicCAS(p.status, 0u, code) == 0u) { p.status[1] = a; p.status[2] = b; p.status[3] = blockIdx.x; }
}

// Slow path of every mbarrier wait: bounded by %globaltimer.  Returns false when the watchdog fired or another CTA
// aborted; the caller then walks the rest of its (static) schedule without waiting and without issuing copies, so the
// kernel terminates and the host sees THK_E_TIMEOUT instead of a wedged GPU.  Inlined: a CALL inside the tile loop forces
// everything that is live across it (accumulators, activations) through the few callee-saved registers of the ABI --
// ptxas then keeps the accumulators in local memory.
__device__ __forceinline__ bool mbar_wait_slow(unsigned* status, unsigned long long timeout_ns, uint32_t bar, uint32_t parity, unsigned tag) {
    if (ld_volatile_u32(status) != 0) return false;      // an abort is sticky: after it every wait gives up at once
    const unsigned long long t0 = gtimer();
    unsigned it = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++it & 255u) == 0u) {
            if (ld_volatile_u32(status) != 0) return false;
            if (gtimer() - t0 > timeout_ns) {
                if (atomicCAS(status, 0u, 0x100u | tag) == 0u) { status[1] = bar; status[2] = parity; status[3] = blockIdx.x; }
                return false;
            }
        }
    }
    return true;
}
// watchdog step of the global-memory polls; returns true when the poller must give up
__device__ __forceinline__ bool poll_watchdog(unsigned* status, unsigned long long timeout_ns, unsigned long long& t0, unsigned code, unsigned a, unsigned b) {
    if (ld_volatile_u32(status) != 0) return true;
    if (t0 == 0) { t0 = gtimer(); return false; }
    if (gtimer() - t0 > timeout_ns) {
        if (atomicCAS(status, 0u, code) == 0u) { status[1] = a; status[2] = b; status[3] = blockIdx.x; }
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------
// shared memory carve-up: | ring slots | SmemMisc | hand-off ring + attention scratch | xs (activation vector) |
// ------------------------------------------------------------------------------------------
struct SmemMisc {
    unsigned long long full[kNumSlots];
    unsigned long long empty[kNumSlots];
    unsigned long long red_full[kDumpBufs];   // row-group sums `buf` written by all math warps
    unsigned long long red_free[kDumpBufs];   // ... and consumed by the epilogue warp
    float2 rope[kMaxHeadDim / 2];       // (cos, sin) of n_past * theta_i for this token
    int range[5][2];                    // this CTA's row range [r, r_end) of every matvec phase (computed once per launch)
    int niter[5];                       // ... and its number of (row group, segment-of-a-pair) iterations: all a math warp needs to know
    unsigned long long wstat[kMathWarps + kSvcWarps][4];   // -DTHK_PROFILE: cycles a warp spent waiting, see WaitSlot
};
constexpr int kMiscBytes = 2048;
static_assert(sizeof(SmemMisc) <= kMiscBytes, "SmemMisc too large");
constexpr int kRecFloats = kRows + 1;                          // per warp: 8 row sums + its share of sum(v^2) of the phase input
constexpr int kRedFloats = kMathWarps * kRecFloats;            // one row-group hand-off record
constexpr int kAttScratchOff = kDumpBufs * kRedFloats;         // attention scratch (floats) behind the hand-off ring
constexpr int kRedBytes = 8 * 1024;                            // ring (2.3 KB) + attention scratch (<= 5.1 KB)
static_assert(kAttScratchOff * 4 + (kMathWarps * kMaxHeadDim + 2 * kMathWarps + 2 * kMaxHeadDim) * 4 <= kRedBytes, "attention scratch does not fit");
constexpr int kXsOffset = kNumSlots * kSlotBytes + kMiscBytes + kRedBytes;

struct Smem {
    unsigned char* slots;
    SmemMisc* misc;
    float* red;       // [kDumpBufs][kMathWarps][kRecFloats] hand-off ring, then the attention scratch
    float* xs;
    uint32_t slots_a, full_a, empty_a, red_full_a, red_free_a, red_a, xs_a;   // shared-window addresses
};
__device__ __forceinline__ Smem carve(unsigned char* base) {
    Smem s;
    s.slots = base;
    s.misc = (SmemMisc*)(base + kNumSlots * kSlotBytes);
    s.red = (float*)(base + kNumSlots * kSlotBytes + kMiscBytes);
    s.xs = (float*)(base + kXsOffset);
    // The shared-window address of the base is read ONCE and made opaque: otherwise every use re-derives it from the
    // generic pointer (S2UR SR_CgaCtaId + ULEA, ~35 cycles of latency each -- three times per tile in the hot loop).
    uint32_t a = smem_u32(base);
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    s.slots_a = a;
    s.full_a = a + kNumSlots * kSlotBytes + (uint32_t)offsetof(SmemMisc, full);
    s.empty_a = a + kNumSlots * kSlotBytes + (uint32_t)offsetof(SmemMisc, empty);
    s.red_full_a = a + kNumSlots * kSlotBytes + (uint32_t)offsetof(SmemMisc, red_full);
    s.red_free_a = a + kNumSlots * kSlotBytes + (uint32_t)offsetof(SmemMisc, red_free);
    s.red_a = a + kNumSlots * kSlotBytes + kMiscBytes;
    s.xs_a = a + kXsOffset;
    return s;
}
__device__ __forceinline__ Smem smem_view() {
    extern __shared__ __align__(1024) unsigned char smem_base[];
    return carve(smem_base);
}

// -DTHK_PROFILE: who waits for whom.  Every warp accumulates (lane 0, in shared memory) the cycles it spent in the slow
// path of its waits and copies its four counters to the timeline buffer when it finishes.
enum WaitSlot { WS_RING = 0,     // math: ring slot not full yet; producer: no empty slot; epilogue: no row group handed over yet
                WS_HANDOFF = 1,  // math: hand-off record not consumed yet; epilogue: residual operand (flagged vector)
                WS_POLL = 2,     // math: prologue (flagged vector / staging); epilogue: grid barrier
                WS_TOTAL = 3 };  // lifetime of the warp
constexpr int kProfWarps = kMathWarps + kSvcWarps;
#ifdef THK_PROFILE
__device__ __forceinline__ long long wstat_t0() { return clock64(); }
__device__ __forceinline__ void wstat_add(int slot, long long t0) {
    if ((threadIdx.x & 31) == 0) smem_view().misc->wstat[threadIdx.x >> 5][slot] += (unsigned long long)(clock64() - t0);
}
__device__ __noinline__ void wstat_flush(unsigned long long* prof, long long t_start) {
    if ((threadIdx.x & 31) != 0 || prof == nullptr) return;
    SmemMisc* misc = smem_view().misc;
    const int w = threadIdx.x >> 5;
    misc->wstat[w][WS_TOTAL] = (unsigned long long)(clock64() - t_start);
    unsigned long long* dst = prof + (size_t)gridDim.x * (kProfPhases * 8 + 4 + 2 * kProfTiles) + ((size_t)blockIdx.x * kProfWarps + w) * 4;
    for (int i = 0; i < 4; ++i) dst[i] = misc->wstat[w][i];
}
#else
__device__ __forceinline__ long long wstat_t0() { return 0; }
__device__ __forceinline__ void wstat_add(int, long long) {}
__device__ __forceinline__ void wstat_flush(unsigned long long*, long long) {}
#endif

// activation layout in xs: 256-col chunk c, lane l owns cols 8l..8l+7; its first float4 (cols 8l..8l+3) sits at
// (c*64 + l)*16 B and its second at (c*64 + 32 + l)*16 B -- consecutive lanes hit consecutive banks (conflict free).
// A chunk's values are staged by the two warps that own the chunk (K tiles alternate between them) and read by both.

// ring position shared by producer and consumers (kept incrementally: no modulo per tile)
struct Ring {
    uint32_t sl, par;           // slot, phase parity of its barriers
    __device__ __forceinline__ void advance() { if (++sl == kNumSlots) { sl = 0; par ^= 1u; } }
};

// per-thread consumer state
struct Cons {
    Ring ring;
    bool dead;            // watchdog fired somewhere: stop waiting, keep walking the schedule
};
// Wait for the tile in the current ring slot (the math warps consume the tiles in order, so a parity wait cannot alias).
__device__ __forceinline__ void wait_full(const DecParams& p, const Smem& S, Cons& c, unsigned tag) {
    if (c.dead) return;
    const uint32_t bar = S.full_a + c.ring.sl * 8;
    if (!mbar_try_wait(bar, c.ring.par)) {
        const long long t0 = wstat_t0();
        c.dead = !mbar_wait_slow(p.status, p.timeout_ns, bar, c.ring.par, tag);
        wstat_add(WS_RING, t0);
    }
}
__device__ __forceinline__ void release_slot(const Smem& S, Cons& c, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(S.empty_a + c.ring.sl * 8);
    c.ring.advance();
}
// ------------------------------------------------------------------------------------------
// this CTA's share of a phase: contiguous (even-aligned) row range of the concatenated segments,
// walked in groups of up to kRows rows; the last group of a range (or of a segment) may be short.
// ------------------------------------------------------------------------------------------
struct RowIt {
    int r, r_end;              // current / end row in the concatenated row space (paired: rows of segment 0)
    int si, row0, nrows;       // segment, first row within it, rows in this group
    int e0, e1;                // end of segment 0 / segment 1 in the concatenated row space (registers: an indexed
                               // constant-bank read per segment test costs ~60 cycles of latency at every row-group end)
    // even-aligned contiguous share of the phase's rows; the two divisions cost ~0.3 us of dependent latency, so the
    // kernel evaluates this once per launch and phase kind (SmemMisc::range) instead of at every phase start
    static __host__ __device__ __forceinline__ void share(const PhaseDesc& d, unsigned n, unsigned b, int& r0, int& r1) {
        const unsigned total = (unsigned)(d.paired ? d.rows[0] : d.rows[0] + d.rows[1] + d.rows[2]);
        r0 = (int)(((total * b) / n) & ~1u);
        r1 = (b + 1 == n) ? (int)total : (int)(((total * (b + 1)) / n) & ~1u);
    }
    __host__ __device__ __forceinline__ void init_range(int r0, int r1, const PhaseDesc& d) {
        r = r0;
        r_end = r1;
        e0 = d.paired ? 0x7fffffff : d.rows[0];
        e1 = d.paired ? 0x7fffffff : d.rows[0] + d.rows[1];
        place();
    }
    __device__ __forceinline__ void init(const SmemMisc* misc, int ph, const PhaseDesc& d) { init_range(misc->range[ph][0], misc->range[ph][1], d); }
    __host__ __device__ __forceinline__ bool valid() const { return r < r_end; }
    __host__ __device__ __forceinline__ void place() {
        // a group never crosses a segment or the end of this CTA's share
        si = r < e0 ? 0 : (r < e1 ? 1 : 2);
        const int seg_begin = r < e0 ? 0 : (r < e1 ? e0 : e1);
        const int seg_end = r < e0 ? e0 : (r < e1 ? e1 : 0x7fffffff);
        row0 = r - seg_begin;
        const int a = seg_end - r < kRows ? seg_end - r : kRows, b = r_end - r;
        nrows = a < b ? a : b;
    }
    __host__ __device__ __forceinline__ void next() { r += nrows; place(); }
};

// attention work split (shared by producer and consumers)
struct AttSched { int S, N; };
__device__ __forceinline__ AttSched make_att(const DecParams& p) {
    AttSched a;
    a.N = p.n_past + 1;
    int S = (int)gridDim.x / p.Hl;
    const int by_len = (a.N + 15) / 16;          // at least ~16 positions per split
    if (S > by_len) S = by_len;
    if (S > p.att_max_split) S = p.att_max_split;
    if (S > a.N) S = a.N;
    if (S < 1) S = 1;
    a.S = S;
    return a;
}
__device__ __forceinline__ int part_stride(const DecParams& p) { return p.head_dim + 4; }   // {m, l, -, -, o[D]}

// ------------------------------------------------------------------------------------------
// PRODUCER
// ------------------------------------------------------------------------------------------
struct Prod {
    Ring ring;
    bool dead;
    unsigned tiles;
    long long wait_cyc;
    uint64_t pol;
};
__device__ __forceinline__ void wait_empty(const DecParams& p, const Smem& S, Prod& c, unsigned tag) {
    if (c.dead) return;
    const uint32_t bar = S.empty_a + c.ring.sl * 8;
    if (!mbar_try_wait(bar, c.ring.par ^ 1u)) {
        const long long t0 = PROF(p) ? clock64() : 0;
        c.dead = !mbar_wait_slow(p.status, p.timeout_ns, bar, c.ring.par ^ 1u, tag);
        if (PROF(p)) c.wait_cyc += clock64() - t0;
        wstat_add(WS_RING, t0);
    }
}

// bulk L2 prefetch of [base, base + bytes) spread over the lanes (SASS UBLKPF); bytes % 16 == 0
__device__ __forceinline__ void l2_prefetch_span(const unsigned char* base, uint32_t bytes, int lane) {
    constexpr uint32_t kPiece = 8192u;
#pragma unroll 1
    for (uint32_t off = (uint32_t)lane * kPiece; off < bytes; off += 32u * kPiece)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(min(kPiece, bytes - off)) : "memory");
}

__device__ __forceinline__ void produce_mat_phase(const DecParams& p, const Smem& S, Prod& c, int ph, const uint16_t* w0p,
                                               const uint16_t* w1p, const uint16_t* w2p, unsigned phase_idx) {
    const PhaseDesc& d = p.ph[ph];
    const int lane = threadIdx.x & 31;
    const int C = d.C, KT = d.KT, CT = d.CT, nsub = d.paired ? 2 : 1;
    bool first = true;
    int j = 0;                                  // tiles of this phase issued so far
    RowIt it;
    for (it.init(S.misc, ph, d); it.valid(); it.next()) {
        for (int sub = 0; sub < nsub; ++sub) {
            const int si = d.paired ? sub : it.si;
            const uint16_t* wrow = (si == 0 ? w0p : si == 1 ? w1p : w2p) + (size_t)it.row0 * C;
            for (int kt = 0; kt < KT; ++kt, ++j) {
                const int col0 = kt * CT;
                const int ncols = min(CT, C - col0);
                if (j == kNumSlots && p.l2_ahead != 0u && !c.dead && !mbar_try_wait(S.empty_a + c.ring.sl * 8, c.ring.par ^ 1u)) {
                    // The ring now holds nothing but tiles of this phase and its first tile has not been consumed: the
                    // consumers are still in the previous phase's tail / the grid barrier / the prologue, and the ring
                    // (128 KB, ~2.4 us of this SM's HBM share) cannot cover that stall.  Ask L2 for this CTA's next rows
                    // so HBM keeps streaming; after the stall the ring refills from L2 faster than HBM could feed it.
                    const uint32_t row_bytes_full = (uint32_t)C * 2u;
                    if (d.paired) {
                        const uint32_t bytes = min((uint32_t)(it.r_end - it.r) * row_bytes_full, p.l2_ahead >> 1);
                        l2_prefetch_span((const unsigned char*)(w0p + (size_t)it.row0 * C), bytes, lane);
                        l2_prefetch_span((const unsigned char*)(w1p + (size_t)it.row0 * C), bytes, lane);
                    } else {
                        const uint32_t rows_left = (uint32_t)min(d.rows[si] - it.row0, it.r_end - it.r);
                        l2_prefetch_span((const unsigned char*)wrow, min(rows_left * row_bytes_full, p.l2_ahead), lane);
                    }
                }
                wait_empty(p, S, c, 1);
                if (PROF(p) && first && lane == 0) { prof_mark(PROF(p), phase_idx, PROF_PROD_FIRST); first = false; }
                if (!c.dead) {       // after an abort nothing is issued any more: slots may still be in use and tx counts must not pile up
                    const uint32_t fb = S.full_a + c.ring.sl * 8, dst = S.slots_a + c.ring.sl * kSlotBytes;
                    const uint32_t row_bytes = (uint32_t)ncols * 2u;
                    if (lane == 0) mbar_expect_tx(fb, (uint32_t)it.nrows * row_bytes);
                    __syncwarp();
                    if (lane < it.nrows) bulk_g2s(dst + (uint32_t)lane * row_bytes, wrow + (size_t)lane * C + col0, row_bytes, fb, c.pol);
                }
                if (PROF(p) && (int)phase_idx == p.prof_phase && j < kProfTiles && lane == 0) prof_tile(PROF(p), 1, j);
                c.ring.advance();
                ++c.tiles;
            }
        }
    }
    if (PROF(p) && lane == 0) prof_mark(PROF(p), phase_idx, PROF_PROD_LAST);
}

__device__ __forceinline__ void produce_att_phase(const DecParams& p, const Smem& S, Prod& c, const thk_llama_layer& L, unsigned phase_idx) {
    const int lane = threadIdx.x & 31;
    const AttSched a = make_att(p);
    const int D = p.head_dim;
    bool first = true;
    for (int u = blockIdx.x; u < p.Hl * a.S; u += gridDim.x) {
        const int h = u / a.S, sp = u % a.S;
        const int pa = (int)(((long long)a.N * sp) / a.S);
        int pb = (int)(((long long)a.N * (sp + 1)) / a.S);
        if (pb > p.n_past) pb = p.n_past;            // position n_past is read from global by the consumers
        for (int pos = pa; pos < pb; pos += p.att_tpos) {
            const int np = min(p.att_tpos, pb - pos);
            const uint32_t esz = p.kv_f16 ? 2u : 4u;
            const uint32_t bytes = (uint32_t)np * D * esz;
            for (int kv = 0; kv < 2; ++kv) {
                wait_empty(p, S, c, 2);
                if (PROF(p) && first && lane == 0) { prof_mark(PROF(p), phase_idx, PROF_PROD_FIRST); first = false; }
                const unsigned char* base = (const unsigned char*)(kv == 0 ? L.key_cache : L.value_cache) + ((size_t)h * p.n_ctx + pos) * D * esz;
                const uint32_t fb = S.full_a + c.ring.sl * 8;
                if (lane == 0 && !c.dead) { mbar_expect_tx(fb, bytes); bulk_g2s(S.slots_a + c.ring.sl * kSlotBytes, base, bytes, fb, c.pol); }
                __syncwarp();
                c.ring.advance();
                ++c.tiles;
            }
        }
    }
    if (PROF(p) && lane == 0) prof_mark(PROF(p), phase_idx, PROF_PROD_LAST);
}

__device__ __forceinline__ int phase_of(int k) { return k == K_QKV ? PH_QKV : k == K_WO ? PH_WO : k == K_W13 ? PH_W13 : k == K_W2 ? PH_W2 : PH_OUT; }

__device__ void producer_main(const DecParams& p, const Smem& S) {
    const long long t0 = clock64();
    Prod c{{0u, 0u}, false, 0u, 0, l2_evict_first_policy()};
    const int nsteps = 5 * p.n_layer + 1;
    int l = 0, k = K_QKV;                                 // same step order as the consumers
    for (int i = 0; i < nsteps; ++i) {
        const thk_llama_layer* L = p.layers + (l < p.n_layer ? l : p.n_layer - 1);
        if (k == K_ATT) {
            produce_att_phase(p, S, c, *L, (unsigned)i);
        } else {
            const uint16_t* w0 = k == K_QKV ? L->wq : k == K_WO ? L->wo : k == K_W13 ? L->w1 : k == K_W2 ? L->w2 : p.out_w;
            const uint16_t* w1 = k == K_QKV ? L->wk : k == K_W13 ? L->w3 : nullptr;
            const uint16_t* w2 = k == K_QKV ? L->wv : nullptr;
            produce_mat_phase(p, S, c, phase_of(k), w0, w1, w2, (unsigned)i);
        }
        if (k == K_W2) { ++l; k = (l < p.n_layer) ? K_QKV : K_OUT; } else ++k;
    }
    if (PROF(p) && (threadIdx.x & 31) == 0) {
        unsigned long long* stt = PROF(p) + (size_t)gridDim.x * kProfPhases * 8 + (size_t)blockIdx.x * 4;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        stt[0] = (unsigned long long)c.wait_cyc; stt[1] = (unsigned long long)(clock64() - t0); stt[2] = c.tiles; stt[3] = smid;
    }
    wstat_flush(PROF(p), t0);
}

// ------------------------------------------------------------------------------------------
// MATH WARPS
// ------------------------------------------------------------------------------------------
// 8 weights (one uint4 of f16) times 8 activations into an accumulator PAIR, with the packed f32x2 FMA of sm_100
// (FFMA2: two independent round-to-nearest FMAs per instruction -- bit-identical to two fmaf, half the issue slots).
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
__device__ __forceinline__ unsigned long long cvt2(uint32_t h2) {     // exact f16x2 -> f32x2 (== th.cpp:312-333)
    const float2 f = h2_to_f2(h2);
    return pack2(f.x, f.y);
}
__device__ __forceinline__ void fma8(const uint4& w, const unsigned long long* x, unsigned long long& acc) {
    acc = ffma2(x[0], cvt2(w.x), acc);
    acc = ffma2(x[1], cvt2(w.y), acc);
    acc = ffma2(x[2], cvt2(w.z), acc);
    acc = ffma2(x[3], cvt2(w.w), acc);
}

// Hand the eight row sums of a finished row group to the epilogue warp.  A transposing shuffle reduction (4 + 2 + 1
// exchanges halve the rows a lane holds while doubling the lanes summed, then 2 plain steps) leaves the warp total of row
// (lane >> 2) & 7 in every lane; 8 floats per warp go into a ring of kDumpBufs hand-off records, so the math warps can run
// kDumpBufs row groups ahead of the epilogue warp.  Fixed order of additions: deterministic.
__device__ __forceinline__ void flush_group(const DecParams& p, const Smem& S, Cons& c, unsigned& gq, float (&v)[kRows], float ss,
                                            uint32_t dump_a, int lane) {
    const bool u4 = (lane & 16) != 0, u3 = (lane & 8) != 0, u2 = (lane & 4) != 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float recv = __shfl_xor_sync(0xffffffffu, u4 ? v[r] : v[r + 4], 16);
        v[r] = (u4 ? v[r + 4] : v[r]) + recv;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const float recv = __shfl_xor_sync(0xffffffffu, u3 ? v[r] : v[r + 2], 8);
        v[r] = (u3 ? v[r + 2] : v[r]) + recv;
    }
    const float recv = __shfl_xor_sync(0xffffffffu, u2 ? v[0] : v[1], 4);
    v[0] = (u2 ? v[1] : v[0]) + recv;
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    const unsigned buf = gq & (kDumpBufs - 1), use = gq / kDumpBufs;
    if (use > 0 && !c.dead) {
        const uint32_t fb = S.red_free_a + buf * 8;
        if (!mbar_try_wait(fb, (use - 1) & 1u)) {
            const long long t0 = wstat_t0();
            c.dead = !mbar_wait_slow(p.status, p.timeout_ns, fb, (use - 1) & 1u, 6);
            wstat_add(WS_HANDOFF, t0);
        }
    }
    if ((lane & 3) == 0) sts32f(dump_a + (buf * kRedFloats + (unsigned)((lane >> 2) & 7)) * 4u, v[0]);
    if (lane == 1) sts32f(dump_a + (buf * kRedFloats + kRows) * 4u, ss);          // this warp's share of sum(v^2) (norm phases)
    __syncwarp();
    if (lane == 0) mbar_arrive(S.red_full_a + buf * 8);       // ONE arrival per warp: several lanes arriving on one mbarrier in
                                                              // the same instruction are not counted per lane (measured: hangs / faults)
    ++gq;
}

// consumer state carried from phase to phase: row groups handed over so far and the ring position
__device__ __forceinline__ unsigned long long pack_state(unsigned gq, const Ring& r) {
    return ((unsigned long long)(r.sl | (r.par << 8)) << 32) | gq;
}
__device__ __forceinline__ void unpack_state(unsigned long long st, unsigned& gq, Ring& r) {
    gq = (unsigned)st;
    r.sl = (unsigned)(st >> 32) & 0xffu;
    r.par = (unsigned)(st >> 40) & 1u;
}

// one tile's weights of this lane: 8 rows x 8 columns of f16
struct TileW { uint4 w[kRows]; };
__device__ __forceinline__ void tile_ldw(uint32_t wa, uint32_t stride, TileW& t) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) t.w[r] = lds128(wa + (uint32_t)r * stride);
}
__device__ __forceinline__ void xs_ld(uint32_t xa, bool ok, unsigned long long (&xp)[4]) {
    float4 x0 = lds128f(xa), x1 = lds128f(xa + 512u);
    if (!ok) { x0 = make_float4(0.f, 0.f, 0.f, 0.f); x1 = x0; }
    xp[0] = pack2(x0.x, x0.y); xp[1] = pack2(x0.z, x0.w); xp[2] = pack2(x1.x, x1.y); xp[3] = pack2(x1.z, x1.w);
}

// Stream this CTA's tiles of one matvec phase.  Per tile a warp owns one 256-column chunk of all 8 rows (rows past a
// short group's end hold stale bytes; their sums are never read): eight 128-bit shared loads by shared-window address,
// exact f16 -> f32 conversion, packed FFMA2 against its lane's 8 activations of that chunk.
//   * The serial chain of a tile -- wait, shared loads, convert, FMA, release -- is what bounds a CTA's tile rate, not any
//     pipe (ncu: issue slots 34 % busy, every pipe < 25 %), so everything that is not math is kept off it.
//   * XREG: the phase has <= 2 K tiles (every 4096-column matrix): this lane's activations (packed for FFMA2; zero for
//     columns past the end) stay in registers for the whole phase; else they are re-read from xs with the weights.
//   (A non-blocking look-ahead on the next tile's barrier -- mbarrier.test_wait issued under the math -- was measured and
//   dropped: 2.67 -> 2.64 ms/token WITHOUT it.)
// The warp hands its row sums to the epilogue warp once per row group (`gq` counts them since kernel start).
// (Everything on the math warps' path is inlined: an out-of-line phase function may only use the ABI's scratch registers
// freely, and ptxas then kept half of the accumulators in local memory; the math warpgroups get their registers from
// setmaxnreg instead.)
template <bool XREG>
__device__ __forceinline__ unsigned long long math_mat_phase(const DecParams& p, int ph, unsigned long long state, float ss, unsigned phase_idx) {
    const Smem S = smem_view();
    const int ct = (int)threadIdx.x - kMathBase, cw = ct >> 5, lane = ct & 31;
    unsigned long long* const pm = (PROF(p) != nullptr && ct == 0) ? PROF(p) + (size_t)blockIdx.x * kProfPhases * 8 : nullptr;
    Cons c{{0u, 0u}, false};
    unsigned gq;
    unpack_state(state, gq, c.ring);
    const PhaseDesc& d = p.ph[ph];
    const int C = d.C, KT = d.KT, CT = d.CT;
    const int n_iters = S.misc->niter[ph];            // (row group, segment-of-a-pair) iterations of this CTA: the rows themselves do not matter here
    const int col = (cw << 8) + (lane << 3);                                  // this lane's first column inside a tile
    const uint32_t xlane_a = S.xs_a + (uint32_t)(((cw << 8) + (lane << 2)) << 2);   // + col0 * 4: first float4; second 512 B on
    const uint32_t dump_a = S.red_a + (uint32_t)(cw * kRecFloats * 4);            // this warp's record inside a hand-off buffer
    unsigned long long xr[2][4];
    if (XREG) {
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            const bool ok = kt < KT && col < CT && kt * CT + col < C;
            xs_ld(ok ? xlane_a + (uint32_t)((kt * CT) << 2) : S.xs_a + ((uint32_t)lane << 4), ok, xr[kt]);   // (never address past xs)
        }
    }
    mark(pm, p, phase_idx, PROF_PROLOGUE);
    bool first = pm != nullptr;
    int ntile = 0;
    mark(pm, p, phase_idx, PROF_WAIT_FULL);     // schedule computed, about to wait for the first tile
#pragma unroll 1
    for (int it = 0; it < n_iters; ++it) {
        unsigned long long acc[kRows];                    // (even-column sum, odd-column sum) per row
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0ull;
        if (XREG) {
            // <= 2 K tiles, activations in registers: straight-line code, no column bookkeeping
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                if (kt < KT) {
                    const int ncols = min(CT, C - kt * CT);
                    const uint32_t stride = (uint32_t)ncols * 2u;
                    const uint32_t wa = S.slots_a + c.ring.sl * kSlotBytes + (col < ncols ? (uint32_t)col << 1 : 0u);
                    wait_full(p, S, c, 3);
                    if (first) { mark(pm, p, phase_idx, PROF_FIRST_TILE); first = false; }
                    const uint32_t eb = S.empty_a + c.ring.sl * 8;
                    c.ring.advance();
#ifndef THK_EXP_NOMATH      // (measurement build, wrong results: tiles are waited for and released, not multiplied)
                    TileW t;
                    tile_ldw(wa, stride, t);
#pragma unroll
                    for (int r = 0; r < kRows; ++r) fma8(t.w[r], xr[kt], acc[r]);
#endif
                    __syncwarp();                                 // every lane has read the tile (lanes may have diverged at the waits)
                    if (lane == 0) mbar_arrive(eb);
                    if (pm != nullptr && (int)phase_idx == p.prof_phase && ntile < kProfTiles) prof_tile(PROF(p), 0, ntile++);
                }
            }
        } else {
            int col0 = 0;
#pragma unroll 1
            for (int kt = 0; kt < KT; ++kt, col0 += CT) {
                const int ncols = min(CT, C - col0);
                const bool ok = col < ncols;                          // lanes past a short tile's last column read column 0 against zeros
                const uint32_t stride = (uint32_t)ncols * 2u;
                const uint32_t wa = S.slots_a + c.ring.sl * kSlotBytes + (ok ? (uint32_t)col << 1 : 0u);
                wait_full(p, S, c, 3);
                if (first) { mark(pm, p, phase_idx, PROF_FIRST_TILE); first = false; }
                const uint32_t eb = S.empty_a + c.ring.sl * 8;
                c.ring.advance();
#ifndef THK_EXP_NOMATH
                TileW t;
                unsigned long long xp[4];
                xs_ld(ok ? xlane_a + ((uint32_t)col0 << 2) : S.xs_a + ((uint32_t)lane << 4), ok, xp);
                tile_ldw(wa, stride, t);
#pragma unroll
                for (int r = 0; r < kRows; ++r) fma8(t.w[r], xp, acc[r]);
#endif
                __syncwarp();
                if (lane == 0) mbar_arrive(eb);
                if (pm != nullptr && (int)phase_idx == p.prof_phase && ntile < kProfTiles) prof_tile(PROF(p), 0, ntile++);
            }
        }
        float v[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[r]));
            v[r] = lo + hi;
        }
        flush_group(p, S, c, gq, v, ss, dump_a, lane);
    }
    mark(pm, p, phase_idx, PROF_LAST_TILE);
    return pack_state(gq, c.ring);
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    float s = a.x * b.x;
    s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    return s;
}

// Flagged-vector read: this lane's 8 columns of NCH chunks (chunk u starts at element col[u]; col[u] < 0: no such chunk),
// all loads in flight together, repeated until every element carries `epoch` -- that IS the synchronisation with the
// CTAs that produce the vector (no grid barrier on these transitions).  After a stale read only the stale chunks are read
// again (tune poll_single: 0 = everything again, 1 = spin on the one stale element, then everything again).  Bounded by
// the watchdog.
template <int NCH>
__device__ __forceinline__ void load_flagged(const DecParams& p, const unsigned long long* vec, const int (&col)[NCH], unsigned epoch,
                                             unsigned long long (&e)[NCH * 8], bool& dead) {
    unsigned long long t0 = 0;
    unsigned it = 0;
    bool fresh[NCH];
#pragma unroll
    for (int u = 0; u < NCH; ++u) fresh[u] = col[u] < 0;
    while (true) {
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
            if (!fresh[u]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ld_flagged2(vec + col[u] + 2 * j, e[u * 8 + 2 * j], e[u * 8 + 2 * j + 1]);
            }
        }
        int stale = -1;
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
            if (!fresh[u]) {
                bool f = true;
#pragma unroll
                for (int j = 0; j < 8; ++j) if ((unsigned)(e[u * 8 + j] >> 32) != epoch) { f = false; if (stale < 0) stale = col[u] + j; }
                // poll_single == 2: a chunk that has arrived is kept; only the stale chunks are read again -- every re-read
                // costs a whole L2 round trip BEHIND the weight stream's queue (1.3 us + 0.4 us per ring slot in flight,
                // profiles/r2_timeline_and_experiments.md), so the last poll must also be the read that delivers the data
                fresh[u] = f && p.poll_single == 2;
            }
        }
        if (stale < 0 || dead || p.nosync) break;
        if (p.poll_single == 1) {
            unsigned long long w;
            do {
                w = ld_flagged1(vec + stale);
                if ((++it & 63u) == 0u && poll_watchdog(p.status, p.timeout_ns, t0, 0x500u, (unsigned)stale, epoch)) { dead = true; break; }
            } while ((unsigned)(w >> 32) != epoch);
        } else if ((++it & 63u) == 0u && poll_watchdog(p.status, p.timeout_ns, t0, 0x500u, (unsigned)stale, epoch)) {
            dead = true;
        }
    }
}
__device__ __forceinline__ float4 flagged_f4(const unsigned long long* e) {
    return make_float4(__uint_as_float((unsigned)e[0]), __uint_as_float((unsigned)e[1]), __uint_as_float((unsigned)e[2]), __uint_as_float((unsigned)e[3]));
}
// one flagged element, polled until it carries `epoch`
__device__ __forceinline__ float wait_flagged1(const DecParams& p, const unsigned long long* ptr, unsigned epoch, bool& dead) {
    unsigned long long t0 = 0, w;
    unsigned it = 0;
    while (true) {
        w = ld_flagged1(ptr);
        if ((unsigned)(w >> 32) == epoch || dead || p.nosync) break;
        if ((++it & 63u) == 0u && poll_watchdog(p.status, p.timeout_ns, t0, 0x501u, 0u, epoch)) { dead = true; break; }
    }
    return __uint_as_float((unsigned)w);
}

// four consecutive flagged elements (32-byte aligned), polled until all carry `epoch`
__device__ __forceinline__ float4 wait_flagged4(const DecParams& p, const unsigned long long* ptr, unsigned epoch, bool& dead) {
    unsigned long long e0, e1, e2, e3, t0 = 0;
    unsigned it = 0;
    while (true) {
        ld_flagged2(ptr, e0, e1);
        ld_flagged2(ptr + 2, e2, e3);
        const bool ok = (unsigned)(e0 >> 32) == epoch && (unsigned)(e1 >> 32) == epoch && (unsigned)(e2 >> 32) == epoch && (unsigned)(e3 >> 32) == epoch;
        if (ok || dead || p.nosync) break;
        if ((++it & 63u) == 0u && poll_watchdog(p.status, p.timeout_ns, t0, 0x502u, 0u, epoch)) { dead = true; break; }
    }
    return make_float4(__uint_as_float((unsigned)e0), __uint_as_float((unsigned)e1), __uint_as_float((unsigned)e2), __uint_as_float((unsigned)e3));
}

// Single-query attention over this CTA's (head, KV split) units (cmdbuf_mat_mul QK^T * 1/sqrt(D),
// cmdbuf_row_softmax, cmdbuf_mat_mul P*V; th-llama.cpp:365-380).  Every warp runs its own online
// softmax over the positions j = warp (mod 8) of each K/V tile pair -- no CTA barrier per tile -- and
// the 8 warps are merged once per unit.  The split results {m, l, o[D]} go to p.part; the Wo
// prologue merges the splits of a head.
// this lane's 4 dims of KV row j of a tile (f32 or f16 rows)
__device__ __forceinline__ float4 kv_row4(const void* tile, int j, int D, int lane, bool f16) {
    if (!f16) return *(const float4*)((const float*)tile + j * D + (lane << 2));
    const uint2 u = *(const uint2*)((const uint16_t*)tile + j * D + (lane << 2));
    const float2 a = h2_to_f2(u.x), b = h2_to_f2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ unsigned long long math_att_phase(const DecParams& p, unsigned long long state, int layer) {
    const Smem S = smem_view();
    const int ct = (int)threadIdx.x - kMathBase, mw = ct >> 5, lane = ct & 31;
    Cons c{{0u, 0u}, false};
    unsigned gq;
    unpack_state(state, gq, c.ring);
    const thk_llama_layer L = p.layers[layer];
    const AttSched a = make_att(p);
    const int D = p.head_dim;
    const float scale = 1.0f / sqrtf((float)D);
    const bool act = lane < (D >> 2);
    const bool kvh = p.kv_f16 != 0;
    float* sc_o = S.red + kAttScratchOff;             // [8][128]
    float* sc_m = sc_o + kMathWarps * kMaxHeadDim;    // [8]
    float* sc_l = sc_m + kMathWarps;                  // [8]
    float* sc_kn = sc_l + kMathWarps;                 // [128] the new position's K row (warp 0 only)
    float* sc_vn = sc_kn + kMaxHeadDim;               // [128] ... and V row
    for (int u = blockIdx.x; u < p.Hl * a.S; u += gridDim.x) {
        const int h = u / a.S, sp = u % a.S;
        const int pa = (int)(((long long)a.N * sp) / a.S);
        const int pb_full = (int)(((long long)a.N * (sp + 1)) / a.S);
        const int pb = min(pb_full, p.n_past);
        const bool has_new = (pb_full == a.N);
        float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act && p.att_flagged) {
            // no grid barrier behind the QKV phase: the read of this head's q (and K / V row) is the wait
            const unsigned epoch = p.flag_epoch + (unsigned)layer + 1u;
            q4 = wait_flagged4(p, p.qf + h * D + (lane << 2), epoch, c.dead);
            if (has_new && mw == 0) {
                *(float4*)(sc_kn + (lane << 2)) = wait_flagged4(p, p.knf + h * D + (lane << 2), epoch, c.dead);
                *(float4*)(sc_vn + (lane << 2)) = wait_flagged4(p, p.vnf + h * D + (lane << 2), epoch, c.dead);
            }
        } else if (act) {
            q4 = __ldcg((const float4*)(p.q + h * D) + lane);
            if (has_new && mw == 0) {   // the token's own K/V row (QKV epilogue of this launch): same L2 round trip as q,
                const size_t off = ((size_t)h * p.n_ctx + p.n_past) * D;     // parked in shared memory until the tiles are done
                float4 kn4, vn4;
                if (!kvh) {
                    kn4 = __ldcg((const float4*)(L.key_cache + off) + lane);
                    vn4 = __ldcg((const float4*)(L.value_cache + off) + lane);
                } else {
                    const uint2 ku = __ldcg((const uint2*)((const uint16_t*)L.key_cache + off) + lane);
                    const uint2 vu = __ldcg((const uint2*)((const uint16_t*)L.value_cache + off) + lane);
                    const float2 ka = h2_to_f2(ku.x), kb = h2_to_f2(ku.y), va = h2_to_f2(vu.x), vb = h2_to_f2(vu.y);
                    kn4 = make_float4(ka.x, ka.y, kb.x, kb.y);
                    vn4 = make_float4(va.x, va.y, vb.x, vb.y);
                }
                *(float4*)(sc_kn + (lane << 2)) = kn4;
                *(float4*)(sc_vn + (lane << 2)) = vn4;
            }
        }
        float m = -INFINITY, lsum = 0.f;
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int pos = pa; pos < pb; pos += p.att_tpos) {
            const int np = min(p.att_tpos, pb - pos);
            // the K tile and the V tile of these positions sit in consecutive slots
            wait_full(p, S, c, 4);
            const void* kt = (const void*)(S.slots + c.ring.sl * kSlotBytes);
            Cons cv = c;
            cv.ring.advance();
            wait_full(p, S, cv, 5);
            const void* vt = (const void*)(S.slots + cv.ring.sl * kSlotBytes);
            c.dead = c.dead || cv.dead;
            for (int j0 = mw; j0 < np; j0 += kAttPerRound * kMathWarps) {      // kAttPerRound positions of this warp per round
                float s[kAttPerRound];
#pragma unroll
                for (int t = 0; t < kAttPerRound; ++t) {
                    const int j = j0 + t * kMathWarps;
                    s[t] = (act && j < np) ? dot4(q4, kv_row4(kt, j, D, lane, kvh)) : 0.f;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                    for (int t = 0; t < kAttPerRound; ++t) s[t] += __shfl_xor_sync(0xffffffffu, s[t], off);
                }
                float mx = -INFINITY;
#pragma unroll
                for (int t = 0; t < kAttPerRound; ++t) { s[t] = (j0 + t * kMathWarps < np) ? s[t] * scale : -INFINITY; mx = fmaxf(mx, s[t]); }
                const float m_new = fmaxf(m, mx);
                const float corr = expf(m - m_new);
                lsum *= corr; o4.x *= corr; o4.y *= corr; o4.z *= corr; o4.w *= corr;
                m = m_new;
#pragma unroll
                for (int t = 0; t < kAttPerRound; ++t) {
                    const int j = j0 + t * kMathWarps;
                    if (j < np) {
                        const float pj = expf(s[t] - m);
                        lsum += pj;
                        if (act) {
                            const float4 v4 = kv_row4(vt, j, D, lane, kvh);
                            o4.x = fmaf(pj, v4.x, o4.x); o4.y = fmaf(pj, v4.y, o4.y);
                            o4.z = fmaf(pj, v4.z, o4.z); o4.w = fmaf(pj, v4.w, o4.w);
                        }
                    }
                }
            }
            release_slot(S, c, lane);
            release_slot(S, c, lane);
        }
        if (has_new && mw == 0) {   // the token's own position
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float sdot = 0.f;
            if (act) {
                sdot = dot4(q4, *(const float4*)(sc_kn + (lane << 2)));
                v4 = *(const float4*)(sc_vn + (lane << 2));
            }
            sdot = warp_sum(sdot) * scale;
            const float m_new = fmaxf(m, sdot);
            const float corr = expf(m - m_new);
            m = m_new;
            const float pj = expf(sdot - m);
            lsum = lsum * corr + pj;
            o4.x = fmaf(pj, v4.x, o4.x * corr); o4.y = fmaf(pj, v4.y, o4.y * corr);
            o4.z = fmaf(pj, v4.z, o4.z * corr); o4.w = fmaf(pj, v4.w, o4.w * corr);
        }
        // merge the 8 warps (fixed order -> deterministic)
        if (act) *(float4*)(sc_o + mw * kMaxHeadDim + (lane << 2)) = o4;
        if (lane == 0) { sc_m[mw] = m; sc_l[mw] = lsum; }
        bar_sync(BAR_MATH, kMathThreads);
        if (ct < D) {
            float M = sc_m[0];
#pragma unroll
            for (int w = 1; w < kMathWarps; ++w) M = fmaxf(M, sc_m[w]);
            float od = 0.f, lt = 0.f;
#pragma unroll
            for (int w = 0; w < kMathWarps; ++w) {
                const float e = expf(sc_m[w] - M);          // warps without positions have m = -inf -> 0
                od = fmaf(sc_o[w * kMaxHeadDim + ct], e, od);
                lt = fmaf(sc_l[w], e, lt);
            }
            float* part = p.part + (size_t)(h * a.S + sp) * part_stride(p);
            part[4 + ct] = od;
            if (ct == 0) { part[0] = M; part[1] = lt; }
        }
        bar_sync(BAR_MATH, kMathThreads);   // scratch reuse by the next unit
    }
    return pack_state(gq, c.ring);
}

// ---- prologues: stage the phase's activation vector in shared memory (permuted layout) ----
// A math warp only ever multiplies against ITS columns of the activation vector: chunk cw of every K tile, and within a
// chunk a lane only its own 8 columns.  So the staged vector xs is lane-private storage: every lane loads, transforms and
// stores exactly the values it will read back -- no CTA barrier, no cross-warp traffic, and a warp starts on its first
// tile as soon as its own loads have landed.
__device__ __forceinline__ float4 emb_f4(const uint16_t* row, int i) {
    const uint2 u = __ldg((const uint2*)(row + i));
    const float2 a = h2_to_f2(u.x), b = h2_to_f2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
}

// xs <- v * gain, v = the residual stream (cmdbuf_rms_norm + cmdbuf_row_element_multiply, th.cpp:1153-1200,1298-1315).
// The RMS scale 1/sqrt(mean(v^2) + 1e-6) is a scalar, so it is applied to the finished row sums by the epilogue warp
// instead: here every warp only leaves its share of sum(v^2) (returned; it travels to the epilogue warp inside every
// hand-off record).  fsrc != nullptr: v comes from a flagged vector (waits for `epoch`); else v is the embedding row
// (first phase of a launch), which is also published as the flagged residual stream `xf_out` (epoch `epoch`), every
// element by exactly one CTA.  The gain was L2-prefetched during the previous phase.
__device__ __forceinline__ float prologue_norm(const DecParams& p, int ph, const uint16_t* emb_row, const unsigned long long* fsrc, unsigned epoch,
                                            const float* gain, unsigned long long* xf_out) {
    const Smem S = smem_view();
    const int ct = (int)threadIdx.x - kMathBase, cw = ct >> 5, lane = ct & 31;
    const PhaseDesc& d = p.ph[ph];
    bool dead = false;
    const int n = d.C;
    const int lcol = (cw << 8) + (lane << 3);                 // this lane's first column inside a K tile
    const uint32_t xa = S.xs_a + (uint32_t)(((cw << 8) + (lane << 2)) << 2);
    float ss = 0.f;
    for (int kt0 = 0; kt0 < d.KT; kt0 += 2) {                 // two K tiles per L2 round trip
        float4 v[4], g[4];
        int col[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int kt = kt0 + u;
            const int c = kt * d.CT + lcol;
            col[u] = (kt < d.KT && lcol < d.CT && c < n) ? c : -1;
            if (col[u] >= 0) {
                g[2 * u] = __ldg((const float4*)(gain + c));
                g[2 * u + 1] = __ldg((const float4*)(gain + c + 4));
            }
        }
        if (fsrc != nullptr) {
            unsigned long long e[16];
            load_flagged<2>(p, fsrc, col, epoch, e, dead);
#pragma unroll
            for (int u = 0; u < 2; ++u) if (col[u] >= 0) { v[2 * u] = flagged_f4(e + 8 * u); v[2 * u + 1] = flagged_f4(e + 8 * u + 4); }
        } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) if (col[u] >= 0) { v[2 * u] = emb_f4(emb_row, col[u]); v[2 * u + 1] = emb_f4(emb_row, col[u] + 4); }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (col[u] >= 0) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const float4 t = v[2 * u + hh], gg = g[2 * u + hh];
                    ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
                    const int i = col[u] + 4 * hh;
                    if (xf_out && (unsigned)(i >> 2) % gridDim.x == blockIdx.x) {
                        st_flagged(xf_out + i, t.x, epoch); st_flagged(xf_out + i + 1, t.y, epoch);
                        st_flagged(xf_out + i + 2, t.z, epoch); st_flagged(xf_out + i + 3, t.w, epoch);
                    }
                    sts128f(xa + (uint32_t)(((kt0 + u) * d.CT) << 2) + 512u * hh, make_float4(t.x * gg.x, t.y * gg.y, t.z * gg.z, t.w * gg.w));
                }
            }
        }
    }
    return warp_sum(ss);
}
// xs <- the FFN hidden vector for W2, from the flagged vector (waits for `epoch`)
__device__ __forceinline__ void prologue_copy(const DecParams& p, int ph, const unsigned long long* fsrc, unsigned epoch) {
    const Smem S = smem_view();
    const int ct = (int)threadIdx.x - kMathBase, cw = ct >> 5, lane = ct & 31;
    const PhaseDesc& d = p.ph[ph];
    bool dead = false;
    const int n = d.C;
    const int lcol = (cw << 8) + (lane << 3);
    const uint32_t xa = S.xs_a + (uint32_t)(((cw << 8) + (lane << 2)) << 2);
    constexpr int NCH = THK_COPY_CHUNKS;                      // K tiles per L2 round trip (6 would cover the 7B FFN vector in one, but spills:
                                                              // local memory behind the weight stream cost 1.2 ms/token, measured)
    for (int kt0 = 0; kt0 < d.KT; kt0 += NCH) {
        int col[NCH];
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
            const int kt = kt0 + u;
            const int c = kt * d.CT + lcol;
            col[u] = (kt < d.KT && lcol < d.CT && c < n) ? c : -1;
        }
        unsigned long long e[NCH * 8];
        load_flagged<NCH>(p, fsrc, col, epoch, e, dead);
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
            if (col[u] >= 0) {
                sts128f(xa + (uint32_t)(((kt0 + u) * d.CT) << 2), flagged_f4(e + 8 * u));
                sts128f(xa + (uint32_t)(((kt0 + u) * d.CT) << 2) + 512u, flagged_f4(e + 8 * u + 4));
            }
        }
    }
}
// xs <- attention output of the local heads: merge the KV splits of every head (softmax denominators included).  A lane
// merges the two float4 of one chunk per round (they lie in the same head, so they share {m, l}); the loads of a round are
// in flight together (~0.6 us per L2 round trip under the weight stream).  (A flagged, barrier-free version of the
// QKV -> attention -> Wo hand-overs was built and measured slower: 2.69 -> 2.96 ms/token -- the polling of 128 split
// records by every CTA costs more than the two grid barriers it removes.)
__device__ __forceinline__ void prologue_att_merge(const DecParams& p, int ph) {
    const Smem S = smem_view();
    const int ct = (int)threadIdx.x - kMathBase, cw = ct >> 5, lane = ct & 31;
    const PhaseDesc& d = p.ph[ph];
    const AttSched a = make_att(p);
    const int D = p.head_dim, ps = part_stride(p);
    const int lcol = (cw << 8) + (lane << 3);
    const uint32_t xa = S.xs_a + (uint32_t)(((cw << 8) + (lane << 2)) << 2);
    for (int kt = 0; kt < d.KT; ++kt) {
        const int c = kt * d.CT + lcol;
        if (lcol >= d.CT || c >= d.C) continue;
        const int h = c / D, dd = c - h * D;                  // 8 consecutive columns never straddle a head (D % 8 == 0)
        const size_t pbase = (size_t)h * a.S * ps;
        float M = -INFINITY, lt = 0.f;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        for (int s0 = 0; s0 < a.S; s0 += 4) {
            float2 ml[4];
            float4 v0[4], v1[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) if (s0 + t < a.S) {
                const float* r = p.part + pbase + (size_t)(s0 + t) * ps;
                ml[t] = __ldcg((const float2*)r);
                v0[t] = __ldcg((const float4*)(r + 4 + dd));
                v1[t] = __ldcg((const float4*)(r + 8 + dd));
            }
            float Mn = M;
#pragma unroll
            for (int t = 0; t < 4; ++t) if (s0 + t < a.S) Mn = fmaxf(Mn, ml[t].x);
            const float c0 = expf(M - Mn);                   // first round: exp(-inf) = 0
            lt *= c0;
            o0.x *= c0; o0.y *= c0; o0.z *= c0; o0.w *= c0;
            o1.x *= c0; o1.y *= c0; o1.z *= c0; o1.w *= c0;
            M = Mn;
#pragma unroll
            for (int t = 0; t < 4; ++t) if (s0 + t < a.S) {
                const float e = expf(ml[t].x - M);
                lt = fmaf(ml[t].y, e, lt);
                o0.x = fmaf(v0[t].x, e, o0.x); o0.y = fmaf(v0[t].y, e, o0.y); o0.z = fmaf(v0[t].z, e, o0.z); o0.w = fmaf(v0[t].w, e, o0.w);
                o1.x = fmaf(v1[t].x, e, o1.x); o1.y = fmaf(v1[t].y, e, o1.y); o1.z = fmaf(v1[t].z, e, o1.z); o1.w = fmaf(v1[t].w, e, o1.w);
            }
        }
        o0.x /= lt; o0.y /= lt; o0.z /= lt; o0.w /= lt;
        o1.x /= lt; o1.y /= lt; o1.z /= lt; o1.w /= lt;
        sts128f(xa + (uint32_t)((kt * d.CT) << 2), o0);
        sts128f(xa + (uint32_t)((kt * d.CT) << 2) + 512u, o1);
    }
}
__device__ __forceinline__ void prefetch_gain(const float* gain, int n, int ct) {
    for (int off = ct * 32; off < n; off += kMathThreads * 32)       // one 128-byte line per thread
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gain + off));
}

__device__ void math_main(const DecParams& p, int tok) {
    const long long t_life = wstat_t0();
    const int ct = (int)threadIdx.x - kMathBase;
    unsigned long long* const pm = (PROF(p) != nullptr && ct == 0) ? PROF(p) + (size_t)blockIdx.x * kProfPhases * 8 : nullptr;
    unsigned long long state = 0ull;      // row groups handed over, ring position
    const uint16_t* emb_row = p.emb + (size_t)tok * p.n_embd;
    mark(pm, p, 0, PROF_START);
    const int nsteps = 5 * p.n_layer + 1;
    int l = 0, k = K_QKV;
    float ss = 0.f;                       // this warp's share of sum(v^2) of the current norm phase's input
    for (int i = 0; i < nsteps; ++i) {
        const thk_llama_layer* L = p.layers + (l < p.n_layer ? l : p.n_layer - 1);
        // ---- prologue ----
        const unsigned epoch = p.flag_epoch + (unsigned)l + 1u;          // flagged vectors written in layer l (K_QKV / K_OUT read layer l-1's x)
        const long long tp0 = wstat_t0();
#ifdef THK_EXP_NOPRO      // measurement build (wrong results): no prologue, profiles/r2_timeline_and_experiments.md section 1
        if (i == 0) {
#else
        if (k == K_QKV || k == K_W13 || k == K_OUT) {
#endif
            const float* gain = (k == K_QKV) ? L->attention_norm : (k == K_W13) ? L->ffn_norm : p.norm;
            // step 0: x <- f32(tok_embeddings[token]) (th-llama.cpp:577-585; device-side like :552-575): every CTA converts
            // the row itself and publishes its share as the flagged residual stream (the Wo epilogue / the reducer adds to it)
            const unsigned long long* fsrc = i == 0 ? nullptr : (k == K_W13 ? p.h1f : p.xf);
            ss = prologue_norm(p, phase_of(k), emb_row, fsrc, k == K_W13 ? epoch : epoch - 1u, gain, i == 0 ? p.xf : nullptr);
        }
#ifndef THK_EXP_NOPRO
        else if (k == K_WO) {
            prologue_att_merge(p, PH_WO);
        } else if (k == K_W2) {
            prologue_copy(p, PH_W2, p.fff, epoch);
        }
#endif
        wstat_add(WS_POLL, tp0);
        // ---- tiles ----
        if (k == K_ATT) {
            mark(pm, p, (unsigned)i, PROF_PROLOGUE);
            state = math_att_phase(p, state, l);
        } else if (p.ph[phase_of(k)].KT <= 2) {
            state = math_mat_phase<true>(p, phase_of(k), state, ss, (unsigned)i);
        } else {
            state = math_mat_phase<false>(p, phase_of(k), state, ss, (unsigned)i);
        }
        if (k == K_OUT) break;
        // ---- next phase's gain while the grid barrier forms ----
        if (k == K_WO) prefetch_gain(L->ffn_norm, p.n_embd, ct);
        else if (k == K_W2) prefetch_gain(l + 1 < p.n_layer ? p.layers[l + 1].attention_norm : p.norm, p.n_embd, ct);
        mark(pm, p, (unsigned)i, PROF_ARRIVE);
        if ((k == K_QKV && !p.att_flagged) || k == K_ATT) {   // (Wo -> W13 -> W2 -> QKV: the next prologue waits on the flagged vector itself)
            if (k == K_ATT) bar_sync(BAR_PRE, kMathThreads + 32);    // our global writes (split results) precede the epilogue warp's arrive
            bar_sync(BAR_ALL, kMathThreads + 32);                     // the epilogue warp has passed the grid barrier
        }
        mark(pm, p, (unsigned)i + 1, PROF_START);
        if (k == K_W2) { ++l; k = (l < p.n_layer) ? K_QKV : K_OUT; } else ++k;
    }
    wstat_flush(PROF(p), t_life);
}

// ------------------------------------------------------------------------------------------
// EPILOGUE WARP
// ------------------------------------------------------------------------------------------
// Grid-wide barrier between data-dependent phases, driven by lane 0 of the epilogue warp.  Only this warp
// (and, in the attention phase, the math warps that met it at BAR_PRE) wrote global data in the phase;
// __syncwarp / bar.sync order those writes before lane 0's gpu-scope release (cumulativity), and the
// acquire fence + BAR_ALL make the other CTAs' writes visible to every thread here (read with .cg).
__device__ __forceinline__ void grid_barrier(const DecParams& p, unsigned nbar, int lane, bool& dead) {
    const long long tw0 = wstat_t0();
    __syncwarp();
    if (lane == 0 && !dead && !p.nosync) {
        unsigned* const ctr = p.bar_ctr;
        red_release_add(ctr, 1u);
        const unsigned target = nbar * gridDim.x;
        unsigned long long t0 = 0;
        unsigned it = 0;
        while (ld_relaxed_u32(ctr) < target) {    // (pipelined polling, 4 loads in flight, was measured slower: 2.80 -> 2.87 ms/token)
            if ((++it & 63u) == 0u && poll_watchdog(p.status, p.timeout_ns, t0, 0x200u, nbar, target)) { dead = true; break; }
        }
        asm volatile("fence.acquire.gpu;" ::: "memory");
    }
    dead = __shfl_sync(0xffffffffu, dead ? 1 : 0, 0) != 0;
    wstat_add(WS_POLL, tw0);
}

enum EpiKind { EPI_QKV, EPI_WO, EPI_W13, EPI_W2, EPI_OUT };
struct EpiState { float gate; float best; int best_idx; };

template <bool kTP>
__device__ __forceinline__ void epi_mat_phase(const DecParams& p, const Smem& S, int ph, int kind, const thk_llama_layer* L, unsigned epoch,
                                              EpiState& es, unsigned& gq, bool& dead, int lane) {
    const PhaseDesc& d = p.ph[ph];
    const int D = p.head_dim, tp_size = kTP ? p.tp_size : 1, tp_rank = kTP ? p.tp_rank : 0, n_ctx = p.n_ctx, n_past = p.n_past;
    const int nsub = d.paired ? 2 : 1;
    // lane (r, q): row r = lane & 7 of the row group, math warps 2q and 2q + 1
    const int rr = lane & 7, qq = lane >> 3;
    const uint32_t my_a = S.red_a + (uint32_t)(((2 * qq) * kRecFloats + rr) * 4);
    const bool has_resid = (kind == EPI_WO || kind == EPI_W2) && !kTP;
    const bool need_scale = (kind == EPI_QKV || kind == EPI_W13 || kind == EPI_OUT);
    bool have_scale = false;
    float scale = 1.0f;
    // The residual operand comes from the flagged vector the phase input was derived from (x entering the layer for Wo,
    // h1 for W2): 32 rows per poll, one per lane, refilled every fourth row group -- its L2 round trip is off the
    // per-group path, and the epoch check makes the read self-synchronising.
    const unsigned long long* rsrc = (kind == EPI_WO) ? p.xf : p.h1f;
    const unsigned repoch = (kind == EPI_WO) ? epoch - 1u : epoch;
    int rbase = -(1 << 30);
    float rval = 0.f;
    RowIt it;
    it.init(S.misc, ph, d);
    while (it.valid()) {
        const int row0 = it.row0, nrows = it.nrows, si_cur = it.si;
        it.next();
        for (int sub = 0; sub < nsub; ++sub) {
            const int si = d.paired ? sub : si_cur;
            float resid = 0.f;
            if (has_resid) {
                if (row0 + nrows > rbase + 32) {
                    rbase = row0;
                    const long long t0 = wstat_t0();
                    rval = wait_flagged1(p, rsrc + min(rbase + lane, p.n_embd - 1), repoch, dead);
                    wstat_add(WS_HANDOFF, t0);
                }
                resid = __shfl_sync(0xffffffffu, rval, (row0 - rbase + rr) & 31);
            }
            const unsigned buf = gq & (kDumpBufs - 1), use = gq / kDumpBufs;
            if (!dead) {
                const uint32_t fb = S.red_full_a + buf * 8;
                if (!mbar_try_wait(fb, use & 1u)) {
                    const long long t0 = wstat_t0();
                    dead = !mbar_wait_slow(p.status, p.timeout_ns, fb, use & 1u, 7);
                    wstat_add(WS_RING, t0);
                }
            }
            const uint32_t src = my_a + buf * (uint32_t)(kRedFloats * 4);
            float y = lds32f(src) + lds32f(src + kRecFloats * 4);
            y += __shfl_xor_sync(0xffffffffu, y, 8);
            y += __shfl_xor_sync(0xffffffffu, y, 16);      // every lane: total of row (lane & 7)
            if (need_scale && !have_scale) {
                // RMSNorm scale of this phase's input (th.cpp:1153-1200): every hand-off record carries the eight warps'
                // shares of sum(v^2); added in warp order: deterministic.
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < kMathWarps; ++w) tot += lds32f(S.red_a + (buf * kRedFloats + w * kRecFloats + kRows) * 4u);
                scale = 1.0f / sqrtf(tot / (float)d.C + 1e-6f);
                have_scale = true;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(S.red_free_a + buf * 8);
            ++gq;
            if (need_scale) y *= scale;
            const float y1 = __shfl_down_sync(0xffffffffu, y, 1);     // the pair partner for RoPE
            if (lane < nrows) {
                const int r = row0 + lane;
                switch (kind) {
                case EPI_QKV:
                    if (si == 2) {                                   // V: append (th-llama.cpp:338)
                        const size_t idx = ((size_t)(r / D) * n_ctx + n_past) * D + (r % D);
                        if (!p.kv_f16) L->value_cache[idx] = y;
                        else ((__half*)L->value_cache)[idx] = __float2half_rn(y);
                        if (p.att_flagged) st_flagged(p.vnf + r, p.kv_f16 ? __half2float(__float2half_rn(y)) : y, epoch);
                    } else if ((lane & 1) == 0) {                    // Q / K: rotate the pair (r, r+1), th.cpp:1457-1492
                        const float2 cs = S.misc->rope[(r % D) >> 1];
                        const float a = y * cs.x - y1 * cs.y, b = y * cs.y + y1 * cs.x;
                        if (si == 0) {
                            *(float2*)(p.q + r) = make_float2(a, b);
                            if (p.att_flagged) { st_flagged(p.qf + r, a, epoch); st_flagged(p.qf + r + 1, b, epoch); }
                        } else {                                                                          // th-llama.cpp:337
                            const size_t idx = ((size_t)(r / D) * n_ctx + n_past) * D + (r % D);
                            if (!p.kv_f16) *(float2*)(L->key_cache + idx) = make_float2(a, b);
                            else *(__half2*)((__half*)L->key_cache + idx) = __floats2half2_rn(a, b);
                            if (p.att_flagged) {
                                st_flagged(p.knf + r, p.kv_f16 ? __half2float(__float2half_rn(a)) : a, epoch);
                                st_flagged(p.knf + r + 1, p.kv_f16 ? __half2float(__float2half_rn(b)) : b, epoch);
                            }
                        }
                    }
                    break;
                case EPI_WO:                                                           // th-llama.cpp:409
                    if (!kTP) st_flagged(p.h1f + r, resid + y, epoch);
                    else for (int dst = 0; dst < tp_size; ++dst) st_flagged_sys(xbf_ptr(p, dst, 0, tp_rank) + r, y, epoch);   // partial -> every rank
                    break;
                case EPI_W13:
                    if (sub == 0) es.gate = y;
                    else {                                                              // :436,:438
                        const float gv = es.gate;
                        st_flagged(p.fff + r, (gv / (1.0f + expf(-gv))) * y, epoch);
                    }
                    break;
                case EPI_W2:                                                           // th-llama.cpp:447
                    if (!kTP) { const float v = resid + y; p.x[r] = v; st_flagged(p.xf + r, v, epoch); }
                    else for (int dst = 0; dst < tp_size; ++dst) st_flagged_sys(xbf_ptr(p, dst, 1, tp_rank) + r, y, epoch);
                    break;
                default: {                                                             // EPI_OUT
                    if (p.logits) p.logits[r] = y;
                    const int gid = tp_rank * p.Vl + r;
                    if (es.best_idx < 0 || y > es.best) { es.best = y; es.best_idx = gid; }
                } break;
                }
            }
        }
    }
}

template <bool kTP>
__device__ void epi_main(const DecParams& p, const Smem& S) {
    const long long t_life = wstat_t0();
    const int lane = (int)threadIdx.x & 31;
    EpiState es{0.f, 0.f, -1};
    bool dead = false;
    unsigned gq = 0, nbar = 0;
    // RoPE table for this token: cos/sin(n_past * 10000^(-2i/D)) (th.cpp:1478-1482), once per CTA
    for (int i = lane; i < (p.head_dim >> 1); i += 32) {
        const float theta = powf(10000.0f, (-(float)(2 * i)) / (float)p.head_dim);
        float sn, cs;
        sincosf((float)p.n_past * theta, &sn, &cs);
        S.misc->rope[i] = make_float2(cs, sn);
    }
    __syncwarp();
    const int nsteps = 5 * p.n_layer + 1;
    int l = 0, k = K_QKV;
    for (int i = 0; i < nsteps; ++i) {
        if (k != K_ATT) {
            const int kind = k == K_QKV ? EPI_QKV : k == K_WO ? EPI_WO : k == K_W13 ? EPI_W13 : k == K_W2 ? EPI_W2 : EPI_OUT;
            epi_mat_phase<kTP>(p, S, phase_of(k), kind, k == K_OUT ? nullptr : p.layers + l, p.flag_epoch + (unsigned)l + 1u, es, gq, dead, lane);
        } else {
            bar_sync(BAR_PRE, kMathThreads + 32);
        }
        if (k == K_OUT) break;
        if ((k == K_QKV && !p.att_flagged) || k == K_ATT) {
            ++nbar;
            grid_barrier(p, nbar, lane, dead);
            bar_sync(BAR_ALL, kMathThreads + 32);
        }
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
        ++nbar;
        grid_barrier(p, nbar, lane, dead);
        if (blockIdx.x == 0) {
            // every lane takes CTAs lane, lane+32, ... (loads in flight together), then the warp reduces
            float gv = 0.f; int gi = -1;
            for (unsigned b0 = 0; b0 < gridDim.x; b0 += 128) {
                int idx[4]; float v[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const unsigned b = b0 + (unsigned)(t * 32 + lane);
                    idx[t] = b < gridDim.x ? __ldcg(p.amax_idx + b) : -1;
                    v[t] = b < gridDim.x ? __ldcg(p.amax_val + b) : 0.f;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (idx[t] >= 0 && (gi < 0 || v[t] > gv || (v[t] == gv && idx[t] < gi))) { gv = v[t]; gi = idx[t]; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, gv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, gi, off);
                if (oi >= 0 && (gi < 0 || ov > gv || (ov == gv && oi < gi))) { gv = ov; gi = oi; }
            }
            if (lane == 0) {
                if (kTP) {
                    // cross-rank argmax: push this rank's candidate to every rank, wait for all, lowest id wins ties
                    const unsigned epoch = p.epoch_base + 1u;
                    for (int d = 0; d < p.tp_size; ++d) { xamax_val(p, d)[p.tp_rank] = gv; xamax_idx(p, d)[p.tp_rank] = gi; }
                    __threadfence_system();
                    for (int d = 0; d < p.tp_size; ++d) st_release_sys(xflags(p, d, 1) + p.tp_rank, epoch);
                    unsigned long long t0 = 0; unsigned it = 0;
                    gv = 0.f; gi = -1;
                    for (int src = 0; src < p.tp_size; ++src) {
                        const unsigned* f = xflags(p, p.tp_rank, 1) + src;
                        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
                            if ((++it & 63u) == 0u && poll_watchdog(p.status, p.timeout_ns, t0, 0x401u, (unsigned)src, epoch)) break;
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
    if (PROF(p) && lane == 0) prof_mark(PROF(p), 5u * (unsigned)p.n_layer + 1u, PROF_START);
    wstat_flush(PROF(p), t_life);
}

// ------------------------------------------------------------------------------------------
// REDUCER WARP (tensor parallel)
// ------------------------------------------------------------------------------------------
// The all-reduce after Wo and W2 (th-llama.cpp:402-409, 441-447 under Megatron-style sharding) in one NVLink hop plus one
// local hop: every rank's epilogue pushed its partial rows into the slot [which][source rank] of every rank's exchange
// region; here CTA b owns elements [E b / grid, E (b+1) / grid) of the residual stream, one or a few per lane, IN
// REGISTERS for the whole token.  Per exchange a lane polls the tp partials of its elements (all loads in flight; the
// epoch stamp makes the read the wait), adds them in rank order -- bitwise identical on every rank -- and publishes the
// new residual value in the local flagged vector the next prologue (and nobody else) polls.
__device__ void reducer_main(const DecParams& p, int tok) {
    const long long t_life = wstat_t0();
    const int lane = (int)threadIdx.x & 31;
    const int E = p.n_embd, tp = p.tp_size;
    const int i0 = (int)(((long long)E * blockIdx.x) / gridDim.x), i1 = (int)(((long long)E * (blockIdx.x + 1)) / gridDim.x);
    const uint16_t* emb_row = p.emb + (size_t)tok * E;
    float resid[kMaxOwn];
#pragma unroll
    for (int j = 0; j < kMaxOwn; ++j) {
        const int i = i0 + lane + 32 * j;
        resid[j] = i < i1 ? __half2float(__ushort_as_half(__ldg(emb_row + i))) : 0.f;
    }
    const unsigned long long* xbf = (const unsigned long long*)p.xchg[p.tp_rank];
    bool dead = false;
    for (int l = 0; l < p.n_layer; ++l) {
        const unsigned epoch = p.flag_epoch + (unsigned)l + 1u;
        for (int which = 0; which < 2; ++which) {
            unsigned long long* dstv = which == 0 ? p.h1f : p.xf;
#pragma unroll
            for (int j = 0; j < kMaxOwn; ++j) {
                const int i = i0 + lane + 32 * j;
                if (i >= i1) continue;
                const unsigned long long* src = xbf + (size_t)which * tp * E + i;
                unsigned long long e[8];
                unsigned long long t0 = 0;
                unsigned it = 0;
                const long long tw0 = wstat_t0();
                while (true) {
                    bool ok = true;
#pragma unroll
                    for (int r = 0; r < 8; ++r) if (r < tp) e[r] = ld_flagged1_sys(src + (size_t)r * E);
#pragma unroll
                    for (int r = 0; r < 8; ++r) if (r < tp && (unsigned)(e[r] >> 32) != epoch) ok = false;
                    if (ok || dead || p.nosync) break;
                    if ((++it & 63u) == 0u && poll_watchdog(p.status, p.timeout_ns, t0, 0x600u | (unsigned)which, (unsigned)i, epoch)) dead = true;
                }
                wstat_add(WS_POLL, tw0);
                float v = resid[j];
#pragma unroll
                for (int r = 0; r < 8; ++r) if (r < tp) v += __uint_as_float((unsigned)e[r]);
                resid[j] = v;
                st_flagged(dstv + i, v, epoch);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxOwn; ++j) {
        const int i = i0 + lane + 32 * j;
        if (i < i1) p.x[i] = resid[j];
    }
    wstat_flush(PROF(p), t_life);
}

template <bool kTP>
__global__ void __launch_bounds__(kThreads, 1) decode_kernel(const __grid_constant__ DecParams p) {
    const Smem S = smem_view();
    const int tok = *p.token;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kNumSlots; ++i) {
            mbar_init(S.full_a + i * 8, 1);
            mbar_init(S.empty_a + i * 8, kMathWarps);
        }
        for (int i = 0; i < kDumpBufs; ++i) {
            mbar_init(S.red_full_a + i * 8, kMathWarps);
            mbar_init(S.red_free_a + i * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x == 0) *p.bar_next = 0u;   // arm the next launch's barrier counter
    }
    if (PROF(p) && threadIdx.x < (kMathWarps + kSvcWarps) * 4) (&S.misc->wstat[0][0])[threadIdx.x] = 0ull;
    // A token id outside the vocabulary (th-llama.cpp:606 asserts) is the same on every CTA (and every rank): the whole
    // grid leaves before any phase starts, nothing waits on anything, and the host reads THK_E_INVALID from the status word.
    if (tok < 0 || tok >= p.n_vocab) {
        if (threadIdx.x == 0 && blockIdx.x == 0) raise_abort(p, 0x300u, (unsigned)tok, 0u);
        return;
    }
    if ((int)threadIdx.x >= kSvcBase + 64 && (int)threadIdx.x < kSvcBase + 69) {  // (lanes of the reducer warp)
        const int ph = (int)threadIdx.x - (kSvcBase + 64);
        RowIt::share(p.ph[ph], gridDim.x, blockIdx.x, S.misc->range[ph][0], S.misc->range[ph][1]);
        RowIt it;
        int n = 0;
        for (it.init_range(S.misc->range[ph][0], S.misc->range[ph][1], p.ph[ph]); it.valid(); it.next()) ++n;
        S.misc->niter[ph] = n * (p.ph[ph].paired ? 2 : 1);
    }
    __syncthreads();
    if ((int)threadIdx.x >= kSvcBase) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kServiceRegs));
        const int sw = ((int)threadIdx.x - kSvcBase) >> 5;         // service warp index
        if (sw == 0) producer_main(p, S);
        else if (sw == 1) epi_main<kTP>(p, S);
        else if (sw == 2 && kTP) reducer_main(p, tok);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kMathRegs));
        math_main(p, tok);
    }
}

size_t decode_smem_bytes(int max_vec) {
    return (size_t)kXsOffset + (size_t)((max_vec + 255) & ~255) * sizeof(float);
}

PhaseDesc make_phase(int nseg, const int* rows, int C, bool paired) {
    PhaseDesc d{};
    d.C = C; d.nseg = nseg; d.paired = paired ? 1 : 0;
    for (int i = 0; i < 3; ++i) d.rows[i] = i < nseg ? rows[i] : 0;
    const int chunks = (C + 255) / 256;
    d.KT = (chunks + kChunks - 1) / kChunks;
    d.CT = ((chunks + d.KT - 1) / d.KT) * 256;          // K split evenly over the tiles
    return d;
}

}  // namespace

// the five matvec phases of a (possibly tensor-parallel) model: segments, rows and K tiling
static void fill_phases(DecParams& p) {
    { const int r[3] = {p.Eh, p.Eh, p.Eh}; p.ph[PH_QKV] = make_phase(3, r, p.n_embd, false); }
    { const int r[3] = {p.n_embd, 0, 0}; p.ph[PH_WO] = make_phase(1, r, p.Eh, false); }
    { const int r[3] = {p.Fh, p.Fh, 0}; p.ph[PH_W13] = make_phase(2, r, p.n_embd, true); }
    { const int r[3] = {p.n_embd, 0, 0}; p.ph[PH_W2] = make_phase(1, r, p.Fh, false); }
    { const int r[3] = {p.Vl, 0, 0}; p.ph[PH_OUT] = make_phase(1, r, p.n_embd, false); }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct thk_decoder {
    thk_ctx* ctx = nullptr;
    DecParams p{};
    thk_llama_layer* d_layers = nullptr;
    float* scratch = nullptr;
    unsigned long long* flagged = nullptr;   // h1f [n_embd] | xf [n_embd] | fff [Fh]
    unsigned flag_seq = 0;
    unsigned* ctrl = nullptr;       // [64 barrier counters][4 status]
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
        thk_set_error("decode kernel aborted: code 0x%x a=%u b=%u cta=%u (0x1xx mbarrier wait, 0x200 grid barrier, 0x300 bad token, 0x4xx argmax exchange, "
                      "0x5xx flagged vector, 0x6xx tensor-parallel partials)", st[0], st[1], st[2], st[3]);
        cudaMemsetAsync(d->p.status, 0, sizeof st, d->ctx->stream);
        return st[0] == 0x300u ? THK_E_INVALID : THK_E_TIMEOUT;
    }
    return THK_OK;
}

extern "C" int thk_decoder_destroy(thk_decoder* d) {
    if (!d) return THK_OK;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    cudaFree(d->d_layers); cudaFree(d->scratch); cudaFree(d->flagged); cudaFree(d->ctrl); cudaFree(d->d_tok); cudaFree(d->d_prof); cudaFree(d->xchg);
    delete d;
    return THK_OK;
}

static int decoder_alloc(thk_decoder* d, const thk_llama_layer* layers) {
    DecParams& p = d->p;
    const int tp = p.tp_size, D = p.head_dim;
    if (tp > 1) THK_CUDA(cudaFuncSetAttribute(decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
    else THK_CUDA(cudaFuncSetAttribute(decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
    THK_CUDA(cudaMalloc(&d->d_layers, sizeof(thk_llama_layer) * p.n_layer));
    THK_CUDA(cudaMemcpy(d->d_layers, layers, sizeof(thk_llama_layer) * p.n_layer, cudaMemcpyHostToDevice));
    p.layers = d->d_layers;
    // scratch: x [E]; q [Eh]; part [Hl*S*(D+4)] split-KV results {m, l, -, -, o[D]}; amax [grid] x2
    const size_t part_n = (size_t)p.Hl * p.att_max_split * (D + 4);
    const size_t nfl = (size_t)p.n_embd + p.Eh + part_n + 2 * d->grid + 64;
    THK_CUDA(cudaMalloc(&d->scratch, nfl * sizeof(float)));
    THK_CUDA(cudaMemset(d->scratch, 0, nfl * sizeof(float)));
    float* f = d->scratch;
    p.x = f; f += p.n_embd; p.q = f; f += p.Eh;
    p.part = f; f += part_n; p.amax_val = f; f += d->grid; p.amax_idx = (int*)f;
    const size_t nflag = (size_t)2 * p.n_embd + (size_t)((p.Fh + 1) & ~1) + (size_t)3 * p.Eh;
    THK_CUDA(cudaMalloc(&d->flagged, nflag * sizeof(unsigned long long)));
    THK_CUDA(cudaMemset(d->flagged, 0, nflag * sizeof(unsigned long long)));     // epoch 0 is never used
    p.h1f = d->flagged; p.xf = p.h1f + p.n_embd; p.fff = p.xf + p.n_embd;
    p.qf = p.fff + ((p.Fh + 1) & ~1); p.knf = p.qf + p.Eh; p.vnf = p.knf + p.Eh;
    const size_t nctrl = 64 + 4;
    THK_CUDA(cudaMalloc(&d->ctrl, nctrl * sizeof(unsigned)));
    THK_CUDA(cudaMemset(d->ctrl, 0, nctrl * sizeof(unsigned)));
    p.status = d->ctrl + 64;
    THK_CUDA(cudaMalloc(&d->d_tok, sizeof(int) * 2));
    if (tp > 1) {
        d->xchg_bytes = ((size_t)2 * tp * p.n_embd * sizeof(unsigned long long) + (size_t)4 * tp * sizeof(unsigned) + 255) & ~(size_t)255;
        THK_CUDA(cudaMalloc(&d->xchg, d->xchg_bytes));
        THK_CUDA(cudaMemset(d->xchg, 0, d->xchg_bytes));
        p.xchg[p.tp_rank] = d->xchg;
    }
    THK_CUDA(cudaDeviceSynchronize());
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
    THK_CHECK_ARG(D % 16 == 0 && D <= kMaxHeadDim, "head_dim %d unsupported (need a multiple of 16, <= %d)", D, kMaxHeadDim);
    THK_CHECK_ARG(dims->n_embd % 8 == 0 && dims->n_ff % 8 == 0, "n_embd and n_ff must be multiples of 8 (16-byte f16 rows)");
    THK_CHECK_ARG(dims->n_head % tp == 0 && dims->n_ff % tp == 0 && dims->n_vocab % tp == 0, "tp_size must divide n_head, n_ff, n_vocab");
    THK_CHECK_ARG((dims->n_ff / tp) % 8 == 0, "n_ff / tp_size = %d must be a multiple of 8 (16-byte rows of the W2 shard)", dims->n_ff / tp);
    THK_CHECK_ARG(dims->n_ctx > 0 && dims->n_layer > 0 && dims->n_vocab > 0, "bad dims");
    THK_CHECK_ARG(dims->n_layer <= 254, "n_layer %d unsupported (epoch-stamped vectors encode the layer in 8 bits)", dims->n_layer);
    THK_CHECK_ARG(tp == 1 || dims->n_embd <= 32 * kMaxOwn * ctx->sm_count, "n_embd %d too large for the tensor-parallel reducer (<= %d)", dims->n_embd, 32 * kMaxOwn * ctx->sm_count);
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
    fill_phases(p);
    if (getenv("THK_DEBUG"))
        for (int i = 0; i < 5; ++i)
            fprintf(stderr, "phase %d: C=%d CT=%d KT=%d rows=%d,%d,%d paired=%d\n", i, p.ph[i].C, p.ph[i].CT, p.ph[i].KT,
                    p.ph[i].rows[0], p.ph[i].rows[1], p.ph[i].rows[2], p.ph[i].paired);
    p.kv_f16 = dims->kv_f16 ? 1 : 0;
    p.att_tpos = kSlotBytes / (D * (p.kv_f16 ? 2 : 4));
    if (p.att_tpos > kMaxTilePos) p.att_tpos = kMaxTilePos;
    p.att_max_split = d->grid / p.Hl > 0 ? d->grid / p.Hl : 1;
    if (p.att_max_split > kMaxSplit) p.att_max_split = kMaxSplit;
    p.timeout_ns = 4000000000ull;
    p.prof_phase = -1;
    // tuning knobs (defaults = the measured best)
    p.l2_ahead = (unsigned)(getenv("THK_L2_AHEAD_KB") ? atoi(getenv("THK_L2_AHEAD_KB")) : 64) * 1024u;
    p.poll_single = getenv("THK_POLL_SINGLE") ? atoi(getenv("THK_POLL_SINGLE")) : 2;
    p.att_flagged = getenv("THK_ATT_FLAGGED") ? atoi(getenv("THK_ATT_FLAGGED")) : 0;
    const int max_vec = p.n_embd > p.Fh ? p.n_embd : p.Fh;
    d->smem = decode_smem_bytes(max_vec);
    if (d->smem > 227 * 1024) {
        thk_set_error("thk_decoder_create: needs %zu bytes of shared memory (> 227 KB); n_ff/tp=%d too large", d->smem, p.Fh);
        delete d;
        return THK_E_UNSUPPORTED;
    }
    const int rc = decoder_alloc(d, layers);
    if (rc != THK_OK) { thk_decoder_destroy(d); return rc; }     // one cleanup path: frees whatever was allocated
    *out = d;
    return THK_OK;
}

static int launch_step(thk_decoder* d, const int32_t* token, int32_t n_past, float* logits, int32_t* next_token, float* next_logit) {
    THK_ENTER(d->ctx);
    DecParams p = d->p;
    p.token = token; p.n_past = n_past; p.logits = logits; p.next_token = next_token; p.next_logit = next_logit;
    if (p.tp_size > 1) {
        if (!d->peers_set) { thk_set_error("thk_decoder_step: tensor-parallel decoder has no peers (call thk_decoder_set_peers)"); return THK_E_INVALID; }
        // the end-of-launch argmax exchange always runs: every rank embeds the token the ranks agreed on
        if (!p.next_token && !p.next_logit) p.next_token = d->d_tok + 1;
        p.epoch_base = d->epoch;
        d->epoch += 2u;
    }
    p.flag_epoch = (++d->flag_seq) << 8;                 // layer l of this launch stamps its vectors with flag_epoch + l + 1
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
    if (p.tp_size > 1) THK_CUDA(cudaLaunchKernelEx(&cfg, decode_kernel<true>, p));
    else THK_CUDA(cudaLaunchKernelEx(&cfg, decode_kernel<false>, p));
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

// timeline profile: per CTA, per phase (= grid barriers passed so far: 5*layer + {qkv,att,wo,w13,w2}, then logits),
// 8 u64 slots of %globaltimer ns (ProfSlot), followed by per-CTA producer stats [empty-wait cycles, total cycles, tiles, smid]
extern "C" int thk_decoder_profile(thk_decoder* d, int enable, unsigned long long* host_out, int n) {
    THK_CHECK_ARG(d, "thk_decoder_profile: null argument");
#ifndef THK_PROFILE
    if (enable) { thk_set_error("thk_decoder_profile: this library was built without -DTHK_PROFILE (load token_hawk_b200/lib_prof: THK_LIBDIR=lib_prof)"); return THK_E_UNSUPPORTED; }
#endif
    THK_ENTER(d->ctx);
    if (enable && !d->d_prof) {
        const size_t nprof = (size_t)d->grid * (kProfPhases * 8 + 4 + 2 * kProfTiles + kProfWarps * 4);
        THK_CUDA(cudaMalloc(&d->d_prof, nprof * sizeof(unsigned long long)));
        THK_CUDA(cudaMemset(d->d_prof, 0, nprof * sizeof(unsigned long long)));
    }
    d->p.prof = enable ? d->d_prof : nullptr;
    if (host_out && n > 0 && d->d_prof) {
        THK_CUDA(cudaStreamSynchronize(d->ctx->stream));
        const size_t nprof = (size_t)d->grid * (kProfPhases * 8 + 4 + 2 * kProfTiles + kProfWarps * 4);
        THK_CUDA(cudaMemcpy(host_out, d->d_prof, sizeof(unsigned long long) * ((size_t)n > nprof ? nprof : (size_t)n), cudaMemcpyDeviceToHost));
    }
    return THK_OK;
}

extern "C" int thk_decoder_tune(thk_decoder* d, const char* key, int value) {
    THK_CHECK_ARG(d && key, "thk_decoder_tune: null argument");
    if (!strcmp(key, "l2_ahead_kb")) { THK_CHECK_ARG(value >= 0 && value <= 1024, "l2_ahead_kb out of range"); d->p.l2_ahead = (unsigned)value * 1024u; }
    else if (!strcmp(key, "prof_phase")) d->p.prof_phase = value;
    else if (!strcmp(key, "att_flagged")) d->p.att_flagged = value != 0;
    else if (!strcmp(key, "poll_single")) { THK_CHECK_ARG(value >= 0 && value <= 2, "poll_single: 0, 1 or 2"); d->p.poll_single = value; }
    else if (!strcmp(key, "nosync")) d->p.nosync = value != 0;
    else if (!strcmp(key, "timeout_ms")) { THK_CHECK_ARG(value > 0, "timeout_ms must be positive"); d->p.timeout_ns = (unsigned long long)value * 1000000ull; }
    else { thk_set_error("thk_decoder_tune: unknown key %s", key); return THK_E_INVALID; }
    return THK_OK;
}

// The static schedule of one CTA in one matvec phase, computed on the host with the very code the kernel runs
// (RowIt): out = {C, KT, CT, paired, n_groups, then per group: segment, first row, rows}.  Lets CPU tests check that the
// partition covers every row exactly once for any shape / grid / tensor-parallel split without a GPU.
extern "C" int thk_decoder_plan(const thk_llama_dims* dims, int phase, int grid, int cta, int32_t* out, int cap) {
    THK_CHECK_ARG(dims && out && phase >= 0 && phase < 5 && grid > 0 && cta >= 0 && cta < grid && cap >= 5, "thk_decoder_plan: bad argument");
    const int tp = dims->tp_size > 0 ? dims->tp_size : 1;
    THK_CHECK_ARG(dims->n_embd > 0 && dims->n_ff > 0 && dims->n_vocab > 0 && dims->n_embd % tp == 0 && dims->n_ff % tp == 0 && dims->n_vocab % tp == 0,
                  "thk_decoder_plan: bad dims");
    DecParams pp{};
    pp.n_embd = dims->n_embd; pp.Eh = dims->n_embd / tp; pp.Fh = dims->n_ff / tp; pp.Vl = dims->n_vocab / tp;
    fill_phases(pp);
    const PhaseDesc& ph = pp.ph[phase];
    int r0, r1;
    RowIt::share(ph, (unsigned)grid, (unsigned)cta, r0, r1);
    RowIt it;
    it.init_range(r0, r1, ph);
    out[0] = ph.C; out[1] = ph.KT; out[2] = ph.CT; out[3] = ph.paired;
    int n = 0;
    for (; it.valid(); it.next()) {
        if (5 + 3 * (n + 1) > cap) { thk_set_error("thk_decoder_plan: output too small"); return THK_E_INVALID; }
        out[5 + 3 * n] = it.si; out[6 + 3 * n] = it.row0; out[7 + 3 * n] = it.nrows;
        ++n;
    }
    out[4] = n;
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
    if (flags) *flags = d->xchg + (size_t)2 * d->p.tp_size * d->p.n_embd * sizeof(unsigned long long) + (size_t)2 * d->p.tp_size * sizeof(unsigned);
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
