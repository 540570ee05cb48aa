// th.cpp -- host side of the TokenHawk op surface (include/th/th.hpp) over the CUDA C ABI.
// Each cmdbuf_* keeps the reference op's validation rules and two-phase pipeline behaviour
// (kayvr/token-hawk th.cpp, cited per function) and ends in exactly one thk_* launch.
#include "th/th.hpp"

#include <math.h>
#include <string.h>
#include <string>
#include <vector>

#include <chrono>

namespace th {

std::string get_TensorType_name(TensorType dt) {
    switch (dt) {
    case TensorType_F16: return "f16";
    case TensorType_F32: return "f32";
    default: return "unknown";
    }
}

double get_time_seconds() {   // th.cpp:23-28
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// fp16 helpers (th.cpp:291-359): host-side only (loader, embedding backup)
static inline float bits_f(uint32_t w) { float f; memcpy(&f, &w, 4); return f; }
static inline uint32_t f_bits(float f) { uint32_t w; memcpy(&w, &f, 4); return w; }
float ggml_compute_fp16_to_fp32(ggml_fp16_t h) {
    const uint32_t w = (uint32_t)h << 16, sign = w & 0x80000000u, two_w = w + w;
    const float normalized = bits_f((two_w >> 4) + (0xE0u << 23)) * bits_f(0x7800000u);
    const float denormalized = bits_f((two_w >> 17) | (126u << 23)) - 0.5f;
    return bits_f(sign | (two_w < (1u << 27) ? f_bits(denormalized) : f_bits(normalized)));
}
ggml_fp16_t ggml_compute_fp32_to_fp16(float f) {
    float base = (fabsf(f) * bits_f(0x77800000u)) * bits_f(0x08800000u);
    const uint32_t w = f_bits(f), shl1_w = w + w, sign = w & 0x80000000u;
    uint32_t bias = shl1_w & 0xFF000000u;
    if (bias < 0x71000000u) bias = 0x71000000u;
    base = bits_f((bias >> 1) + 0x07800000u) + base;
    const uint32_t bits = f_bits(base);
    const uint32_t nonsign = ((bits >> 13) & 0x00007C00u) + (bits & 0x00000FFFu);
    return (ggml_fp16_t)((sign >> 16) | (shl1_w > 0xFF000000u ? 0x7E00u : nonsign));
}

// ---------------------------------------------------------------------------------------------
// TensorBuffer (th.cpp:150-229)
// ---------------------------------------------------------------------------------------------
TensorBuffer::TensorBuffer(TensorShape shapeIn, TensorType typeIn, WGPUDevice device, WGPUBufferUsageFlags usage) {
    shape = shapeIn;
    type = typeIn;
    originalShape = shape;
    if (device) allocate_gpu_memory(device, usage);
}

TensorBuffer::TensorBuffer(const void* data, TensorShape shapeIn, TensorType typeIn, bool backup, WGPUDevice device,
                           WGPUQueue queue, WGPUBufferUsageFlags usage) {
    shape = shapeIn;
    type = typeIn;
    originalShape = shape;
    if (backup) {
        cpuBackup.resize(get_size_bytes());
        memcpy(cpuBackup.data(), data, cpuBackup.size());
    }
    if (device) allocate_gpu_memory(device, usage);
    if (queue) upload_data_to_gpu(queue, data);
}

TensorBuffer& TensorBuffer::operator=(TensorBuffer&& o) noexcept {
    if (this != &o) {
        free_buffers();
        shape = o.shape; type = o.type; cpuOnly = o.cpuOnly; gpu = o.gpu; owner = o.owner;
        originalShape = o.originalShape; cpuBackup = std::move(o.cpuBackup); name = std::move(o.name);
        o.gpu = nullptr; o.cpuBackup.clear();
    }
    return *this;
}
TensorBuffer::TensorBuffer(TensorBuffer&& o) noexcept
    : shape(o.shape), type(o.type), cpuOnly(o.cpuOnly), cpuBackup(std::move(o.cpuBackup)), gpu(o.gpu),
      originalShape(o.originalShape), name(std::move(o.name)), owner(o.owner) {
    o.gpu = nullptr;
    o.cpuBackup.clear();
}

size_t TensorBuffer::get_size_bytes() const {
    assert(type != TensorType_Unknown);
    return (size_t)shape.get_total_num_elements() * get_TensorType_size(type);
}

void TensorBuffer::allocate_gpu_memory(WGPUDevice device, WGPUBufferUsageFlags) {
    assert(!gpu);
    assert(type != TensorType_Unknown);
    void* p = nullptr;
    if (thk_malloc(device, get_size_bytes(), &p) != THK_OK) {
        fprintf(stderr, "TensorBuffer: %s\n", thk_last_error());
        return;
    }
    gpu = p;
    owner = device;
}

void TensorBuffer::upload_data_to_gpu(WGPUQueue queue, const void* data) {
    assert(gpu);
    if (thk_upload(queue, gpu, 0, data, get_size_bytes()) != THK_OK || thk_sync(queue) != THK_OK)
        fprintf(stderr, "TensorBuffer upload: %s\n", thk_last_error());
}

void TensorBuffer::free_buffers() {
    if (gpu != nullptr) {
        thk_free(owner, gpu);
        gpu = nullptr;
    }
}

void print_TensorBuffer(TensorBuffer* buffer, const char* bufferName) {   // th.cpp:280-289
    printf("%s: %s %s gpu=%p\n", bufferName, get_TensorType_name(buffer->type).c_str(), buffer->shape.to_string().c_str(), buffer->gpu);
}

// ---------------------------------------------------------------------------------------------
// pipeline validation cache (th.cpp:89-148)
// ---------------------------------------------------------------------------------------------
static void store_pipeline_validation(ComputePipeline& p, const TensorBuffer* A, const TensorBuffer* B = nullptr,
                                      const TensorBuffer* C = nullptr) {
    if (A) { p.sa = A->shape; p.ta = A->type; }
    if (B) { p.sb = B->shape; p.tb = B->type; }
    if (C) { p.sc = C->shape; p.tc = C->type; }
    p.built = true;
}
static bool validate_pipeline(const ComputePipeline& p, const TensorBuffer* A, const TensorBuffer* B = nullptr,
                              const TensorBuffer* C = nullptr) {
    if (A && (p.sa != A->shape || p.ta != A->type)) return false;
    if (B && (p.sb != B->shape || p.tb != B->type)) return false;
    if (C && (p.sc != C->shape || p.tc != C->type)) return false;
    return true;
}
bool are_pipelines_similar(const ComputePipeline& a, const ComputePipeline& b) {
    return a.sa == b.sa && a.sb == b.sb && a.sc == b.sc && a.ta == b.ta && a.tb == b.tb && a.tc == b.tc;
}

#define TH_FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return {}; } while (0)

// Shared two-phase logic.  Returns: 0 = go on and launch, 1 = pipeline was only built (return {}),
// -1 = cached shapes disagree.
static int pipeline_gate(ComputePipeline* pipeline, bool checkShapes, const TensorBuffer* A, const TensorBuffer* B = nullptr,
                         const TensorBuffer* C = nullptr) {
    if (!pipeline) return 0;
    if (!pipeline->is_valid()) {
        store_pipeline_validation(*pipeline, A, B, C);
        if (pipeline->buildPipelineFlag) return 1;
        return 0;
    }
    if (checkShapes && !validate_pipeline(*pipeline, A, B, C)) return -1;
    return 0;
}
// With an encoder or pass the launch joins the caller's submission: it is enqueued on the stream now.
// Without either, the reference hands back a finished command buffer for the caller to submit
// (th.cpp:836-856); here that is a deferred launch executed by queue_submit().
int64_t g_launch_count = 0;   // kernels launched through the op surface (bench's gpu_launches)
// Test hook: while non-null, every command issued through the op surface appends its label (cmdbuf_* name, "copy" for
// the buffer-to-buffer copies of build_layer_cmdbuf).  tests/test_gpu_graph_trace.py compares the sequence with the
// command stream the reference's own th_eval_gpu encodes (tests/golden/graph_trace_tiny.json).
std::vector<std::string>* g_op_trace = nullptr;
void trace_command(const char* label) { if (g_op_trace) g_op_trace->push_back(label); }

template <typename F>
static CommandBuffer emit(WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, const char* op, F launch) {
    trace_command(op);
    if (!encoder && !pass) return CommandBuffer(std::function<int()>(launch), op);
    const int rc = launch();
    if (rc != THK_OK) { fprintf(stderr, "%s: %s\n", op, thk_last_error()); return {}; }
    ++g_launch_count;
    return CommandBuffer(true);
}
bool queue_submit(WGPUQueue, CommandBuffer& cb) {
    if (!cb.is_valid()) return false;
    if (cb.deferred) {
        const int rc = cb.deferred();
        cb.deferred = nullptr;
        if (rc == THK_OK) ++g_launch_count;
        if (rc != THK_OK) { fprintf(stderr, "%s: %s\n", cb.label, thk_last_error()); cb.cmdBuffer = false; return false; }
    }
    return true;
}
static bool have_gpu(const TensorBuffer& t, const char* op, const char* which) {
    if (t.gpu) return true;
    fprintf(stderr, "%s: tensor %s has no device buffer\n", op, which);
    return false;
}
static int64_t dim1(int64_t v) { return v == 0 ? 1 : v; }

// ---------------------------------------------------------------------------------------------
// ops
// ---------------------------------------------------------------------------------------------
// th.cpp:750-861, validation 541-608
CommandBuffer cmdbuf_mat_mul(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                             const TensorBuffer& A, const TensorBuffer& B, const TensorBuffer& C, int transposeB, WGPUBuffer uniforms) {
    if (A.type != TensorType_F32) TH_FAIL("cmdbuf_mat_mul: A.type != TensorType_F32");
    if (!(B.type == TensorType_F32 || B.type == TensorType_F16)) TH_FAIL("cmdbuf_mat_mul: B.type != TensorType_F32 or TensorType_F16");
    if (C.type != TensorType_F32) TH_FAIL("cmdbuf_mat_mul: C.type != TensorType_F32");
    if (A.shape.c != B.shape.r) TH_FAIL("cmdbuf_mat_mul: Num rows in A not equal to columns in B");
    if (A.shape.r == 0 || A.shape.c == 0) TH_FAIL("cmdbuf_mat_mul: one of the dimensions of A is zero.");
    if (B.shape.r == 0 || B.shape.c == 0) TH_FAIL("cmdbuf_mat_mul: one of the dimensions of B is zero.");
    if (A.shape.b > 1) {
        if (A.shape.b != B.shape.b) TH_FAIL("cmdbuf_mat_mul: A.shape.b != B.shape.b");
        if (A.shape.b != C.shape.b) TH_FAIL("cmdbuf_mat_mul: A.shape.b != C.shape.b");
    }
    if (C.shape.r != A.shape.r) TH_FAIL("cmdbuf_mat_mul: C's number of rows do not match A's");
    if (C.shape.c != B.shape.c) TH_FAIL("cmdbuf_mat_mul: C's number of columns do not match B's");
    const bool useUniforms = uniforms != nullptr;
    const int g = pipeline_gate(pipeline, !useUniforms, &A, &B, &C);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_mat_mul: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_mat_mul", "A") || !have_gpu(B, "cmdbuf_mat_mul", "B") || !have_gpu(C, "cmdbuf_mat_mul", "C")) return {};
    const float* a = (const float*)A.gpu; const void* b = B.gpu; float* c = (float*)C.gpu;
    const int64_t batch = dim1(A.shape.b), M = A.shape.r, K = A.shape.c, N = B.shape.c;
    const int f16 = B.type == TensorType_F16;
    // batched-prompt weight matmul (th-llama.cpp:308-310,404,429-430,444): dense contraction -> tensor cores
    const bool tc = f16 && transposeB && batch == 1 && !useUniforms && M >= 8 && N % 32 == 0 && K % 64 == 0;
    return emit(encoder, pass, "cmdbuf_mat_mul", [=]() {
        if (tc) return thk_gemm_f16_tc(device, a, (const uint16_t*)b, c, M, N, K);
        return thk_mat_mul(device, a, b, c, batch, M, K, N, transposeB, f16, (const thk_dims_uniforms*)uniforms); });
}

// th.cpp:1036-1151, validation 914-937
CommandBuffer cmdbuf_transpose(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                               const TensorBuffer& A, const TensorBuffer& B, bool zy, WGPUBuffer dimBuffer) {
    if (zy) {
        if (A.shape.r != B.shape.b || A.shape.c != B.shape.c || A.shape.b != B.shape.r)
            TH_FAIL("cmdbuf_transpose: zy: input shape doesn't match expected transpose of output.");
    } else {
        if (A.shape.r != B.shape.c || A.shape.c != B.shape.r || A.shape.b != B.shape.b)
            TH_FAIL("cmdbuf_transpose: yx: input shape doesn't match expected transpose of output.");
    }
    if (A.type != TensorType_F32 || B.type != TensorType_F32) TH_FAIL("cmdbuf_transpose: tensors must be f32");
    const bool useUniforms = dimBuffer != nullptr;
    const int g = pipeline_gate(pipeline, !useUniforms, &A, &B);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_transpose: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_transpose", "A") || !have_gpu(B, "cmdbuf_transpose", "B")) return {};
    const float* in = (const float*)A.gpu; float* out = (float*)B.gpu;
    const int64_t Bn = dim1(A.shape.b), M = dim1(A.shape.r), N = A.shape.c;
    const int izy = zy ? 1 : 0;
    return emit(encoder, pass, "cmdbuf_transpose", [=]() {
        return thk_transpose(device, in, out, Bn, M, N, izy, (const thk_dims_uniforms*)dimBuffer); });
}

// th.cpp:1229-1296
CommandBuffer cmdbuf_rms_norm(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                              const TensorBuffer& A) {
    if (A.type != TensorType_F32) TH_FAIL("cmdbuf_rms_norm: A.type != TensorType_F32");
    if (A.shape.c == 0) TH_FAIL("cmdbuf_rms_norm: zero columns");
    const int g = pipeline_gate(pipeline, true, &A);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("create_rms_norm: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_rms_norm", "A")) return {};
    float* x = (float*)A.gpu; const int64_t rows = dim1(A.shape.b) * dim1(A.shape.r), N = A.shape.c;
    return emit(encoder, pass, "cmdbuf_rms_norm", [=]() { return thk_rms_norm(device, x, rows, N); });
}

// th.cpp:1368-1449, validation 1317-1327
CommandBuffer cmdbuf_row_element_multiply(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                          ComputePipeline* pipeline, const TensorBuffer& A, const TensorBuffer& B) {
    if (B.shape.r != 1) TH_FAIL("create_row_element_multiply: Expected B to have 1 row, got %lld", (long long)B.shape.r);
    if (A.shape.c != B.shape.c) TH_FAIL("create_row_element_multiply: column counts differ (%lld vs %lld)", (long long)A.shape.c, (long long)B.shape.c);
    if (A.type != TensorType_F32 || B.type != TensorType_F32) TH_FAIL("create_row_element_multiply: tensors must be f32");
    const int g = pipeline_gate(pipeline, true, &A, &B);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("create_row_element_multiply: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_row_element_multiply", "A") || !have_gpu(B, "cmdbuf_row_element_multiply", "B")) return {};
    float* x = (float*)A.gpu; const float* gn = (const float*)B.gpu;
    const int64_t rows = dim1(A.shape.b) * dim1(A.shape.r), N = A.shape.c;
    return emit(encoder, pass, "cmdbuf_row_element_multiply", [=]() { return thk_row_element_multiply(device, x, gn, rows, N); });
}

// th.cpp:1536-1616.  A viewed [b = tokens][r = heads][c = head_dim]
CommandBuffer cmdbuf_RoPE(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                          const TensorBuffer& A, WGPUBuffer networkUniforms) {
    if (A.type != TensorType_F32) TH_FAIL("cmdbuf_RoPE: A.type != TensorType_F32");
    if (!networkUniforms) TH_FAIL("cmdbuf_RoPE: networkUniforms is null");
    if (A.shape.c % 2 != 0 || A.shape.c == 0) TH_FAIL("cmdbuf_RoPE: head dimension must be even");
    const int g = pipeline_gate(pipeline, true, &A);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_RoPE: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_RoPE", "A")) return {};
    float* x = (float*)A.gpu; const int64_t T = dim1(A.shape.b), H = dim1(A.shape.r), D = A.shape.c;
    return emit(encoder, pass, "cmdbuf_RoPE", [=]() { return thk_rope(device, x, T, H, D, (const thk_network_uniforms*)networkUniforms); });
}

static CommandBuffer softmax_common(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                                    const TensorBuffer& A, WGPUBuffer dimBuffer, bool masked) {
    const char* op = masked ? "cmdbuf_masked_softmax" : "cmdbuf_row_softmax";
    if (A.type != TensorType_F32) TH_FAIL("%s: tensor must be f32", op);
    const bool useUniforms = dimBuffer != nullptr;
    const int g = pipeline_gate(pipeline, !useUniforms, &A);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("%s: Pipeline validation failed.", op);
    if (!have_gpu(A, op, "A")) return {};
    float* a = (float*)A.gpu; const int64_t batch = dim1(A.shape.b), M = dim1(A.shape.r), N = A.shape.c;
    return emit(encoder, pass, op, [=]() {
        return masked ? thk_masked_softmax(device, a, batch, M, N, (const thk_dims_uniforms*)dimBuffer)
                      : thk_row_softmax(device, a, batch, M, N, (const thk_dims_uniforms*)dimBuffer); });
}
// th.cpp:1774-1863 / 2034-2119
CommandBuffer cmdbuf_masked_softmax(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                                    const TensorBuffer& A, WGPUBuffer dimBuffer) { return softmax_common(device, encoder, pass, pipeline, A, dimBuffer, true); }
CommandBuffer cmdbuf_row_softmax(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                                 const TensorBuffer& A, WGPUBuffer dimBuffer) { return softmax_common(device, encoder, pass, pipeline, A, dimBuffer, false); }

// th.cpp:2235-2335, validation 2154-2171
CommandBuffer cmdbuf_addition(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                              const TensorBuffer& a, const TensorBuffer& b, const TensorBuffer& c) {
    if (a.shape != b.shape) TH_FAIL("create_addition: a.shape not equal to b.shape");
    if (a.shape != c.shape) TH_FAIL("create_addition: a.shape not equal to c.shape");
    if (a.type != TensorType_F32 || b.type != TensorType_F32 || c.type != TensorType_F32) TH_FAIL("create_addition: tensors must be f32");
    const int g = pipeline_gate(pipeline, true, &a, &b, &c);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("create_addition: Pipeline validation failed.");
    if (!have_gpu(a, "cmdbuf_addition", "a") || !have_gpu(b, "cmdbuf_addition", "b") || !have_gpu(c, "cmdbuf_addition", "c")) return {};
    const float* pa = (const float*)a.gpu; const float* pb = (const float*)b.gpu; float* pc = (float*)c.gpu;
    const int64_t n = a.shape.get_total_num_elements();
    return emit(encoder, pass, "cmdbuf_addition", [=]() { return thk_addition(device, pa, pb, pc, n); });
}

// th.cpp:2589-2677, validation 2532-2542
CommandBuffer cmdbuf_element_mult_in_place(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                           ComputePipeline* pipeline, const TensorBuffer& a, const TensorBuffer& b) {
    if (a.shape != b.shape) TH_FAIL("cmdbuf_element_mult_in_place: a.shape not equal to b.shape");
    if (a.type != TensorType_F32 || b.type != TensorType_F32) TH_FAIL("cmdbuf_element_mult_in_place: tensors must be f32");
    const int g = pipeline_gate(pipeline, true, &a, &b);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_element_mult_in_place: Pipeline validation failed.");
    if (!have_gpu(a, "cmdbuf_element_mult_in_place", "a") || !have_gpu(b, "cmdbuf_element_mult_in_place", "b")) return {};
    float* pa = (float*)a.gpu; const float* pb = (const float*)b.gpu; const int64_t n = a.shape.get_total_num_elements();
    return emit(encoder, pass, "cmdbuf_element_mult_in_place", [=]() { return thk_element_mult_in_place(device, pa, pb, n); });
}

// th.cpp:2754-2837
CommandBuffer cmdbuf_silu(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                          const TensorBuffer& a) {
    if (a.type != TensorType_F32) TH_FAIL("cmdbuf_silu: tensor must be f32");
    const int g = pipeline_gate(pipeline, true, &a);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_silu: Pipeline validation failed.");
    if (!have_gpu(a, "cmdbuf_silu", "a")) return {};
    float* pa = (float*)a.gpu; const int64_t n = a.shape.get_total_num_elements();
    return emit(encoder, pass, "cmdbuf_silu", [=]() { return thk_silu(device, pa, n); });
}

static bool validate_vector_mat_mul_trans(const TensorBuffer& A, const TensorBuffer& B, const TensorBuffer& C, int64_t bCols) {
    if (A.type != TensorType_F32) { fprintf(stderr, "create_vector_mat_mul_trans: A.type != TensorType_F32\n"); return false; }
    if (!(B.type == TensorType_F32 || B.type == TensorType_F16)) { fprintf(stderr, "create_vector_mat_mul_trans: B.type != TensorType_F32 or TensorType_F16\n"); return false; }
    if (C.type != TensorType_F32) { fprintf(stderr, "create_vector_mat_mul_trans: C.type != TensorType_F32\n"); return false; }
    if (A.shape.c != bCols) { fprintf(stderr, "create_vector_mat_mul_trans: Expecting number of columns in A to match B\n"); return false; }
    if (B.shape.r == 0 || B.shape.c == 0) { fprintf(stderr, "create_vector_mat_mul_trans: one of the dimensions of B is zero.\n"); return false; }
    if (A.shape.b > 1 && (A.shape.b != B.shape.b || A.shape.b != C.shape.b)) { fprintf(stderr, "create_vector_mat_mul_trans: batch dimensions differ\n"); return false; }
    if (C.shape.c != B.shape.r) { fprintf(stderr, "C's number of columns do not match B's\n"); return false; }
    return true;
}

// th.cpp:3046-3139, validation 2894-2949.  The reference additionally needs C >= 256 and C % 256 == 0
// (th.cpp:2996-3006, a workgroup artefact); here rows only need 16-byte alignment (C % 8 for f16).
CommandBuffer cmdbuf_vector_mat_mul_trans(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                          ComputePipeline* pipeline, const TensorBuffer& A, const TensorBuffer& B, const TensorBuffer& C,
                                          int64_t aOffset) {
    if (!validate_vector_mat_mul_trans(A, B, C, B.shape.c)) return {};
    const int g = pipeline_gate(pipeline, true, &A, &B, &C);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("Pipeline shapes not equal to input shapes.");
    if (!have_gpu(A, "cmdbuf_vector_mat_mul_trans", "A") || !have_gpu(B, "cmdbuf_vector_mat_mul_trans", "B") ||
        !have_gpu(C, "cmdbuf_vector_mat_mul_trans", "C")) return {};
    const float* a = (const float*)A.gpu; const void* b = B.gpu; float* c = (float*)C.gpu;
    const int64_t R = B.shape.r, Cc = B.shape.c, batch = dim1(A.shape.b);
    const int f16 = B.type == TensorType_F16;
    return emit(encoder, pass, "cmdbuf_vector_mat_mul_trans", [=]() {
        return thk_vector_mat_mul_trans(device, a, (size_t)aOffset, b, c, R, Cc, batch, f16); });
}

// th.cpp:3795-3912 (validation 3588-3667).  splitBuffers (per-split uniform blocks) are accepted for
// signature compatibility and ignored: the split index is a launch argument here.
CommandBuffer cmdbuf_vector_multi_mat_mul_split_trans(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                                      ComputePipeline* pipeline, const TensorBuffer& A, const std::vector<TensorBuffer*> B,
                                                      const TensorBuffer& C, const std::vector<TensorBuffer*>& scratchBuffers,
                                                      int64_t aOffset, const std::vector<WGPUBuffer>&, bool) {
    if (B.empty() || !B[0]) TH_FAIL("cmdbuf_vector_multi_mat_mul_split_trans: no B buffers");
    if (scratchBuffers.size() + 1 < B.size()) TH_FAIL("cmdbuf_vector_multi_mat_mul_split_trans: need %zu scratch buffers", B.size() - 1);
    for (auto* b : B) {
        if (!b || b->shape != B[0]->shape || b->type != B[0]->type) TH_FAIL("cmdbuf_vector_multi_mat_mul_split_trans: split buffers differ in shape/type");
        if (!have_gpu(*b, "cmdbuf_vector_multi_mat_mul_split_trans", "B[i]")) return {};
    }
    const int64_t Ctot = B[0]->shape.c * (int64_t)B.size();
    if (!validate_vector_mat_mul_trans(A, *B[0], C, Ctot)) return {};
    const int g = pipeline_gate(pipeline, true, &A, B[0], &C);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_vector_multi_mat_mul_split_trans: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_vector_multi_mat_mul_split_trans", "A") || !have_gpu(C, "cmdbuf_vector_multi_mat_mul_split_trans", "C")) return {};
    std::vector<const void*> ptrs;
    for (auto* b : B) ptrs.push_back(b->gpu);
    float* scratch = scratchBuffers.empty() ? nullptr : (float*)scratchBuffers[0]->gpu;
    const float* a = (const float*)A.gpu; float* c = (float*)C.gpu;
    const int64_t R = B[0]->shape.r; const int f16 = B[0]->type == TensorType_F16;
    return emit(encoder, pass, "cmdbuf_vector_multi_mat_mul_split_trans", [=]() {
        return thk_vector_multi_mat_mul_split_trans(device, a, (size_t)aOffset, ptrs.data(), (int)ptrs.size(), c, scratch, R, Ctot, f16); });
}

// th.cpp:4042-4127 (validation 3947-3957).  numSplits only shaped the reference's dispatch (and its
// coverage bug); the sum here covers every element.
CommandBuffer cmdbuf_vector_reduce(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                                   const TensorBuffer& A, const TensorBuffer& B, int) {
    if (A.shape != B.shape) TH_FAIL("cmdbuf_vector_reduce: shapes differ");
    if (A.type != TensorType_F32 || B.type != TensorType_F32) TH_FAIL("cmdbuf_vector_reduce: tensors must be f32");
    const int g = pipeline_gate(pipeline, true, &A, &B);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_vector_reduce: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_vector_reduce", "A") || !have_gpu(B, "cmdbuf_vector_reduce", "B")) return {};
    float* a = (float*)A.gpu; const float* b = (const float*)B.gpu; const int64_t n = A.shape.get_total_num_elements();
    return emit(encoder, pass, "cmdbuf_vector_reduce", [=]() { return thk_vector_reduce(device, a, b, n); });
}

// th.cpp:4264-4351 (validation 4167-4195): one f16 row of B -> f32 row of A; offsets in bytes
CommandBuffer cmdbuf_f16_f32_conversion(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                                        const TensorBuffer& A, const TensorBuffer& B, int, const int aOffset, const int bOffset) {
    if (A.type != TensorType_F32) TH_FAIL("cmdbuf_f16_f32_conversion: A.type != TensorType_F32");
    if (B.type != TensorType_F16) TH_FAIL("cmdbuf_f16_f32_conversion: B.type != TensorType_F16");
    if (A.shape.c != B.shape.c) TH_FAIL("cmdbuf_f16_f32_conversion: column counts differ");
    const int g = pipeline_gate(pipeline, true, &A, &B);
    if (g == 1) return {};
    if (g < 0) TH_FAIL("cmdbuf_f16_f32_conversion: Pipeline validation failed.");
    if (!have_gpu(A, "cmdbuf_f16_f32_conversion", "A") || !have_gpu(B, "cmdbuf_f16_f32_conversion", "B")) return {};
    float* out = (float*)A.gpu; const uint16_t* in = (const uint16_t*)B.gpu; const int64_t n = A.shape.c;
    return emit(encoder, pass, "cmdbuf_f16_f32_conversion", [=]() {
        return thk_f16_f32_conversion(device, out, (size_t)aOffset, in, (size_t)bOffset, n); });
}

}  // namespace th
