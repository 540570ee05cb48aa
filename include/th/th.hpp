// th.hpp -- the TokenHawk op-dispatch surface (tensor + kernel launch API) over CUDA.
//
// Mirrors the reference's th.hpp (kayvr/token-hawk th.hpp:20-452): same type names, members,
// op names, argument order and error behaviour, so th-llama style graph code compiles against it.
// What changed underneath: WGPUDevice/WGPUQueue -> thk_ctx (a CUDA device + stream),
// WGPUBuffer -> device pointer, WGSL pipelines -> precompiled sm_100a kernels reached through the
// C ABI in thk_cabi.h.  There is no CPU fallback.
//
// Recording model.  The reference encodes commands into a WGPUCommandEncoder and submits later.
// Here an op with a non-null `encoder` is enqueued on the context's stream immediately (a CUDA
// stream IS an in-order command queue); with a null encoder the op also enqueues and returns a
// valid CommandBuffer token, where the reference would return a command buffer for the caller to
// submit.  The two-phase pipeline idiom is kept: an op handed a ComputePipeline whose
// buildPipelineFlag is set and which is not yet valid only stores the validation shapes and
// returns without launching (th.cpp:767-775); later calls validate cached shapes (th.cpp:99-124).
#pragma once

#include <assert.h>
#include <stdint.h>
#include <stdio.h>

#include <functional>
#include <string>
#include <vector>

#include "thk_cabi.h"

namespace th {

// ---- handle types standing in for the WebGPU ones (same spelling, CUDA meaning) ----
using WGPUDevice = thk_ctx*;              // device + stream
using WGPUQueue = thk_ctx*;               // the same in-order stream
using WGPUBuffer = void*;                 // device pointer
struct EncoderTag;                        // any non-null value means "enqueue as part of a larger submission"
using WGPUCommandEncoder = EncoderTag*;
using WGPUComputePassEncoder = EncoderTag*;
using WGPUBufferUsageFlags = uint32_t;

enum TensorType { TensorType_Unknown, TensorType_F16, TensorType_F32 };   // th.hpp:20-24

std::string get_TensorType_name(TensorType dt);

inline size_t get_TensorType_size(TensorType dt) {
    if (dt == TensorType_F16) return 2;
    if (dt == TensorType_F32) return 4;
    assert(false);
    return 0;
}

// th.hpp:37-73.  A zero dimension means "absent".
struct TensorShape {
    int64_t l{}, b{}, r{}, c{};

    int64_t get_total_num_elements() const {
        if (l == 0 && b == 0 && r == 0 && c == 0) return 0;
        int64_t n = 1;
        if (l > 0) n *= l;
        if (b > 0) n *= b;
        if (r > 0) n *= r;
        if (c > 0) n *= c;
        return n;
    }
    std::string to_string() const {
        return "L:" + std::to_string(l) + " B:" + std::to_string(b) + " R:" + std::to_string(r) + " C:" + std::to_string(c);
    }
    void print() const { printf("Shape: %s\n", to_string().c_str()); }
    void canonicalize() {
        if (l == 1) l = 0;
        if (b == 1) b = 0;
        if (r == 0) r = 1;
    }
};
inline bool operator==(const TensorShape& a, const TensorShape& b) { return a.l == b.l && a.b == b.b && a.r == b.r && a.c == b.c; }
inline bool operator!=(const TensorShape& a, const TensorShape& b) { return !(a == b); }

// Move-only RAII wrapper of one device allocation (th.hpp:79-148, th.cpp:150-229).
struct TensorBuffer {
    static const WGPUBufferUsageFlags k_default_usage = 0;

    TensorBuffer() = default;
    TensorBuffer(TensorShape shape, TensorType type, WGPUDevice device = nullptr, WGPUBufferUsageFlags usage = k_default_usage);
    TensorBuffer(const void* data, TensorShape shape, TensorType type, bool backup, WGPUDevice device = nullptr,
                 WGPUQueue queue = nullptr, WGPUBufferUsageFlags usage = k_default_usage);
    ~TensorBuffer() { free_buffers(); }
    TensorBuffer(const TensorBuffer&) = delete;
    TensorBuffer& operator=(const TensorBuffer&) = delete;
    TensorBuffer& operator=(TensorBuffer&& other) noexcept;
    TensorBuffer(TensorBuffer&& other) noexcept;

    size_t get_size_bytes() const;
    void allocate_gpu_memory(WGPUDevice device, WGPUBufferUsageFlags usage);
    void upload_data_to_gpu(WGPUQueue queue, const void* data);
    void reset_shape() { shape = originalShape; }
    int64_t get_num_dims() const { return (shape.l != 0) + (shape.b != 0) + (shape.r != 0) + (shape.c != 0); }
    void free_buffers();
    bool is_valid() const { return gpu != nullptr || !cpuBackup.empty(); }

    TensorShape shape{};
    TensorType type = TensorType_Unknown;
    bool cpuOnly = false;
    std::vector<uint8_t> cpuBackup{};
    WGPUBuffer gpu{};
    TensorShape originalShape{};
    std::string name{};
    WGPUDevice owner{};      // context the allocation belongs to (needed to free it)
};

// th.hpp:150-245.  No JIT here: a "pipeline" is the cached shape/type contract of one op site.
struct ComputePipeline {
    ComputePipeline() = default;
    ComputePipeline(bool buildPipelineFlagIn) : buildPipelineFlag(buildPipelineFlagIn) {}
    ComputePipeline(const ComputePipeline&) = delete;
    ComputePipeline& operator=(const ComputePipeline&) = delete;
    ComputePipeline(ComputePipeline&&) noexcept = default;
    ComputePipeline& operator=(ComputePipeline&&) noexcept = default;

    bool is_valid() const { return built; }
    void free_buffers() { built = false; }

    bool built = false;              // stands in for bindGroupLayout/pipeline being non-null
    bool buildPipelineFlag = false;
    TensorShape sa{}; TensorType ta = TensorType_Unknown;
    TensorShape sb{}; TensorType tb = TensorType_Unknown;
    TensorShape sc{}; TensorType tc = TensorType_Unknown;
};

// th.hpp:247-290.  Valid = the op was accepted.  Ops called with an encoder/pass are already
// enqueued; ops called with neither carry a deferred launch that queue_submit() executes
// (dropping the CommandBuffer drops the work, like an unsubmitted WGPUCommandBuffer).
struct CommandBuffer {
    CommandBuffer() = default;
    explicit CommandBuffer(bool enqueued) : cmdBuffer(enqueued) {}
    CommandBuffer(std::function<int()> launch, const char* opLabel) : cmdBuffer(true), deferred(std::move(launch)), label(opLabel) {}
    CommandBuffer(const CommandBuffer&) = delete;
    CommandBuffer& operator=(const CommandBuffer&) = delete;
    CommandBuffer(CommandBuffer&& o) noexcept : cmdBuffer(o.cmdBuffer), deferred(std::move(o.deferred)), label(o.label) { o.cmdBuffer = false; }
    CommandBuffer& operator=(CommandBuffer&& o) noexcept {
        cmdBuffer = o.cmdBuffer; deferred = std::move(o.deferred); label = o.label; o.cmdBuffer = false;
        return *this;
    }
    void free_buffers() { cmdBuffer = false; deferred = nullptr; }
    bool is_valid() const { return cmdBuffer; }
    bool cmdBuffer = false;
    std::function<int()> deferred{};
    const char* label = "";
};
// wgpuQueueSubmit for one CommandBuffer (th-llama.cpp:640): runs a deferred launch on the stream.
bool queue_submit(WGPUQueue queue, CommandBuffer& cb);

void print_TensorBuffer(TensorBuffer* buffer, const char* bufferName);
bool are_pipelines_similar(const ComputePipeline& a, const ComputePipeline& b);
double get_time_seconds();

typedef uint16_t ggml_fp16_t;
ggml_fp16_t ggml_compute_fp32_to_fp16(float f);
float ggml_compute_fp16_to_fp32(ggml_fp16_t h);

// ---- ops (th.hpp:302-452).  `uniforms`/`dimBuffer` are device pointers to the 32-byte blocks. ----
CommandBuffer cmdbuf_mat_mul(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                             ComputePipeline* pipeline, const TensorBuffer& A, const TensorBuffer& B,
                             const TensorBuffer& C, int transposeB = 0, WGPUBuffer uniforms = nullptr);
CommandBuffer cmdbuf_transpose(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                               ComputePipeline* pipeline, const TensorBuffer& inputF32, const TensorBuffer& outputF32,
                               bool zy, WGPUBuffer dimBuffer);
CommandBuffer cmdbuf_rms_norm(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                              ComputePipeline* pipeline, const TensorBuffer& A);
CommandBuffer cmdbuf_row_element_multiply(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                          ComputePipeline* pipeline, const TensorBuffer& inputOutputBuffer,
                                          const TensorBuffer& rowMultBuffer);
CommandBuffer cmdbuf_RoPE(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                          ComputePipeline* pipeline, const TensorBuffer& A, WGPUBuffer networkUniforms);
CommandBuffer cmdbuf_masked_softmax(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                    ComputePipeline* pipeline, const TensorBuffer& inputOutputBuffer, WGPUBuffer dimBuffer);
CommandBuffer cmdbuf_row_softmax(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                 ComputePipeline* pipeline, const TensorBuffer& inputOutputBuffer, WGPUBuffer dimBuffer);
CommandBuffer cmdbuf_addition(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                              ComputePipeline* pipeline, const TensorBuffer& a, const TensorBuffer& b, const TensorBuffer& c);
CommandBuffer cmdbuf_element_mult_in_place(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                           ComputePipeline* pipeline, const TensorBuffer& a, const TensorBuffer& b);
CommandBuffer cmdbuf_silu(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                          ComputePipeline* pipeline, const TensorBuffer& a);
CommandBuffer cmdbuf_vector_mat_mul_trans(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                          ComputePipeline* pipeline, const TensorBuffer& A, const TensorBuffer& B,
                                          const TensorBuffer& C, int64_t aOffset);
CommandBuffer cmdbuf_vector_multi_mat_mul_split_trans(WGPUDevice device, WGPUCommandEncoder encoder,
                                                      WGPUComputePassEncoder pass, ComputePipeline* pipeline,
                                                      const TensorBuffer& A, const std::vector<TensorBuffer*> B,
                                                      const TensorBuffer& C, const std::vector<TensorBuffer*>& scratchBuffers,
                                                      int64_t aOffset, const std::vector<WGPUBuffer>& splitBuffers,
                                                      bool useDimsFromUniforms);
CommandBuffer cmdbuf_vector_reduce(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                   ComputePipeline* pipeline, const TensorBuffer& A, const TensorBuffer& B, int numSplits);
CommandBuffer cmdbuf_f16_f32_conversion(WGPUDevice device, WGPUCommandEncoder encoder, WGPUComputePassEncoder pass,
                                        ComputePipeline* pipeline, const TensorBuffer& A, const TensorBuffer& B,
                                        int numSplits, const int aOffset, const int bOffset);

}  // namespace th
