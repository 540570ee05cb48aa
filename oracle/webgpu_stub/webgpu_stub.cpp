// Host-memory implementation of the stub in webgpu/webgpu.h.  TEST INFRASTRUCTURE ONLY.
// Buffers are calloc'd host blocks; copies are memcpy executed immediately; dispatches are
// counted and otherwise ignored (WGSL cannot execute here).
#include <webgpu/webgpu.h>
#include <cstdlib>
#include <cstring>

struct WGPUBufferImpl { void* data; uint64_t size; };
struct WGPUDeviceImpl { int unused; };
struct WGPUQueueImpl { int unused; };
struct WGPUCommandEncoderImpl { int unused; };
struct WGPUComputePassEncoderImpl { int unused; };
struct WGPUCommandBufferImpl { int unused; };
struct WGPUShaderModuleImpl { int unused; };
struct WGPUBindGroupLayoutImpl { int unused; };
struct WGPUPipelineLayoutImpl { int unused; };
struct WGPUComputePipelineImpl { int unused; };
struct WGPUBindGroupImpl { int unused; };

static WGPUDeviceImpl g_device; static WGPUQueueImpl g_queue;
static WGPUCommandEncoderImpl g_enc; static WGPUComputePassEncoderImpl g_pass;
static WGPUCommandBufferImpl g_cb; static WGPUShaderModuleImpl g_sm;
static WGPUBindGroupLayoutImpl g_bgl; static WGPUPipelineLayoutImpl g_pl;
static WGPUComputePipelineImpl g_cp; static WGPUBindGroupImpl g_bg;
static uint64_t g_dispatches = 0;

extern "C" {
WGPUDevice thstub_device(void) { return &g_device; }
WGPUQueue thstub_queue(void) { return &g_queue; }
void* thstub_buffer_data(WGPUBuffer b) { return b ? b->data : nullptr; }
uint64_t thstub_buffer_size(WGPUBuffer b) { return b ? b->size : 0; }
uint64_t thstub_dispatch_count(void) { return g_dispatches; }

WGPUBuffer wgpuDeviceCreateBuffer(WGPUDevice, WGPUBufferDescriptor const* d) {
    WGPUBufferImpl* b = new WGPUBufferImpl;
    b->size = d->size; b->data = std::calloc(1, d->size ? d->size : 1);
    return b;
}
void wgpuBufferRelease(WGPUBuffer b) { if (b) { std::free(b->data); delete b; } }
void wgpuQueueWriteBuffer(WGPUQueue, WGPUBuffer b, uint64_t off, void const* data, size_t size) {
    if (b && off + size <= b->size) std::memcpy((char*)b->data + off, data, size);
}
void wgpuCommandEncoderCopyBufferToBuffer(WGPUCommandEncoder, WGPUBuffer s, uint64_t so, WGPUBuffer d, uint64_t doff, uint64_t size) {
    if (s && d && so + size <= s->size && doff + size <= d->size) std::memmove((char*)d->data + doff, (char*)s->data + so, size);
}
WGPUShaderModule wgpuDeviceCreateShaderModule(WGPUDevice, WGPUShaderModuleDescriptor const*) { return &g_sm; }
void wgpuShaderModuleRelease(WGPUShaderModule) {}
WGPUBindGroupLayout wgpuDeviceCreateBindGroupLayout(WGPUDevice, WGPUBindGroupLayoutDescriptor const*) { return &g_bgl; }
void wgpuBindGroupLayoutReference(WGPUBindGroupLayout) {}
WGPUPipelineLayout wgpuDeviceCreatePipelineLayout(WGPUDevice, WGPUPipelineLayoutDescriptor const*) { return &g_pl; }
void wgpuPipelineLayoutRelease(WGPUPipelineLayout) {}
WGPUComputePipeline wgpuDeviceCreateComputePipeline(WGPUDevice, WGPUComputePipelineDescriptor const*) { return &g_cp; }
void wgpuComputePipelineRelease(WGPUComputePipeline) {}
WGPUBindGroup wgpuDeviceCreateBindGroup(WGPUDevice, WGPUBindGroupDescriptor const*) { return &g_bg; }
void wgpuBindGroupRelease(WGPUBindGroup) {}
WGPUCommandEncoder wgpuDeviceCreateCommandEncoder(WGPUDevice, WGPUCommandEncoderDescriptor const*) { return &g_enc; }
WGPUComputePassEncoder wgpuCommandEncoderBeginComputePass(WGPUCommandEncoder, WGPUComputePassDescriptor const*) { return &g_pass; }
void wgpuComputePassEncoderSetPipeline(WGPUComputePassEncoder, WGPUComputePipeline) {}
void wgpuComputePassEncoderSetBindGroup(WGPUComputePassEncoder, uint32_t, WGPUBindGroup, size_t, uint32_t const*) {}
void wgpuComputePassEncoderDispatchWorkgroups(WGPUComputePassEncoder, uint32_t, uint32_t, uint32_t) { ++g_dispatches; }
void wgpuComputePassEncoderEnd(WGPUComputePassEncoder) {}
void wgpuComputePassEncoderRelease(WGPUComputePassEncoder) {}
WGPUCommandBuffer wgpuCommandEncoderFinish(WGPUCommandEncoder, WGPUCommandBufferDescriptor const*) { return &g_cb; }
void wgpuCommandEncoderRelease(WGPUCommandEncoder) {}
void wgpuCommandBufferRelease(WGPUCommandBuffer) {}
void wgpuQueueSubmit(WGPUQueue, size_t, WGPUCommandBuffer const*) {}
void wgpuBufferMapAsync(WGPUBuffer, WGPUMapModeFlags, size_t, size_t, WGPUBufferMapCallback cb, void* ud) { if (cb) cb(WGPUBufferMapAsyncStatus_Success, ud); }
void const* wgpuBufferGetConstMappedRange(WGPUBuffer b, size_t off, size_t) { return b ? (char*)b->data + off : nullptr; }
void wgpuBufferUnmap(WGPUBuffer) {}
void wgpuDeviceTick(WGPUDevice) {}
}
