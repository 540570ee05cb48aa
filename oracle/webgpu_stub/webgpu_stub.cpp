// Host-memory implementation of the stub in webgpu/webgpu.h.  TEST INFRASTRUCTURE ONLY.
// Buffers are calloc'd host blocks; copies are memcpy executed immediately; dispatches do nothing
// (WGSL cannot execute here) but are COUNTED and, while a trace is open (thstub_trace_begin),
// RECORDED: pipeline label, workgroup counts, the bound buffers (id, offset, size) and the contents
// of every bound buffer small enough to be a uniform block.  Buffer-to-buffer copies are recorded
// too.  That is the command stream the reference's own th_eval_gpu encodes
// (/root/reference/th-llama.cpp:464-660), which tests/test_graph_trace.py pins the oracle's and the
// product's op order against.
#include <webgpu/webgpu.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct WGPUBufferImpl { void* data; uint64_t size; uint32_t id; };
struct WGPUDeviceImpl { int unused; };
struct WGPUQueueImpl { int unused; };
struct WGPUCommandEncoderImpl { int unused; };
struct WGPUCommandBufferImpl { int unused; };
struct WGPUShaderModuleImpl { int unused; };
struct WGPUBindGroupLayoutImpl { int unused; };
struct WGPUPipelineLayoutImpl { int unused; };
struct WGPUComputePipelineImpl { std::string label; };
struct BoundBuffer { uint32_t binding; uint32_t id; uint64_t offset, size; WGPUBuffer buf; };
struct WGPUBindGroupImpl { std::vector<BoundBuffer> entries; };
struct WGPUComputePassEncoderImpl { std::string label; std::vector<BoundBuffer> bound; };

static WGPUDeviceImpl g_device; static WGPUQueueImpl g_queue;
static WGPUCommandEncoderImpl g_enc; static WGPUComputePassEncoderImpl g_pass;
static WGPUCommandBufferImpl g_cb; static WGPUShaderModuleImpl g_sm;
static WGPUBindGroupLayoutImpl g_bgl; static WGPUPipelineLayoutImpl g_pl;
static uint64_t g_dispatches = 0;
static uint32_t g_next_buffer_id = 1;
static bool g_tracing = false;
static std::string g_trace;          // one JSON object per line
static const uint64_t kUniformMax = 256;

static void append_bound(std::string& o, const std::vector<BoundBuffer>& b) {
    char t[160];
    o += "\"binds\": [";
    for (size_t i = 0; i < b.size(); ++i) {
        snprintf(t, sizeof t, "%s[%u, %u, %llu, %llu]", i ? ", " : "", b[i].binding, b[i].id, (unsigned long long)b[i].offset,
                 (unsigned long long)b[i].size);
        o += t;
    }
    o += "], \"uniforms\": {";
    bool first = true;
    for (const BoundBuffer& e : b) {
        if (!e.buf || e.buf->size > kUniformMax) continue;
        snprintf(t, sizeof t, "%s\"%u\": [", first ? "" : ", ", e.binding);
        o += t;
        first = false;
        const uint32_t* w = (const uint32_t*)((const char*)e.buf->data + e.offset);
        const uint64_t n = (e.buf->size - e.offset) / 4;
        for (uint64_t i = 0; i < n; ++i) { snprintf(t, sizeof t, "%s%u", i ? ", " : "", w[i]); o += t; }
        o += "]";
    }
    o += "}";
}

extern "C" {
WGPUDevice thstub_device(void) { return &g_device; }
WGPUQueue thstub_queue(void) { return &g_queue; }
void* thstub_buffer_data(WGPUBuffer b) { return b ? b->data : nullptr; }
uint64_t thstub_buffer_size(WGPUBuffer b) { return b ? b->size : 0; }
uint32_t thstub_buffer_id(WGPUBuffer b) { return b ? b->id : 0; }
uint64_t thstub_dispatch_count(void) { return g_dispatches; }
void thstub_trace_begin(void) { g_trace.clear(); g_tracing = true; }
const char* thstub_trace_end(void) { g_tracing = false; return g_trace.c_str(); }

WGPUBuffer wgpuDeviceCreateBuffer(WGPUDevice, WGPUBufferDescriptor const* d) {
    WGPUBufferImpl* b = new WGPUBufferImpl;
    b->size = d->size; b->data = std::calloc(1, d->size ? d->size : 1); b->id = g_next_buffer_id++;
    return b;
}
void wgpuBufferRelease(WGPUBuffer b) { if (b) { std::free(b->data); delete b; } }
void wgpuQueueWriteBuffer(WGPUQueue, WGPUBuffer b, uint64_t off, void const* data, size_t size) {
    if (b && off + size <= b->size) std::memcpy((char*)b->data + off, data, size);
    if (g_tracing && b) {
        char t[160];
        snprintf(t, sizeof t, "{\"kind\": \"write\", \"dst\": %u, \"dst_off\": %llu, \"size\": %llu}\n", b->id, (unsigned long long)off,
                 (unsigned long long)size);
        g_trace += t;
    }
}
void wgpuCommandEncoderCopyBufferToBuffer(WGPUCommandEncoder, WGPUBuffer s, uint64_t so, WGPUBuffer d, uint64_t doff, uint64_t size) {
    if (s && d && so + size <= s->size && doff + size <= d->size) std::memmove((char*)d->data + doff, (char*)s->data + so, size);
    if (g_tracing) {
        char t[200];
        snprintf(t, sizeof t, "{\"kind\": \"copy\", \"src\": %u, \"src_off\": %llu, \"dst\": %u, \"dst_off\": %llu, \"size\": %llu}\n",
                 s ? s->id : 0, (unsigned long long)so, d ? d->id : 0, (unsigned long long)doff, (unsigned long long)size);
        g_trace += t;
    }
}
WGPUShaderModule wgpuDeviceCreateShaderModule(WGPUDevice, WGPUShaderModuleDescriptor const*) { return &g_sm; }
void wgpuShaderModuleRelease(WGPUShaderModule) {}
WGPUBindGroupLayout wgpuDeviceCreateBindGroupLayout(WGPUDevice, WGPUBindGroupLayoutDescriptor const*) { return &g_bgl; }
void wgpuBindGroupLayoutReference(WGPUBindGroupLayout) {}
WGPUPipelineLayout wgpuDeviceCreatePipelineLayout(WGPUDevice, WGPUPipelineLayoutDescriptor const*) { return &g_pl; }
void wgpuPipelineLayoutRelease(WGPUPipelineLayout) {}
// Pipelines and bind groups are never freed: the reference keeps handles in caches and releases others right after
// encoding; a few hundred small objects per evaluated token in a test process do not matter.
WGPUComputePipeline wgpuDeviceCreateComputePipeline(WGPUDevice, WGPUComputePipelineDescriptor const* d) {
    WGPUComputePipelineImpl* p = new WGPUComputePipelineImpl;
    p->label = (d && d->label) ? d->label : "";
    return p;
}
void wgpuComputePipelineRelease(WGPUComputePipeline) {}
WGPUBindGroup wgpuDeviceCreateBindGroup(WGPUDevice, WGPUBindGroupDescriptor const* d) {
    WGPUBindGroupImpl* g = new WGPUBindGroupImpl;
    for (size_t i = 0; d && i < d->entryCount; ++i) {
        const WGPUBindGroupEntry& e = d->entries[i];
        g->entries.push_back(BoundBuffer{e.binding, e.buffer ? e.buffer->id : 0u, e.offset, e.size, e.buffer});
    }
    return g;
}
void wgpuBindGroupRelease(WGPUBindGroup) {}
WGPUCommandEncoder wgpuDeviceCreateCommandEncoder(WGPUDevice, WGPUCommandEncoderDescriptor const*) { return &g_enc; }
WGPUComputePassEncoder wgpuCommandEncoderBeginComputePass(WGPUCommandEncoder, WGPUComputePassDescriptor const*) {
    g_pass.label.clear(); g_pass.bound.clear();
    return &g_pass;
}
void wgpuComputePassEncoderSetPipeline(WGPUComputePassEncoder p, WGPUComputePipeline pl) { if (p && pl) p->label = pl->label; }
void wgpuComputePassEncoderSetBindGroup(WGPUComputePassEncoder p, uint32_t, WGPUBindGroup g, size_t, uint32_t const*) {
    if (p && g) p->bound = g->entries;
}
void wgpuComputePassEncoderDispatchWorkgroups(WGPUComputePassEncoder p, uint32_t x, uint32_t y, uint32_t z) {
    ++g_dispatches;
    if (g_tracing && p) {
        char t[160];
        snprintf(t, sizeof t, "{\"kind\": \"dispatch\", \"label\": \"%s\", \"wg\": [%u, %u, %u], ", p->label.c_str(), x, y, z);
        g_trace += t;
        append_bound(g_trace, p->bound);
        g_trace += "}\n";
    }
}
void wgpuComputePassEncoderEnd(WGPUComputePassEncoder) {}
void wgpuComputePassEncoderRelease(WGPUComputePassEncoder) {}
WGPUCommandBuffer wgpuCommandEncoderFinish(WGPUCommandEncoder, WGPUCommandBufferDescriptor const*) { return &g_cb; }
void wgpuCommandEncoderRelease(WGPUCommandEncoder) {}
void wgpuCommandBufferRelease(WGPUCommandBuffer) {}
void wgpuQueueSubmit(WGPUQueue, size_t, WGPUCommandBuffer const*) {
    if (g_tracing) g_trace += "{\"kind\": \"submit\"}\n";
}
void wgpuBufferMapAsync(WGPUBuffer, WGPUMapModeFlags, size_t, size_t, WGPUBufferMapCallback cb, void* ud) { if (cb) cb(WGPUBufferMapAsyncStatus_Success, ud); }
void const* wgpuBufferGetConstMappedRange(WGPUBuffer b, size_t off, size_t) { return b ? (char*)b->data + off : nullptr; }
void wgpuBufferUnmap(WGPUBuffer) {}
void wgpuDeviceTick(WGPUDevice) {}
}
