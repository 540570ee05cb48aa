/*
 * Minimal host-memory stand-in for <webgpu/webgpu.h>, written for this repo (NOT Dawn's header).
 * TEST INFRASTRUCTURE ONLY.  It declares exactly the handles, enums, structs and the 29 entry
 * points the reference's core translation units use so that those files
 * can be compiled, unmodified and from where they lie under /root/reference, into
 * oracle/_ref/libthref_host.so.  Buffers are plain host allocations; compute dispatches do
 * nothing (the WGSL cannot run here), so only the reference's HOST logic becomes executable:
 * fp16 helpers, sampler, tokenizer, the ggjt loader -- and th_eval_gpu's command encoding, which the
 * stub can record (thstub_trace_begin / _end).
 */
#ifndef TH_ORACLE_WEBGPU_STUB_H
#define TH_ORACLE_WEBGPU_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct WGPUDeviceImpl* WGPUDevice;
typedef struct WGPUQueueImpl* WGPUQueue;
typedef struct WGPUBufferImpl* WGPUBuffer;
typedef struct WGPUCommandEncoderImpl* WGPUCommandEncoder;
typedef struct WGPUComputePassEncoderImpl* WGPUComputePassEncoder;
typedef struct WGPUCommandBufferImpl* WGPUCommandBuffer;
typedef struct WGPUShaderModuleImpl* WGPUShaderModule;
typedef struct WGPUBindGroupLayoutImpl* WGPUBindGroupLayout;
typedef struct WGPUPipelineLayoutImpl* WGPUPipelineLayout;
typedef struct WGPUComputePipelineImpl* WGPUComputePipeline;
typedef struct WGPUBindGroupImpl* WGPUBindGroup;

typedef uint32_t WGPUFlags;
typedef WGPUFlags WGPUBufferUsageFlags;
typedef WGPUFlags WGPUShaderStageFlags;
typedef WGPUFlags WGPUMapModeFlags;

enum {
    WGPUBufferUsage_MapRead = 0x1, WGPUBufferUsage_MapWrite = 0x2, WGPUBufferUsage_CopySrc = 0x4,
    WGPUBufferUsage_CopyDst = 0x8, WGPUBufferUsage_Uniform = 0x40, WGPUBufferUsage_Storage = 0x80,
};
enum { WGPUShaderStage_Compute = 0x4 };
enum { WGPUMapMode_Read = 0x1, WGPUMapMode_Write = 0x2 };
typedef enum {
    WGPUBufferBindingType_Undefined = 0, WGPUBufferBindingType_Uniform = 1,
    WGPUBufferBindingType_Storage = 2, WGPUBufferBindingType_ReadOnlyStorage = 3,
} WGPUBufferBindingType;
typedef enum { WGPUBufferMapAsyncStatus_Success = 0, WGPUBufferMapAsyncStatus_Error = 1 } WGPUBufferMapAsyncStatus;
typedef enum { WGPUSType_Invalid = 0, WGPUSType_ShaderModuleWGSLDescriptor = 6 } WGPUSType;

typedef struct WGPUChainedStruct { struct WGPUChainedStruct const* next; WGPUSType sType; } WGPUChainedStruct;

typedef struct { WGPUChainedStruct const* nextInChain; char const* label; WGPUBufferUsageFlags usage; uint64_t size; bool mappedAtCreation; } WGPUBufferDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; WGPUBufferBindingType type; bool hasDynamicOffset; uint64_t minBindingSize; } WGPUBufferBindingLayout;
typedef struct { WGPUChainedStruct const* nextInChain; uint32_t binding; WGPUShaderStageFlags visibility; WGPUBufferBindingLayout buffer; } WGPUBindGroupLayoutEntry;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; size_t entryCount; WGPUBindGroupLayoutEntry const* entries; } WGPUBindGroupLayoutDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; uint32_t binding; WGPUBuffer buffer; uint64_t offset; uint64_t size; } WGPUBindGroupEntry;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; WGPUBindGroupLayout layout; size_t entryCount; WGPUBindGroupEntry const* entries; } WGPUBindGroupDescriptor;
typedef struct { WGPUChainedStruct chain; char const* source; } WGPUShaderModuleWGSLDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; } WGPUShaderModuleDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; size_t bindGroupLayoutCount; WGPUBindGroupLayout const* bindGroupLayouts; } WGPUPipelineLayoutDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; WGPUShaderModule module; char const* entryPoint; size_t constantCount; void const* constants; } WGPUProgrammableStageDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; WGPUPipelineLayout layout; WGPUProgrammableStageDescriptor compute; } WGPUComputePipelineDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; } WGPUCommandEncoderDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; } WGPUComputePassDescriptor;
typedef struct { WGPUChainedStruct const* nextInChain; char const* label; } WGPUCommandBufferDescriptor;

typedef void (*WGPUBufferMapCallback)(WGPUBufferMapAsyncStatus status, void* userdata);

WGPUBuffer wgpuDeviceCreateBuffer(WGPUDevice, WGPUBufferDescriptor const*);
void wgpuBufferRelease(WGPUBuffer);
void wgpuQueueWriteBuffer(WGPUQueue, WGPUBuffer, uint64_t offset, void const* data, size_t size);
void wgpuCommandEncoderCopyBufferToBuffer(WGPUCommandEncoder, WGPUBuffer src, uint64_t srcOff, WGPUBuffer dst, uint64_t dstOff, uint64_t size);
WGPUShaderModule wgpuDeviceCreateShaderModule(WGPUDevice, WGPUShaderModuleDescriptor const*);
void wgpuShaderModuleRelease(WGPUShaderModule);
WGPUBindGroupLayout wgpuDeviceCreateBindGroupLayout(WGPUDevice, WGPUBindGroupLayoutDescriptor const*);
void wgpuBindGroupLayoutReference(WGPUBindGroupLayout);
WGPUPipelineLayout wgpuDeviceCreatePipelineLayout(WGPUDevice, WGPUPipelineLayoutDescriptor const*);
void wgpuPipelineLayoutRelease(WGPUPipelineLayout);
WGPUComputePipeline wgpuDeviceCreateComputePipeline(WGPUDevice, WGPUComputePipelineDescriptor const*);
void wgpuComputePipelineRelease(WGPUComputePipeline);
WGPUBindGroup wgpuDeviceCreateBindGroup(WGPUDevice, WGPUBindGroupDescriptor const*);
void wgpuBindGroupRelease(WGPUBindGroup);
WGPUCommandEncoder wgpuDeviceCreateCommandEncoder(WGPUDevice, WGPUCommandEncoderDescriptor const*);
WGPUComputePassEncoder wgpuCommandEncoderBeginComputePass(WGPUCommandEncoder, WGPUComputePassDescriptor const*);
void wgpuComputePassEncoderSetPipeline(WGPUComputePassEncoder, WGPUComputePipeline);
void wgpuComputePassEncoderSetBindGroup(WGPUComputePassEncoder, uint32_t, WGPUBindGroup, size_t, uint32_t const*);
void wgpuComputePassEncoderDispatchWorkgroups(WGPUComputePassEncoder, uint32_t, uint32_t, uint32_t);
void wgpuComputePassEncoderEnd(WGPUComputePassEncoder);
void wgpuComputePassEncoderRelease(WGPUComputePassEncoder);
WGPUCommandBuffer wgpuCommandEncoderFinish(WGPUCommandEncoder, WGPUCommandBufferDescriptor const*);
void wgpuCommandEncoderRelease(WGPUCommandEncoder);
void wgpuCommandBufferRelease(WGPUCommandBuffer);
void wgpuQueueSubmit(WGPUQueue, size_t, WGPUCommandBuffer const*);
void wgpuBufferMapAsync(WGPUBuffer, WGPUMapModeFlags, size_t offset, size_t size, WGPUBufferMapCallback, void* userdata);
void const* wgpuBufferGetConstMappedRange(WGPUBuffer, size_t offset, size_t size);
void wgpuBufferUnmap(WGPUBuffer);
void wgpuDeviceTick(WGPUDevice);

/* stub-only accessors used by oracle/ref_shim.cpp */
void* thstub_buffer_data(WGPUBuffer);
uint64_t thstub_buffer_size(WGPUBuffer);
WGPUDevice thstub_device(void);
WGPUQueue thstub_queue(void);
uint64_t thstub_dispatch_count(void);
uint32_t thstub_buffer_id(WGPUBuffer);
void thstub_trace_begin(void);          /* start recording dispatches / copies / writes / submits */
const char* thstub_trace_end(void);     /* stop; one JSON object per line, valid until the next begin */

#ifdef __cplusplus
}
#endif
#endif
