// context.cu -- device context, buffers and copies behind the C ABI (include/thk_cabi.h).
// Replaces the WGPUDevice/WGPUQueue/WGPUBuffer plumbing of the reference.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void thk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char* thk_last_error(void) { return g_err; }
extern "C" const char* thk_version(void) { return "thk-sm100a 0.1 (token_hawk_b200)"; }

static int init_common(int device, cudaStream_t stream, bool own, thk_ctx** out) {
    THK_CHECK_ARG(out != nullptr, "thk_init: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        thk_set_error("thk_init: no CUDA device (%s); this library has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
        return THK_E_CUDA;
    }
    THK_CHECK_ARG(device >= 0 && device < n, "thk_init: device %d out of range [0,%d)", device, n);
    THK_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    THK_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) {
        thk_set_error("thk_init: device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
        return THK_E_UNSUPPORTED;
    }
    thk_ctx* c = new thk_ctx;
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    c->cc_major = p.major;
    c->cc_minor = p.minor;
    c->total_mem = p.totalGlobalMem;
    if (own) {
        cudaError_t es = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (es != cudaSuccess) { delete c; thk_set_error("cudaStreamCreate: %s", cudaGetErrorString(es)); return THK_E_CUDA; }
        c->own_stream = true;
    } else {
        c->stream = stream;
    }
    *out = c;
    return THK_OK;
}

extern "C" int thk_init(int device, thk_ctx** out) { return init_common(device, nullptr, true, out); }
extern "C" int thk_init_on_stream(int device, void* stream, thk_ctx** out) {
    return init_common(device, (cudaStream_t)stream, false, out);
}
extern "C" int thk_destroy(thk_ctx* ctx) {
    THK_ENTER(ctx);
    if (!ctx) return THK_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->gemm_ws);
    cudaFree(ctx->gemm_status);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return THK_OK;
}
extern "C" int thk_device_info(thk_ctx* ctx, int* sm, int* maj, int* min, size_t* mem) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx, "thk_device_info: null ctx");
    if (sm) *sm = ctx->sm_count;
    if (maj) *maj = ctx->cc_major;
    if (min) *min = ctx->cc_minor;
    if (mem) *mem = ctx->total_mem;
    return THK_OK;
}
extern "C" void* thk_stream(thk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int thk_malloc(thk_ctx* ctx, size_t bytes, void** dptr) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dptr, "thk_malloc: null argument");
    *dptr = nullptr;
    THK_CUDA(cudaSetDevice(ctx->device));
    THK_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return THK_OK;
}
extern "C" int thk_free(thk_ctx* ctx, void* dptr) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx, "thk_free: null ctx");
    if (!dptr) return THK_OK;
    THK_CUDA(cudaSetDevice(ctx->device));
    THK_CUDA(cudaFree(dptr));
    return THK_OK;
}
extern "C" int thk_memset(thk_ctx* ctx, void* dptr, int value, size_t bytes) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dptr, "thk_memset: null argument");
    THK_CUDA(cudaMemsetAsync(dptr, value, bytes, ctx->stream));
    return THK_OK;
}
extern "C" int thk_upload(thk_ctx* ctx, void* dst, size_t off, const void* src, size_t bytes) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dst && src, "thk_upload: null argument");
    THK_CUDA(cudaMemcpyAsync((char*)dst + off, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return THK_OK;
}
extern "C" int thk_download(thk_ctx* ctx, void* dst, const void* src, size_t off, size_t bytes) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dst && src, "thk_download: null argument");
    THK_CUDA(cudaMemcpyAsync(dst, (const char*)src + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    THK_CUDA(cudaStreamSynchronize(ctx->stream));
    return THK_OK;
}
extern "C" int thk_copy(thk_ctx* ctx, void* dst, size_t doff, const void* src, size_t soff, size_t bytes) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dst && src, "thk_copy: null argument");
    THK_CUDA(cudaMemcpyAsync((char*)dst + doff, (const char*)src + soff, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return THK_OK;
}
extern "C" int thk_sync(thk_ctx* ctx) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx, "thk_sync: null ctx");
    THK_CUDA(cudaStreamSynchronize(ctx->stream));
    return THK_OK;
}
extern "C" int thk_host_alloc(thk_ctx* ctx, size_t bytes, void** hptr) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && hptr, "thk_host_alloc: null argument");
    THK_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return THK_OK;
}
extern "C" int thk_host_free(thk_ctx* ctx, void* hptr) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx, "thk_host_free: null ctx");
    if (hptr) THK_CUDA(cudaFreeHost(hptr));
    return THK_OK;
}

// ---- cross-process / cross-device mapping of exchange regions (tensor parallel wiring) ----
extern "C" int thk_ipc_export(thk_ctx* ctx, void* dptr, unsigned char* handle64) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && dptr && handle64, "thk_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    THK_CUDA(cudaSetDevice(ctx->device));
    THK_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, 64);
    return THK_OK;
}
extern "C" int thk_ipc_import(thk_ctx* ctx, const unsigned char* handle64, void** dptr) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && handle64 && dptr, "thk_ipc_import: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    THK_CUDA(cudaSetDevice(ctx->device));
    THK_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return THK_OK;
}
extern "C" int thk_ipc_close(thk_ctx* ctx, void* dptr) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx, "thk_ipc_close: null ctx");
    if (dptr) THK_CUDA(cudaIpcCloseMemHandle(dptr));
    return THK_OK;
}
// same-process multi-GPU: let `ctx`'s device read/write memory of `peer`'s device
extern "C" int thk_enable_peer_access(thk_ctx* ctx, thk_ctx* peer) {
    THK_ENTER(ctx);
    THK_CHECK_ARG(ctx && peer, "thk_enable_peer_access: null argument");
    if (ctx->device == peer->device) return THK_OK;
    int can = 0;
    THK_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
    if (!can) { thk_set_error("device %d cannot access device %d", ctx->device, peer->device); return THK_E_UNSUPPORTED; }
    THK_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { thk_set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return THK_E_CUDA; }
    cudaGetLastError();
    return THK_OK;
}
