// api.cu — library management entry points of include/svt_av1_b200.h.
#include <cstdarg>

#include "common.cuh"

namespace svtb200 {
std::atomic<uint64_t> g_launches{0};
static thread_local char t_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
[[noreturn]] void fatal(const char *what, cudaError_t e) {
    fprintf(stderr, "libsvtav1_b200: fatal CUDA error in %s: %s (no CPU fallback)\n", what, cudaGetErrorString(e));
    abort();
}
void ThreadCtx::drop() {
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    if (stream) cudaStreamDestroy(stream);
    h = d = nullptr;
    stream = nullptr;
    cap = 0;
}
void ThreadCtx::reserve(size_t bytes) {
    int dev = 0;
    SVTB_CUDA_FATAL(cudaGetDevice(&dev));
    if (dev != device) { // this thread now serves another GPU: its stream and staging belong to the old one
        if (device >= 0) {
            cudaSetDevice(device);
            drop();
            cudaSetDevice(dev);
        }
        device = dev;
    }
    if (!stream) SVTB_CUDA_FATAL(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (bytes <= cap) return;
    size_t ncap = cap ? cap : (1u << 20);
    while (ncap < bytes) ncap *= 2;
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    SVTB_CUDA_FATAL(cudaMallocHost(&h, ncap));
    SVTB_CUDA_FATAL(cudaMalloc(&d, ncap));
    cap = ncap;
}
ThreadCtx::~ThreadCtx() {
    // the CUDA context may already be gone at thread/process exit; ignore errors
    drop();
}
ThreadCtx &tls() {
    static thread_local ThreadCtx ctx;
    return ctx;
}
} // namespace svtb200

using namespace svtb200;

extern "C" {
int svt_b200_version(void) { return 100; }
int svt_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return e == cudaErrorNoDevice ? 0 : SVT_B200_ERR_CUDA;
    }
    return n;
}
int svt_b200_set_device(int device) {
    SVTB_CUDA_TRY(cudaSetDevice(device));
    return SVT_B200_OK;
}
const char *svt_b200_last_error(void) { return t_err; }
uint64_t svt_b200_launch_count(void) { return g_launches.load(); }
void *svt_b200_malloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void svt_b200_free(void *p) { cudaFree(p); }
void *svt_b200_malloc_host(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void svt_b200_free_host(void *p) { cudaFreeHost(p); }
int svt_b200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    SVTB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return SVT_B200_OK;
}
int svt_b200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    SVTB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SVT_B200_OK;
}
int svt_b200_stream_sync(void *stream) {
    SVTB_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return SVT_B200_OK;
}
}
