// common.cuh — shared infrastructure of libsvtav1_b200.so (sm_100a only).
//  * error capture (thread-local message, negative SvtB200Status codes, no CPU fallback)
//  * launch counter (svt_b200_launch_count)
//  * ThreadCtx: per-calling-thread stream + growable pinned/device staging for the RTCD drop-ins, which —
//    like the reference's function pointers — are re-entrant and called concurrently from
//    every pipeline thread with raw host pointers.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/svt_av1_b200.h"

namespace svtb200 {

extern std::atomic<uint64_t> g_launches;
void set_error(const char *fmt, ...);

#define SVTB_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            svtb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SVT_B200_ERR_CUDA;                                                         \
        }                                                                                     \
    } while (0)

// Launch + count. Use as: SVTB_LAUNCH(kernel, grid, block, smem, stream, args...)
#define SVTB_LAUNCH(kern, grid, block, smem, stream, ...)           \
    do {                                                            \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);   \
        svtb200::g_launches.fetch_add(1, std::memory_order_relaxed); \
    } while (0)

// A drop-in has no error return (the reference's kernels return void): a CUDA failure is fatal, exactly
// like EB_ErrorMax through lib_svt_encoder_send_error_exit — log and abort; never fall back to the CPU.
[[noreturn]] void fatal(const char *what, cudaError_t e);
#define SVTB_CUDA_FATAL(expr)                            \
    do {                                                 \
        cudaError_t e__ = (expr);                        \
        if (e__ != cudaSuccess) svtb200::fatal(#expr, e__); \
    } while (0)

struct ThreadCtx {
    cudaStream_t stream = nullptr;
    uint8_t *h = nullptr; // pinned
    uint8_t *d = nullptr; // device
    size_t cap = 0;
    int device = -1; // the device stream / d belong to: rebuilt when the calling thread's current device changes
    void reserve(size_t bytes);
    void drop();
    ~ThreadCtx();
};

// Per-DEVICE one-time set-up (cudaFuncSetAttribute, occupancy queries and table uploads are per device, a process may
// use several - svt_b200_set_device): `fn` runs once for the calling thread's current device, under a lock, and its
// CUDA errors are fatal (no CPU fallback).  Usage: static PerDeviceOnce once; once.run([]{ ... });
struct PerDeviceOnce {
    std::atomic<int> done[64];
    PerDeviceOnce() {
        for (auto &d : done) d.store(0);
    }
    template <typename F>
    void run(F fn) {
        int dev = 0;
        SVTB_CUDA_FATAL(cudaGetDevice(&dev));
        dev &= 63;
        if (done[dev].load(std::memory_order_acquire) == 2) return;
        int expect = 0;
        if (done[dev].compare_exchange_strong(expect, 1)) {
            fn();
            done[dev].store(2, std::memory_order_release);
        } else {
            while (done[dev].load(std::memory_order_acquire) != 2) {
            } // another pipeline thread is setting this device up
        }
    }
};
#define SVTB_ATTR(kernel, bytes) SVTB_CUDA_FATAL(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
ThreadCtx &tls();

// ---- device helpers -------------------------------------------------------------------------------
// 4 byte-wise absolute differences accumulated into acc: one VABSDIFF4.U8.ACC on sm_100a.
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

} // namespace svtb200
