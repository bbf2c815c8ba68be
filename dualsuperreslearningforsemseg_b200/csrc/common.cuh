// Shared host/device helpers for libdsrl_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/dsrl_b200.h"

namespace dsrl {

// ---- error plumbing: never throw across the C boundary ---------------------------------------------------
void set_last_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define DSRL_CUDA_TRY(expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::dsrl::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return DSRL_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

#define DSRL_FAIL(code, ...)                  \
    do {                                      \
        ::dsrl::set_last_error(__VA_ARGS__);  \
        return (code);                        \
    } while (0)

// Checks the launch itself (configuration errors); execution errors surface at the caller's next sync.
#define DSRL_LAUNCH_CHECK()                                                                         \
    do {                                                                                            \
        ::dsrl::count_launch();                                                                     \
        DSRL_CUDA_TRY(cudaPeekAtLastError());                                                       \
    } while (0)

int device_sm_count();          // cached
// A zero-initialised device word from a per-device pool for kernels that elect their last CTA with a self-resetting
// atomicInc: round robin over 4096 words for eager launches, a word of its own (never handed out again) for a launch that
// `st` is capturing into a CUDA graph.  Returns nullptr (and sets the error) on failure.  The pool is allocated on
// first use, which must not happen inside a stream capture: run one eager call before capturing a CUDA graph.
unsigned *next_ticket_slot(cudaStream_t st);
int require_device();           // DSRL_OK or DSRL_ERR_CUDA (no device / wrong arch): there is no CPU fallback

// ---- device helpers ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide sum; result valid in every thread.  `scratch` needs >= 33 elements.  Deterministic order.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        T t = lane < nw ? scratch[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// Streaming (read-once) loads: bypass L1 allocation, keep 128-bit width.
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint16_t ldg_stream_u16(const void *p) {
    uint16_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint8_t ldg_stream_u8(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
    return (uint8_t)r;
}
#endif  // __CUDACC__

}  // namespace dsrl
