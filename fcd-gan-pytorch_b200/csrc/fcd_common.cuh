// fcd_common.cuh — shared helpers for every kernel family in libfcd_b200.so.
//
// Data-layout conventions:
//   * "split" activation  = two bf16 NHWC planes (hi, lo) with hi = bf16_rn(v), lo = bf16_rn(v - hi);
//                           hi + lo reproduces an fp32 value to ~2^-17 relative. Every convolution
//                           INPUT (forward activations, backward output-gradients) is a split tensor,
//                           so the tcgen05 kernels can TMA it straight into MMA-ready shared memory
//                           and recover fp32-class accuracy with three bf16 MMAs (hi*hi + lo*hi + hi*lo).
//   * raw conv OUTPUT     = fp32 NHWC.
//   * every NHWC tensor carries an explicit pixel pitch `ld` (elements between consecutive pixels) so
//     channel slices of a concatenation buffer can be read / written in place (virtual concat).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fcd_b200.h"

namespace fcd {

// Last error text, per host thread (C-ABI entry points return the code, fcd_last_error() the text).
void set_error(int code, const char* fmt, ...);

#define FCD_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            ::fcd::set_error(FCD_ERR_ARG, __VA_ARGS__); \
            return FCD_ERR_ARG;                       \
        }                                             \
    } while (0)

#define FCD_CUDA_OK(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::fcd::set_error(FCD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                 \
            return FCD_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define FCD_LAUNCH_OK() FCD_CUDA_OK(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int FCD_MAX_DEVICES = 64;
int current_device();     // cudaGetDevice, clamped to [0, FCD_MAX_DEVICES)
int sm_count();           // of the current device

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: remember it per device ordinal
struct PerDeviceOnce {
    bool done[FCD_MAX_DEVICES] = {};
    bool& operator()() { return done[current_device()]; }
};

// ---- split-bf16 helpers --------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
    return __bfloat162float(hi) + __bfloat162float(lo);
}

// Pointer pair for a split tensor; lo may be null ("fast" single-plane bf16 mode).
struct SplitPtr {
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
};
struct SplitCPtr {
    const __nv_bfloat16* hi;
    const __nv_bfloat16* lo;
};
__device__ __forceinline__ float load_split(const SplitCPtr& p, size_t i) {
    float v = __bfloat162float(p.hi[i]);
    if (p.lo) v += __bfloat162float(p.lo[i]);
    return v;
}
__device__ __forceinline__ void store_split(const SplitPtr& p, size_t i, float v) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    p.hi[i] = h;
    if (p.lo) p.lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---- reductions ----------------------------------------------------------------------------
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

// Compensated fp32 accumulator (Kahan): s - c carries the running sum to ~2^-48 relative without touching the FP64 pipe.
// (-fmad must not contract these: the operations below are written so that no a*b+c pattern exists.)
struct KahanF {
    float s = 0.f, c = 0.f;
    __device__ __forceinline__ void add(float v) {
        const float y = __fsub_rn(v, c);
        const float t = __fadd_rn(s, y);
        c = __fsub_rn(__fsub_rn(t, s), y);
        s = t;
    }
    __device__ __forceinline__ double value() const { return static_cast<double>(s) - static_cast<double>(c); }
};

// Activation kinds shared by the BN/activation kernels and the conv epilogues.
__device__ __forceinline__ float act_fwd(int kind, float u, float slope) {
    switch (kind) {
        case FCD_ACT_RELU: return u > 0.f ? u : 0.f;
        case FCD_ACT_PRELU:
        case FCD_ACT_LEAKY: return u > 0.f ? u : slope * u;
        default: return u;
    }
}
// derivative w.r.t. the pre-activation u
__device__ __forceinline__ float act_grad(int kind, float u, float slope) {
    switch (kind) {
        case FCD_ACT_RELU: return u > 0.f ? 1.f : 0.f;
        case FCD_ACT_PRELU:
        case FCD_ACT_LEAKY: return u > 0.f ? 1.f : slope;
        default: return 1.f;
    }
}

static inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace fcd
