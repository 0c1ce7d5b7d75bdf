// Shared helpers for the dgnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dgnn_b200.h"

namespace dgnn {

extern thread_local char g_err[512];

inline int fail(const char* what, const char* detail) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, detail);
    return 1;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(what, cudaGetErrorString(e));
    return 0;
}

#define DGNN_REQUIRE(cond, what)                                   \
    do {                                                           \
        if (!(cond)) return ::dgnn::fail(__func__, what);          \
    } while (0)

int sm_count();          // SMs of the CURRENT device (cached per device)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: done once per (kernel, device)
int ensure_dyn_smem(const void* kernel, int bytes, const char* what);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// Loads that stay where they are written: ptxas is free to sink an ordinary read-only load down to its first use (it does,
// under register pressure), which turns a software prefetch into a blocking load.  `asm volatile` keeps the issue point.
__device__ __forceinline__ float4 ldg4_pinned(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ldg4i_pinned(const int32_t* p) {
    int4 v;
    asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// 256-bit read-only load (LDG.E.256, sm_100+); p must be 32-byte aligned
__device__ __forceinline__ void ldg8(const float* p, float4& a, float4& b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// 256-bit store (STG.E.256, sm_100+); p must be 32-byte aligned
__device__ __forceinline__ void stg8(float* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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

// h = relu?(x*scale + shift) applied on load ("producer affine on load")
__device__ __forceinline__ float act(float x, float sc, float sh, bool relu) {
    float y = fmaf(x, sc, sh);
    return relu ? fmaxf(y, 0.f) : y;
}

}  // namespace dgnn
