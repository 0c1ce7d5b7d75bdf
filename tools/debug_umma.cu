// Development probe: run ONE tcgen05.mma (kind::tf32, cta_group::1, M = 128) on caller-provided
// shared-memory images and descriptor bits, and return the [128 x n] accumulator.  Used by
// tools/probe_umma.py to validate operand layouts on hardware; not part of the product library:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared tools/debug_umma.cu \
//        dgnn_b200/csrc/head.o -o gpurun_variants/libdebug_umma.so -lcudart
#include "../dgnn_b200/csrc/umma.cuh"
#include "../dgnn_b200/csrc/common.cuh"

namespace dgnn {
using namespace umma;

__global__ void __launch_bounds__(128, 1) debug_umma_kernel(const uint8_t* a_img, int a_bytes, const uint8_t* b_img,
                                                            int b_bytes, unsigned long long a_desc, unsigned long long b_desc,
                                                            uint32_t idesc, int n, float* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + ((a_bytes + 1023) & ~1023);
    const int tid = threadIdx.x;
    for (int i = tid; i < a_bytes; i += 128) a_s[i] = a_img[i];
    for (int i = tid; i < b_bytes; i += 128) b_s[i] = b_img[i];
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (tid < 32) tmem_alloc(&slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = slot;
    if (tid == 0) {
        uint64_t ad = a_desc | (uint64_t)((smem_u32(a_s) & 0x3FFFFu) >> 4);
        uint64_t bd = b_desc | (uint64_t)((smem_u32(b_s) & 0x3FFFFu) >> 4);
        mma_tf32(tb, ad, bd, idesc, 0u);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < n; c0 += 32) {
        float v[32];
        tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 32; ++i) out[(size_t)(warp * 32 + lane) * n + c0 + i] = v[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tb, 256);
}
}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_debug_umma(const uint8_t* a_img, int a_bytes, const uint8_t* b_img, int b_bytes, uint64_t a_desc,
                               uint64_t b_desc, uint32_t idesc, int n, float* out, void* stream) {
    DGNN_REQUIRE(n % 32 == 0 && n <= 256, "n must be a multiple of 32, <= 256");
    size_t smem = (size_t)((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023) + 1024;
    DGNN_REQUIRE(smem <= 200 * 1024, "images too large");
    cudaError_t e = cudaFuncSetAttribute(debug_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail("dgnn_debug_umma", cudaGetErrorString(e));
    debug_umma_kernel<<<1, 128, smem, as_stream(stream)>>>(a_img, a_bytes, b_img, b_bytes, a_desc, b_desc, idesc, n, out);
    return check_launch("dgnn_debug_umma");
}
