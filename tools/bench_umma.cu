// Development probe: how long does ONE tcgen05.mma take when a single thread issues a long stream of them?  One CTA per SM,
// static (zero) operands in shared memory, `reps` MMAs back to back, clock64 around the stream + its commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared tools/bench_umma.cu \
//        dgnn_b200/csrc/head.o -o gpurun_variants/libbench_umma.so -lcudart
//   python tools/bench_umma.py
// variant bits: 0-1 operand layout (0 = SWIZZLE_128B atoms [rows x 128 B], 1 = SWIZZLE_32B atoms [rows x 32 B], 2 = SWIZZLE_64B,
// 3 = no swizzle), bit 2: bf16 (kind::f16, K = 16) instead of tf32 (K = 8), bit 3: A operand in tensor memory,
// bits 4-5: log2 of the number of accumulators used round-robin (independent dependency chains),
// bits 6-7: log2 of the number of ISSUING warps (each its own accumulator of n <= 128 columns; lane 0 of the warp issues).
#include "../dgnn_b200/csrc/umma.cuh"
#include "../dgnn_b200/csrc/common.cuh"

namespace dgnn {
using namespace umma;

__device__ __forceinline__ uint64_t desc_of(uint32_t addr, int layout) {
    // K-major; SBO = bytes between 8-row groups
    const uint32_t row_bytes = layout == 0 ? 128u : (layout == 1 ? 32u : (layout == 2 ? 64u : 32u));
    const uint64_t swz = layout == 0 ? 2 : (layout == 1 ? 6 : (layout == 2 ? 4 : 0));
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= swz << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) bench_umma_kernel(int variant, int n, int reps, float* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) { mbar_init(&bar, 1 << ((variant >> 6) & 3)); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (tid < 32) tmem_alloc(&slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = slot;
    const int layout = variant & 3, bf16 = (variant >> 2) & 1, a_tmem = (variant >> 3) & 1, nacc = 1 << ((variant >> 4) & 3);
    const int nwarp = 1 << ((variant >> 6) & 3);
    __shared__ long long t_begin[4], t_end[4];
    if ((tid & 31) == 0 && (tid >> 5) < nwarp) {
        const int w = tid >> 5;
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
        const uint32_t fmt = bf16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t kstep = layout == 0 ? 32u : (layout == 2 ? 32u : 8u * 128u * 32u / 8u);   // bytes to the next k-step
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t kk = (uint32_t)(r & 3);
            const uint32_t d = nwarp > 1 ? tb + (uint32_t)(w * 112) : tb + (uint32_t)((r & (nacc - 1)) * (512 / nacc));
            const uint64_t bd = desc_of(b0 + (layout == 0 ? kk * 32u : kk * 8192u), layout);
            if (a_tmem) {
                const uint32_t at = tb + 448u + kk * 8u;
                if (bf16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                                 ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            } else {
                const uint64_t ad = desc_of(a0 + (layout == 0 ? kk * 32u : kk * 4096u), layout);
                if (bf16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                else
                    mma_tf32(d, ad, bd, idesc, 1u);
            }
            (void)kstep;
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        t_begin[w] = t0; t_end[w] = t1;
    }
    __syncthreads();
    if (tid == 0) {
        long long b = t_begin[0], e = t_end[0];
        for (int w = 1; w < nwarp; ++w) { b = t_begin[w] < b ? t_begin[w] : b; e = t_end[w] > e ? t_end[w] : e; }
        out[blockIdx.x] = (float)(e - b) / (float)(reps * nwarp);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tb, 512);
}
}  // namespace dgnn

using namespace dgnn;

extern "C" int bench_umma(int variant, int n, int reps, float* out, int grid, void* stream) {
    cudaError_t e = cudaFuncSetAttribute(bench_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return fail("bench_umma", cudaGetErrorString(e));
    bench_umma_kernel<<<grid, 128, 97 * 1024, as_stream(stream)>>>(variant, n, reps, out);
    return check_launch("bench_umma");
}
