// tcgen05 / TMEM / mbarrier PTX wrappers and the shared-memory operand layout used by the
// tensor-core kernels (sm_100a).
//
// Operand layout: K-major, 128-byte swizzle, TF32.  One "K-atom" is [rows x 32 floats]:
// row r at byte r*128, 8-row groups 1024 B apart (SBO), 16-byte chunk c of a row stored at
// chunk (c ^ (r & 7)).  Atom bases are 1024-byte aligned.  An MMA k-step (K = 8 floats = 32 B)
// inside an atom advances the descriptor start address by 32 B.
//
// FP32 fidelity: every product is evaluated as hi*hi + lo*hi + hi*lo with
// hi = rna_tf32(x), lo = x - hi truncated to TF32 by the tensor core (3xTF32, fp32 accumulation in TMEM); the
// dropped lo*lo term and the truncation of lo are ~2^-21 relative (indistinguishable from fp32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgnn {
namespace umma {

constexpr int ATOM_K = 32;              // floats per K-atom row (128 B)
constexpr int ATOM_ROW_BYTES = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of float element (row r, k in [0,32)) inside a K-atom
__device__ __forceinline__ uint32_t atom_off(int r, int k) {
    return (uint32_t)r * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)k & 3u) << 2);
}

// Round to TF32 (10 mantissa bits), nearest with ties away from zero, for finite inputs: add half an ulp of the
// TF32 grid to the bit pattern and clear the 13 low bits.  (cvt.rna.tf32.f32 compiles to the same two operations plus
// an Inf/NaN guard of three more instructions on sm_100a; an Inf/NaN input is garbage for the product anyway.)
__device__ __forceinline__ float tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// 3xTF32 split: hi = rna_tf32(x); lo = x - hi is exact in fp32 and is handed to the tensor core as is (kind::tf32
// reads the 19 high bits of an operand, i.e. truncates lo to TF32: |x - hi - trunc(lo)| <= 2^-21 |x|).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rna(x);
    lo = x - hi;
}

// shared-memory accesses through 32-bit shared-space addresses (generic pointers derived from the aligned dynamic
// shared base compile to 64-bit generic LD / ST)
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, const float2& v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Bounded wait: a pipeline bug must trap (surfacing as a CUDA error) instead of hanging the GPU.
// A failed probe backs off with nanosleep so that waiting warps do not steal issue slots from working ones.
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try(addr, parity)) return;
    uint32_t spins = 0;
    for (;;) {
#ifdef DGNN_MBAR_BACKOFF
        __nanosleep(DGNN_MBAR_BACKOFF);
#endif
        if (mbar_try(addr, parity)) return;
        if (++spins > (1u << 28)) __trap();
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive and report how many arrivals the phase still needed BEFORE this one (1 => this arrival completed the phase).
// mbarrier.arrive has release semantics at CTA scope: no __threadfence_block() (MEMBAR, which also waits for the
// thread's in-flight cp.async copies) is needed between a warp's operand stores and the count.
__device__ __forceinline__ uint32_t mbar_arrive_pending(uint64_t* bar) {
    uint64_t state;
    uint32_t pending;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
    return pending;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit; completion (bytes) is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- proxies / fences -------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
// whole warp; writes the base address to *slot (shared)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread i = lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- MMA ----------------------------------------------------------------------------------------
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, SBO = 1024 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;               // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T, one K = 8 step; single thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// MN-major operand (the contiguous dimension is M / N): 128B-swizzled blocks of 32 mn x 8 k;
// lbo = bytes between 32-element mn blocks, sbo = bytes between 8-k groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32_mn(int m, int n) {
    return make_idesc_tf32(m, n) | (1u << 15) | (1u << 16);
}
// make an mbarrier track completion of all previously issued MMAs of this thread
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

}  // namespace umma
}  // namespace dgnn
