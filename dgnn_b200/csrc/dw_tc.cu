// Weight gradient of a layer on tensor cores with a TMEM-resident accumulator:
//
//     dW[f_out, k_total] = sum over cells  dz[cell]^T . [agg | h][cell]
//
// The contraction runs over cells, the slow dimension of the row-major activations, so both
// operands must be transposed on their way into shared memory.  (tcgen05.mma kind::tf32 silently
// yields zeros for MN-major descriptors on this part — measured with tools/probe_umma.py — so the
// hardware transpose is not an option.)  The producer warps transpose in registers: 4 lanes hold
// 4 channels x 1 cell each, two shuffle rounds turn that into 1 channel x 4 cells, i.e. one
// 16-byte chunk of the K-major (K = cell) swizzled operand row of that channel.
// The [128 x N] fp32 accumulator (N = k_total <= 256 TMEM columns) stays in TMEM for the whole
// lifetime of the persistent CTA; it is read out once at the end as this CTA's partial, and the
// partials of the 148 CTAs are summed by dgnn_reduce_partials_f32 (deterministic, no atomics).
// 3xTF32 split as in layer_tc.cu.
#include "umma.cuh"
#include "common.cuh"

namespace dgnn {

using namespace umma;

constexpr int DW_NPW = 16;
constexpr int DW_THREADS = (DW_NPW + 1) * 32;
constexpr int DW_CELLS = 32;                       // cells (K) per stage: 2 per producer warp
constexpr int DW_A_BYTES = 128 * 128;               // M = 128 channel rows x 32 cells (one K-atom): 16 KB

struct DwTcArgs {
    const float* dy;
    const float* z;
    const float* ng;
    const float* na;
    const float* nb;
    const float* nmean;
    const float* nrstd;
    const float* agg;
    const float* x_in;
    const float* in_scale;
    const float* in_shift;
    int relu_in;
    int64_t n_tgt;
    int f_in, f_out, k_total, np, stages;
    float* partials;  // [grid][f_out][k_total]
};

// 4x4 transpose across the 4 lanes 4q..4q+3: in: lane j holds (c0..c3) of cell j; out: lane j holds
// channel j of cells 0..3
__device__ __forceinline__ float4 transpose4(float4 v, int j) {
    {
        const bool up = (j & 2) != 0;
        float a = up ? v.x : v.z, b = up ? v.y : v.w;
        a = __shfl_xor_sync(0xffffffffu, a, 2);
        b = __shfl_xor_sync(0xffffffffu, b, 2);
        if (up) { v.x = a; v.y = b; } else { v.z = a; v.w = b; }
    }
    {
        const bool up = (j & 1) != 0;
        float a = up ? v.x : v.y, b = up ? v.z : v.w;
        a = __shfl_xor_sync(0xffffffffu, a, 1);
        b = __shfl_xor_sync(0xffffffffu, b, 1);
        if (up) { v.x = a; v.z = b; } else { v.y = a; v.w = b; }
    }
    return v;
}

__device__ __forceinline__ void put_split4(uint8_t* hi, uint8_t* lo, uint32_t off, float4 v) {
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

__global__ void __launch_bounds__(DW_THREADS, 1) dw_tc_kernel(const DwTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[2], bar_empty[2], bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b_bytes = p.np * 128;
    const int stage_bytes = 2 * DW_A_BYTES + 2 * b_bytes;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_full[s], DW_NPW); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    // zero all stages once: channel blocks beyond f_out / k_total are never written but are read by the MMA
    for (int i = tid; i < p.stages * stage_bytes / 16; i += DW_THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    if (warp == DW_NPW) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_groups = (p.n_tgt + DW_CELLS - 1) / DW_CELLS;
    // contiguous range of 32-cell groups per CTA
    const int64_t per = (n_groups + gridDim.x - 1) / gridDim.x;
    const int64_t g_begin = (int64_t)blockIdx.x * per;
    const int64_t g_end = g_begin + per < n_groups ? g_begin + per : n_groups;
    const int64_t my_groups = g_end > g_begin ? g_end - g_begin : 0;

    if (warp == DW_NPW) {
        const uint32_t idesc = make_idesc_tf32(128, p.np);
        for (int64_t i = 0; i < my_groups; ++i) {
            const uint32_t s = (uint32_t)(i % p.stages), use = (uint32_t)(i / p.stages);
            if (lane == 0) {
                mbar_wait(&bar_full[s], use & 1);
                tc_fence_after_sync();
                const uint32_t ah = smem_u32(smem + (size_t)s * stage_bytes), al = ah + DW_A_BYTES;
                const uint32_t bh = al + DW_A_BYTES, bl = bh + b_bytes;
#pragma unroll
                for (int kk = 0; kk < DW_CELLS / 8; ++kk) {
                    const uint32_t ko = kk * 32;
                    mma_tf32(tmem_base, make_desc(ah + ko), make_desc(bh + ko), idesc, (i > 0 || kk > 0) ? 1u : 0u);
                    mma_tf32(tmem_base, make_desc(al + ko), make_desc(bh + ko), idesc, 1u);
                    mma_tf32(tmem_base, make_desc(ah + ko), make_desc(bl + ko), idesc, 1u);
                }
                mma_commit(&bar_empty[s]);
                if (i == my_groups - 1) mma_commit(&bar_done);
            }
            __syncwarp();
        }
    } else {
        const bool relu = p.relu_in != 0;
        const int j = lane & 3, c = lane >> 2;       // cell within a group of 4, 16-byte chunk within a 32-channel block
        const int nb_a = (p.f_out + 31) >> 5;        // 32-channel blocks of dz
        const int n_items = 8 * (nb_a + (p.np >> 5));  // (cell group, channel block) pairs per stage
        for (int64_t i = 0; i < my_groups; ++i) {
            const uint32_t s = (uint32_t)(i % p.stages), use = (uint32_t)(i / p.stages);
            uint8_t* a_hi = smem + (size_t)s * stage_bytes;
            uint8_t* a_lo = a_hi + DW_A_BYTES;
            uint8_t* b_hi = a_lo + DW_A_BYTES;
            uint8_t* b_lo = b_hi + b_bytes;
            // phase 1: all global loads of this warp's items (up to 6) go in flight together
            constexpr int MAXI = 6;   // 8 * (4 + 8) items / 16 warps
            float4 r0[MAXI], r1[MAXI];
#pragma unroll
            for (int u = 0; u < MAXI; ++u) {
                const int item = warp + u * DW_NPW;
                r0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                r1[u] = r0[u];
                if (item >= n_items) continue;
                const int g = item & 7, blk = item >> 3;
                const int64_t t = (g_begin + i) * DW_CELLS + g * 4 + j;
                if (t >= p.n_tgt) continue;
                if (blk < nb_a) {
                    const int ch = blk * 32 + c * 4;
                    if (ch < p.f_out) {
                        r0[u] = ldg4(p.dy + (size_t)t * p.f_out + ch);
                        if (p.ng != nullptr) r1[u] = ldg4(p.z + (size_t)t * p.f_out + ch);
                    }
                } else {
                    const int n = (blk - nb_a) * 32 + c * 4;
                    if (n < p.k_total) {
                        if (p.agg != nullptr && n < p.f_in) r0[u] = ldg4(p.agg + (size_t)t * p.f_in + n);
                        else r0[u] = ldg4(p.x_in + (size_t)t * p.f_in + (p.agg != nullptr ? n - p.f_in : n));
                    }
                }
            }
            mbar_wait(&bar_empty[s], (use & 1) ^ 1);
            // phase 2: transform, transpose (4 channels x 1 cell -> 1 channel x 4 cells), split, store
#pragma unroll
            for (int u = 0; u < MAXI; ++u) {
                const int item = warp + u * DW_NPW;
                if (item >= n_items) continue;
                const int g = item & 7, blk = item >> 3;
                const int64_t t = (g_begin + i) * DW_CELLS + g * 4 + j;
                const bool tv = t < p.n_tgt;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                uint8_t *dst_hi, *dst_lo;
                int row;
                if (blk < nb_a) {                       // dz block -> A operand rows (channels)
                    const int ch = blk * 32 + c * 4;
                    if (tv && ch < p.f_out) {
                        const float4 d = r0[u];
                        if (p.ng != nullptr) {
                            const float4 zv = r1[u];
                            float4 gg = ldg4(p.ng + ch), a = ldg4(p.na + ch), b = ldg4(p.nb + ch), m = ldg4(p.nmean + ch),
                                   rs = ldg4(p.nrstd + ch);
                            v.x = gg.x * d.x - (a.x + (zv.x - m.x) * rs.x * b.x);
                            v.y = gg.y * d.y - (a.y + (zv.y - m.y) * rs.y * b.y);
                            v.z = gg.z * d.z - (a.z + (zv.z - m.z) * rs.z * b.z);
                            v.w = gg.w * d.w - (a.w + (zv.w - m.w) * rs.w * b.w);
                        } else {
                            v = d;
                        }
                    }
                    dst_hi = a_hi; dst_lo = a_lo;
                    row = blk * 32 + c * 4 + j;
                } else {                                // [agg | h] block -> B operand rows (columns of dW)
                    const int n = (blk - nb_a) * 32 + c * 4;
                    if (tv && n < p.k_total) {
                        v = r0[u];
                        if (!(p.agg != nullptr && n < p.f_in)) {
                            const int col = p.agg != nullptr ? n - p.f_in : n;
                            if (p.in_scale != nullptr) {
                                float4 sc = ldg4(p.in_scale + col), sh = ldg4(p.in_shift + col);
                                v.x = act(v.x, sc.x, sh.x, relu); v.y = act(v.y, sc.y, sh.y, relu);
                                v.z = act(v.z, sc.z, sh.z, relu); v.w = act(v.w, sc.w, sh.w, relu);
                            } else if (relu) {
                                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                            }
                        }
                    }
                    dst_hi = b_hi; dst_lo = b_lo;
                    row = (blk - nb_a) * 32 + c * 4 + j;
                }
                v = transpose4(v, j);                   // now: channel `row`, cells 4g .. 4g+3
                put_split4(dst_hi, dst_lo, (uint32_t)row * 128u + (uint32_t)((g ^ (row & 7)) << 4), v);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[s]);
        }
        // read this CTA's partial out of TMEM
        float* out = p.partials + (size_t)blockIdx.x * p.f_out * p.k_total;
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        if (my_groups > 0) {
            mbar_wait(&bar_done, 0);
            tc_fence_after_sync();
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c0 = (grp + 4 * j) * 32;
            if (c0 >= p.np) break;
            float v[32];
            if (my_groups > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (row < p.f_out) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int n = c0 + i;
                    if (n >= p.k_total) continue;
                    *reinterpret_cast<float4*>(out + (size_t)row * p.k_total + n) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == DW_NPW) tmem_dealloc(tmem_base, 256);
}

}  // namespace dgnn

using namespace dgnn;

static inline int ceil32i(int x) { return (x + 31) / 32 * 32; }

extern "C" int dgnn_dw_tc_supported(int f_out, int k_total) {
    return (f_out % 4 == 0 && k_total % 4 == 0 && f_out <= 128 && ceil32i(k_total) <= 256) ? 1 : 0;
}

extern "C" int dgnn_dw_bwd_tc(const float* dy, const float* z, const float* g, const float* a, const float* b,
                              const float* mean, const float* rstd, const float* agg, const float* x_in,
                              const float* in_scale, const float* in_shift, int relu_in, int64_t n_tgt, int f_in,
                              int f_out, int k_total, float* partials, void* stream) {
    DGNN_REQUIRE(dgnn_dw_tc_supported(f_out, k_total), "widths not supported by the tensor-core dW kernel");
    DGNN_REQUIRE(k_total == (agg ? 2 * f_in : f_in), "k_total mismatch");
    DGNN_REQUIRE(dy && x_in && partials, "null pointer");
    DwTcArgs p;
    p.dy = dy; p.z = z; p.ng = g; p.na = a; p.nb = b; p.nmean = mean; p.nrstd = rstd;
    p.agg = agg; p.x_in = x_in; p.in_scale = in_scale; p.in_shift = in_shift; p.relu_in = relu_in;
    p.n_tgt = n_tgt; p.f_in = f_in; p.f_out = f_out; p.k_total = k_total;
    p.np = ceil32i(k_total);
    p.stages = 2;
    p.partials = partials;
    size_t smem = (size_t)p.stages * (2 * DW_A_BYTES + 2 * p.np * 128) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
        if (e != cudaSuccess) return fail("dgnn_dw_bwd_tc", cudaGetErrorString(e));
        configured = true;
    }
    dw_tc_kernel<<<sm_count(), DW_THREADS, smem, as_stream(stream)>>>(p);
    return check_launch("dgnn_dw_bwd_tc");
}
