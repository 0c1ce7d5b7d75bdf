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
// The [128 x N] fp32 accumulator (N <= 256 TMEM columns) stays in TMEM for the whole
// lifetime of the persistent CTA; it is read out once at the end as this CTA's partial, and the
// partials of the 148 CTAs are summed by dgnn_reduce_partials_f32 (deterministic, no atomics).
// One launch covers a [128 rows of dW] x [256 columns] block; wider layers (f_out > 128 or k_total > 256:
// modelnet.yaml's 512 / 1024) are run block by block by the launch wrapper.
// 3xTF32 split as in layer_tc.cu.
#include "umma.cuh"
#include "common.cuh"

namespace dgnn {

using namespace umma;

#ifndef DGNN_DW_NPW
#define DGNN_DW_NPW 12
#endif
constexpr int DW_NPW = DGNN_DW_NPW;                // producer warps: 12 = one per 32-channel operand block at 128 -> 128 (128
constexpr int DW_THREADS = (DW_NPW + 1) * 32;      // registers).  24 (two per block, 72 registers, spills) measured 2x slower on B200
constexpr int DW_CELLS = 32;                       // cells (K) per stage = 8 groups of 4 cells
constexpr int DW_A_BYTES = 128 * 128;               // M = 128 channel rows x 32 cells (one K-atom): 16 KB
constexpr int DW_MAX_STAGES = 4;                    // narrow layers have small stages: more of them in flight

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
    int f_in, f_out, k_total, np, stages;   // f_out / k_total: rows / columns of THIS launch's dW block; f_in: full input width
    int dz_ld;        // row stride of dy / z (the layer's full f_out; dy, z and the norm coefficients are pre-offset)
    int k_off;        // first column of the block inside [agg | h]
    int out_ld;       // row stride of the partials (the layer's full k_total; `partials` is pre-offset to the block)
    int64_t part_stride;   // floats between the partials of consecutive CTAs (full f_out * full k_total)
    int db_stride;    // doubles between the db partials of consecutive CTAs (full f_out)
    int wpb, active_warps;   // producer warps per 32-channel block; warps that produce (the rest idle)
    float* partials;  // [grid][f_out][k_total]
    double* db_partials;  // [grid][f_out] column sums of dz (bias gradient), may be NULL
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
    sts128(smem_u32(hi) + off, h);
    sts128(smem_u32(lo) + off, l);
}

__global__ void __launch_bounds__(DW_THREADS, 1) dw_tc_kernel(const DwTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[DW_MAX_STAGES], bar_empty[DW_MAX_STAGES], bar_done;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float red_db[DW_NPW][32];                 // per producer warp: column sums of its 32-channel dz block
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b_bytes = p.np * 128;
    const int stage_bytes = 2 * DW_A_BYTES + 2 * b_bytes;
    if (tid == 0) {
        for (int s = 0; s < DW_MAX_STAGES; ++s) { mbar_init(&bar_full[s], p.active_warps); mbar_init(&bar_empty[s], 1); }
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
        uint32_t s = 0, ph = 0;
        for (int64_t i = 0; i < my_groups; ++i) {
            if (lane == 0) {
                mbar_wait(&bar_full[s], ph);
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
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
    } else if (warp < p.active_warps) {
        // Every producer thread owns ONE 4-channel chunk of one 32-channel operand block for the whole kernel
        // (block = warp / wpb), so source pointer, row stride, activation / normalisation coefficients and the
        // destination row are loop invariants; per 32-cell stage it handles the cell groups gsub, gsub + wpb, ...
        const bool relu = p.relu_in != 0;
        const int j = lane & 3, c = lane >> 2;       // cell within a group of 4, 16-byte chunk within the block
        const int nb_a = (p.f_out + 31) >> 5;        // 32-channel blocks of dz (A operand); then np/32 blocks of [agg | h]
        const int wpb = p.wpb;
        const int blk = warp / wpb, gsub = warp % wpb;
        const bool is_a = blk < nb_a;
        const int ch = (is_a ? blk : blk - nb_a) * 32 + c * 4;        // first channel of the chunk in its operand
        // kind: 0 = dz, 1 = agg (raw), 2 = h(x) ; -1 = padding chunk (zeros)
        int kind = -1, stride = 0;
        const float* src = nullptr;
        const float* src2 = nullptr;
        float4 k0 = make_float4(1.f, 1.f, 1.f, 1.f), k1 = make_float4(0.f, 0.f, 0.f, 0.f), k2 = k1, k3 = k1, k4 = k0;
        if (is_a) {
            if (ch < p.f_out) {
                kind = 0; stride = p.dz_ld; src = p.dy + ch;
                if (p.ng != nullptr) {
                    src2 = p.z + ch;
                    k0 = ldg4(p.ng + ch); k1 = ldg4(p.na + ch); k2 = ldg4(p.nb + ch); k3 = ldg4(p.nmean + ch); k4 = ldg4(p.nrstd + ch);
                    // dz = g*dy - (a + (z - m)*rs*b) = g*dy - z*(rs*b) - (a - m*rs*b)
                    k2 = make_float4(k4.x * k2.x, k4.y * k2.y, k4.z * k2.z, k4.w * k2.w);
                    k1 = make_float4(k1.x - k3.x * k2.x, k1.y - k3.y * k2.y, k1.z - k3.z * k2.z, k1.w - k3.w * k2.w);
                }
            }
        } else if (ch < p.k_total) {
            stride = p.f_in;
            const int cha = p.k_off + ch;             // column inside [agg | h]
            if (p.agg != nullptr && cha < p.f_in) { kind = 1; src = p.agg + cha; }
            else {
                const int col = p.agg != nullptr ? cha - p.f_in : cha;
                kind = 2; src = p.x_in + col;
                if (p.in_scale != nullptr) { k0 = ldg4(p.in_scale + col); k1 = ldg4(p.in_shift + col); }
            }
        }
        const bool affine = kind == 2 && p.in_scale != nullptr;
        const bool norm = kind == 0 && p.ng != nullptr;
        const int row = (is_a ? blk : blk - nb_a) * 32 + c * 4 + j;   // operand row this thread writes after the transpose
        const uint32_t dst0 = (is_a ? 0u : 2u * DW_A_BYTES) + (uint32_t)row * 128u;
        const uint32_t lo_off = is_a ? (uint32_t)DW_A_BYTES : (uint32_t)b_bytes;
        const uint32_t rsw = (uint32_t)(row & 7);
        // items (cell groups of 4) per stage: 8 / wpb, processed in batches of <= 4 whose global loads are issued one
        // batch ahead of their use (software pipeline across batches and stages)
        const int n_it = 8 / wpb;
        const int bs = n_it < 4 ? n_it : 4;          // items per batch
        const int bps = n_it / bs;                   // batches per stage (1 or 2)
        const int64_t n_batches = my_groups * bps;
        constexpr int MAXB = 4;
        struct Batch { float4 r0[MAXB], r1[MAXB]; };
        const int n_tgt32 = (int)p.n_tgt;            // < 2^31 cells per GPU (checked by the launcher)
        const int cell_first = (int)(g_begin * DW_CELLS) + j + gsub * 4;
        const int ustep = wpb * 4;                   // cells between consecutive items of a thread
        // batch (stage i, half h): first cell = cell_first + i*32 + h*bs*ustep
        auto load_batch = [&](int tb, bool live, Batch& r) {
#pragma unroll
            for (int u = 0; u < MAXB; ++u) {
                r.r0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                r.r1[u] = r.r0[u];
                if (u < bs && live) {
                    const int t = tb + u * ustep;
                    if (t < n_tgt32) {
                        r.r0[u] = ldg4(src + (size_t)t * stride);
                        if (norm) r.r1[u] = ldg4(src2 + (size_t)t * stride);
                    }
                }
            }
        };
        uint32_t s = 0, ph = 0;
        float4 dbv = make_float4(0.f, 0.f, 0.f, 0.f);  // running sum of this thread's dz values (its 4 channels): db
        Batch cur, nxt;
        const bool on = kind >= 0;
        load_batch(cell_first, on && n_batches > 0, cur);
        int h = 0;
        int tb = cell_first;                         // first cell of the current batch; hstep to the next one
        const int hstep = bs * ustep, sstep = DW_CELLS - (bps - 1) * hstep;
        for (int64_t bi = 0; bi < n_batches; ++bi) {
            const int tb_next = tb + (h + 1 == bps ? sstep : hstep);
            load_batch(tb_next, on && bi + 1 < n_batches, nxt);
            uint8_t* stg = smem + (size_t)s * stage_bytes;
            if (h == 0) mbar_wait(&bar_empty[s], ph ^ 1);
#pragma unroll
            for (int u = 0; u < MAXB; ++u) {
                if (u < bs) {
                    const int g = gsub + (h * bs + u) * wpb;
                    float4 v = cur.r0[u];
                    if (norm) {
                        const float4 zv = cur.r1[u];
                        const bool tv = tb + u * ustep < n_tgt32;  // padding rows must stay zero (k1 != 0)
                        v.x = tv ? fmaf(k0.x, v.x, -fmaf(zv.x, k2.x, k1.x)) : 0.f;
                        v.y = tv ? fmaf(k0.y, v.y, -fmaf(zv.y, k2.y, k1.y)) : 0.f;
                        v.z = tv ? fmaf(k0.z, v.z, -fmaf(zv.z, k2.z, k1.z)) : 0.f;
                        v.w = tv ? fmaf(k0.w, v.w, -fmaf(zv.w, k2.w, k1.w)) : 0.f;
                        dbv.x += v.x; dbv.y += v.y; dbv.z += v.z; dbv.w += v.w;
                    } else if (affine) {
                        const bool tv = tb + u * ustep < n_tgt32;
                        v.x = tv ? act(v.x, k0.x, k1.x, relu) : 0.f; v.y = tv ? act(v.y, k0.y, k1.y, relu) : 0.f;
                        v.z = tv ? act(v.z, k0.z, k1.z, relu) : 0.f; v.w = tv ? act(v.w, k0.w, k1.w, relu) : 0.f;
                    } else if (kind == 2 && relu) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    } else if (kind == 0) {                 // dz = dy (no normalisation): padding rows are zero already
                        dbv.x += v.x; dbv.y += v.y; dbv.z += v.z; dbv.w += v.w;
                    }
                    v = transpose4(v, j);                   // now: channel `row`, cells 4g .. 4g+3
                    const uint32_t off = dst0 + ((((uint32_t)g) ^ rsw) << 4);
                    put_split4(stg + off, stg + off + lo_off, 0u, v);
                }
            }
            tb = tb_next;
            if (++h == bps) {
                h = 0;
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[s]);
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
            cur = nxt;
        }
        // db: sum over the 4 cells of a lane group, then one float4 per (warp, chunk); warps of a block are added below
        dbv.x += __shfl_xor_sync(0xffffffffu, dbv.x, 1); dbv.y += __shfl_xor_sync(0xffffffffu, dbv.y, 1);
        dbv.z += __shfl_xor_sync(0xffffffffu, dbv.z, 1); dbv.w += __shfl_xor_sync(0xffffffffu, dbv.w, 1);
        dbv.x += __shfl_xor_sync(0xffffffffu, dbv.x, 2); dbv.y += __shfl_xor_sync(0xffffffffu, dbv.y, 2);
        dbv.z += __shfl_xor_sync(0xffffffffu, dbv.z, 2); dbv.w += __shfl_xor_sync(0xffffffffu, dbv.w, 2);
        if (j == 0) *reinterpret_cast<float4*>(&red_db[warp][c * 4]) = is_a ? dbv : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (warp < DW_NPW) {
        if (lane < 8) *reinterpret_cast<float4*>(&red_db[warp][lane * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (warp < DW_NPW) {
        // read this CTA's partial out of TMEM
        float* out = p.partials + (size_t)blockIdx.x * p.part_stride;
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        if (my_groups > 0) {
            mbar_wait(&bar_done, 0);
            tc_fence_after_sync();
        }
#pragma unroll 1
        for (int chunk = grp; chunk * 32 < p.np; chunk += DW_NPW / 4) {
            const int c0 = chunk * 32;
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
                    *reinterpret_cast<float4*>(out + (size_t)row * p.out_ld + n) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (p.db_partials != nullptr) {                      // channel ch lives in block ch / 32, produced by wpb warps
        for (int ch = tid; ch < p.f_out; ch += DW_THREADS) {
            double a = 0.0;
            for (int w = 0; w < p.wpb; ++w) a += (double)red_db[(ch >> 5) * p.wpb + w][ch & 31];
            p.db_partials[(size_t)blockIdx.x * p.db_stride + ch] = a;
        }
    }
    if (warp == DW_NPW) tmem_dealloc(tmem_base, 256);
}


// -----------------------------------------------------------------------------------------------------------------------
// dW, second formulation ("row-block" kernel) for layers whose block is the whole matrix (f_out <= 128, k_total <= 256).
//
// A stage is 32 consecutive cells.  Their rows of dy, z, agg and x are CONTIGUOUS in global memory (full-width rows), so
// one warp fetches a whole raw stage with four 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) into a shared-memory
// ring - no per-thread global loads, no address arithmetic in the producers.  The transposition the contraction over
// cells needs is done by ADDRESSING: producer thread = one operand row (= one channel); it reads its channel of 16 cells
// with 16 LDS.32 (lanes = consecutive channels: conflict-free), applies the normalisation backward / the producer
// affine + ReLU with per-thread-constant coefficients, splits hi / lo and writes four 16-byte chunks per image (swizzled:
// conflict-free).  No shuffles, no selects.  The single operand buffer works as two half-stages of 16 cells (k-steps 0-1
// and 2-3 of the 32-cell K-atom): while the tensor core consumes one half the producers fill the other.
// The raw ring is as deep as the shared memory allows (2 stages of 64 KB at 128 -> 128, 3 - 4 for the narrower layers).
// db = column sums of dz: the thread that owns a dz channel keeps its sum in a register for the whole kernel.
struct Dw2Args {
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
    int f_in, f_out, k_total, np;
    float* partials;      // [grid][f_out][k_total]
    double* db_partials;  // [grid][f_out], may be NULL
    int ring;             // raw stages in flight: 2 at 128 -> 128 (64 KB each), up to DW2_MAX_RING for narrower layers
};

constexpr int DW2_MAX_RING = 8;
constexpr int DW2_PW = 12;                          // producer warps: thread = operand row (128 dz rows + up to 256 [agg | h] rows)
constexpr int DW2_THREADS = (DW2_PW + 2) * 32;      // + MMA warp + bulk-copy warp
constexpr int DW2_CELLS = 32;

// FULL = 0: ONE operand buffer worked as two half-stages of 16 cells (all that fits next to the raw ring at 128 -> 128).
// FULL = 1: TWO whole operand buffers, one hand-off per 32-cell stage: a hand-off (operand stores -> fence -> MMA ->
// commit -> producers) measures ~0.6 us whatever the operand size, and at two per stage it, not the bytes, set the pace
// of the narrower layers (5 - 5.7 us per 128 cells at 64 -> 128 and 28 -> 64 against 7.5 us at 128 -> 128).
template <int FULL>
__global__ void __launch_bounds__(DW2_THREADS, 1) dw2_tc_kernel(const Dw2Args p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t raw_full[DW2_MAX_RING], raw_empty[DW2_MAX_RING], op_full[2], op_empty[2], bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool norm = p.ng != nullptr;
    const int b_bytes = p.np * 128;
    // operand buffer: A hi | A lo | B hi | B lo, then the raw ring: per stage dy | z | agg | x
    uint8_t* a_hi = smem;
    uint8_t* b_hi = smem + 2 * DW_A_BYTES;
    const uint32_t dy_bytes = (uint32_t)DW2_CELLS * p.f_out * 4u, x_bytes = (uint32_t)DW2_CELLS * p.f_in * 4u;
    const uint32_t off_z = dy_bytes, off_agg = off_z + (norm ? dy_bytes : 0u), off_x = off_agg + (p.agg ? x_bytes : 0u);
    const uint32_t raw_stage = (off_x + x_bytes + 127u) & ~127u;
    const uint32_t opb = 2u * DW_A_BYTES + 2u * (uint32_t)b_bytes;       // one operand buffer
    uint8_t* raw0 = smem + (FULL ? 2u : 1u) * opb;
    if (tid == 0) {
        for (int i = 0; i < DW2_MAX_RING; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], DW2_PW); }
        for (int i = 0; i < 2; ++i) { mbar_init(&op_full[i], DW2_PW); mbar_init(&op_empty[i], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    // operand rows beyond f_out / k_total are never written but are read by the MMA: zero the whole buffer once
    for (int i = tid; i < (int)((FULL ? 2u : 1u) * opb / 16u); i += DW2_THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    if (warp == DW2_PW) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_groups = (p.n_tgt + DW2_CELLS - 1) / DW2_CELLS;
    const int64_t per = (n_groups + gridDim.x - 1) / gridDim.x;           // contiguous range of 32-cell stages per CTA
    const int64_t g_begin = (int64_t)blockIdx.x * per;
    const int64_t g_end = g_begin + per < n_groups ? g_begin + per : n_groups;
    const int64_t my = g_end > g_begin ? g_end - g_begin : 0;

    if (warp == DW2_PW + 1) {
        // ------------------------------------------------------------ bulk-copy warp: raw stages, one stage ahead
        if (lane == 0) {
            for (int64_t s = 0; s < my; ++s) {
                const int rs = (int)(s % p.ring);
                mbar_wait(&raw_empty[rs], (uint32_t)(((s / p.ring) & 1) ^ 1));
                const int64_t c0 = (g_begin + s) * DW2_CELLS;
                const int64_t left = p.n_tgt - c0;
                const uint32_t rows = (uint32_t)(left < DW2_CELLS ? left : DW2_CELLS);
                uint8_t* dst = raw0 + (size_t)rs * raw_stage;
                const uint32_t nd = rows * (uint32_t)p.f_out * 4u, nx = rows * (uint32_t)p.f_in * 4u;
                mbar_arrive_expect_tx(&raw_full[rs], nd * (norm ? 2u : 1u) + nx * (p.agg ? 2u : 1u));
                bulk_g2s(dst, p.dy + (size_t)c0 * p.f_out, nd, &raw_full[rs]);
                if (norm) bulk_g2s(dst + off_z, p.z + (size_t)c0 * p.f_out, nd, &raw_full[rs]);
                if (p.agg) bulk_g2s(dst + off_agg, p.agg + (size_t)c0 * p.f_in, nx, &raw_full[rs]);
                bulk_g2s(dst + off_x, p.x_in + (size_t)c0 * p.f_in, nx, &raw_full[rs]);
            }
        }
    } else if (warp == DW2_PW) {
        // ------------------------------------------------------------ MMA warp: two half-stages (2 k-steps each) per stage
        const uint32_t idesc = make_idesc_tf32(128, p.np);
        if (lane == 0) {
            const uint32_t ah = smem_u32(a_hi), al = ah + DW_A_BYTES, bh = smem_u32(b_hi), bl = bh + b_bytes;
            for (int64_t s = 0; s < my; ++s) {
                if (FULL) {
                    const uint32_t o = (uint32_t)(s & 1), ob = o * opb;
                    mbar_wait(&op_full[o], (uint32_t)((s >> 1) & 1));
                    tc_fence_after_sync();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t ko = ob + kk * 32;
                        mma_tf32(tmem_base, make_desc(ah + ko), make_desc(bh + ko), idesc, (s > 0 || kk > 0) ? 1u : 0u);
                        mma_tf32(tmem_base, make_desc(al + ko), make_desc(bh + ko), idesc, 1u);
                        mma_tf32(tmem_base, make_desc(ah + ko), make_desc(bl + ko), idesc, 1u);
                    }
                    mma_commit(&op_empty[o]);
                } else {
#pragma unroll
                    for (int hs = 0; hs < 2; ++hs) {
                        mbar_wait(&op_full[hs], (uint32_t)(s & 1));
                        tc_fence_after_sync();
#pragma unroll
                        for (int kk = 2 * hs; kk < 2 * hs + 2; ++kk) {
                            const uint32_t ko = kk * 32;
                            mma_tf32(tmem_base, make_desc(ah + ko), make_desc(bh + ko), idesc, (s > 0 || kk > 0) ? 1u : 0u);
                            mma_tf32(tmem_base, make_desc(al + ko), make_desc(bh + ko), idesc, 1u);
                            mma_tf32(tmem_base, make_desc(ah + ko), make_desc(bl + ko), idesc, 1u);
                        }
                        mma_commit(&op_empty[hs]);
                    }
                }
                if (s == my - 1) mma_commit(&bar_done);
            }
        }
    } else {
        // ------------------------------------------------------------ producers: thread = operand row
        const int row_a = tid;                               // dz channel (A rows) for tid < 128
        const int row_b = tid - 128;                         // [agg | h] channel (B rows) for tid >= 128
        const bool is_a = tid < 128;
        const bool live = is_a ? row_a < p.f_out : row_b < p.k_total;
        // source of this row inside a raw stage, and its coefficients (loop invariants)
        uint32_t src_off = 0, src2_off = 0, stride = 0;
        float c0 = 1.f, c1 = 0.f, c2 = 0.f;                  // dz = c0*dy - z*c2 - c1   |   h = relu?(x*c0 + c1)
        int kind = -1;                                       // 0 dz, 1 agg (raw), 2 h(x)
        if (live) {
            if (is_a) {
                kind = 0; stride = (uint32_t)p.f_out * 4u; src_off = (uint32_t)row_a * 4u; src2_off = off_z + (uint32_t)row_a * 4u;
                if (norm) {
                    const float g = __ldg(p.ng + row_a), a = __ldg(p.na + row_a), b = __ldg(p.nb + row_a);
                    const float m = __ldg(p.nmean + row_a), rs = __ldg(p.nrstd + row_a);
                    c0 = g; c2 = rs * b; c1 = a - m * c2;    // g*dy - (a + (z - m)*rs*b)
                }
            } else {
                stride = (uint32_t)p.f_in * 4u;
                if (p.agg != nullptr && row_b < p.f_in) { kind = 1; src_off = off_agg + (uint32_t)row_b * 4u; }
                else {
                    const int col = p.agg != nullptr ? row_b - p.f_in : row_b;
                    kind = 2; src_off = off_x + (uint32_t)col * 4u;
                    if (p.in_scale != nullptr) { c0 = __ldg(p.in_scale + col); c1 = __ldg(p.in_shift + col); }
                }
            }
        }
        const bool relu = kind == 2 && p.relu_in != 0;
        const bool affine = kind == 2 && p.in_scale != nullptr;
        const int row = is_a ? row_a : row_b;
        const uint32_t dst_hi = smem_u32(is_a ? a_hi : b_hi) + (uint32_t)row * 128u;
        const uint32_t lo_off = is_a ? (uint32_t)DW_A_BYTES : (uint32_t)b_bytes;
        const uint32_t rsw = (uint32_t)(row & 7);
        const uint32_t raw_u32 = smem_u32(raw0);
        float dbsum = 0.f;
        for (int64_t s = 0; s < my; ++s) {
            const int rs = (int)(s % p.ring);
            const int64_t cbase = (g_begin + s) * DW2_CELLS;
            const int64_t left = p.n_tgt - cbase;
            const int nvalid = (int)(left < DW2_CELLS ? left : DW2_CELLS);
            mbar_wait(&raw_full[rs], (uint32_t)((s / p.ring) & 1));
            const uint32_t rbase = raw_u32 + (uint32_t)rs * raw_stage;
#pragma unroll
            for (int hs = 0; hs < 2; ++hs) {
                float v[16];
                if (live) {
                    const uint32_t a0 = rbase + src_off + (uint32_t)(16 * hs) * stride;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float d;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d) : "r"(a0 + (uint32_t)i * stride));
                        v[i] = d;
                    }
                    if (kind == 0 && norm) {
                        const uint32_t z0 = rbase + src2_off + (uint32_t)(16 * hs) * stride;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float zz;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(zz) : "r"(z0 + (uint32_t)i * stride));
                            v[i] = fmaf(c0, v[i], -fmaf(zz, c2, c1));
                        }
                    } else if (affine) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = act(v[i], c0, c1, relu);
                    } else if (relu) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (nvalid < 16 * hs + 16) {             // tail stage: cells beyond n_tgt contribute nothing
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (16 * hs + i >= nvalid) v[i] = 0.f;
                    }
                    if (kind == 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) dbsum += v[i];
                    }
                }
                const uint32_t ob = FULL ? (uint32_t)(s & 1) * opb : 0u;
                if (FULL) { if (hs == 0) mbar_wait(&op_empty[s & 1], (uint32_t)(((s >> 1) & 1) ^ 1)); }
                else mbar_wait(&op_empty[hs], (uint32_t)((s & 1) ^ 1));
                if (live) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float4 h, l;
                        split_tf32(v[4 * c], h.x, l.x); split_tf32(v[4 * c + 1], h.y, l.y);
                        split_tf32(v[4 * c + 2], h.z, l.z); split_tf32(v[4 * c + 3], h.w, l.w);
                        const uint32_t a = dst_hi + ob + ((((uint32_t)(4 * hs + c)) ^ rsw) << 4);
                        sts128(a, h);
                        sts128(a + lo_off, l);
                    }
                }
                if (!FULL || hs == 1) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&op_full[FULL ? (int)(s & 1) : hs]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[rs]);      // every lane has read its share of the raw stage
        }
        // db partial of this CTA (one thread per dz channel, no reduction needed)
        if (p.db_partials != nullptr && is_a && row_a < p.f_out)
            p.db_partials[(size_t)blockIdx.x * p.f_out + row_a] = (double)dbsum;
        // read this CTA's partial out of TMEM: warp w reads TMEM lanes 32 (w & 3) .., column chunks (w >> 2), + 3, ..
        float* out = p.partials + (size_t)blockIdx.x * p.f_out * p.k_total;
        const int q = warp & 3, grp = warp >> 2;
        const int orow = q * 32 + lane;
        if (my > 0) {
            mbar_wait(&bar_done, 0);
            tc_fence_after_sync();
        }
#pragma unroll 1
        for (int chunk = grp; chunk * 32 < p.np; chunk += DW2_PW / 4) {
            const int col0 = chunk * 32;
            float v[32];
            if (my > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (orow < p.f_out) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int n = col0 + i;
                    if (n >= p.k_total) continue;
                    *reinterpret_cast<float4*>(out + (size_t)orow * p.k_total + n) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == DW2_PW) tmem_dealloc(tmem_base, 256);
}

}  // namespace dgnn

using namespace dgnn;

static inline int ceil32i(int x) { return (x + 31) / 32 * 32; }

extern "C" int dgnn_dw_tc_supported(int f_out, int k_total) {
    return (f_out % 4 == 0 && k_total % 4 == 0 && f_out > 0 && k_total > 0) ? 1 : 0;   // any width: [128 x 256] blocks
}

extern "C" int dgnn_dw_bwd_tc(const float* dy, const float* z, const float* g, const float* a, const float* b,
                              const float* mean, const float* rstd, const float* agg, const float* x_in,
                              const float* in_scale, const float* in_shift, int relu_in, int64_t n_tgt, int f_in,
                              int f_out, int k_total, float* partials, double* db_partials, void* stream) {
    DGNN_REQUIRE(dgnn_dw_tc_supported(f_out, k_total), "widths not supported by the tensor-core dW kernel");
    DGNN_REQUIRE(k_total == (agg ? 2 * f_in : f_in), "k_total mismatch");
    DGNN_REQUIRE(dy && x_in && partials, "null pointer");
    DGNN_REQUIRE(n_tgt < (int64_t)1 << 31, "more than 2^31 cells on one GPU");
#ifndef DGNN_DW_OLD
    if (f_out <= 128 && k_total <= 256 && (f_out % 4) == 0 && (f_in % 4) == 0) {
        // row-block kernel: the block is the whole matrix, rows are fetched as contiguous bulk copies
        Dw2Args q;
        memset(&q, 0, sizeof(q));
        q.dy = dy; q.z = g ? z : nullptr; q.ng = g; q.na = a; q.nb = b; q.nmean = mean; q.nrstd = rstd;
        q.agg = agg; q.x_in = x_in; q.in_scale = in_scale; q.in_shift = in_shift; q.relu_in = relu_in;
        q.n_tgt = n_tgt; q.f_in = f_in; q.f_out = f_out; q.k_total = k_total; q.np = ceil32i(k_total);
        q.partials = partials; q.db_partials = db_partials;
        const size_t raw = (((size_t)DW2_CELLS * 4 * ((size_t)f_out * (g ? 2 : 1) + (size_t)f_in * (agg ? 2 : 1))) + 127) & ~(size_t)127;
        // the kernel runs at (bytes in flight) / (copy round trip): as many raw stages as the shared memory holds
        const size_t opb = 2 * DW_A_BYTES + 2 * (size_t)q.np * 128;
#ifndef DGNN_DW2_HALF_ONLY
        const bool full = 2 * opb + 2 * raw + 1024 <= 226 * 1024;         // two whole operand buffers + >= 2 raw stages fit
#else
        const bool full = false;
#endif
        const size_t fixed = (full ? 2 : 1) * opb + 1024;
        int ring = fixed + 2 * raw <= 226 * 1024 ? (int)((226 * 1024 - fixed) / raw) : 0;
        if (ring > DW2_MAX_RING) ring = DW2_MAX_RING;
        q.ring = ring;
        const size_t smem = fixed + (size_t)ring * raw;
        if (ring >= 2) {
            if (full) {
                if (int rc_ = ensure_dyn_smem((const void*)dw2_tc_kernel<1>, 226 * 1024, "dgnn_dw_bwd_tc")) return rc_;
                dw2_tc_kernel<1><<<sm_count(), DW2_THREADS, smem, as_stream(stream)>>>(q);
            } else {
                if (int rc_ = ensure_dyn_smem((const void*)dw2_tc_kernel<0>, 226 * 1024, "dgnn_dw_bwd_tc")) return rc_;
                dw2_tc_kernel<0><<<sm_count(), DW2_THREADS, smem, as_stream(stream)>>>(q);
            }
            return check_launch("dgnn_dw_bwd_tc");
        }
    }
#endif
    if (int rc_ = ensure_dyn_smem((const void*)dw_tc_kernel, 210 * 1024, "dgnn_dw_bwd_tc")) return rc_;
    DwTcArgs p;
    memset(&p, 0, sizeof(p));
    p.agg = agg; p.x_in = x_in; p.in_scale = in_scale; p.in_shift = in_shift; p.relu_in = relu_in;
    p.n_tgt = n_tgt; p.f_in = f_in;
    p.dz_ld = f_out; p.out_ld = k_total; p.part_stride = (int64_t)f_out * k_total; p.db_stride = f_out;
    for (int m0 = 0; m0 < f_out; m0 += 128) {
        const int mw = f_out - m0 < 128 ? f_out - m0 : 128;
        p.f_out = mw;
        p.dy = dy + m0; p.z = z ? z + m0 : nullptr;
        p.ng = g ? g + m0 : nullptr; p.na = a ? a + m0 : nullptr; p.nb = b ? b + m0 : nullptr;
        p.nmean = mean ? mean + m0 : nullptr; p.nrstd = rstd ? rstd + m0 : nullptr;
        for (int n0 = 0; n0 < k_total; n0 += 256) {
            const int nw = k_total - n0 < 256 ? k_total - n0 : 256;
            p.k_total = nw; p.k_off = n0;
            p.np = ceil32i(nw);
            {   // as many 32-cell stages as fit 200 KB (2 at 128 -> 128, 4 for the narrow layers)
                const int stage_bytes = 2 * DW_A_BYTES + 2 * p.np * 128;
                int st = (200 * 1024) / stage_bytes;
                p.stages = st > DW_MAX_STAGES ? DW_MAX_STAGES : (st < 2 ? 2 : st);
            }
            p.partials = partials + (size_t)m0 * k_total + n0;
            p.db_partials = (db_partials && n0 == 0) ? db_partials + m0 : nullptr;
            const int nb = (mw + 31) / 32 + p.np / 32;      // operand blocks of 32 channels (<= 12)
            p.wpb = nb * 4 <= DW_NPW ? 4 : (nb * 2 <= DW_NPW ? 2 : 1);
            p.active_warps = nb * p.wpb;
            size_t smem = (size_t)p.stages * (2 * DW_A_BYTES + 2 * p.np * 128) + 1024;
            dw_tc_kernel<<<sm_count(), DW_THREADS, smem, as_stream(stream)>>>(p);
            if (int rc = check_launch("dgnn_dw_bwd_tc")) return rc;
        }
    }
    return 0;
}
