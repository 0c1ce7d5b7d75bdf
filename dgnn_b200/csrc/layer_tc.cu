// Tensor-core (tcgen05 + TMEM) version of the fused layer kernels.  One launch covers an N slice of at most
// 256 output columns (one UMMA tile, two TMEM accumulator buffers); wider layers (modelnet.yaml's 512 / 1024) are
// run as several column slices by the launch wrappers, the contraction length K (streamed K-atoms) is unbounded.
//
//   forward :  z = [agg | h] . [W_j | W_i]^T       (dgnn_layer_fwd_tc, gather or dense mode)
//   backward:  [d_agg | d_self] = dz . [W_j | W_i]  (dgnn_dense_bwd_tc)
//
// One persistent CTA per SM, warp-specialised, no CTA-wide barrier in the main loop:
//   * 16 producer warps own tiles of 128 cells (UMMA M = 128).  The A operand never exists in
//     global memory: per K-atom (32 features) each warp produces its 8 rows of the [128 x 32]
//     slice — gathering neighbour rows and applying the edge filter (forward), or applying the
//     normalisation backward to dy (backward) — splits it into TF32 hi/lo parts and stores it
//     straight into the 128B-swizzled UMMA layout of a ring stage, then arrives on the stage's
//     `full` mbarrier.  Fast warps run ahead by the ring depth.
//   * 1 TMA warp bulk-copies (cp.async.bulk) every atom's pre-packed weight slice next to the A slice,
//     running a ring ahead; 1 MMA warp waits for `full`, issues the 12 tcgen05.mma of the stage
//     (3xTF32: hi*hi + lo*hi + hi*lo, 4 k-steps) accumulating in TMEM, and tcgen05.commit's the
//     stage's `empty` mbarrier.
//   * accumulators are double-buffered in TMEM (2 x N columns): the producer warps run the
//     epilogue of tile i (tcgen05.ld -> bias / affine / ReLU / BN partials -> global) after they
//     have produced tile i+1, and hand the buffer back through `acc_free`.
#include <cuda.h>
#include <stdlib.h>

#include "umma.cuh"
#include "common.cuh"

namespace dgnn {

using namespace umma;

constexpr int TC_M = 128;
constexpr int NPW = 16;                        // producer warps
constexpr int TC_THREADS = (NPW + 2) * 32;     // + 1 MMA warp + 1 TMA (weight slice) warp
constexpr int A_ATOM_BYTES = TC_M * ATOM_ROW_BYTES;  // 16 KB
constexpr int MAX_STAGES = 4;
constexpr int RAW_DEPTH = 4;                   // dense forward: raw input atoms staged by cp.async this many atoms ahead

enum { MODE_FWD_DENSE = 0, MODE_FWD_GATHER = 1, MODE_BWD = 2 };

struct TcArgs {
    // forward producer
    const float* agg_in;   // dense mode: first ka_agg atoms come from this matrix (no activation), may be NULL
    const float* x_in;
    const float* in_scale;
    const float* in_shift;
    int relu_in;
    const int32_t* nbr;
    const float* ea;
    const float* w_e;
    const float* b_e;
    // backward producer: dz = g*dy - (a + xhat*b)
    const float* dy;
    const float* z;
    const float* ng;
    const float* na;
    const float* nb;
    const float* nmean;
    const float* nrstd;
    // operand B
    const float* b_packed;  // [KA][2][NP][32] swizzled atoms (hi, lo)
    int ka;                 // K-atoms
    int ka_agg;             // atoms of the agg segment (gather mode), 0 otherwise
    int np;                 // padded N of this launch's column slice (multiple of 32, <= 256)
    int stages;
    int n_off;              // backward: first column of the slice inside [d_agg | d_self]
    int n_total;            // backward: real columns of [d_agg | d_self] (2 f_in or f_in)
    int out_ld;             // forward: row stride of `out` (the layer's full f_out; `out`, bias, out_scale ... are pre-offset)
    int stats_ld;           // forward: full f_out (stride between the sum and sum^2 blocks of `stats`, pre-offset)
    // sizes
    int64_t n_tgt;
    int f_in;   // forward: input width; backward: width of d_agg / d_self
    int f_out;  // forward: output width of the slice (N); backward: K (width of dy)
    // forward epilogue
    const float* bias;
    const float* out_scale;
    const float* out_shift;
    int relu_out;
    float* out;
    float* agg_save;
    double* stats;
    // backward epilogue
    float* d_agg;
    float* d_self;
    double* db_partials;
};

__device__ __forceinline__ void store_split2(uint8_t* a_hi, uint8_t* a_lo, int r, int k, float v0, float v1) {
    float h0, l0, h1, l1;
    split_tf32(v0, h0, l0);
    split_tf32(v1, h1, l1);
    uint32_t off = atom_off(r, k);
    sts64(smem_u32(a_hi) + off, make_float2(h0, h1));
    sts64(smem_u32(a_lo) + off, make_float2(l0, l1));
}
__device__ __forceinline__ void store_split4(uint8_t* a_hi, uint8_t* a_lo, int r, int k, float4 v) {
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    uint32_t off = atom_off(r, k);
    sts128(smem_u32(a_hi) + off, h);
    sts128(smem_u32(a_lo) + off, l);
}

// warp `w` fills its 8 rows of the atom with h(x_in[row, f0 .. f0+32))
__device__ __forceinline__ void produce_rows(const TcArgs& p, uint8_t* a_hi, uint8_t* a_lo, int64_t tile0, int f0,
                                             int warp, int lane, bool raw_agg = false) {
    const bool relu = p.relu_in != 0 && !raw_agg;
    const int c = (lane & 7) * 4;
    const int f = f0 + c;
    const float* src = raw_agg ? p.agg_in : p.x_in;
    const bool affine = p.in_scale != nullptr && !raw_agg;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (affine && f < p.f_in) { sc = ldg4(p.in_scale + f); sh = ldg4(p.in_shift + f); }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int r = warp * 8 + (lane >> 3) + it * 4;
        const int64_t t = tile0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < p.n_tgt && f < p.f_in) {
            v = ldg4(src + (size_t)t * p.f_in + f);
            if (affine) {
                v.x = act(v.x, sc.x, sh.x, relu); v.y = act(v.y, sc.y, sh.y, relu);
                v.z = act(v.z, sc.z, sh.z, relu); v.w = act(v.w, sc.w, sh.w, relu);
            } else if (relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
        }
        store_split4(a_hi, a_lo, r, c, v);
    }
}

// ---- two-phase producers (dense / dz): loads are issued one atom ahead of their use -------------------
struct RawAtom {
    float4 v[2];   // x / agg / dy for the thread's two rows
    float4 z[2];   // z (backward with a normalisation)
};

template <int MODE>
__device__ __forceinline__ void load_raw(const TcArgs& p, int64_t tile0, int a, int warp, int lane, RawAtom& r) {
    const int c = (lane & 7) * 4;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int64_t t = tile0 + warp * 8 + (lane >> 3) + it * 4;
        r.v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        r.z[it] = r.v[it];
        if (t >= p.n_tgt) continue;
        if (MODE == MODE_BWD) {
            const int f = a * ATOM_K + c;
            if (f < p.f_out) {
                r.v[it] = ldg4(p.dy + (size_t)t * p.f_out + f);
                if (p.ng != nullptr) r.z[it] = ldg4(p.z + (size_t)t * p.f_out + f);
            }
        } else {
            const bool raw_agg = a < p.ka_agg;
            const int f = (raw_agg ? a : a - p.ka_agg) * ATOM_K + c;
            if (f < p.f_in) r.v[it] = ldg4((raw_agg ? p.agg_in : p.x_in) + (size_t)t * p.f_in + f);
        }
    }
}

// dense forward: the thread's two 16-byte pieces of raw atom `a` go global -> shared by cp.async into the slot the same
// thread reads back later (thread-private: no cross-thread synchronisation, no registers held while in flight)
__device__ __forceinline__ void issue_raw_fwd(const TcArgs& p, uint32_t slot, int64_t tile0, int a, int warp, int lane) {
    const int c = (lane & 7) * 4;
    const bool raw_agg = a < p.ka_agg;
    const int f = (raw_agg ? a : a - p.ka_agg) * ATOM_K + c;
    const float* src = raw_agg ? p.agg_in : p.x_in;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int row = warp * 8 + (lane >> 3) + it * 4;
        const int64_t t = tile0 + row;
        if (t < p.n_tgt && f < p.f_in) {
            const uint32_t dst = slot + atom_off(row, c);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + (size_t)t * p.f_in + f));
        }
    }
    cp_async_commit();
}
__device__ __forceinline__ void read_raw_fwd(uint32_t slot, int warp, int lane, RawAtom& r) {
    const int c = (lane & 7) * 4;
#pragma unroll
    for (int it = 0; it < 2; ++it) r.v[it] = lds128(slot + atom_off(warp * 8 + (lane >> 3) + it * 4, c));
}

template <int MODE>
__device__ __forceinline__ void store_raw(const TcArgs& p, uint8_t* a_hi, uint8_t* a_lo, int64_t tile0, int a, int warp,
                                          int lane, const RawAtom& r, float* red_db) {
    const int c = (lane & 7) * 4;
    if (MODE == MODE_BWD) {
        const int f = a * ATOM_K + c;
        const bool norm = p.ng != nullptr && f < p.f_out;
        float4 g = make_float4(1.f, 1.f, 1.f, 1.f), aa = make_float4(0.f, 0.f, 0.f, 0.f), b = aa, m = aa, rs = g;
        if (norm) { g = ldg4(p.ng + f); aa = ldg4(p.na + f); b = ldg4(p.nb + f); m = ldg4(p.nmean + f); rs = ldg4(p.nrstd + f); }
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int row = warp * 8 + (lane >> 3) + it * 4;
            const int64_t t = tile0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < p.n_tgt && f < p.f_out) {
                const float4 d = r.v[it], zv = r.z[it];
                if (norm) {
                    v.x = g.x * d.x - (aa.x + (zv.x - m.x) * rs.x * b.x);
                    v.y = g.y * d.y - (aa.y + (zv.y - m.y) * rs.y * b.y);
                    v.z = g.z * d.z - (aa.z + (zv.z - m.z) * rs.z * b.z);
                    v.w = g.w * d.w - (aa.w + (zv.w - m.w) * rs.w * b.w);
                } else {
                    v = d;
                }
                cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
            }
            store_split4(a_hi, a_lo, row, c, v);
        }
        if (red_db != nullptr) {
            // column sums of dz (-> db).  Only used when the dW kernel (which owns db at f_out <= 128, dw_tc.cu) is not.
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
            if (lane < 8 && f < p.f_out) {
                atomicAdd(&red_db[f], cs.x); atomicAdd(&red_db[f + 1], cs.y);
                atomicAdd(&red_db[f + 2], cs.z); atomicAdd(&red_db[f + 3], cs.w);
            }
        }
    } else {
        const bool raw_agg = a < p.ka_agg;
        const int f = (raw_agg ? a : a - p.ka_agg) * ATOM_K + c;
        const bool relu = p.relu_in != 0 && !raw_agg;
        const bool affine = p.in_scale != nullptr && !raw_agg && f < p.f_in;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (affine) { sc = ldg4(p.in_scale + f); sh = ldg4(p.in_shift + f); }
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int row = warp * 8 + (lane >> 3) + it * 4;
            const int64_t t = tile0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < p.n_tgt && f < p.f_in) {
                v = r.v[it];
                if (affine) {
                    v.x = act(v.x, sc.x, sh.x, relu); v.y = act(v.y, sc.y, sh.y, relu);
                    v.z = act(v.z, sc.z, sh.z, relu); v.w = act(v.w, sc.w, sh.w, relu);
                } else if (relu) {
                    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                }
            }
            store_split4(a_hi, a_lo, row, c, v);
        }
    }
}

// warp `w` fills its 8 rows of the atom with agg[row, f0 .. f0+32): 16 lanes per cell, 2 features per
// lane; the 4 passes are software-pipelined (indices for all passes up front, neighbour rows one pass ahead)
template <int FE>
__device__ __forceinline__ void produce_agg(const TcArgs& p, uint8_t* a_hi, uint8_t* a_lo, int64_t tile0, int f0,
                                            int warp, int lane) {
    const int sub = lane >> 4, li = lane & 15;
    const int F = p.f_in;
    const int f = f0 + li * 2;
    const bool fv = f < F;
    const bool relu = p.relu_in != 0;
    float we0[FE > 0 ? FE : 1], we1[FE > 0 ? FE : 1];
    float be0 = 1.f, be1 = 1.f;
    if (FE > 0) {
#pragma unroll
        for (int j = 0; j < FE; ++j) {
            we0[j] = fv ? __ldg(p.w_e + (size_t)f * FE + j) : 0.f;
            we1[j] = fv ? __ldg(p.w_e + (size_t)(f + 1) * FE + j) : 0.f;
        }
        be0 = fv ? __ldg(p.b_e + f) : 0.f;
        be1 = fv ? __ldg(p.b_e + f + 1) : 0.f;
    }
    float sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f;
    if (p.in_scale != nullptr && fv) {
        sc0 = __ldg(p.in_scale + f); sc1 = __ldg(p.in_scale + f + 1);
        sh0 = __ldg(p.in_shift + f); sh1 = __ldg(p.in_shift + f + 1);
    }
    // rows of this warp: 8w .. 8w+7; pass i handles rows 8w + 2i + sub
    int4 nbs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t t = tile0 + warp * 8 + i * 2 + sub;
        nbs[i] = make_int4(-1, -1, -1, -1);
        if (t < p.n_tgt) nbs[i] = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
    }
    float2 xs[4], xn[4];
    {
        const int nbv[4] = {nbs[0].x, nbs[0].y, nbs[0].z, nbs[0].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xs[k] = make_float2(0.f, 0.f);
            if (nbv[k] >= 0 && fv) xs[k] = ldg2(p.x_in + (size_t)nbv[k] * F + f);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int cell = warp * 8 + i * 2 + sub;
        const int64_t t = tile0 + cell;
        const int nbv[4] = {nbs[i].x, nbs[i].y, nbs[i].z, nbs[i].w};
        if (i < 3) {
            const int nn[4] = {nbs[i + 1].x, nbs[i + 1].y, nbs[i + 1].z, nbs[i + 1].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xn[k] = make_float2(0.f, 0.f);
                if (nn[k] >= 0 && fv) xn[k] = ldg2(p.x_in + (size_t)nn[k] * F + f);
            }
        }
        float a0 = 0.f, a1 = 0.f;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (nbv[k] < 0) continue;
            ++cnt;
            float ph0 = be0, ph1 = be1;
            if (FE > 0) {
                const float* er = p.ea + ((size_t)t * 4 + k) * FE;
#pragma unroll
                for (int j = 0; j < FE; j += 4) {
                    float4 e = ldg4(er + j);
                    ph0 = fmaf(we0[j], e.x, ph0); ph1 = fmaf(we1[j], e.x, ph1);
                    ph0 = fmaf(we0[j + 1], e.y, ph0); ph1 = fmaf(we1[j + 1], e.y, ph1);
                    ph0 = fmaf(we0[j + 2], e.z, ph0); ph1 = fmaf(we1[j + 2], e.z, ph1);
                    ph0 = fmaf(we0[j + 3], e.w, ph0); ph1 = fmaf(we1[j + 3], e.w, ph1);
                }
            }
            float h0 = act(xs[k].x, sc0, sh0, relu), h1 = act(xs[k].y, sc1, sh1, relu);
            a0 = fmaf(h0, ph0, a0);
            a1 = fmaf(h1, ph1, a1);
        }
        float d = (float)(cnt > 0 ? cnt : 1);
        a0 = fv ? a0 / d : 0.f;
        a1 = fv ? a1 / d : 0.f;
        store_split2(a_hi, a_lo, cell, li * 2, a0, a1);
        if (p.agg_save != nullptr && t < p.n_tgt && fv)
            *reinterpret_cast<float2*>(p.agg_save + (size_t)t * F + f) = make_float2(a0, a1);
        if (i < 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) xs[k] = xn[k];
        }
    }
}

// butterfly transpose-reduce: on return lane l holds the sum over the warp's 32 rows of column l
__device__ __forceinline__ float warp_colsum32(float (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float mine = up ? v[i + off] : v[i];
            float theirs = up ? v[i] : v[i + off];
            v[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, off);
        }
    }
    return v[0];
}

// one producer warp's share of a tile epilogue: TMEM lanes 32q..32q+31 (rows), column chunks grp, grp+4, ..
// st_sum / st_sq: per-lane running column sums (chunk slot j = (chunk - grp)/4), kept across tiles
// (q = TMEM lane quarter of the calling warp, grp of NGRP = which of the warps sharing the quarter: chunks grp, grp + NGRP, ..)
template <int MODE, int NGRP = 4>
__device__ __forceinline__ void epilogue(const TcArgs& p, uint32_t tmem_acc, int64_t tile0, int warp, int lane,
                                         double (&st_sum)[8 / NGRP], double (&st_sq)[8 / NGRP], const int4 nb) {
    const int q = warp & 3, grp = (warp >> 2) % NGRP;
    const int r = q * 32 + lane;
    const int64_t t = tile0 + r;
    const bool tv = t < p.n_tgt;
    const int n_real = MODE == MODE_BWD ? p.n_total - p.n_off : p.f_out;   // real columns from the slice start
    float icnt = 1.f;
    if (MODE == MODE_BWD && p.nbr != nullptr && tv) {   // nb = nbr[t], loaded by the caller a tile earlier
        int cnt = (nb.x >= 0) + (nb.y >= 0) + (nb.z >= 0) + (nb.w >= 0);
        icnt = 1.f / (float)(cnt > 0 ? cnt : 1);
    }
    const bool want_stats = MODE != MODE_BWD && p.stats != nullptr;
#pragma unroll
    for (int j = 0; j < 8 / NGRP; ++j) {
        const int chunk = grp + NGRP * j;
        const int c0 = chunk * 32;
        if (c0 >= p.np) break;
        float v[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (MODE != MODE_BWD) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const int n = c0 + i;
                if (n >= n_real || p.bias == nullptr) continue;
                float4 bi = ldg4(p.bias + n);
                v[i] += bi.x; v[i + 1] += bi.y; v[i + 2] += bi.z; v[i + 3] += bi.w;
            }
            if (tv) {
                const bool al8 = (p.out_ld & 7) == 0;
                float4 held = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int n = c0 + i;
                    if (n >= n_real) continue;
                    float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    if (p.out_scale != nullptr) {
                        float4 os = ldg4(p.out_scale + n), oh = ldg4(p.out_shift + n);
                        o.x = fmaf(o.x, os.x, oh.x); o.y = fmaf(o.y, os.y, oh.y);
                        o.z = fmaf(o.z, os.z, oh.z); o.w = fmaf(o.w, os.w, oh.w);
                    }
                    if (p.relu_out) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    if (!al8) *reinterpret_cast<float4*>(p.out + (size_t)t * p.out_ld + n) = o;
                    else if ((i & 4) == 0) held = o;                      // pair two float4 into one 32-byte store
                    else stg8(p.out + (size_t)t * p.out_ld + n - 4, held, o);
                }
            }
            if (want_stats) {
                float sq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!tv) v[i] = 0.f;
                    sq[i] = v[i] * v[i];
                }
                st_sum[j] += (double)warp_colsum32(v);
                st_sq[j] += (double)warp_colsum32(sq);
            }
        } else if (tv) {
            // columns [0, f_in) of the tile are d_agg (x 1/cnt), [f_in, 2 f_in) d_self.  A 32-column chunk lies entirely in one
            // of them at the 32-multiple widths; its row pointer is formed once (not re-read from the constant bank per store)
            const bool has_agg = p.nbr != nullptr;
            const int ca = p.n_off + c0;                // column of the chunk inside [d_agg | d_self]
            const bool all_agg = has_agg && ca + 32 <= p.f_in;
            const bool all_self = !has_agg || ca >= p.f_in;
            if ((all_agg || all_self) && c0 + 32 <= n_real && (p.f_in & 7) == 0) {
                float* dst = (all_agg ? p.d_agg : p.d_self) + (size_t)t * p.f_in + (all_agg ? ca : (has_agg ? ca - p.f_in : ca));
                const float s = all_agg ? icnt : 1.f;
#pragma unroll
                for (int i = 0; i < 32; i += 8)
                    stg8(dst + i, make_float4(v[i] * s, v[i + 1] * s, v[i + 2] * s, v[i + 3] * s),
                         make_float4(v[i + 4] * s, v[i + 5] * s, v[i + 6] * s, v[i + 7] * s));
            } else {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int n = c0 + i;
                    if (n >= n_real) continue;
                    const int na = p.n_off + n;
                    const bool is_agg = has_agg && na < p.f_in;
                    float s = is_agg ? icnt : 1.f;
                    float* dst = is_agg ? p.d_agg : p.d_self;
                    int col = is_agg ? na : (has_agg ? na - p.f_in : na);
                    *reinterpret_cast<float4*>(dst + (size_t)t * p.f_in + col) =
                        make_float4(v[i] * s, v[i + 1] * s, v[i + 2] * s, v[i + 3] * s);
                }
            }
        }
    }
}

template <int MODE, int FE>
__global__ void __launch_bounds__(TC_THREADS, 1) layer_tc_kernel(const TcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES];
    __shared__ uint64_t bar_acc_full[2], bar_acc_free[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float red_db[256];
    __shared__ double red_st[2 * 256];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int b_atom_bytes = p.np * ATOM_ROW_BYTES;
    const int stage_bytes = 2 * A_ATOM_BYTES + 2 * b_atom_bytes;

    if (tid == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(&bar_full[s], NPW + 1);   // 16 producer warps + the TMA arrive.expect_tx
            mbar_init(&bar_empty[s], 1);        // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_acc_full[b], 1);     // tcgen05.commit
            mbar_init(&bar_acc_free[b], NPW);   // every producer warp after its epilogue share
        }
        fence_barrier_init();
    }
    for (int c = tid; c < 256; c += TC_THREADS) red_db[c] = 0.f;
    for (int c = tid; c < 512; c += TC_THREADS) red_st[c] = 0.0;
    if (warp == NPW) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_tiles = (p.n_tgt + TC_M - 1) / TC_M;

    if (warp == NPW) {
        // ------------------------------------------------------------------ MMA / TMA warp
        const uint32_t idesc = make_idesc_tf32(TC_M, p.np);
        uint32_t tile_cnt = 0, s = 0, use = 0;       // ring stage and how often it has been used (parity = use & 1)
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            const uint32_t acc = tile_cnt & 1;
            const uint32_t tmem_acc = tmem_base + acc * (uint32_t)p.np;
            for (int a = 0; a < p.ka; ++a) {
                if (lane == 0) {
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    uint8_t* b_hi = st + 2 * A_ATOM_BYTES;
                    if (a == 0 && tile_cnt >= 2) mbar_wait(&bar_acc_free[acc], ((tile_cnt >> 1) - 1) & 1);
                    mbar_wait(&bar_full[s], use & 1);
                    tc_fence_after_sync();
                    const uint32_t ah = smem_u32(st), al = ah + A_ATOM_BYTES, bh = smem_u32(b_hi), bl = bh + b_atom_bytes;
#pragma unroll
                    for (int kk = 0; kk < ATOM_K / 8; ++kk) {
                        const uint32_t ko = kk * 32;
                        mma_tf32(tmem_acc, make_desc(ah + ko), make_desc(bh + ko), idesc, (a > 0 || kk > 0) ? 1u : 0u);
                        mma_tf32(tmem_acc, make_desc(al + ko), make_desc(bh + ko), idesc, 1u);
                        mma_tf32(tmem_acc, make_desc(ah + ko), make_desc(bl + ko), idesc, 1u);
                    }
                    mma_commit(&bar_empty[s]);
                    if (a == p.ka - 1) mma_commit(&bar_acc_full[acc]);
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ++use; }
            }
        }
    } else if (warp == NPW + 1) {
        // ------------------------------------------------------------------ TMA warp: weight slices, a ring ahead
        uint32_t s = 0, use = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int a = 0; a < p.ka; ++a) {
                if (lane == 0) {
                    uint8_t* b_hi = smem + (size_t)s * stage_bytes + 2 * A_ATOM_BYTES;
                    mbar_wait(&bar_empty[s], (use & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar_full[s], 2u * (uint32_t)b_atom_bytes);
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.b_packed) + (size_t)a * 2 * b_atom_bytes;
                    bulk_g2s(b_hi, src, (uint32_t)b_atom_bytes, &bar_full[s]);
                    bulk_g2s(b_hi + b_atom_bytes, src + b_atom_bytes, (uint32_t)b_atom_bytes, &bar_full[s]);
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ++use; }
            }
        }
    } else {
        // ------------------------------------------------------------------ producer warps
        double st_sum[2] = {0.0, 0.0}, st_sq[2] = {0.0, 0.0};
        uint32_t tile_cnt = 0, s = 0, use = 0;
        int64_t prev_tile0 = -1;
        RawAtom cur, nxt;                              // raw rows of the atom being stored / the next one
        if (MODE == MODE_BWD && (int64_t)blockIdx.x < n_tiles) load_raw<MODE>(p, (int64_t)blockIdx.x * TC_M, 0, warp, lane, cur);
        // dense forward: cp.async ring of raw atoms behind the operand stages, RAW_DEPTH - 1 atoms ahead of the stores
        const uint32_t raw_base = smem_u32(smem) + (uint32_t)p.stages * (uint32_t)stage_bytes;
        int64_t ld_tile = blockIdx.x;
        int ld_a = 0;
        uint32_t ld_i = 0, st_i = 0;
        auto issue_next = [&]() {
            if (ld_tile < n_tiles) issue_raw_fwd(p, raw_base + (ld_i % RAW_DEPTH) * (uint32_t)A_ATOM_BYTES, ld_tile * TC_M, ld_a, warp, lane);
            else cp_async_commit();
            ++ld_i;
            if (++ld_a == p.ka) { ld_a = 0; ld_tile += gridDim.x; }
        };
        if (MODE == MODE_FWD_DENSE)
            for (int i = 0; i < RAW_DEPTH - 1; ++i) issue_next();
        int4 nb_epi = make_int4(-1, -1, -1, -1);
        auto load_nb_epi = [&](int64_t t0) {           // the epilogue's row of the ELL table (1/cnt of d_agg)
            if (MODE == MODE_BWD && p.nbr != nullptr) {
                const int64_t t = t0 + (warp & 3) * 32 + lane;
                if (t < p.n_tgt) nb_epi = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
            }
        };
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            const int64_t tile0 = tile * TC_M;
#ifndef DGNN_NO_NB_PREFETCH
            if (prev_tile0 >= 0) load_nb_epi(prev_tile0);   // in flight while this tile is produced
#endif
            for (int a = 0; a < p.ka; ++a) {
                uint8_t* a_hi = smem + (size_t)s * stage_bytes;
                uint8_t* a_lo = a_hi + A_ATOM_BYTES;
                if (MODE == MODE_FWD_DENSE) {
#ifdef DGNN_RAW_LATE
                    cp_async_wait<RAW_DEPTH - 2>();      // the oldest of the RAW_DEPTH - 1 groups in flight has landed
#else
                    issue_next();
                    cp_async_wait<RAW_DEPTH - 1>();
#endif
                    read_raw_fwd(raw_base + (st_i % RAW_DEPTH) * (uint32_t)A_ATOM_BYTES, warp, lane, cur);
                    ++st_i;
                }
                if (MODE == MODE_BWD) {
                    // loads of the next atom (possibly of the next tile) go in flight before this one is processed
                    const bool last = a + 1 == p.ka;
                    const int64_t ntile = last ? tile + gridDim.x : tile;
                    if (ntile < n_tiles) load_raw<MODE>(p, ntile * TC_M, last ? 0 : a + 1, warp, lane, nxt);
                }
                mbar_wait(&bar_empty[s], (use & 1) ^ 1);
                if (MODE == MODE_FWD_GATHER) {
                    if (a < p.ka_agg) produce_agg<FE>(p, a_hi, a_lo, tile0, a * ATOM_K, warp, lane);
                    else produce_rows(p, a_hi, a_lo, tile0, (a - p.ka_agg) * ATOM_K, warp, lane);
                } else {
                    store_raw<MODE>(p, a_hi, a_lo, tile0, a, warp, lane, cur, (MODE == MODE_BWD && p.db_partials) ? red_db : nullptr);
                    if (MODE == MODE_BWD) cur = nxt;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[s]);
#ifdef DGNN_RAW_LATE     // variant: next raw atom issued after the fence.proxy.async (= MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC); measured: no gain
                if (MODE == MODE_FWD_DENSE) issue_next();
#endif
                if (++s == (uint32_t)p.stages) { s = 0; ++use; }
            }
            if (prev_tile0 >= 0) {
                const uint32_t pacc = (tile_cnt - 1) & 1;
                mbar_wait(&bar_acc_full[pacc], ((tile_cnt - 1) >> 1) & 1);
                tc_fence_after_sync();
#ifdef DGNN_NO_NB_PREFETCH
                load_nb_epi(prev_tile0);
#endif
                epilogue<MODE>(p, tmem_base + pacc * (uint32_t)p.np, prev_tile0, warp, lane, st_sum, st_sq, nb_epi);
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_acc_free[pacc]);
            }
            prev_tile0 = tile0;
        }
        if (prev_tile0 >= 0) {
            const uint32_t pacc = (tile_cnt - 1) & 1;
            load_nb_epi(prev_tile0);
            mbar_wait(&bar_acc_full[pacc], ((tile_cnt - 1) >> 1) & 1);
            tc_fence_after_sync();
            epilogue<MODE>(p, tmem_base + pacc * (uint32_t)p.np, prev_tile0, warp, lane, st_sum, st_sq, nb_epi);
        }
        if (MODE != MODE_BWD && p.stats != nullptr) {
            const int grp = warp >> 2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = (grp + 4 * j) * 32 + lane;
                if (c < p.f_out) {
                    atomicAdd(&red_st[c], st_sum[j]);
                    atomicAdd(&red_st[256 + c], st_sq[j]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (MODE != MODE_BWD && p.stats != nullptr) {
        double* my = p.stats + (size_t)blockIdx.x * 2 * p.stats_ld;
        for (int c = tid; c < p.f_out; c += TC_THREADS) {
            my[c] = red_st[c];
            my[p.stats_ld + c] = red_st[256 + c];
        }
    }
    if (MODE == MODE_BWD && p.db_partials != nullptr) {
        double* my = p.db_partials + (size_t)blockIdx.x * p.f_out;
        for (int c = tid; c < p.f_out; c += TC_THREADS) my[c] = (double)red_db[c];
    }
    if (warp == NPW) tmem_dealloc(tmem_base, 512);
}

// -----------------------------------------------------------------------------------------------------------------------
// Dense forward / backward, second formulation: the raw atoms arrive by TMA TENSOR loads.
//
// The producers of layer_tc_kernel fetch every raw atom themselves (cp.async or register prefetch, one 16-byte piece per
// thread with its own address arithmetic) and the kernels measured latency-bound on exactly those loads (ncu: long
// scoreboard 2.6 / 6.7 stalled warps per issue).  Here one thread issues, per K-atom, ONE cp.async.bulk.tensor.2d
// (SASS UTMALDG) of the box [128 rows x 32 floats] of agg / x / dy / z through a tensor map with the 128-byte swizzle:
// the atom lands in shared memory already in the UMMA operand layout, rows and columns beyond the matrix zero-filled,
// several atoms ahead of its consumers and without a single instruction in the producer warps.
//   * The landing slot IS the hi operand: kind::tf32 reads the 19 high bits of an operand, so an atom that needs no
//     arithmetic (the agg half of the forward) is used as it lies and the producers only add lo = v - trunc_tf32(v); atoms
//     that need the producer affine + ReLU (forward) or the normalisation backward (dy, z -> dz) are rewritten in place
//     (hi) and into the neighbouring slot (lo; in the backward that slot held z).
//   * Two rings: A stages (raw -> hi | lo, 32 KB) filled by the TMA-A warp, weight-slice stages filled by the TMA-B warp
//     from L2; the rings advance independently, so a slow weight slice does not hold back the activation stream.
//   * MMA warp, double-buffered TMEM accumulators and the epilogue are those of layer_tc_kernel.
constexpr int D2_NPW = 8, D2_NEW = 8;         // producer warps (operand conversion), epilogue warps (2 per TMEM lane quarter)
constexpr int D2_WARPS = D2_NPW + D2_NEW;
constexpr int D2_THREADS = (D2_WARPS + 3) * 32;     // + MMA warp + TMA-A warp + TMA-B warp
constexpr int D2_MAX_A = 4, D2_MAX_B = 4;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Epilogue of dense2_tc_kernel: one warp's share (TMEM lane quarter q, column chunks grp, grp + 2, ..) of a tile goes
// TMEM -> registers -> (bias / eval affine + ReLU / BN partials | 1/cnt) -> a 4 KB swizzled staging block in shared memory
// -> ONE TMA tensor store of the [32 rows x 32 columns] box (SASS UTMASTG): full 128-byte lines instead of one 32-byte
// sector per lane and instruction (the per-row stores of layer_tc_kernel's epilogue cost 4096 separate requests per
// backward tile and bounded the kernel); rows / columns beyond the matrix are clipped by the tensor map.
template <int MODE>
__device__ __forceinline__ void epilogue2(const TcArgs& p, const CUtensorMap* tmo0, const CUtensorMap* tmo1, uint32_t tmem_acc,
                                          int64_t tile0, int warp, int lane, uint32_t stage, double (&st_sum)[4],
                                          double (&st_sq)[4], const int4 nb) {
    const int q = warp & 3, grp = (warp >> 2) & 1;
    const int64_t t = tile0 + q * 32 + lane;
    const bool tv = t < p.n_tgt;
    float icnt = 1.f;
    if (MODE == MODE_BWD && p.nbr != nullptr && tv) {
        int cnt = (nb.x >= 0) + (nb.y >= 0) + (nb.z >= 0) + (nb.w >= 0);
        icnt = 1.f / (float)(cnt > 0 ? cnt : 1);
    }
    const bool want_stats = MODE != MODE_BWD && p.stats != nullptr;
    const uint32_t srow = stage + (uint32_t)lane * 128u, sw = (uint32_t)lane & 7u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c0 = (grp + 2 * j) * 32;
        if (c0 >= p.np) break;
        float v[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        const CUtensorMap* tm = tmo0;
        int col = c0;
        if (MODE != MODE_BWD) {
            if (p.bias != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (c0 + i >= p.f_out) continue;
                    const float4 bi = ldg4(p.bias + c0 + i);
                    v[i] += bi.x; v[i + 1] += bi.y; v[i + 2] += bi.z; v[i + 3] += bi.w;
                }
            }
        } else {
            const int ca = p.n_off + c0;               // column of the chunk inside [d_agg | d_self]
            const bool is_agg = p.nbr != nullptr && ca < p.f_in;
            if (is_agg) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= icnt;
            } else {
                tm = tmo1;
                col = p.nbr != nullptr ? ca - p.f_in : ca;
            }
            if (is_agg) col = ca;
        }
        if (lane == 0) tma_store_wait_read();          // the previous box has left the staging block
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (MODE != MODE_BWD) {
                if (p.out_scale != nullptr && c0 + i < p.f_out) {
                    const float4 os = ldg4(p.out_scale + c0 + i), oh = ldg4(p.out_shift + c0 + i);
                    o.x = fmaf(o.x, os.x, oh.x); o.y = fmaf(o.y, os.y, oh.y);
                    o.z = fmaf(o.z, os.z, oh.z); o.w = fmaf(o.w, os.w, oh.w);
                }
                if (p.relu_out) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            }
            sts128(srow + ((((uint32_t)i >> 2) ^ sw) << 4), o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tma_store_2d(tm, col, (int)(tile0 + q * 32), stage);
        if (want_stats) {
            float sq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (!tv) v[i] = 0.f;
                sq[i] = v[i] * v[i];
            }
            st_sum[j] += (double)warp_colsum32(v);
            st_sq[j] += (double)warp_colsum32(sq);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(D2_THREADS, 1) dense2_tc_kernel(const TcArgs p, const __grid_constant__ CUtensorMap tm0,
                                                                  const __grid_constant__ CUtensorMap tm1,
                                                                  const __grid_constant__ CUtensorMap tmo0,
                                                                  const __grid_constant__ CUtensorMap tmo1, int a_stages,
                                                                  int b_stages) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[D2_MAX_A], a_ready[D2_MAX_A], a_empty[D2_MAX_A], b_full[D2_MAX_B], b_empty[D2_MAX_B];
    __shared__ uint64_t bar_acc_full[2], bar_acc_free[2];
    __shared__ uint32_t tmem_slot;
    __shared__ double red_st[MODE == MODE_BWD ? 2 : 2 * 256];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int b_atom_bytes = p.np * ATOM_ROW_BYTES;
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = a_base + (uint32_t)a_stages * 2u * A_ATOM_BYTES;
    // coefficient table behind the rings: forward (scale | shift) per column of the x segment, backward (c0 | c1 | c2)
    // per channel with dz = c0 * dy - (z * c2 + c1); padded columns are the identity
    // behind the rings: the epilogue warps' staging blocks (4 KB each), then the coefficient table
    const uint32_t rings = (uint32_t)a_stages * 2u * A_ATOM_BYTES + (uint32_t)b_stages * 2u * (uint32_t)b_atom_bytes;
    const uint32_t stage_base = a_base + rings;
    float* tab = reinterpret_cast<float*>(smem + rings + D2_NEW * 4096);
    const int ka_x = p.ka - p.ka_agg;
    const int tab_n = (MODE == MODE_BWD ? p.ka : ka_x) * ATOM_K;
    const bool norm = MODE == MODE_BWD && p.ng != nullptr;

    if (tid == 0) {
        for (int s = 0; s < D2_MAX_A; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_ready[s], D2_NPW); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < D2_MAX_B; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_free[b], D2_NEW); }
        fence_barrier_init();
    }
    if (MODE != MODE_BWD)
        for (int c = tid; c < 512; c += D2_THREADS) red_st[c] = 0.0;
    for (int c = tid; c < tab_n; c += D2_THREADS) {
        if (MODE == MODE_BWD) {
            float c0 = 1.f, c1 = 0.f, c2 = 0.f;
            if (norm && c < p.f_out) {
                const float g = __ldg(p.ng + c), a = __ldg(p.na + c), b = __ldg(p.nb + c);
                const float m = __ldg(p.nmean + c), rs = __ldg(p.nrstd + c);
                c0 = g; c2 = rs * b; c1 = a - m * c2;
            }
            tab[c] = c0; tab[tab_n + c] = c1; tab[2 * tab_n + c] = c2;
        } else {
            const bool on = p.in_scale != nullptr && c < p.f_in;
            tab[c] = on ? __ldg(p.in_scale + c) : 1.f;
            tab[tab_n + c] = on ? __ldg(p.in_shift + c) : 0.f;
        }
    }
    if (warp == D2_WARPS) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_tiles = (p.n_tgt + TC_M - 1) / TC_M;

    if (warp == D2_WARPS) {
        // ------------------------------------------------------------------ MMA warp
        const uint32_t idesc = make_idesc_tf32(TC_M, p.np);
        uint32_t tile_cnt = 0, sa = 0, ua = 0, sb = 0, ub = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            const uint32_t acc = tile_cnt & 1;
            const uint32_t tmem_acc = tmem_base + acc * (uint32_t)p.np;
            for (int a = 0; a < p.ka; ++a) {
                if (lane == 0) {
                    if (a == 0 && tile_cnt >= 2) mbar_wait(&bar_acc_free[acc], ((tile_cnt >> 1) - 1) & 1);
                    mbar_wait(&b_full[sb], ub & 1);
                    mbar_wait(&a_ready[sa], ua & 1);
                    tc_fence_after_sync();
                    const uint32_t ah = a_base + sa * 2u * A_ATOM_BYTES, al = ah + A_ATOM_BYTES;
                    const uint32_t bh = b_base + sb * 2u * (uint32_t)b_atom_bytes, bl = bh + (uint32_t)b_atom_bytes;
#pragma unroll
                    for (int kk = 0; kk < ATOM_K / 8; ++kk) {
                        const uint32_t ko = kk * 32;
                        mma_tf32(tmem_acc, make_desc(ah + ko), make_desc(bh + ko), idesc, (a > 0 || kk > 0) ? 1u : 0u);
                        mma_tf32(tmem_acc, make_desc(al + ko), make_desc(bh + ko), idesc, 1u);
                        mma_tf32(tmem_acc, make_desc(ah + ko), make_desc(bl + ko), idesc, 1u);
                    }
                    mma_commit(&a_empty[sa]);
                    mma_commit(&b_empty[sb]);
                    if (a == p.ka - 1) mma_commit(&bar_acc_full[acc]);
                }
                __syncwarp();
                if (++sa == (uint32_t)a_stages) { sa = 0; ++ua; }
                if (++sb == (uint32_t)b_stages) { sb = 0; ++ub; }
            }
        }
    } else if (warp == D2_WARPS + 1) {
        // ------------------------------------------------------------------ TMA-A warp: raw atoms, a ring ahead
        if (lane == 0) {
            uint32_t sa = 0, ua = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int row0 = (int)(tile * TC_M);
                for (int a = 0; a < p.ka; ++a) {
                    const uint32_t slot = a_base + sa * 2u * A_ATOM_BYTES;
                    mbar_wait(&a_empty[sa], (ua & 1) ^ 1);
                    if (MODE == MODE_BWD) {
                        mbar_arrive_expect_tx(&a_full[sa], norm ? 2u * A_ATOM_BYTES : (uint32_t)A_ATOM_BYTES);
                        tma_load_2d(slot, &tm0, a * ATOM_K, row0, &a_full[sa]);
                        if (norm) tma_load_2d(slot + A_ATOM_BYTES, &tm1, a * ATOM_K, row0, &a_full[sa]);
                    } else {
                        mbar_arrive_expect_tx(&a_full[sa], (uint32_t)A_ATOM_BYTES);
                        if (a < p.ka_agg) tma_load_2d(slot, &tm0, a * ATOM_K, row0, &a_full[sa]);
                        else tma_load_2d(slot, &tm1, (a - p.ka_agg) * ATOM_K, row0, &a_full[sa]);
                    }
                    if (++sa == (uint32_t)a_stages) { sa = 0; ++ua; }
                }
            }
        }
    } else if (warp == D2_WARPS + 2) {
        // ------------------------------------------------------------------ TMA-B warp: weight slices (L2), a ring ahead
        if (lane == 0) {
            uint32_t sb = 0, ub = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int a = 0; a < p.ka; ++a) {
                    uint8_t* b_hi = smem + (size_t)a_stages * 2 * A_ATOM_BYTES + (size_t)sb * 2 * b_atom_bytes;
                    mbar_wait(&b_empty[sb], (ub & 1) ^ 1);
#ifdef DGNN_D2_NOB      // timing experiment: weight slices loaded once per ring slot only (wrong results)
                    if (ub > 0) { mbar_arrive(&b_full[sb]); if (++sb == (uint32_t)b_stages) { sb = 0; ++ub; } continue; }
#endif
                    mbar_arrive_expect_tx(&b_full[sb], 2u * (uint32_t)b_atom_bytes);
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.b_packed) + (size_t)a * 2 * b_atom_bytes;
                    bulk_g2s(b_hi, src, 2u * (uint32_t)b_atom_bytes, &b_full[sb]);
                    if (++sb == (uint32_t)b_stages) { sb = 0; ++ub; }
                }
            }
        }
    } else if (warp >= D2_NPW) {
        // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter warp & 3)
        // tile i's accumulator is read out while the producers and the tensor core are busy with tile i + 1
        double st_sum[4] = {0.0, 0.0, 0.0, 0.0}, st_sq[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t tile_cnt = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            const int64_t tile0 = tile * TC_M;
            int4 nb_epi = make_int4(-1, -1, -1, -1);
            if (MODE == MODE_BWD && p.nbr != nullptr) {  // the row of the ELL table (1 / cnt of d_agg), in flight during the wait
                const int64_t t = tile0 + (warp & 3) * 32 + lane;
                if (t < p.n_tgt) nb_epi = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
            }
            const uint32_t acc = tile_cnt & 1;
            mbar_wait(&bar_acc_full[acc], (tile_cnt >> 1) & 1);
            tc_fence_after_sync();
#ifndef DGNN_D2_NOEPI   // timing experiment: no epilogue at all (no results)
            epilogue2<MODE>(p, &tmo0, &tmo1, tmem_base + acc * (uint32_t)p.np, tile0, warp, lane,
                            stage_base + (uint32_t)(warp - D2_NPW) * 4096u, st_sum, st_sq, nb_epi);
#endif
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_free[acc]);
        }
        if (lane == 0) tma_store_wait_all();
        if (MODE != MODE_BWD && p.stats != nullptr) {
            const int grp = (warp >> 2) & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = (grp + 2 * j) * 32 + lane;
                if (c < p.f_out) {
                    atomicAdd(&red_st[c], st_sum[j]);
                    atomicAdd(&red_st[256 + c], st_sq[j]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ producer warps
        uint32_t sa = 0, ua = 0;
        // this thread's four 16-byte pieces of every atom: rows 16 warp + (lane >> 3) + 4 i, chunk lane & 7
        const uint32_t r0 = (uint32_t)(warp * 16 + (lane >> 3)), ch = (uint32_t)(lane & 7);
        uint32_t off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) off[i] = (r0 + 4u * i) * 128u + ((ch ^ ((r0 + 4u * i) & 7u)) << 4);
        const uint32_t tab_u32 = smem_u32(tab) + ch * 16u;
        const bool relu = p.relu_in != 0;
        const bool affine = p.in_scale != nullptr;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int a = 0; a < p.ka; ++a) {
                const uint32_t hi = a_base + sa * 2u * A_ATOM_BYTES, lo = hi + A_ATOM_BYTES;
                mbar_wait(&a_full[sa], ua & 1);
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = lds128(hi + off[i]);
                if (MODE == MODE_BWD) {
                    if (norm) {
                        const uint32_t t = tab_u32 + (uint32_t)a * 128u;
                        const float4 c0 = lds128(t), c1 = lds128(t + (uint32_t)tab_n * 4u), c2 = lds128(t + (uint32_t)tab_n * 8u);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 z = lds128(lo + off[i]);
                            v[i].x = fmaf(c0.x, v[i].x, -fmaf(z.x, c2.x, c1.x)); v[i].y = fmaf(c0.y, v[i].y, -fmaf(z.y, c2.y, c1.y));
                            v[i].z = fmaf(c0.z, v[i].z, -fmaf(z.z, c2.z, c1.z)); v[i].w = fmaf(c0.w, v[i].w, -fmaf(z.w, c2.w, c1.w));
                        }
                    }
                } else if (a >= p.ka_agg) {
                    if (affine) {
                        const uint32_t t = tab_u32 + (uint32_t)(a - p.ka_agg) * 128u;
                        const float4 sc = lds128(t), sh = lds128(t + (uint32_t)tab_n * 4u);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            v[i].x = act(v[i].x, sc.x, sh.x, relu); v[i].y = act(v[i].y, sc.y, sh.y, relu);
                            v[i].z = act(v[i].z, sc.z, sh.z, relu); v[i].w = act(v[i].w, sc.w, sh.w, relu);
                        }
                    } else if (relu) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f);
                        }
                    }
                }
                const bool in_place = MODE != MODE_BWD ? (a < p.ka_agg || (!affine && !relu)) : !norm;
                if (in_place) {
                    // the raw atom is the hi operand (the tensor core drops the 13 low mantissa bits): only lo is written
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 l;
                        // lo = v - trunc_tf32(v) is up to 2^-10 |v|: rounded (not truncated by the tensor core) to TF32, so the
                        // pair carries v to 2^-21 like the hi = rna_tf32(v) split does
                        l.x = tf32_rna(v[i].x - __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u));
                        l.y = tf32_rna(v[i].y - __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u));
                        l.z = tf32_rna(v[i].z - __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u));
                        l.w = tf32_rna(v[i].w - __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u));
                        sts128(lo + off[i], l);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 h, l;
                        split_tf32(v[i].x, h.x, l.x); split_tf32(v[i].y, h.y, l.y); split_tf32(v[i].z, h.z, l.z); split_tf32(v[i].w, h.w, l.w);
                        sts128(hi + off[i], h);
                        sts128(lo + off[i], l);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[sa]);
                if (++sa == (uint32_t)a_stages) { sa = 0; ++ua; }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (MODE != MODE_BWD && p.stats != nullptr) {
        double* my = p.stats + (size_t)blockIdx.x * 2 * p.stats_ld;
        for (int c = tid; c < p.f_out; c += D2_THREADS) {
            my[c] = red_st[c];
            my[p.stats_ld + c] = red_st[256 + c];
        }
    }
    if (warp == D2_WARPS) tmem_dealloc(tmem_base, 512);
}

// -----------------------------------------------------------------------------------------------------------------------
// Dense forward, third formulation: the A operand lives in TENSOR MEMORY.  (Built with -DDGNN_DENSE3; not the default.)
//
// Measured on B200 (128 -> 128, 604 913 cells): 7.9 us per 128-cell tile against 8.2 us for dense2_tc_kernel, and the
// knock-out runs of THIS kernel (macros below) say why it is not more: no raw-atom loads 7.9 us, no weight-slice loads
// 7.9 us, no epilogue 7.1 us, ONE tcgen05.mma per K-atom instead of twelve 5.0 us.  The MMA stream of the one issuing
// thread is the bound: tools/bench_umma measures 216 - 245 cycles per tcgen05.mma issued from one thread whatever N, data
// type, swizzle or A source (122 / 61 from two / four issuing warps), and a tile needs 96 of them.  Several issuing warps
// were tried in dense2_tc_kernel (7.7 us) and dropped: MMAs of different threads are unordered, so the fp32 accumulation
// order - and the last bits of the result - changed from run to run, and the repo guarantees bit-identical launches.
//
// dense2_tc_kernel is bound by shared-memory bandwidth: a tf32 MMA with both operands in shared memory reads 128 B/clk at
// N = 128 - all there is - and 3xTF32 triples those reads.  tcgen05.mma takes A from TMEM instead: the producers read the
// TMA-landed raw atom once (thread = one row of the atom and half of its 32 columns), apply the affine + ReLU, split hi / lo
// and write both with tcgen05.st (lane = row, column = k).  The raw slot is released as soon as the producers hold it in
// registers, the MMAs read only the weight slice from shared memory, nothing is written back to it.
//   TMEM (512 columns): two accumulators of <= 128 columns, then a ring of 4 A stages of 64 columns (hi | lo).
//   Shared memory: raw landing ring (16 KB per atom), weight-slice ring, epilogue staging, coefficient table.
constexpr int D3_TSTAGES = 4;                  // A stages in TMEM
constexpr int D3_MAX_RAW = 8;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, one K = 8 step; single thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(D2_THREADS, 1) dense3_fwd_kernel(const TcArgs p, const __grid_constant__ CUtensorMap tm0,
                                                                   const __grid_constant__ CUtensorMap tm1,
                                                                   const __grid_constant__ CUtensorMap tmo0, int raw_stages,
                                                                   int b_stages) {
    constexpr int MODE = MODE_FWD_DENSE;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t raw_full[D3_MAX_RAW], raw_empty[D3_MAX_RAW], at_full[D3_TSTAGES], at_empty[D3_TSTAGES];
    __shared__ uint64_t b_full[D2_MAX_B], b_empty[D2_MAX_B], bar_acc_full[2], bar_acc_free[2];
    __shared__ uint32_t tmem_slot;
    __shared__ double red_st[2 * 256];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int b_atom_bytes = p.np * ATOM_ROW_BYTES;
    const uint32_t raw_base = smem_u32(smem);
    const uint32_t b_base = raw_base + (uint32_t)raw_stages * A_ATOM_BYTES;
    const uint32_t rings = (uint32_t)raw_stages * A_ATOM_BYTES + (uint32_t)b_stages * 2u * (uint32_t)b_atom_bytes;
    const uint32_t stage_base = raw_base + rings;
    float* tab = reinterpret_cast<float*>(smem + rings + D2_NEW * 4096);
    const int ka_x = p.ka - p.ka_agg;
    const int tab_n = ka_x * ATOM_K;

    if (tid == 0) {
        for (int s = 0; s < D3_MAX_RAW; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], D2_NPW); }
        for (int s = 0; s < D3_TSTAGES; ++s) { mbar_init(&at_full[s], D2_NPW); mbar_init(&at_empty[s], 1); }
        for (int s = 0; s < D2_MAX_B; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_free[b], D2_NEW); }
        fence_barrier_init();
    }
    for (int c = tid; c < 512; c += D2_THREADS) red_st[c] = 0.0;
    for (int c = tid; c < tab_n; c += D2_THREADS) {
        const bool on = p.in_scale != nullptr && c < p.f_in;
        tab[c] = on ? __ldg(p.in_scale + c) : 1.f;
        tab[tab_n + c] = on ? __ldg(p.in_shift + c) : 0.f;
    }
    if (warp == D2_WARPS) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem_a0 = tmem_base + 256u;                 // A stages behind the two accumulators
    const int64_t n_tiles = (p.n_tgt + TC_M - 1) / TC_M;

    if (warp == D2_WARPS) {
        // ------------------------------------------------------------------ MMA warp
        const uint32_t idesc = make_idesc_tf32(TC_M, p.np);
        uint32_t tile_cnt = 0, ts = 0, ut = 0, sb = 0, ub = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            const uint32_t acc = tile_cnt & 1;
            const uint32_t tmem_acc = tmem_base + acc * 128u;
            for (int a = 0; a < p.ka; ++a) {
                if (lane == 0) {
                    if (a == 0 && tile_cnt >= 2) mbar_wait(&bar_acc_free[acc], ((tile_cnt >> 1) - 1) & 1);
                    mbar_wait(&b_full[sb], ub & 1);
                    mbar_wait(&at_full[ts], ut & 1);
                    tc_fence_after_sync();
                    const uint32_t ah = tmem_a0 + ts * 64u, al = ah + 32u;
                    const uint32_t bh = b_base + sb * 2u * (uint32_t)b_atom_bytes, bl = bh + (uint32_t)b_atom_bytes;
#ifdef DGNN_D3_NOMMA    // timing experiment: one MMA per atom instead of twelve (wrong results)
                    for (int kk = 0; kk < 1; ++kk) {
                        const uint32_t ko = kk * 32, kc = kk * 8;
                        mma_tf32_ts(tmem_acc, ah + kc, make_desc(bh + ko), idesc, (a > 0 || kk > 0) ? 1u : 0u);
                    }
                    if (false)
#endif
#pragma unroll
                    for (int kk = 0; kk < ATOM_K / 8; ++kk) {
                        const uint32_t ko = kk * 32, kc = kk * 8;
                        mma_tf32_ts(tmem_acc, ah + kc, make_desc(bh + ko), idesc, (a > 0 || kk > 0) ? 1u : 0u);
                        mma_tf32_ts(tmem_acc, al + kc, make_desc(bh + ko), idesc, 1u);
                        mma_tf32_ts(tmem_acc, ah + kc, make_desc(bl + ko), idesc, 1u);
                    }
                    mma_commit(&at_empty[ts]);
                    mma_commit(&b_empty[sb]);
                    if (a == p.ka - 1) mma_commit(&bar_acc_full[acc]);
                }
                __syncwarp();
                if (++ts == D3_TSTAGES) { ts = 0; ++ut; }
                if (++sb == (uint32_t)b_stages) { sb = 0; ++ub; }
            }
        }
    } else if (warp == D2_WARPS + 1) {
        // ------------------------------------------------------------------ TMA-A warp: raw atoms, a ring ahead
        if (lane == 0) {
            uint32_t rs = 0, ur = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int row0 = (int)(tile * TC_M);
                for (int a = 0; a < p.ka; ++a) {
                    mbar_wait(&raw_empty[rs], (ur & 1) ^ 1);
#ifdef DGNN_D3_NOA      // timing experiment: no raw atoms loaded (wrong results)
                    mbar_arrive(&raw_full[rs]);
#else
                    mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)A_ATOM_BYTES);
                    if (a < p.ka_agg) tma_load_2d(raw_base + rs * A_ATOM_BYTES, &tm0, a * ATOM_K, row0, &raw_full[rs]);
                    else tma_load_2d(raw_base + rs * A_ATOM_BYTES, &tm1, (a - p.ka_agg) * ATOM_K, row0, &raw_full[rs]);
#endif
                    if (++rs == (uint32_t)raw_stages) { rs = 0; ++ur; }
                }
            }
        }
    } else if (warp == D2_WARPS + 2) {
        // ------------------------------------------------------------------ TMA-B warp: weight slices (L2), a ring ahead
        if (lane == 0) {
            uint32_t sb = 0, ub = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int a = 0; a < p.ka; ++a) {
                    uint8_t* b_hi = smem + (size_t)raw_stages * A_ATOM_BYTES + (size_t)sb * 2 * b_atom_bytes;
                    mbar_wait(&b_empty[sb], (ub & 1) ^ 1);
#ifdef DGNN_D3_NOB      // timing experiment: no weight slices loaded (wrong results)
                    mbar_arrive(&b_full[sb]); (void)b_hi;
#else
                    mbar_arrive_expect_tx(&b_full[sb], 2u * (uint32_t)b_atom_bytes);
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.b_packed) + (size_t)a * 2 * b_atom_bytes;
                    bulk_g2s(b_hi, src, 2u * (uint32_t)b_atom_bytes, &b_full[sb]);
#endif
                    if (++sb == (uint32_t)b_stages) { sb = 0; ++ub; }
                }
            }
        }
    } else if (warp >= D2_NPW) {
        // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter warp & 3)
        double st_sum[4] = {0.0, 0.0, 0.0, 0.0}, st_sq[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t tile_cnt = 0;
        const int4 nb_epi = make_int4(-1, -1, -1, -1);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            const int64_t tile0 = tile * TC_M;
            const uint32_t acc = tile_cnt & 1;
            mbar_wait(&bar_acc_full[acc], (tile_cnt >> 1) & 1);
            tc_fence_after_sync();
#ifndef DGNN_D3_NOEPI   // timing experiment: no epilogue (no results)
            epilogue2<MODE>(p, &tmo0, &tmo0, tmem_base + acc * 128u, tile0, warp, lane,
                            stage_base + (uint32_t)(warp - D2_NPW) * 4096u, st_sum, st_sq, nb_epi);
#endif
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_free[acc]);
        }
        if (lane == 0) tma_store_wait_all();
        if (p.stats != nullptr) {
            const int grp = (warp >> 2) & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = (grp + 2 * j) * 32 + lane;
                if (c < p.f_out) {
                    atomicAdd(&red_st[c], st_sum[j]);
                    atomicAdd(&red_st[256 + c], st_sq[j]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ producer warps: raw atom -> TMEM (hi | lo)
        // thread = row 32 q + lane of the atom (q = warp & 3 = the TMEM lane quarter this warp may touch) and the 16 columns
        // of half = warp >> 2
        const int q = warp & 3, half = warp >> 2;
        const uint32_t r = (uint32_t)(q * 32 + lane);
        uint32_t off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) off[i] = r * 128u + ((((uint32_t)(4 * half + i)) ^ (r & 7u)) << 4);
        const uint32_t tab_u32 = smem_u32(tab) + (uint32_t)half * 64u;
        const bool relu = p.relu_in != 0;
        const bool affine = p.in_scale != nullptr;
        const uint32_t t_lane = tmem_a0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * half);
        uint32_t rs = 0, ur = 0, ts = 0, ut = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int a = 0; a < p.ka; ++a) {
                const uint32_t slot = raw_base + rs * A_ATOM_BYTES;
                mbar_wait(&raw_full[rs], ur & 1);
                float v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 x4 = lds128(slot + off[i]);
                    v[4 * i] = x4.x; v[4 * i + 1] = x4.y; v[4 * i + 2] = x4.z; v[4 * i + 3] = x4.w;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_empty[rs]);          // the atom is in registers: the slot may be refilled
                if (a >= p.ka_agg) {
                    if (affine) {
                        const uint32_t t = tab_u32 + (uint32_t)(a - p.ka_agg) * 128u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 sc = lds128(t + 16u * i), sh = lds128(t + (uint32_t)tab_n * 4u + 16u * i);
                            v[4 * i] = act(v[4 * i], sc.x, sh.x, relu); v[4 * i + 1] = act(v[4 * i + 1], sc.y, sh.y, relu);
                            v[4 * i + 2] = act(v[4 * i + 2], sc.z, sh.z, relu); v[4 * i + 3] = act(v[4 * i + 3], sc.w, sh.w, relu);
                        }
                    } else if (relu) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                }
                float lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split_tf32(v[i], v[i], lo[i]);
                mbar_wait(&at_empty[ts], (ut & 1) ^ 1);               // the MMAs that read this TMEM stage are done
                tc_fence_after_sync();
                tmem_st16(t_lane + ts * 64u, v);
                tmem_st16(t_lane + ts * 64u + 32u, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&at_full[ts]);
                if (++rs == (uint32_t)raw_stages) { rs = 0; ++ur; }
                if (++ts == D3_TSTAGES) { ts = 0; ++ut; }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (p.stats != nullptr) {
        double* my = p.stats + (size_t)blockIdx.x * 2 * p.stats_ld;
        for (int c = tid; c < p.f_out; c += D2_THREADS) {
            my[c] = red_st[c];
            my[p.stats_ld + c] = red_st[256 + c];
        }
    }
    if (warp == D2_WARPS) tmem_dealloc(tmem_base, 512);
}

// ---- weight packing: w[n, k] (row stride ld) -> per K-atom swizzled hi / lo images ---------------
// Rows are packed in slices of `slice` rows (= one launch's N): slice s holds [KA][2][np_s][32] with np_s = its padded rows.
constexpr int TC_NSLICE = 256;      // backward: N = columns of [d_agg | d_self] per launch (2 stages of 96 KB)
constexpr int TC_FWD_SLICE = 128;   // forward: output columns per launch (the raw-atom ring needs the rest of the shared memory)
__global__ void pack_b_kernel(const float* __restrict__ w, int n_rows, int ld, int seg_len, int n_segs, int seg_pad,
                              int np, int slice, float* __restrict__ packed) {
    const int ka = n_segs * seg_pad / ATOM_K;
    const long long total = (long long)ka * np * ATOM_K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int k = (int)(i % ATOM_K);
        int n = (int)((i / ATOM_K) % np);
        int a = (int)(i / ((long long)ATOM_K * np));
        int kp = a * ATOM_K + k;            // padded k
        int seg = kp / seg_pad, kk = kp % seg_pad;
        float v = 0.f;
        if (n < n_rows && kk < seg_len) v = w[(size_t)n * ld + seg * seg_len + kk];
        float hi, lo;
        split_tf32(v, hi, lo);
        const int sl = n / slice, nl = n - sl * slice;
        const int nps = np - sl * slice < slice ? np - sl * slice : slice;
        size_t base = (size_t)sl * ka * 2 * slice * ATOM_K + (size_t)a * 2 * nps * ATOM_K;
        uint32_t off = atom_off(nl, k) / 4;
        packed[base + off] = hi;
        packed[base + (size_t)nps * ATOM_K + off] = lo;
    }
}

}  // namespace dgnn

using namespace dgnn;

static inline int ceil32(int x) { return (x + 31) / 32 * 32; }

extern "C" int dgnn_tc_packed_floats(int n_rows, int seg_len, int n_segs) {
    return (n_segs * ceil32(seg_len) / ATOM_K) * 2 * ceil32(n_rows) * ATOM_K;
}

extern "C" int dgnn_tc_slice(int backward) { return backward ? TC_NSLICE : TC_FWD_SLICE; }

extern "C" int dgnn_pack_b_tf32(const float* w, int n_rows, int ld, int seg_len, int n_segs, int slice, float* packed,
                                void* stream) {
    DGNN_REQUIRE(w && packed, "null pointer");
    DGNN_REQUIRE(slice == TC_NSLICE || slice == TC_FWD_SLICE, "slice must be dgnn_tc_slice(0) or dgnn_tc_slice(1)");
    int np = ceil32(n_rows), seg_pad = ceil32(seg_len);
    long long total = (long long)(n_segs * seg_pad) * np;
    int grid = (int)((total + 255) / 256);
    if (grid > 1024) grid = 1024;
    pack_b_kernel<<<grid, 256, 0, as_stream(stream)>>>(w, n_rows, ld, seg_len, n_segs, seg_pad, np, slice, packed);
    return check_launch("dgnn_pack_b_tf32");
}

// ring depth: the gather mode keeps >= 90 KB of the SM's 228 KB as L1 for the neighbour rows; the streaming
// (dense / backward) modes use all the shared memory they can get for a deeper ring
// raw_ring: bytes reserved behind the operand stages (dense forward: RAW_DEPTH raw atoms)
static int tc_stage_config(int np, bool gather, int* stages, size_t* smem, int raw_ring = 0) {
    int stage_bytes = 2 * A_ATOM_BYTES + 2 * np * ATOM_ROW_BYTES;
    int s = ((gather ? 132 : 200) * 1024 - raw_ring) / stage_bytes;
    if (s > MAX_STAGES) s = MAX_STAGES;
    if (s < 2) s = 2;
    if ((size_t)s * stage_bytes + raw_ring + 1024 > 200 * 1024) return 1;
    *stages = s;
    *smem = (size_t)s * stage_bytes + raw_ring + 1024;
    return 0;
}

template <int MODE, int FE>
static int launch_tc(const TcArgs& p, size_t smem, cudaStream_t st, const char* what) {
    if (int rc_ = ensure_dyn_smem((const void*)layer_tc_kernel<MODE, FE>, 210 * 1024, what)) return rc_;
    layer_tc_kernel<MODE, FE><<<sm_count(), TC_THREADS, smem, st>>>(p);
    return check_launch(what);
}

extern "C" int dgnn_tc_grid(void) { return sm_count(); }

// 1 if the tensor-core path supports these widths (any multiple of 4: wide layers run as column slices)
extern "C" int dgnn_tc_supported(int f_in, int f_out, int gather) {
    (void)gather;
    return (f_in % 4 == 0 && f_out % 4 == 0 && f_in > 0 && f_out > 0) ? 1 : 0;
}

// forward launches, one per slice of <= TC_FWD_SLICE output columns
template <int MODE, int FE>
static int fwd_slices(TcArgs p, int f_out, bool gather, int raw_ring, cudaStream_t st, const char* what) {
    const float* bias = p.bias; const float* osc = p.out_scale; const float* osh = p.out_shift;
    float* out = p.out; double* stats = p.stats; const float* bp = p.b_packed;
    p.out_ld = f_out; p.stats_ld = f_out;
    for (int n0 = 0; n0 < f_out; n0 += TC_FWD_SLICE) {
        const int w = f_out - n0 < TC_FWD_SLICE ? f_out - n0 : TC_FWD_SLICE;
        p.f_out = w; p.np = ceil32(w);
        p.bias = bias ? bias + n0 : nullptr;
        p.out_scale = osc ? osc + n0 : nullptr; p.out_shift = osh ? osh + n0 : nullptr;
        p.out = out + n0; p.stats = stats ? stats + n0 : nullptr;
        p.b_packed = bp + (size_t)(n0 / TC_FWD_SLICE) * p.ka * 2 * TC_FWD_SLICE * ATOM_K;
        if (n0 > 0) p.agg_save = nullptr;          // the aggregate does not depend on the slice
        size_t smem;
        if (tc_stage_config(p.np, gather, &p.stages, &smem, raw_ring)) return fail(what, "tile does not fit shared memory");
        if (int rc = launch_tc<MODE, FE>(p, smem, st, what)) return rc;
    }
    return 0;
}

// ---- tensor maps for the raw atoms (driver entry point resolved through the runtime: the library does not link libcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encode_fn() {
    static TmapEncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TmapEncodeFn>(ptr);
    }
    return fn;
}
// fp32 matrix [rows x cols], row stride ld floats; box = one K-atom of a 128-row tile, 128-byte swizzle, zero fill
static int make_tmap_atoms(CUtensorMap* tm, const float* base, int64_t rows, int cols, int ld, const char* what,
                           int box_rows = TC_M) {
    TmapEncodeFn enc = tmap_encode_fn();
    if (enc == nullptr) return fail(what, "cuTensorMapEncodeTiled is not available from this driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0) return fail(what, "matrix not 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)ATOM_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(what, "cuTensorMapEncodeTiled failed");
    return 0;
}

// ring depths of dense2_tc_kernel: as many raw (A) stages as fit next to >= 2 weight-slice stages (and the staging blocks)
static int d2_stage_config(int np, int tab_floats, int static_bytes, int staging, int* a_stages, int* b_stages, size_t* smem) {
    const int budget = 227 * 1024 - static_bytes - 1024 - tab_floats * 4 - staging;
    const int a_bytes = 2 * A_ATOM_BYTES, b_bytes = 2 * np * ATOM_ROW_BYTES;
    int a = D2_MAX_A, b = 0;
    for (; a >= 2; --a) {
        b = (budget - a * a_bytes) / b_bytes;
        if (b >= 2) break;
    }
    if (a < 2) { a = 2; b = (budget - a * a_bytes) / b_bytes; }
    if (b < 1) return 1;
    if (b > D2_MAX_B) b = D2_MAX_B;
    if (b > a) b = a;
    *a_stages = a; *b_stages = b;
    *smem = (size_t)a * a_bytes + (size_t)b * b_bytes + (size_t)tab_floats * 4 + staging + 1024;
    return 0;
}

// tmo0 / tmo1: tensor maps of the outputs (forward: the z slice; backward: d_agg, d_self)
template <int MODE>
static int launch_d2(const TcArgs& p, const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmo0,
                     const CUtensorMap& tmo1, cudaStream_t st, const char* what) {
    const int tab = (MODE == MODE_BWD ? 3 * p.ka : 2 * (p.ka - p.ka_agg)) * ATOM_K;
    int a_st, b_st;
    size_t smem;
    if (d2_stage_config(p.np, tab, MODE == MODE_BWD ? 512 : 4608, D2_NEW * 4096, &a_st, &b_st, &smem))
        return fail(what, "tile does not fit shared memory");
    if (int rc_ = ensure_dyn_smem((const void*)dense2_tc_kernel<MODE>, 227 * 1024 - (MODE == MODE_BWD ? 512 : 4608), what)) return rc_;
    dense2_tc_kernel<MODE><<<sm_count(), D2_THREADS, smem, st>>>(p, tm0, tm1, tmo0, tmo1, a_st, b_st);
    return check_launch(what);
}

// dense forward through dense2_tc_kernel, one launch per slice of <= TC_FWD_SLICE output columns
static int fwd_slices_d2(TcArgs p, int f_out, cudaStream_t st, const char* what) {
    CUtensorMap tm_agg, tm_x;
    if (int rc = make_tmap_atoms(&tm_x, p.x_in, p.n_tgt, p.f_in, p.f_in, what)) return rc;
    if (p.agg_in != nullptr) { if (int rc = make_tmap_atoms(&tm_agg, p.agg_in, p.n_tgt, p.f_in, p.f_in, what)) return rc; }
    else tm_agg = tm_x;
    const float* bias = p.bias; const float* osc = p.out_scale; const float* osh = p.out_shift;
    float* out = p.out; double* stats = p.stats; const float* bp = p.b_packed;
    p.out_ld = f_out; p.stats_ld = f_out;
    for (int n0 = 0; n0 < f_out; n0 += TC_FWD_SLICE) {
        const int w = f_out - n0 < TC_FWD_SLICE ? f_out - n0 : TC_FWD_SLICE;
        p.f_out = w; p.np = ceil32(w);
        p.bias = bias ? bias + n0 : nullptr;
        p.out_scale = osc ? osc + n0 : nullptr; p.out_shift = osh ? osh + n0 : nullptr;
        p.out = out + n0; p.stats = stats ? stats + n0 : nullptr;
        p.b_packed = bp + (size_t)(n0 / TC_FWD_SLICE) * p.ka * 2 * TC_FWD_SLICE * ATOM_K;
        CUtensorMap tm_out;                            // the slice's columns of z: [n_tgt x w], row stride f_out
        if (int rc = make_tmap_atoms(&tm_out, p.out, p.n_tgt, w, f_out, what, 32)) return rc;
#ifdef DGNN_DENSE3
        {   // A operand in tensor memory (dense3_fwd_kernel): np <= 128 always holds for a forward slice
            const int tab = 2 * (p.ka - p.ka_agg) * ATOM_K;
            const int budget = 227 * 1024 - 4608 - 1024 - tab * 4 - D2_NEW * 4096;
            const int b_bytes = 2 * p.np * ATOM_ROW_BYTES;
            static const int b_env = getenv("DGNN_D3_B") ? atoi(getenv("DGNN_D3_B")) : 3;
            int b_st = b_env, raw_st = (budget - b_st * b_bytes) / A_ATOM_BYTES;
            if (raw_st > D3_MAX_RAW) raw_st = D3_MAX_RAW;
            if (raw_st >= 2) {
                const size_t smem = (size_t)raw_st * A_ATOM_BYTES + (size_t)b_st * b_bytes + (size_t)tab * 4 + D2_NEW * 4096 + 1024;
                if (int rc_ = ensure_dyn_smem((const void*)dense3_fwd_kernel, 227 * 1024 - 4608, what)) return rc_;
                dense3_fwd_kernel<<<sm_count(), D2_THREADS, smem, st>>>(p, tm_agg, tm_x, tm_out, raw_st, b_st);
                if (int rc = check_launch(what)) return rc;
                continue;
            }
        }
#endif
        if (int rc = launch_d2<MODE_FWD_DENSE>(p, tm_agg, tm_x, tm_out, tm_out, st, what)) return rc;
    }
    return 0;
}

extern "C" int dgnn_layer_fwd_tc(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                                 const int32_t* nbr, const float* ea, int fe, const float* w_e, const float* b_e,
                                 const float* b_packed, const float* bias, const float* out_scale,
                                 const float* out_shift, int relu_out, int64_t n_tgt, int f_in, int f_out, float* out,
                                 float* agg_save, double* stats, void* stream) {
    DGNN_REQUIRE(f_in % 4 == 0 && f_out % 4 == 0, "widths must be multiples of 4");
    DGNN_REQUIRE(x_in && b_packed && out, "null pointer");
    if (w_e == nullptr) fe = 0;
    DGNN_REQUIRE(fe % 4 == 0 && fe <= 32, "edge feature width must be a multiple of 4 and <= 32");
    TcArgs p;
    memset(&p, 0, sizeof(p));
    p.x_in = x_in; p.in_scale = in_scale; p.in_shift = in_shift; p.relu_in = relu_in;
    p.nbr = nbr; p.ea = ea; p.w_e = w_e; p.b_e = b_e; p.b_packed = b_packed;
    const int seg = ceil32(f_in) / ATOM_K;
    p.ka_agg = nbr ? seg : 0;
    p.ka = nbr ? 2 * seg : seg;
    p.n_tgt = n_tgt; p.f_in = f_in;
    p.bias = bias; p.out_scale = out_scale; p.out_shift = out_shift; p.relu_out = relu_out;
    p.out = out; p.agg_save = agg_save; p.stats = stats;
    cudaStream_t st = as_stream(stream);
    const char* what = "dgnn_layer_fwd_tc";
#ifndef DGNN_DENSE_OLD
    if (nbr == nullptr) return fwd_slices_d2(p, f_out, st, what);
#endif
    if (nbr == nullptr) return fwd_slices<MODE_FWD_DENSE, 0>(p, f_out, false, RAW_DEPTH * A_ATOM_BYTES, st, what);
    switch (fe) {
        case 0: return fwd_slices<MODE_FWD_GATHER, 0>(p, f_out, true, 0, st, what);
        case 4: return fwd_slices<MODE_FWD_GATHER, 4>(p, f_out, true, 0, st, what);
        case 8: return fwd_slices<MODE_FWD_GATHER, 8>(p, f_out, true, 0, st, what);
        case 12: return fwd_slices<MODE_FWD_GATHER, 12>(p, f_out, true, 0, st, what);
        case 16: return fwd_slices<MODE_FWD_GATHER, 16>(p, f_out, true, 0, st, what);
        case 20: return fwd_slices<MODE_FWD_GATHER, 20>(p, f_out, true, 0, st, what);
        case 24: return fwd_slices<MODE_FWD_GATHER, 24>(p, f_out, true, 0, st, what);
        case 28: return fwd_slices<MODE_FWD_GATHER, 28>(p, f_out, true, 0, st, what);
        case 32: return fwd_slices<MODE_FWD_GATHER, 32>(p, f_out, true, 0, st, what);
    }
    return fail(what, "unsupported edge feature width");
}

// z = [agg | h(x_in)] . W^T with agg read from memory (agg may be NULL: plain dense layer)
extern "C" int dgnn_dense_fwd_tc(const float* agg, const float* x_in, const float* in_scale, const float* in_shift,
                                 int relu_in, const float* b_packed, const float* bias, const float* out_scale,
                                 const float* out_shift, int relu_out, int64_t n_tgt, int f_in, int f_out, float* out,
                                 double* stats, void* stream) {
    DGNN_REQUIRE(f_in % 4 == 0 && f_out % 4 == 0, "widths must be multiples of 4");
    DGNN_REQUIRE(x_in && b_packed && out, "null pointer");
    TcArgs p;
    memset(&p, 0, sizeof(p));
    p.agg_in = agg; p.x_in = x_in; p.in_scale = in_scale; p.in_shift = in_shift; p.relu_in = relu_in;
    p.b_packed = b_packed;
    const int seg = ceil32(f_in) / ATOM_K;
    p.ka_agg = agg ? seg : 0;
    p.ka = agg ? 2 * seg : seg;
    p.n_tgt = n_tgt; p.f_in = f_in;
    p.bias = bias; p.out_scale = out_scale; p.out_shift = out_shift; p.relu_out = relu_out;
    p.out = out; p.stats = stats;
#ifndef DGNN_DENSE_OLD
    return fwd_slices_d2(p, f_out, as_stream(stream), "dgnn_dense_fwd_tc");
#endif
    return fwd_slices<MODE_FWD_DENSE, 0>(p, f_out, false, RAW_DEPTH * A_ATOM_BYTES, as_stream(stream), "dgnn_dense_fwd_tc");
}

extern "C" int dgnn_dense_bwd_tc(const float* dy, const float* z, const float* g, const float* a, const float* b,
                                 const float* mean, const float* rstd, const float* b_packed, const int32_t* nbr,
                                 int64_t n_tgt, int f_in, int f_out, float* d_agg, float* d_self, double* db_partials,
                                 void* stream) {
    DGNN_REQUIRE(f_in % 4 == 0 && f_out % 4 == 0, "widths must be multiples of 4");
    const int n_real = nbr ? 2 * f_in : f_in;
    DGNN_REQUIRE(f_out <= 256 || db_partials == nullptr, "f_out too wide for the shared column sums (take db from the dW kernel)");
    DGNN_REQUIRE(dy && b_packed && d_self && (!nbr || d_agg), "null pointer");
    TcArgs p;
    memset(&p, 0, sizeof(p));
    p.dy = dy; p.z = z; p.ng = g; p.na = a; p.nb = b; p.nmean = mean; p.nrstd = rstd;
    p.nbr = nbr;
    p.ka = ceil32(f_out) / ATOM_K; p.ka_agg = 0;
    p.n_tgt = n_tgt; p.f_in = f_in; p.f_out = f_out;
    p.d_agg = d_agg; p.d_self = d_self;
    p.n_total = n_real;
    cudaStream_t st = as_stream(stream);
#ifndef DGNN_DENSE_OLD
    // dense2_tc_kernel: the column sums of dz (db) come from the dW kernel, and the 32-column chunks of [d_agg | d_self] must
    // not straddle the two matrices (tensor-store epilogue); everything else stays on layer_tc_kernel
    const bool use_d2 = db_partials == nullptr && (f_in % 32 == 0 || nbr == nullptr);
#else
    const bool use_d2 = false;
#endif
    CUtensorMap tm_dy, tm_z, tm_dagg, tm_dself;
    if (use_d2) {
        if (int rc = make_tmap_atoms(&tm_dy, dy, n_tgt, f_out, f_out, "dgnn_dense_bwd_tc")) return rc;
        if (g != nullptr) { if (int rc = make_tmap_atoms(&tm_z, z, n_tgt, f_out, f_out, "dgnn_dense_bwd_tc")) return rc; }
        else tm_z = tm_dy;
        if (int rc = make_tmap_atoms(&tm_dself, d_self, n_tgt, f_in, f_in, "dgnn_dense_bwd_tc", 32)) return rc;
        if (nbr != nullptr) { if (int rc = make_tmap_atoms(&tm_dagg, d_agg, n_tgt, f_in, f_in, "dgnn_dense_bwd_tc", 32)) return rc; }
        else tm_dagg = tm_dself;
    }
    for (int n0 = 0; n0 < n_real; n0 += TC_NSLICE) {
        const int w = n_real - n0 < TC_NSLICE ? n_real - n0 : TC_NSLICE;
        p.np = ceil32(w); p.n_off = n0;
        p.b_packed = b_packed + (size_t)(n0 / TC_NSLICE) * p.ka * 2 * TC_NSLICE * ATOM_K;
        p.db_partials = n0 == 0 ? db_partials : nullptr;
        if (use_d2) {
            if (int rc = launch_d2<MODE_BWD>(p, tm_dy, tm_z, tm_dagg, tm_dself, st, "dgnn_dense_bwd_tc")) return rc;
            continue;
        }
        size_t smem;
        DGNN_REQUIRE(tc_stage_config(p.np, false, &p.stages, &smem) == 0, "tile does not fit shared memory");
        if (int rc = launch_tc<MODE_BWD, 0>(p, smem, st, "dgnn_dense_bwd_tc")) return rc;
    }
    return 0;
}
