// Edge-filtered neighbour aggregation with the edge filter on tensor cores.
//
//   agg[t] = (1/max(cnt,1)) * sum_{k: nbr[t,k]>=0}  h(nbr[t,k]) (*) phi_k,   phi_k = W_e . ea[t,k] + b_e
//
// (learning/surfaceNetStaticEdgeFilters.py:75-96: lin_e, x_j * edge_attr, scatter-mean.)
// Evaluating phi with FMAs costs 80 FMA per feature and cell and made the gather issue-bound
// (profiles/r01_*); here phi_k for a tile of 128 cells is ONE small tcgen05 product
//     PHI_k[128 cells x F] = EA_k[128 x 32] . WE[F x 32]^T        (K = 20 features + bias column, 3xTF32)
// that lands in TMEM.  Each thread then owns one cell row and a strip of F/4 features: it reads its
// strip of PHI_k from TMEM (tcgen05.ld), loads the same strip of the neighbour's row (128-bit loads,
// norm-affine + ReLU of the producer layer applied on load), and accumulates h * phi in registers.
// ~70 warp instructions per cell instead of ~750.
//
// The same kernel computes the backward gather  dh[s] = d_self[s] + sum_k phi(ea_own[s,k]) (*) d_agg[onbr[s,k]]
// (mode 1: no mean, no activation on load, adds `addend`, applies the ReLU mask of the producer layer and
// accumulates the (S1, S2) sums of its normalisation).
//
// Warp roles (one persistent CTA per SM): 16 compute warps (EA staging + strip accumulation),
// 1 MMA warp.  Pipelines: EA ring (smem, 2 stages) and PHI ring (TMEM, 512 / F buffers).
#include "umma.cuh"
#include "common.cuh"

namespace dgnn {

using namespace umma;

constexpr int G_NCW = 16;
constexpr int G_THREADS = (G_NCW + 1) * 32;
constexpr int G_M = 128;
constexpr int G_EA_STAGES = 2;
constexpr int G_ATOM = G_M * 128;  // 16 KB

struct GatherTcArgs {
    const float* x;        // rows to gather: h source (fwd) or d_agg (bwd)
    const float* scale;    // affine on load (fwd), may be NULL
    const float* shift;
    int relu;              // relu on load (fwd)
    const int32_t* nbr;    // [n_rows,4]
    const float* ea;       // [n_rows,4,fe]
    const float* w_e;      // [f, fe]
    const float* b_e;      // [f]
    int fe;
    int64_t n_rows;
    int f;                 // feature width (multiple of 4, <= 128)
    int fp;                // f padded to 32
    int mode;              // 0: forward agg (mean);  1: backward dh
    float* out;            // fwd: agg [n_rows,f];  bwd: dy_prev [n_rows,f] (may be NULL)
    // backward extras
    const float* addend;   // d_self [n_add_rows, f] (rows >= n_add_rows add nothing), may be NULL
    int64_t n_add_rows;
    const float* z_prev;   // pre-norm activations of the producer layer (mask + xhat), may be NULL
    const float* p_scale;  // producer norm affine (y = z*scale + shift), may be NULL
    const float* p_shift;
    const float* p_mean;
    const float* p_rstd;
    int p_relu;
    double* s_partials;    // [grid, 2*f] (S1, S2), may be NULL
    float* dwe_partials;   // [grid, f, 32]: dW_e (cols 0..fe-1) and db_e (col fe), may be NULL
};

constexpr int G_P_BYTES = G_ATOM + 2 * 32 * 128;   // per row-quarter: P_hi [128 x 32 cells] + EA^T hi / lo [32 x 32 cells]

// butterfly transpose-reduce over the warp's 32 rows of CPT columns held in v[0..CPT):
// returns in lane l (l < CPT) ... implemented for CPT = 8, 16, 32 by padding to 32 lanes
template <int CPT>
__device__ __forceinline__ float warp_colsum(float (&v)[CPT], int lane) {
    // reduce rows pairwise until each of the CPT columns has 32/CPT partial copies, then finish by xor-shuffles
#pragma unroll
    for (int off = 16, n = CPT / 2; n >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            float mine = up ? v[i + n] : v[i];
            float theirs = up ? v[i] : v[i + n];
            v[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, off);
        }
    }
    // now v[0] holds, for column c(lane), the sum over a subset of rows; remaining lane bits below are still rows
    float s = v[0];
#pragma unroll
    for (int off = 16 / CPT; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;  // column index = lane / (32 / CPT)   (every 32/CPT consecutive lanes hold the same column)
}

template <int CPT, int MODE>  // CPT: columns (features) per thread = fp / 4;  MODE 0 forward, 1 backward
__global__ void __launch_bounds__(G_THREADS, 1) gather_tc_kernel(const GatherTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ea_full[G_EA_STAGES], ea_empty[G_EA_STAGES];
    __shared__ uint64_t phi_full[8], phi_free[8];
    __shared__ uint64_t p_full[4], p_empty[4], dwe_done;
    __shared__ uint32_t tmem_slot;
    __shared__ double red_s[2 * 128];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int FP = CPT * 4;
    const bool with_dwe = MODE == 1 && p.dwe_partials != nullptr;
    const int phi_cols = with_dwe ? 480 : 512;           // the last 32 TMEM columns hold the dW_e accumulator
    const int n_phi = phi_cols / FP > 8 ? 8 : phi_cols / FP;   // PHI buffers in TMEM
    uint8_t* we_hi = smem;                               // [FP rows x 128 B]
    uint8_t* we_lo = we_hi + FP * 128;
    uint8_t* ea_base = we_lo + FP * 128;                 // stages of (hi 16 KB | lo 16 KB)
    uint8_t* p_base = ea_base + G_EA_STAGES * 2 * G_ATOM;  // 4 x (P_hi | EA^T hi | EA^T lo), backward only

    if (tid == 0) {
        for (int s = 0; s < G_EA_STAGES; ++s) { mbar_init(&ea_full[s], G_NCW); mbar_init(&ea_empty[s], 1); }
        for (int b = 0; b < 8; ++b) { mbar_init(&phi_full[b], 1); mbar_init(&phi_free[b], G_NCW); }
        for (int b = 0; b < 4; ++b) { mbar_init(&p_full[b], 4); mbar_init(&p_empty[b], 1); }
        mbar_init(&dwe_done, 1);
        fence_barrier_init();
    }
    for (int c = tid; c < 256; c += G_THREADS) red_s[c] = 0.0;
    // zero the EA stages (K padding columns stay zero) and build the WE operand:
    // WE[n][e] = w_e[n][e] (e < fe), WE[n][fe] = b_e[n], rest 0
    for (int i = tid; i < (2 * FP * 128 + G_EA_STAGES * 2 * G_ATOM + (with_dwe ? 4 * G_P_BYTES : 0)) / 16; i += G_THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int i = tid; i < p.f * (p.fe + 1); i += G_THREADS) {
        const int n = i / (p.fe + 1), e = i % (p.fe + 1);
        float v = e < p.fe ? __ldg(p.w_e + (size_t)n * p.fe + e) : __ldg(p.b_e + n);
        float hi, lo;
        split_tf32(v, hi, lo);
        const uint32_t off = atom_off(n, e);
        *reinterpret_cast<float*>(we_hi + off) = hi;
        *reinterpret_cast<float*>(we_lo + off) = lo;
    }
    // the bias column of every EA stage: EA[r][fe] = 1
    for (int i = tid; i < G_EA_STAGES * G_M; i += G_THREADS) {
        const int s = i / G_M, r = i % G_M;
        *reinterpret_cast<float*>(ea_base + (size_t)s * 2 * G_ATOM + atom_off(r, p.fe)) = 1.0f;
    }
    if (with_dwe) {
        // row fe of every EA^T operand is all ones: column fe of dW_e accumulates db_e = sum dphi
        for (int i = tid; i < 4 * 32; i += G_THREADS) {
            const int qq = i >> 5, cell = i & 31;
            *reinterpret_cast<float*>(p_base + (size_t)qq * G_P_BYTES + G_ATOM + atom_off(p.fe, cell)) = 1.0f;
        }
    }
    fence_proxy_async_smem();
    if (warp == G_NCW) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_tiles = (p.n_rows + G_M - 1) / G_M;

    if (warp == G_NCW) {
        // ---------------------------------------------------------------- MMA warp
        const uint32_t idesc = make_idesc_tf32(G_M, FP);
        const uint32_t idesc_p = make_idesc_tf32(G_M, 32);
        const uint32_t wh = smem_u32(we_hi), wl = smem_u32(we_lo);
        const uint32_t d_dwe = tmem_base + 480u;
        uint32_t it = 0;
        // dW_e += P^T-stage . EA^T-stage for the slot `slot` (P filled by the 4 warps of every row quarter)
        auto service_p = [&](uint32_t slot) {
#pragma unroll 1
            for (int qq = 0; qq < 4; ++qq) {
                mbar_wait(&p_full[qq], slot & 1);
                tc_fence_after_sync();
                const uint32_t ph = smem_u32(p_base + (size_t)qq * G_P_BYTES), eh = ph + G_ATOM, el = eh + 32 * 128;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t ko = kk * 32;
                    mma_tf32(d_dwe, make_desc(ph + ko), make_desc(eh + ko), idesc_p, (slot > 0 || qq > 0 || kk > 0) ? 1u : 0u);
                    mma_tf32(d_dwe, make_desc(ph + ko), make_desc(el + ko), idesc_p, 1u);
                }
                mma_commit(&p_empty[qq]);
            }
        };
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int k = 0; k < 4; ++k, ++it) {
                if (lane == 0) {
                    const uint32_t s = it % G_EA_STAGES, su = it / G_EA_STAGES;
                    const uint32_t b = it % (uint32_t)n_phi, bu = it / (uint32_t)n_phi;
                    mbar_wait(&ea_full[s], su & 1);
                    if (bu > 0) mbar_wait(&phi_free[b], (bu - 1) & 1);
                    tc_fence_after_sync();
                    const uint32_t ah = smem_u32(ea_base + (size_t)s * 2 * G_ATOM), al = ah + G_ATOM;
                    const uint32_t d = tmem_base + b * (uint32_t)FP;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t ko = kk * 32;
                        mma_tf32(d, make_desc(ah + ko), make_desc(wh + ko), idesc, kk > 0 ? 1u : 0u);
                        mma_tf32(d, make_desc(al + ko), make_desc(wh + ko), idesc, 1u);
                        mma_tf32(d, make_desc(ah + ko), make_desc(wl + ko), idesc, 1u);
                    }
                    mma_commit(&ea_empty[s]);
                    mma_commit(&phi_full[b]);
                    if (with_dwe && it > 0) service_p(it - 1);
                }
                __syncwarp();
            }
        }
        if (lane == 0 && with_dwe && it > 0) {
            service_p(it - 1);
            mma_commit(&dwe_done);
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- compute warps
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        const int c0 = grp * CPT;                      // first feature of this thread's strip
        const bool relu = p.relu != 0;
        const int fe4 = p.fe >> 2;
        double s1d = 0.0, s2d = 0.0;   // running (S1, S2) of column c0 + lane / (32 / CPT)
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t tile0 = tile * G_M;
            const int64_t t = tile0 + row;
            const bool tv = t < p.n_rows;
            int4 nb4 = make_int4(-1, -1, -1, -1);
            if (tv) nb4 = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
            const int nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
            float acc[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i) acc[i] = 0.f;
            for (int k = 0; k <= 4; ++k) {
                if (k < 4) {
                    // ---- stage EA_k: rows of the tile, fe floats each, split into hi / lo
                    const uint32_t itk = it + k;
                    const uint32_t s = itk % G_EA_STAGES, su = itk / G_EA_STAGES;
                    uint8_t* e_hi = ea_base + (size_t)s * 2 * G_ATOM;
                    uint8_t* e_lo = e_hi + G_ATOM;
                    mbar_wait(&ea_empty[s], (su & 1) ^ 1);
                    for (int idx = tid; idx < G_M * fe4; idx += G_NCW * 32) {
                        const int r = idx / fe4, c = idx % fe4;
                        const int64_t tr = tile0 + r;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (tr < p.n_rows) v = ldg4(p.ea + ((size_t)tr * 4 + k) * p.fe + c * 4);
                        float4 h, l;
                        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y);
                        split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                        const uint32_t off = atom_off(r, c * 4);
                        *reinterpret_cast<float4*>(e_hi + off) = h;
                        *reinterpret_cast<float4*>(e_lo + off) = l;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ea_full[s]);
                }
                if (k > 0) {
                    // ---- consume PHI_{k-1}
                    const int kc = k - 1;
                    const uint32_t itc = it + kc;
                    const uint32_t b = itc % (uint32_t)n_phi, bu = itc / (uint32_t)n_phi;
                    mbar_wait(&phi_full[b], bu & 1);
                    tc_fence_after_sync();
                    const int s_row = nbv[kc];
                    if (MODE == 1 && with_dwe) {
                        // the P / EA^T stage of this row quarter must have been consumed (slot itc - 1)
                        mbar_wait(&p_empty[q], (itc & 1) ^ 1);
                        // EA^T rows 8*grp .. 8*grp+7 for this thread's cell (column = lane)
                        uint8_t* eh = p_base + (size_t)q * G_P_BYTES + G_ATOM;
                        uint8_t* el = eh + 32 * 128;
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const int e0 = grp * 8 + h2 * 4;
                            if (e0 >= p.fe) continue;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (tv && s_row >= 0) v = ldg4(p.ea + ((size_t)t * 4 + kc) * p.fe + e0);
                            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float hi, lo;
                                split_tf32(vv[i], hi, lo);
                                const uint32_t off = atom_off(e0 + i, lane);
                                *reinterpret_cast<float*>(eh + off) = hi;
                                *reinterpret_cast<float*>(el + off) = lo;
                            }
                        }
                    }
                    const uint32_t taddr = tmem_base + b * (uint32_t)FP + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
#pragma unroll
                    for (int j = 0; j < CPT; j += 8) {
                        uint32_t ph[8];
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=r"(ph[0]), "=r"(ph[1]), "=r"(ph[2]), "=r"(ph[3]), "=r"(ph[4]), "=r"(ph[5]),
                                       "=r"(ph[6]), "=r"(ph[7])
                                     : "r"(taddr + (uint32_t)j));
                        float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
                        const int f0 = c0 + j;
                        if (s_row >= 0 && f0 < p.f) {
                            xa = ldg4(p.x + (size_t)s_row * p.f + f0);
                            xb = ldg4(p.x + (size_t)s_row * p.f + f0 + 4);
                            if (p.scale != nullptr) {
                                float4 sa = ldg4(p.scale + f0), sb = ldg4(p.scale + f0 + 4);
                                float4 ha = ldg4(p.shift + f0), hb = ldg4(p.shift + f0 + 4);
                                xa.x = act(xa.x, sa.x, ha.x, relu); xa.y = act(xa.y, sa.y, ha.y, relu);
                                xa.z = act(xa.z, sa.z, ha.z, relu); xa.w = act(xa.w, sa.w, ha.w, relu);
                                xb.x = act(xb.x, sb.x, hb.x, relu); xb.y = act(xb.y, sb.y, hb.y, relu);
                                xb.z = act(xb.z, sb.z, hb.z, relu); xb.w = act(xb.w, sb.w, hb.w, relu);
                            } else if (relu) {
                                xa.x = fmaxf(xa.x, 0.f); xa.y = fmaxf(xa.y, 0.f); xa.z = fmaxf(xa.z, 0.f); xa.w = fmaxf(xa.w, 0.f);
                                xb.x = fmaxf(xb.x, 0.f); xb.y = fmaxf(xb.y, 0.f); xb.z = fmaxf(xb.z, 0.f); xb.w = fmaxf(xb.w, 0.f);
                            }
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (MODE == 1 && with_dwe) {
                            // dphi = h(s) * d_agg[t_k]  ->  P^T (rows = features, columns = this quarter's 32 cells), TF32
                            float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
                            if (tv && f0 < p.f) {
                                za = ldg4(p.z_prev + (size_t)t * p.f + f0);
                                zb = ldg4(p.z_prev + (size_t)t * p.f + f0 + 4);
                                if (p.p_scale != nullptr) {
                                    float4 sa = ldg4(p.p_scale + f0), sb = ldg4(p.p_scale + f0 + 4);
                                    float4 ha = ldg4(p.p_shift + f0), hb = ldg4(p.p_shift + f0 + 4);
                                    za.x = fmaf(za.x, sa.x, ha.x); za.y = fmaf(za.y, sa.y, ha.y);
                                    za.z = fmaf(za.z, sa.z, ha.z); za.w = fmaf(za.w, sa.w, ha.w);
                                    zb.x = fmaf(zb.x, sb.x, hb.x); zb.y = fmaf(zb.y, sb.y, hb.y);
                                    zb.z = fmaf(zb.z, sb.z, hb.z); zb.w = fmaf(zb.w, sb.w, hb.w);
                                }
                                if (p.p_relu) {
                                    za.x = fmaxf(za.x, 0.f); za.y = fmaxf(za.y, 0.f); za.z = fmaxf(za.z, 0.f); za.w = fmaxf(za.w, 0.f);
                                    zb.x = fmaxf(zb.x, 0.f); zb.y = fmaxf(zb.y, 0.f); zb.z = fmaxf(zb.z, 0.f); zb.w = fmaxf(zb.w, 0.f);
                                }
                            }
                            const float dp[8] = {za.x * xa.x, za.y * xa.y, za.z * xa.z, za.w * xa.w,
                                                 zb.x * xb.x, zb.y * xb.y, zb.z * xb.z, zb.w * xb.w};
                            uint8_t* pq = p_base + (size_t)q * G_P_BYTES;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                // row = c0 + j + i (c0 + j is a multiple of 8, so row & 7 == i), column = lane
                                const uint32_t off = (uint32_t)(c0 + j + i) * 128u + ((((uint32_t)lane >> 2) ^ (uint32_t)i) << 4) +
                                                     (((uint32_t)lane & 3u) << 2);
                                *reinterpret_cast<float*>(pq + off) = tf32_rna(dp[i]);
                            }
                        }
                        acc[j + 0] = fmaf(xa.x, __uint_as_float(ph[0]), acc[j + 0]);
                        acc[j + 1] = fmaf(xa.y, __uint_as_float(ph[1]), acc[j + 1]);
                        acc[j + 2] = fmaf(xa.z, __uint_as_float(ph[2]), acc[j + 2]);
                        acc[j + 3] = fmaf(xa.w, __uint_as_float(ph[3]), acc[j + 3]);
                        acc[j + 4] = fmaf(xb.x, __uint_as_float(ph[4]), acc[j + 4]);
                        acc[j + 5] = fmaf(xb.y, __uint_as_float(ph[5]), acc[j + 5]);
                        acc[j + 6] = fmaf(xb.z, __uint_as_float(ph[6]), acc[j + 6]);
                        acc[j + 7] = fmaf(xb.w, __uint_as_float(ph[7]), acc[j + 7]);
                    }
                    tc_fence_before_sync();
                    if (MODE == 1 && with_dwe) fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&phi_free[b]);
                        if (MODE == 1 && with_dwe) mbar_arrive(&p_full[q]);
                    }
                }
            }
            it += 4;
            // ---- finish the row strip
            if (MODE == 0) {
                const int cnt = (nbv[0] >= 0) + (nbv[1] >= 0) + (nbv[2] >= 0) + (nbv[3] >= 0);
                const float d = (float)(cnt > 0 ? cnt : 1);
                if (tv) {
#pragma unroll
                    for (int j = 0; j < CPT; j += 4) {
                        if (c0 + j >= p.f) continue;
                        *reinterpret_cast<float4*>(p.out + (size_t)t * p.f + c0 + j) =
                            make_float4(acc[j] / d, acc[j + 1] / d, acc[j + 2] / d, acc[j + 3] / d);
                    }
                }
            } else {
                float dx[CPT];                       // dy * xhat (acc becomes dy)
#pragma unroll
                for (int j = 0; j < CPT; j += 4) {
                    const int f0 = c0 + j;
                    float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                    float4 ox = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!tv || f0 >= p.f) {
                        o = make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        if (p.addend != nullptr && t < p.n_add_rows) {
                            float4 a = ldg4(p.addend + (size_t)t * p.f + f0);
                            o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                        }
                        if (p.z_prev != nullptr) {
                            float4 zv = ldg4(p.z_prev + (size_t)t * p.f + f0);
                            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                            float4 mu = sh, rs = sc;
                            if (p.p_scale != nullptr) { sc = ldg4(p.p_scale + f0); sh = ldg4(p.p_shift + f0); }
                            if (p.p_mean != nullptr) { mu = ldg4(p.p_mean + f0); rs = ldg4(p.p_rstd + f0); }
                            if (p.p_relu) {
                                if (!(fmaf(zv.x, sc.x, sh.x) > 0.f)) o.x = 0.f;
                                if (!(fmaf(zv.y, sc.y, sh.y) > 0.f)) o.y = 0.f;
                                if (!(fmaf(zv.z, sc.z, sh.z) > 0.f)) o.z = 0.f;
                                if (!(fmaf(zv.w, sc.w, sh.w) > 0.f)) o.w = 0.f;
                            }
                            ox = make_float4(o.x * ((zv.x - mu.x) * rs.x), o.y * ((zv.y - mu.y) * rs.y),
                                             o.z * ((zv.z - mu.z) * rs.z), o.w * ((zv.w - mu.w) * rs.w));
                        }
                        if (p.out != nullptr) *reinterpret_cast<float4*>(p.out + (size_t)t * p.f + f0) = o;
                    }
                    acc[j] = o.x; acc[j + 1] = o.y; acc[j + 2] = o.z; acc[j + 3] = o.w;
                    dx[j] = ox.x; dx[j + 1] = ox.y; dx[j + 2] = ox.z; dx[j + 3] = ox.w;
                }
                if (p.s_partials != nullptr) {
                    s1d += (double)warp_colsum<CPT>(acc, lane);
                    s2d += (double)warp_colsum<CPT>(dx, lane);
                }
            }
        }
        if (MODE == 1 && with_dwe && grp == 0) {
            float* outp = p.dwe_partials + (size_t)blockIdx.x * p.f * 32;
            float v[32];
            if (it > 0) {
                mbar_wait(&dwe_done, 0);
                tc_fence_after_sync();
                tmem_ld32(tmem_base + 480u + ((uint32_t)(q * 32) << 16), v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (row < p.f) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(outp + (size_t)row * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
        if (MODE == 1 && p.s_partials != nullptr) {
            const int col = c0 + lane / (32 / CPT);
            if ((lane % (32 / CPT)) == 0 && col < p.f) {
                atomicAdd(&red_s[col], s1d);
                atomicAdd(&red_s[128 + col], s2d);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (MODE == 1 && p.s_partials != nullptr) {
        double* my = p.s_partials + (size_t)blockIdx.x * 2 * p.f;
        for (int c = tid; c < p.f; c += G_THREADS) {
            my[c] = red_s[c];
            my[p.f + c] = red_s[128 + c];
        }
    }
    if (warp == G_NCW) tmem_dealloc(tmem_base, 512);
}

}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_gather_tc_supported(int f, int fe) {
    return (f % 4 == 0 && f >= 4 && f <= 128 && fe % 4 == 0 && fe >= 4 && fe <= 28) ? 1 : 0;
}

static int launch_gather_tc(const GatherTcArgs& p, cudaStream_t st, const char* what) {
    const int cpt = p.fp / 4;
    size_t smem = (size_t)2 * p.fp * 128 + (size_t)G_EA_STAGES * 2 * G_ATOM + 1024 +
                  ((p.mode == 1 && p.dwe_partials) ? 4 * G_P_BYTES : 0);
#define LAUNCH_G(CPT, MODE)                                                                                                \
    do {                                                                                                             \
        static bool configured = false;                                                                              \
        if (!configured) {                                                                                           \
            cudaError_t e = cudaFuncSetAttribute(gather_tc_kernel<CPT, MODE>,                                        \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);           \
            if (e != cudaSuccess) return fail(what, cudaGetErrorString(e));                                          \
            configured = true;                                                                                       \
        }                                                                                                            \
        gather_tc_kernel<CPT, MODE><<<sm_count(), G_THREADS, smem, st>>>(p);                                         \
    } while (0)
    if (p.mode == 0) {
        switch (cpt) {
            case 8: LAUNCH_G(8, 0); break;
            case 16: LAUNCH_G(16, 0); break;
            case 32: LAUNCH_G(32, 0); break;
            default: return fail(what, "unsupported feature width");
        }
    } else {
        switch (cpt) {
            case 8: LAUNCH_G(8, 1); break;
            case 16: LAUNCH_G(16, 1); break;
            case 32: LAUNCH_G(32, 1); break;
            default: return fail(what, "unsupported feature width");
        }
    }
#undef LAUNCH_G
    return check_launch(what);
}

extern "C" int dgnn_gather_tc_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                                  const int32_t* nbr, const float* ea, int fe, const float* w_e, const float* b_e,
                                  int64_t n_tgt, int f_in, float* agg, void* stream) {
    DGNN_REQUIRE(dgnn_gather_tc_supported(f_in, fe), "widths not supported by the tensor-core gather");
    DGNN_REQUIRE(x_in && nbr && ea && w_e && b_e && agg, "null pointer");
    GatherTcArgs p;
    memset(&p, 0, sizeof(p));
    p.x = x_in; p.scale = in_scale; p.shift = in_shift; p.relu = relu_in;
    p.nbr = nbr; p.ea = ea; p.w_e = w_e; p.b_e = b_e; p.fe = fe;
    p.n_rows = n_tgt; p.f = f_in; p.fp = f_in <= 32 ? 32 : (f_in <= 64 ? 64 : 128); p.mode = 0; p.out = agg;
    return launch_gather_tc(p, as_stream(stream), "dgnn_gather_tc_fwd");
}

extern "C" int dgnn_gather_tc_bwd(const float* d_agg, const float* d_self, const int32_t* onbr, const float* ea_own,
                                  int fe, const float* w_e, const float* b_e, const float* z_prev,
                                  const float* p_scale, const float* p_shift, const float* p_mean,
                                  const float* p_rstd, int p_relu, int64_t n_src, int64_t n_tgt, int f_in,
                                  float* dy_prev, double* s_partials, float* dwe_partials, void* stream) {
    DGNN_REQUIRE(dgnn_gather_tc_supported(f_in, fe), "widths not supported by the tensor-core gather");
    DGNN_REQUIRE(d_agg && onbr && ea_own && w_e && b_e, "null pointer");
    GatherTcArgs p;
    memset(&p, 0, sizeof(p));
    p.x = d_agg; p.nbr = onbr; p.ea = ea_own; p.w_e = w_e; p.b_e = b_e; p.fe = fe;
    p.n_rows = n_src; p.f = f_in; p.fp = f_in <= 32 ? 32 : (f_in <= 64 ? 64 : 128); p.mode = 1; p.out = dy_prev;
    p.addend = d_self; p.n_add_rows = n_tgt;
    p.z_prev = z_prev; p.p_scale = p_scale; p.p_shift = p_shift; p.p_mean = p_mean; p.p_rstd = p_rstd;
    p.p_relu = p_relu; p.s_partials = s_partials; p.dwe_partials = dwe_partials;
    DGNN_REQUIRE(dwe_partials == nullptr || z_prev != nullptr, "dW_e needs the layer input (z_prev)");
    return launch_gather_tc(p, as_stream(stream), "dgnn_gather_tc_bwd");
}
