// Edge-filtered neighbour aggregation with the edge filter on tensor cores.
//
//   MODE 0 (forward)   agg[t] = (1/max(cnt,1)) * sum_{k: nbr[t,k]>=0} h(nbr[t,k]) (*) phi_k,
//                      phi_k = W_e . ea[t,k] + b_e
//                      (learning/surfaceNetStaticEdgeFilters.py:75-96: lin_e, x_j * edge_attr, scatter-mean)
//   MODE 1 (backward)  dh[s] = d_self[s] + sum_k phi(ea_own[s,k]) (*) d_agg[onbr[s,k]], then the ReLU mask of
//                      the producer layer -> dy_prev, and the (S1, S2) sums of its normalisation
//   dW_e kernel        dW_e[f,e] = sum_{s,k} h(s)[f] d_agg[onbr[s,k]][f] ea_own[s,k][e],  db_e[f] = sum dphi
//
// Evaluating phi with FMAs costs 80 FMA per feature and cell and made the gather issue-bound
// (profiles/r01_*).  Here PHI_k for a tile of 128 cells is one small tcgen05 product
//     PHI_k[128 cells x F] = EA_k[128 x 32] . WE[F x 32]^T      (K = fe features + a bias column, 3xTF32)
// that lands in TMEM (4 slots x F <= 512 columns).  Each thread then owns one cell row and a strip of
// F/4 features; per 8 features it issues the 8 neighbour-row loads of all four slots (128-bit, the norm
// affine + ReLU of the producer layer applied on load), reads the matching PHI strips from TMEM
// (tcgen05.ld) and accumulates h * phi in registers.
// dW_e is the product P^T . EA with P = dphi [edges x F]: both operands need the edge index contiguous,
// so each thread scatters its strip of P (rounded to TF32) and of EA (hi / lo) transposed into K-major
// shared-memory operands; the [F x 32] accumulator lives in TMEM for the whole kernel.
//
// Warp roles (one persistent CTA per SM): 16 compute warps, 1 MMA warp; no CTA-wide barrier in the loop.
#include "umma.cuh"
#include "common.cuh"

namespace dgnn {

using namespace umma;

constexpr int G_NCW = 16;
constexpr int G_THREADS = (G_NCW + 1) * 32;
constexpr int G_M = 128;
constexpr int G_EA_STAGES = 4;                       // one tile (4 slots) of EA operands ahead
constexpr int G_ATOM = G_M * 128;                    // 16 KB
constexpr int G_P_BYTES = G_ATOM + 2 * 32 * 128;     // P_hi [128 x 32 cells] + EA^T hi / lo [32 x 32 cells]

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
    int fp;                // 32, 64 or 128
    float* out;            // fwd: agg [n_rows,f];  bwd: dy_prev [n_rows,f] (may be NULL)
    // backward extras
    const float* addend;   // d_self [n_add_rows, f] (rows >= n_add_rows add nothing), may be NULL
    int64_t n_add_rows;
    const float* z_prev;   // pre-norm activations of the producer layer = this layer's input (mask, xhat, h)
    const float* p_scale;  // producer norm affine (y = z*scale + shift), may be NULL
    const float* p_shift;
    const float* p_mean;
    const float* p_rstd;
    int p_relu;
    double* s_partials;    // [grid, 2*f] (S1, S2), may be NULL
    float* dwe_partials;   // dW_e kernel: [grid, f, 32]: dW_e (cols 0..fe-1) and db_e (col fe)
};

// butterfly transpose-reduce of 8 columns over the warp's 32 rows: every lane l returns the sum of
// column (l >> 2)
__device__ __forceinline__ float warp_colsum8(float (&v)[8], int lane) {
#pragma unroll
    for (int off = 16, n = 4; n >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            float mine = up ? v[i + n] : v[i];
            float theirs = up ? v[i] : v[i + n];
            v[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, off);
        }
    }
    float s = v[0];
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

__device__ __forceinline__ void act8(float4& a, float4& b, const float4& sa, const float4& sb, const float4& ha,
                                     const float4& hb, bool affine, bool relu) {
    if (affine) {
        a.x = fmaf(a.x, sa.x, ha.x); a.y = fmaf(a.y, sa.y, ha.y); a.z = fmaf(a.z, sa.z, ha.z); a.w = fmaf(a.w, sa.w, ha.w);
        b.x = fmaf(b.x, sb.x, hb.x); b.y = fmaf(b.y, sb.y, hb.y); b.z = fmaf(b.z, sb.z, hb.z); b.w = fmaf(b.w, sb.w, hb.w);
    }
    if (relu) {
        a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
        b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------
// MODE 0 / 1: phi on tensor cores, strips accumulated in registers
template <int CPT, int MODE>  // CPT: features per thread = fp / 4
__global__ void __launch_bounds__(G_THREADS, 1) gather_tc_kernel(const GatherTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ea_full[G_EA_STAGES], ea_empty[G_EA_STAGES];
    __shared__ uint64_t phi_full, phi_free;
    __shared__ uint32_t tmem_slot;
    __shared__ double red_s[2 * 128];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int FP = CPT * 4;
    uint8_t* we_hi = smem;                               // [FP rows x 128 B]
    uint8_t* we_lo = we_hi + FP * 128;
    uint8_t* ea_base = we_lo + FP * 128;                 // stages of (hi 16 KB | lo 16 KB)

    if (tid == 0) {
        for (int s = 0; s < G_EA_STAGES; ++s) { mbar_init(&ea_full[s], G_NCW); mbar_init(&ea_empty[s], 1); }
        mbar_init(&phi_full, 1);
        mbar_init(&phi_free, G_NCW);
        fence_barrier_init();
    }
    for (int c = tid; c < 256; c += G_THREADS) red_s[c] = 0.0;
    // zero the operands (K padding stays zero), then WE[n][e] = w_e[n][e] (e < fe), WE[n][fe] = b_e[n]
    for (int i = tid; i < (2 * FP * 128 + G_EA_STAGES * 2 * G_ATOM) / 16; i += G_THREADS)
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
    for (int i = tid; i < G_EA_STAGES * G_M; i += G_THREADS) {   // bias column of every EA stage
        const int s = i / G_M, r = i % G_M;
        *reinterpret_cast<float*>(ea_base + (size_t)s * 2 * G_ATOM + atom_off(r, p.fe)) = 1.0f;
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
        const uint32_t wh = smem_u32(we_hi), wl = smem_u32(we_lo);
        uint32_t it = 0, tile_cnt = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            for (int k = 0; k < 4; ++k, ++it) {
                if (lane == 0) {
                    const uint32_t s = it % G_EA_STAGES, su = it / G_EA_STAGES;
                    mbar_wait(&ea_full[s], su & 1);
                    if (k == 0 && tile_cnt > 0) mbar_wait(&phi_free, (tile_cnt - 1) & 1);
                    tc_fence_after_sync();
                    const uint32_t ah = smem_u32(ea_base + (size_t)s * 2 * G_ATOM), al = ah + G_ATOM;
                    const uint32_t d = tmem_base + (uint32_t)(k * FP);
                    const int ksteps = (p.fe + 1 + 7) >> 3;     // fe features + bias column, 8 per k-step (3 for fe = 20)
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const uint32_t ko = kk * 32;
                        mma_tf32(d, make_desc(ah + ko), make_desc(wh + ko), idesc, kk > 0 ? 1u : 0u);
                        mma_tf32(d, make_desc(al + ko), make_desc(wh + ko), idesc, 1u);
                        mma_tf32(d, make_desc(ah + ko), make_desc(wl + ko), idesc, 1u);
                    }
                    mma_commit(&ea_empty[s]);
                    if (k == 3) mma_commit(&phi_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------- compute warps
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        const int c0 = grp * CPT;                      // first feature of this thread's strip
        const bool relu = (p.relu & 1) != 0;
        const bool x_skip = (p.relu & 2) != 0, t_skip = (p.relu & 4) != 0, s_skip = (p.relu & 8) != 0;  // dev experiments
        const bool affine = p.scale != nullptr;
        const bool al8 = (p.f & 7) == 0;
        const int fe4 = p.fe >> 2;
        double s1d[CPT / 8], s2d[CPT / 8];             // running (S1, S2) of column c0 + 8*jj + (lane >> 2)
#pragma unroll
        for (int i = 0; i < CPT / 8; ++i) s1d[i] = s2d[i] = 0.0;
        uint32_t it = 0, tile_cnt = 0;
        // stage EA_0..3 of a tile (rows of the tile, fe floats each, split hi / lo); the MMA warp follows
        auto stage_ea = [&](int64_t tile_s) {
            const int64_t base = tile_s * G_M;
            // all loads of the four slots first (<= 2 float4 per thread and slot for fe <= 28) ...
            float4 v[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int idx = tid + u * G_NCW * 32;
                    v[k][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < G_M * fe4) {
                        const int r = idx / fe4, c = idx - r * fe4;
                        const int64_t tr = base + r;
                        if (tr < p.n_rows) v[k][u] = ldg4(p.ea + ((size_t)tr * 4 + k) * p.fe + c * 4);
                    }
                }
            // ... then per slot: wait for its stage, split hi / lo, store, arrive (the MMA warp follows)
#pragma unroll
            for (int k = 0; k < 4; ++k, ++it) {
                const uint32_t s = it % G_EA_STAGES, su = it / G_EA_STAGES;
                uint8_t* e_hi = ea_base + (size_t)s * 2 * G_ATOM;
                uint8_t* e_lo = e_hi + G_ATOM;
                mbar_wait(&ea_empty[s], (su & 1) ^ 1);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int idx = tid + u * G_NCW * 32;
                    if (idx < G_M * fe4) {
                        const int r = idx / fe4, c = idx - r * fe4;
                        float4 h, l;
                        split_tf32(v[k][u].x, h.x, l.x); split_tf32(v[k][u].y, h.y, l.y);
                        split_tf32(v[k][u].z, h.z, l.z); split_tf32(v[k][u].w, h.w, l.w);
                        const uint32_t off = atom_off(r, c * 4);
                        *reinterpret_cast<float4*>(e_hi + off) = h;
                        *reinterpret_cast<float4*>(e_lo + off) = l;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ea_full[s]);
            }
        };
        if ((int64_t)blockIdx.x < n_tiles) stage_ea(blockIdx.x);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
            const int64_t tile0 = tile * G_M;
            const int64_t t = tile0 + row;
            const bool tv = t < p.n_rows;
            int4 nb4 = make_int4(-1, -1, -1, -1);
            if (tv) nb4 = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
            const int nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
            // EA of the next tile goes in flight before this tile is consumed (its MMAs wait for phi_free)
            if (tile + gridDim.x < n_tiles) stage_ea(tile + gridDim.x);
            // ---- consume: all four PHI_k of the tile are in TMEM
            mbar_wait(&phi_full, tile_cnt & 1);
            tc_fence_after_sync();
            const int cnt = (nbv[0] >= 0) + (nbv[1] >= 0) + (nbv[2] >= 0) + (nbv[3] >= 0);
            const float dcnt = (float)(cnt > 0 ? cnt : 1);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
#pragma unroll
            for (int j = 0; j < CPT; j += 8) {
                const int f0 = c0 + j;
                const bool fvalid = f0 < p.f, fvalid_b = f0 + 4 < p.f;   // the two 4-feature halves of the strip step
                uint32_t ph[4][8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!t_skip) tmem_ld8(trow + (uint32_t)(k * FP + j), ph[k]);
                    else { for (int i = 0; i < 8; ++i) ph[k][i] = 0x3f800000u; }
                }
                float4 xa[4], xb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    xa[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    xb[k] = xa[k];
                    if (x_skip) {
                        xa[k] = make_float4(1.f, 1.f, 1.f, 1.f); xb[k] = xa[k];
                    } else if (al8) {   // rows are 32-byte aligned: one 256-bit request instead of two 128-bit ones
                        if (nbv[k] >= 0 && fvalid) ldg8(p.x + (size_t)nbv[k] * p.f + f0, xa[k], xb[k]);
                    } else {
                        if (nbv[k] >= 0 && fvalid) xa[k] = ldg4(p.x + (size_t)nbv[k] * p.f + f0);
                        if (nbv[k] >= 0 && fvalid_b) xb[k] = ldg4(p.x + (size_t)nbv[k] * p.f + f0 + 4);
                    }
                }
                float4 sa = make_float4(1.f, 1.f, 1.f, 1.f), sb = sa, ha = make_float4(0.f, 0.f, 0.f, 0.f), hb = ha;
                if (MODE == 0 && affine && fvalid) { sa = ldg4(p.scale + f0); ha = ldg4(p.shift + f0); }
                if (MODE == 0 && affine && fvalid_b) { sb = ldg4(p.scale + f0 + 4); hb = ldg4(p.shift + f0 + 4); }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (nbv[k] < 0) continue;
                    if (MODE == 0) act8(xa[k], xb[k], sa, sb, ha, hb, affine, relu);
                    acc[0] = fmaf(xa[k].x, __uint_as_float(ph[k][0]), acc[0]);
                    acc[1] = fmaf(xa[k].y, __uint_as_float(ph[k][1]), acc[1]);
                    acc[2] = fmaf(xa[k].z, __uint_as_float(ph[k][2]), acc[2]);
                    acc[3] = fmaf(xa[k].w, __uint_as_float(ph[k][3]), acc[3]);
                    acc[4] = fmaf(xb[k].x, __uint_as_float(ph[k][4]), acc[4]);
                    acc[5] = fmaf(xb[k].y, __uint_as_float(ph[k][5]), acc[5]);
                    acc[6] = fmaf(xb[k].z, __uint_as_float(ph[k][6]), acc[6]);
                    acc[7] = fmaf(xb[k].w, __uint_as_float(ph[k][7]), acc[7]);
                }
                if (MODE == 0) {
                    if (tv && fvalid && !s_skip) {
                        const float4 oa = make_float4(acc[0] / dcnt, acc[1] / dcnt, acc[2] / dcnt, acc[3] / dcnt);
                        const float4 ob = make_float4(acc[4] / dcnt, acc[5] / dcnt, acc[6] / dcnt, acc[7] / dcnt);
                        if (al8) {
                            stg8(p.out + (size_t)t * p.f + f0, oa, ob);      // one full 32-byte sector per thread
                        } else {
                            *reinterpret_cast<float4*>(p.out + (size_t)t * p.f + f0) = oa;
                            if (fvalid_b) *reinterpret_cast<float4*>(p.out + (size_t)t * p.f + f0 + 4) = ob;
                        }
                    }
                } else {
                    float dx[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) dx[i] = 0.f;
                    if (!tv || !fvalid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                    } else {
                        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (!fvalid_b) { acc[4] = acc[5] = acc[6] = acc[7] = 0.f; }
                        if (p.addend != nullptr && t < p.n_add_rows) {
                            float4 a = ldg4(p.addend + (size_t)t * p.f + f0);
                            float4 b = fvalid_b ? ldg4(p.addend + (size_t)t * p.f + f0 + 4) : zero4;
                            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
                        }
                        if (p.z_prev != nullptr) {
                            float4 za = ldg4(p.z_prev + (size_t)t * p.f + f0);
                            float4 zb = fvalid_b ? ldg4(p.z_prev + (size_t)t * p.f + f0 + 4) : zero4;
                            const float zv[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
                            float sc[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, sh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            float mu[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, rs[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
                            if (p.p_scale != nullptr) {
                                float4 a = ldg4(p.p_scale + f0), b = fvalid_b ? ldg4(p.p_scale + f0 + 4) : zero4;
                                float4 c = ldg4(p.p_shift + f0), d = fvalid_b ? ldg4(p.p_shift + f0 + 4) : zero4;
                                sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
                                sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
                            }
                            if (p.p_mean != nullptr) {
                                float4 a = ldg4(p.p_mean + f0), b = fvalid_b ? ldg4(p.p_mean + f0 + 4) : zero4;
                                float4 c = ldg4(p.p_rstd + f0), d = fvalid_b ? ldg4(p.p_rstd + f0 + 4) : zero4;
                                mu[0] = a.x; mu[1] = a.y; mu[2] = a.z; mu[3] = a.w; mu[4] = b.x; mu[5] = b.y; mu[6] = b.z; mu[7] = b.w;
                                rs[0] = c.x; rs[1] = c.y; rs[2] = c.z; rs[3] = c.w; rs[4] = d.x; rs[5] = d.y; rs[6] = d.z; rs[7] = d.w;
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (p.p_relu && !(fmaf(zv[i], sc[i], sh[i]) > 0.f)) acc[i] = 0.f;
                                if (i >= 4 && !fvalid_b) acc[i] = 0.f;
                                dx[i] = acc[i] * ((zv[i] - mu[i]) * rs[i]);
                            }
                        }
                        if (p.out != nullptr) {
                            if (al8) {
                                stg8(p.out + (size_t)t * p.f + f0, make_float4(acc[0], acc[1], acc[2], acc[3]),
                                     make_float4(acc[4], acc[5], acc[6], acc[7]));
                            } else {
                                *reinterpret_cast<float4*>(p.out + (size_t)t * p.f + f0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                                if (fvalid_b)
                                    *reinterpret_cast<float4*>(p.out + (size_t)t * p.f + f0 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                            }
                        }
                    }
                    if (p.s_partials != nullptr) {
                        s1d[j >> 3] += (double)warp_colsum8(acc, lane);
                        s2d[j >> 3] += (double)warp_colsum8(dx, lane);
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&phi_free);
        }
        if (MODE == 1 && p.s_partials != nullptr && (lane & 3) == 0) {
#pragma unroll
            for (int jj = 0; jj < CPT / 8; ++jj) {
                const int col = c0 + jj * 8 + (lane >> 2);
                if (col < p.f) {
                    atomicAdd(&red_s[col], s1d[jj]);
                    atomicAdd(&red_s[128 + col], s2d[jj]);
                }
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

// ---------------------------------------------------------------------------------------------------
// dW_e / db_e.  Per slot and row quarter a double-buffered stage (P^T hi | EA^T hi | EA^T lo).
template <int CPT>
__global__ void __launch_bounds__(G_THREADS, 1) dwe_tc_kernel(const GatherTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t p_full[4][2], p_empty[4][2], dwe_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int qq = 0; qq < 4; ++qq)
            for (int b = 0; b < 2; ++b) { mbar_init(&p_full[qq][b], 4); mbar_init(&p_empty[qq][b], 1); }
        mbar_init(&dwe_done, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < 8 * G_P_BYTES / 16; i += G_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    // row fe of every EA^T operand is all ones: column fe of the accumulator collects db_e = sum dphi
    for (int i = tid; i < 8 * 32; i += G_THREADS)
        *reinterpret_cast<float*>(smem + (size_t)(i >> 5) * G_P_BYTES + G_ATOM + atom_off(p.fe, i & 31)) = 1.0f;
    fence_proxy_async_smem();
    if (warp == G_NCW) tmem_alloc(&tmem_slot, 32);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_tiles = (p.n_rows + G_M - 1) / G_M;

    if (warp == G_NCW) {
        const uint32_t idesc = make_idesc_tf32(G_M, 32);
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int k = 0; k < 4; ++k, ++it) {
                if (lane == 0) {
                    const uint32_t b = it & 1, bu = it >> 1;
#pragma unroll 1
                    for (int qq = 0; qq < 4; ++qq) {
                        mbar_wait(&p_full[qq][b], bu & 1);
                        tc_fence_after_sync();
                        const uint32_t ph = smem_u32(smem + (size_t)(qq * 2 + b) * G_P_BYTES), eh = ph + G_ATOM, el = eh + 32 * 128;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t ko = kk * 32;
                            mma_tf32(tmem_base, make_desc(ph + ko), make_desc(eh + ko), idesc, (it > 0 || qq > 0 || kk > 0) ? 1u : 0u);
                            mma_tf32(tmem_base, make_desc(ph + ko), make_desc(el + ko), idesc, 1u);
                        }
                        mma_commit(&p_empty[qq][b]);
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0 && it > 0) mma_commit(&dwe_done);
        __syncwarp();
    } else {
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        const int c0 = grp * CPT;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t t = tile * G_M + row;
            const bool tv = t < p.n_rows;
            int4 nb4 = make_int4(-1, -1, -1, -1);
            if (tv) nb4 = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
            const int nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
            // h strip of this row
            float h[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j += 4) {
                const int f0 = c0 + j;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tv && f0 < p.f) {
                    v = ldg4(p.z_prev + (size_t)t * p.f + f0);
                    if (p.p_scale != nullptr) {
                        float4 sc = ldg4(p.p_scale + f0), sh = ldg4(p.p_shift + f0);
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                    }
                    if (p.p_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                }
                h[j] = v.x; h[j + 1] = v.y; h[j + 2] = v.z; h[j + 3] = v.w;
            }
            for (int k = 0; k < 4; ++k, ++it) {
                const uint32_t b = it & 1, bu = it >> 1;
                uint8_t* pq = smem + (size_t)(q * 2 + b) * G_P_BYTES;
                uint8_t* eh = pq + G_ATOM;
                uint8_t* el = eh + 32 * 128;
                const int s_row = nbv[k];
                // gathered d_agg strip (all loads first), and this warp's share of the EA row
                float4 da[CPT / 4];
#pragma unroll
                for (int j = 0; j < CPT; j += 8) {
                    da[j >> 2] = make_float4(0.f, 0.f, 0.f, 0.f);
                    da[(j >> 2) + 1] = da[j >> 2];
                    if ((p.f & 7) == 0) {
                        if (s_row >= 0 && c0 + j < p.f) ldg8(p.x + (size_t)s_row * p.f + c0 + j, da[j >> 2], da[(j >> 2) + 1]);
                    } else {
                        if (s_row >= 0 && c0 + j < p.f) da[j >> 2] = ldg4(p.x + (size_t)s_row * p.f + c0 + j);
                        if (s_row >= 0 && c0 + j + 4 < p.f) da[(j >> 2) + 1] = ldg4(p.x + (size_t)s_row * p.f + c0 + j + 4);
                    }
                }
                float4 ev[2];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int e0 = grp * 8 + h2 * 4;
                    ev[h2] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e0 < p.fe && tv && s_row >= 0) ev[h2] = ldg4(p.ea + ((size_t)t * 4 + k) * p.fe + e0);
                }
                mbar_wait(&p_empty[q][b], (bu & 1) ^ 1);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int e0 = grp * 8 + h2 * 4;
                    if (e0 >= p.fe) continue;
                    const float vv[4] = {ev[h2].x, ev[h2].y, ev[h2].z, ev[h2].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float hi, lo;
                        split_tf32(vv[i], hi, lo);
                        const uint32_t off = atom_off(e0 + i, lane);
                        *reinterpret_cast<float*>(eh + off) = hi;
                        *reinterpret_cast<float*>(el + off) = lo;
                    }
                }
#pragma unroll
                for (int j = 0; j < CPT; j += 4) {
                    const float dp[4] = {h[j] * da[j >> 2].x, h[j + 1] * da[j >> 2].y, h[j + 2] * da[j >> 2].z, h[j + 3] * da[j >> 2].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // row = c0 + j + i, column = lane; c0 is a multiple of 8: row & 7 = (j & 4) + i
                        const uint32_t r8 = (uint32_t)((j & 4) + i);
                        const uint32_t off = (uint32_t)(c0 + j + i) * 128u + ((((uint32_t)lane >> 2) ^ r8) << 4) + (((uint32_t)lane & 3u) << 2);
                        *reinterpret_cast<float*>(pq + off) = tf32_rna(dp[i]);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[q][b]);
            }
        }
        if (grp == 0) {
            float* outp = p.dwe_partials + (size_t)blockIdx.x * p.f * 32;
            float v[32];
            if (it > 0) {
                mbar_wait(&dwe_done, 0);
                tc_fence_after_sync();
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
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
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == G_NCW) tmem_dealloc(tmem_base, 32);
}

}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_gather_tc_supported(int f, int fe) {
    return (f % 4 == 0 && f >= 4 && f <= 128 && fe % 4 == 0 && fe >= 4 && fe <= 28) ? 1 : 0;
}

static int fp_of(int f) { return f <= 32 ? 32 : (f <= 64 ? 64 : 128); }

template <int MODE>
static int launch_gather_tc(const GatherTcArgs& p, cudaStream_t st, const char* what) {
    const int cpt = p.fp / 4;
    size_t smem = (size_t)2 * p.fp * 128 + (size_t)G_EA_STAGES * 2 * G_ATOM + 1024;
#define LAUNCH_G(CPT)                                                                                          \
    do {                                                                                                       \
        static bool configured = false;                                                                        \
        if (!configured) {                                                                                     \
            cudaError_t e = cudaFuncSetAttribute(gather_tc_kernel<CPT, MODE>,                                  \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);     \
            if (e != cudaSuccess) return fail(what, cudaGetErrorString(e));                                    \
            configured = true;                                                                                 \
        }                                                                                                      \
        gather_tc_kernel<CPT, MODE><<<sm_count(), G_THREADS, smem, st>>>(p);                                   \
    } while (0)
    switch (cpt) {
        case 8: LAUNCH_G(8); break;
        case 16: LAUNCH_G(16); break;
        case 32: LAUNCH_G(32); break;
        default: return fail(what, "unsupported feature width");
    }
#undef LAUNCH_G
    return check_launch(what);
}

static int launch_dwe_tc(const GatherTcArgs& p, cudaStream_t st, const char* what) {
    const int cpt = p.fp / 4;
    size_t smem = (size_t)8 * G_P_BYTES + 1024;
#define LAUNCH_D(CPT)                                                                                          \
    do {                                                                                                       \
        static bool configured = false;                                                                        \
        if (!configured) {                                                                                     \
            cudaError_t e = cudaFuncSetAttribute(dwe_tc_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 210 * 1024);                                                  \
            if (e != cudaSuccess) return fail(what, cudaGetErrorString(e));                                    \
            configured = true;                                                                                 \
        }                                                                                                      \
        dwe_tc_kernel<CPT><<<sm_count(), G_THREADS, smem, st>>>(p);                                            \
    } while (0)
    switch (cpt) {
        case 8: LAUNCH_D(8); break;
        case 16: LAUNCH_D(16); break;
        case 32: LAUNCH_D(32); break;
        default: return fail(what, "unsupported feature width");
    }
#undef LAUNCH_D
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
    p.n_rows = n_tgt; p.f = f_in; p.fp = fp_of(f_in); p.out = agg;
    return launch_gather_tc<0>(p, as_stream(stream), "dgnn_gather_tc_fwd");
}

extern "C" int dgnn_gather_tc_bwd(const float* d_agg, const float* d_self, const int32_t* onbr, const float* ea_own,
                                  int fe, const float* w_e, const float* b_e, const float* z_prev,
                                  const float* p_scale, const float* p_shift, const float* p_mean,
                                  const float* p_rstd, int p_relu, int64_t n_src, int64_t n_tgt, int f_in,
                                  float* dy_prev, double* s_partials, float* dwe_partials, void* stream) {
    DGNN_REQUIRE(dgnn_gather_tc_supported(f_in, fe), "widths not supported by the tensor-core gather");
    DGNN_REQUIRE(d_agg && onbr && ea_own && w_e && b_e, "null pointer");
    DGNN_REQUIRE(dwe_partials == nullptr || z_prev != nullptr, "dW_e needs the layer input (z_prev)");
    GatherTcArgs p;
    memset(&p, 0, sizeof(p));
    p.x = d_agg; p.nbr = onbr; p.ea = ea_own; p.w_e = w_e; p.b_e = b_e; p.fe = fe;
    p.n_rows = n_src; p.f = f_in; p.fp = fp_of(f_in); p.out = dy_prev;
    p.addend = d_self; p.n_add_rows = n_tgt;
    p.z_prev = z_prev; p.p_scale = p_scale; p.p_shift = p_shift; p.p_mean = p_mean; p.p_rstd = p_rstd;
    p.p_relu = p_relu; p.s_partials = s_partials; p.dwe_partials = dwe_partials;
    cudaStream_t st = as_stream(stream);
    if (dy_prev != nullptr || s_partials != nullptr) {
        int rc = launch_gather_tc<1>(p, st, "dgnn_gather_tc_bwd");
        if (rc) return rc;
    }
    if (dwe_partials != nullptr) return launch_dwe_tc(p, st, "dgnn_gather_tc_bwd(dW_e)");
    return 0;
}
