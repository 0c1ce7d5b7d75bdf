// Backward of the fused message-passing layer (generic-width FP32 path), atomic-free in
// global memory: every kernel writes per-CTA partials that a deterministic reduction sums.
//
// autograd's mirror of learning/surfaceNetStaticEdgeFilters.py:66-96 (+ BatchNorm + ReLU):
//   dgnn_rowdot_bwd  final Linear(F -> out_dim) and the ReLU/norm in front of it
//   dgnn_dense_bwd   dz (norm backward applied on load) . [W_j | W_i]  -> d_agg, d_self, db
//   dgnn_dw_bwd      dW_cat = dz^T . [agg | h]                            (split over cells)
//   dgnn_gather_bwd  dh[s] = d_self[s] + sum_k phi(ea_own[s,k]) * d_agg[onbr[s,k]]  and
//                    dW_e / db_e, through the out-edge ELL table (the symmetric adjacency's
//                    reverse-facet view), plus the ReLU mask and the (S1,S2) sums of the
//                    producer layer's normalisation.
#include "tile_gemm.cuh"

namespace dgnn {

// ---------------------------------------------------------------------------------------------------
// rowdot_bwd: thread owns a 4-column group; partial sums reduced through shared double atomics
template <int OD>
__global__ void __launch_bounds__(256) rowdot_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ z,
                                                          const float* __restrict__ sc, const float* __restrict__ sh,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          int relu, const float* __restrict__ w, long long n, int f,
                                                          float* __restrict__ dy, double* __restrict__ partials) {
    extern __shared__ double sred[];  // OD*f + OD + 2f
    const int plen = OD * f + OD + 2 * f;
    for (int i = threadIdx.x; i < plen; i += blockDim.x) sred[i] = 0.0;
    __syncthreads();
    const int f4 = f >> 2;
    const int rows_per_pass = blockDim.x / f4;  // f <= 1024 -> >= 1
    const int c = (threadIdx.x % f4) * 4;
    const int rsub = threadIdx.x / f4;
    const bool active = rsub < rows_per_pass;
    float4 wv[OD];
    float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (active) {
#pragma unroll
        for (int o = 0; o < OD; ++o) wv[o] = ldg4(w + (size_t)o * f + c);
        if (sc) { s4 = ldg4(sc + c); h4 = ldg4(sh + c); }
        if (mean) { m4 = ldg4(mean + c); r4 = ldg4(rstd + c); }
    }
    float dW[OD][4], db[OD], S1[4], S2[4];
#pragma unroll
    for (int o = 0; o < OD; ++o) { db[o] = 0.f; dW[o][0] = dW[o][1] = dW[o][2] = dW[o][3] = 0.f; }
    S1[0] = S1[1] = S1[2] = S1[3] = 0.f;
    S2[0] = S2[1] = S2[2] = S2[3] = 0.f;
    if (active) {
        for (long long r = (long long)blockIdx.x * rows_per_pass + rsub; r < n; r += (long long)gridDim.x * rows_per_pass) {
            float4 zv = ldg4(z + (size_t)r * f + c);
            float yv[4] = {fmaf(zv.x, s4.x, h4.x), fmaf(zv.y, s4.y, h4.y), fmaf(zv.z, s4.z, h4.z), fmaf(zv.w, s4.w, h4.w)};
            float xh[4] = {(zv.x - m4.x) * r4.x, (zv.y - m4.y) * r4.y, (zv.z - m4.z) * r4.z, (zv.w - m4.w) * r4.w};
            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 0; o < OD; ++o) {
                float d = __ldg(dl + (size_t)r * OD + o);
                const float wo[4] = {wv[o].x, wv[o].y, wv[o].z, wv[o].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j] = fmaf(d, wo[j], g[j]);
                    float av = relu ? fmaxf(yv[j], 0.f) : yv[j];
                    dW[o][j] = fmaf(d, av, dW[o][j]);
                }
                if (c == 0) db[o] += d;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (relu && !(yv[j] > 0.f)) g[j] = 0.f;
                S1[j] += g[j];
                S2[j] = fmaf(g[j], xh[j], S2[j]);
            }
            *reinterpret_cast<float4*>(dy + (size_t)r * f + c) = make_float4(g[0], g[1], g[2], g[3]);
        }
#pragma unroll
        for (int o = 0; o < OD; ++o) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&sred[o * f + c + j], (double)dW[o][j]);
            if (c == 0) atomicAdd(&sred[OD * f + o], (double)db[o]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sred[OD * f + OD + c + j], (double)S1[j]);
            atomicAdd(&sred[OD * f + OD + f + c + j], (double)S2[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plen; i += blockDim.x) partials[(size_t)blockIdx.x * plen + i] = sred[i];
}

// ---------------------------------------------------------------------------------------------------
struct NormBwd {
    const float* g;
    const float* a;
    const float* b;
    const float* mean;
    const float* rstd;
};

// dz = g*dy - (a + xhat*b) for 4 consecutive channels starting at c (identity when g == NULL)
__device__ __forceinline__ float4 dz_of(const NormBwd& nb, float4 dy, float4 z, int c) {
    if (nb.g == nullptr) return dy;
    float4 g = ldg4(nb.g + c), a = ldg4(nb.a + c), b = ldg4(nb.b + c), m = ldg4(nb.mean + c), r = ldg4(nb.rstd + c);
    float4 o;
    o.x = g.x * dy.x - (a.x + (z.x - m.x) * r.x * b.x);
    o.y = g.y * dy.y - (a.y + (z.y - m.y) * r.y * b.y);
    o.z = g.z * dy.z - (a.z + (z.z - m.z) * r.z * b.z);
    o.w = g.w * dy.w - (a.w + (z.w - m.w) * r.w * b.w);
    return o;
}

struct DenseBwdArgs {
    const float* dy;
    const float* z;
    NormBwd nb;
    const float* w_cat;  // [f_out, k_total]
    const int32_t* nbr;
    int64_t n_tgt;
    int f_in, f_out, k_total, lda;
    float* d_agg;
    float* d_self;
    double* db_partials;
};

__global__ void __launch_bounds__(NT, 2) dense_bwd_kernel(const DenseBwdArgs p) {
    extern __shared__ __align__(16) float smem[];
    float* a_s = smem;                  // TM * lda  (dz tile)
    float* w_s = a_s + TM * p.lda;      // 2*TK*TN
    float* red = w_s + 2 * TK * TN;     // f_out (column sums of dz for this tile)
    float* icnt = red + ((p.f_out + 3) & ~3);  // TM (1/max(cnt,1))
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int F = p.f_out, f4 = F >> 2;
    const int padw = p.lda - F;
    for (int idx = tid; idx < TM * padw; idx += NT) a_s[(idx / padw) * p.lda + F + idx % padw] = 0.f;
    double* my_db = p.db_partials ? p.db_partials + (size_t)blockIdx.x * F : nullptr;
    if (my_db)
        for (int c = tid; c < F; c += NT) my_db[c] = 0.0;
    const int64_t n_tiles = (p.n_tgt + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile0 = tile * TM;
        __syncthreads();
        for (int c = tid; c < F; c += NT) red[c] = 0.f;
        if (tid < TM) {
            float ic = 1.f;
            int64_t t = tile0 + tid;
            if (p.nbr && t < p.n_tgt) {
                int4 nb = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
                int cnt = (nb.x >= 0) + (nb.y >= 0) + (nb.z >= 0) + (nb.w >= 0);
                ic = 1.f / (float)(cnt > 0 ? cnt : 1);
            }
            icnt[tid] = ic;
        }
        __syncthreads();
        // when NT % f4 == 0 a thread keeps one column group while striding rows, so its column sums
        // stay in registers; otherwise fall back to one shared atomic per element.
        const bool fixed_col = (NT % f4) == 0;
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int idx = tid; idx < TM * f4; idx += NT) {
            int r = idx / f4, c = (idx % f4) * 4;
            int64_t t = tile0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < p.n_tgt) {
                float4 dyv = ldg4(p.dy + (size_t)t * F + c);
                float4 zv = p.nb.g ? ldg4(p.z + (size_t)t * F + c) : dyv;
                v = dz_of(p.nb, dyv, zv, c);
                if (my_db) {
                    if (fixed_col) {
                        csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
                    } else {
                        atomicAdd(&red[c], v.x); atomicAdd(&red[c + 1], v.y);
                        atomicAdd(&red[c + 2], v.z); atomicAdd(&red[c + 3], v.w);
                    }
                }
            }
            *reinterpret_cast<float4*>(a_s + r * p.lda + c) = v;
        }
        if (my_db && fixed_col && tid < TM * f4) {
            int c = (tid % f4) * 4;
            atomicAdd(&red[c], csum.x); atomicAdd(&red[c + 1], csum.y);
            atomicAdd(&red[c + 2], csum.z); atomicAdd(&red[c + 3], csum.w);
        }
        __syncthreads();
        if (my_db)
            for (int c = tid; c < F; c += NT) my_db[c] += (double)red[c];
        for (int n0 = 0; n0 < p.k_total; n0 += TN) {
            float acc[4][8];
            tile_gemm(acc, a_s, p.lda, p.w_cat, F, p.k_total, n0, w_s);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + h * 64 + tx * 4;
                if (n >= p.k_total) continue;
                const bool is_agg = p.nbr != nullptr && n < p.f_in;
                float* dst = is_agg ? p.d_agg : p.d_self;
                const int col = is_agg ? n : (p.nbr ? n - p.f_in : n);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = ty * 4 + i;
                    const int64_t t = tile0 + r;
                    if (t >= p.n_tgt) continue;
                    float s = is_agg ? icnt[r] : 1.f;
                    float4 o = make_float4(acc[i][h * 4] * s, acc[i][h * 4 + 1] * s, acc[i][h * 4 + 2] * s,
                                           acc[i][h * 4 + 3] * s);
                    *reinterpret_cast<float4*>(dst + (size_t)t * p.f_in + col) = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// dW: C[64 x 128 block of (f_out x k_total)] += dz^T . [agg | h], split over cells
constexpr int DW_BM = 64, DW_BN = 128, DW_KC = 32;

struct DwArgs {
    const float* dy;
    const float* z;
    NormBwd nb;
    const float* agg;
    const float* x_in;
    const float* in_scale;
    const float* in_shift;
    int relu_in;
    int64_t n_tgt;
    int f_in, f_out, k_total;
    int chunks_m, chunks_n, splits;
    float* partials;
};

__global__ void __launch_bounds__(NT, 2) dw_bwd_kernel(const DwArgs p) {
    __shared__ __align__(16) float dz_s[DW_KC][DW_BM + 4];
    __shared__ __align__(16) float a_s[DW_KC][DW_BN + 4];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int chunk = blockIdx.x % (p.chunks_m * p.chunks_n);
    const int split = blockIdx.x / (p.chunks_m * p.chunks_n);
    const int m0 = (chunk / p.chunks_n) * DW_BM, n0 = (chunk % p.chunks_n) * DW_BN;
    const int64_t per = ((p.n_tgt + p.splits - 1) / p.splits + DW_KC - 1) / DW_KC * DW_KC;
    const int64_t c_begin = (int64_t)split * per;
    const int64_t c_end = c_begin + per < p.n_tgt ? c_begin + per : p.n_tgt;
    const bool relu = p.relu_in != 0;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int64_t t0 = c_begin; t0 < c_end; t0 += DW_KC) {
        __syncthreads();
        // dz sub-tile: DW_KC cells x 64 channels
        for (int idx = tid; idx < DW_KC * (DW_BM / 4); idx += NT) {
            int r = idx / (DW_BM / 4), c = m0 + (idx % (DW_BM / 4)) * 4;
            int64_t t = t0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < c_end && c < p.f_out) {
                float4 dyv = ldg4(p.dy + (size_t)t * p.f_out + c);
                float4 zv = p.nb.g ? ldg4(p.z + (size_t)t * p.f_out + c) : dyv;
                v = dz_of(p.nb, dyv, zv, c);
            }
            *reinterpret_cast<float4*>(&dz_s[r][c - m0]) = v;
        }
        // [agg | h] sub-tile: DW_KC cells x 128 columns
        for (int idx = tid; idx < DW_KC * (DW_BN / 4); idx += NT) {
            int r = idx / (DW_BN / 4), n = n0 + (idx % (DW_BN / 4)) * 4;
            int64_t t = t0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < c_end && n < p.k_total) {
                if (p.agg != nullptr && n < p.f_in) {
                    v = ldg4(p.agg + (size_t)t * p.f_in + n);
                } else {
                    int c = p.agg != nullptr ? n - p.f_in : n;
                    v = ldg4(p.x_in + (size_t)t * p.f_in + c);
                    if (p.in_scale) {
                        float4 s4 = ldg4(p.in_scale + c), h4 = ldg4(p.in_shift + c);
                        v.x = act(v.x, s4.x, h4.x, relu); v.y = act(v.y, s4.y, h4.y, relu);
                        v.z = act(v.z, s4.z, h4.z, relu); v.w = act(v.w, s4.w, h4.w, relu);
                    } else if (relu) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    }
                }
            }
            *reinterpret_cast<float4*>(&a_s[r][n - n0]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < DW_KC; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&dz_s[k][ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&a_s[k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&a_s[k][64 + tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(av[i], b0.x, acc[i][0]); acc[i][1] = fmaf(av[i], b0.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b0.z, acc[i][2]); acc[i][3] = fmaf(av[i], b0.w, acc[i][3]);
                acc[i][4] = fmaf(av[i], b1.x, acc[i][4]); acc[i][5] = fmaf(av[i], b1.y, acc[i][5]);
                acc[i][6] = fmaf(av[i], b1.z, acc[i][6]); acc[i][7] = fmaf(av[i], b1.w, acc[i][7]);
            }
        }
    }
    float* out = p.partials + (size_t)split * p.f_out * p.k_total;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.f_out) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int n = n0 + h * 64 + tx * 4;
            if (n >= p.k_total) continue;
            *reinterpret_cast<float4*>(out + (size_t)m * p.k_total + n) =
                make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
struct GatherBwdArgs {
    const float* d_agg;
    const float* d_self;
    const int32_t* onbr;
    const float* ea_own;
    const float* w_e;
    const float* b_e;
    const float* x_in;
    const float* in_scale;
    const float* in_shift;
    const float* in_mean;
    const float* in_rstd;
    int relu_in;
    int64_t n_src, n_tgt;
    int f_in;
    float* dy_prev;
    double* partials;
    int only_dwe;   // 1: only the edge-filter gradients (dW_e, db_e); dh / dy_prev / (S1,S2) are skipped
};

template <int FE>
__global__ void __launch_bounds__(NT, 2) gather_bwd_kernel(const GatherBwdArgs p) {
    extern __shared__ double sred[];  // f_in*(FE+1) + 2*f_in
    const int F = p.f_in;
    const int plen = F * (FE + 1) + 2 * F;
    for (int i = threadIdx.x; i < plen; i += blockDim.x) sred[i] = 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lpc = F > 32 ? 32 : (F > 16 ? 16 : (F > 8 ? 8 : 4));
    const int ipw = 32 / lpc;
    const int sub = lane / lpc, li = lane % lpc;
    const int nch = (F + 63) >> 6;
    const bool relu = p.relu_in != 0;
    const int64_t warp_g = (int64_t)blockIdx.x * (NT / 32) + warp;
    const int64_t n_warps = (int64_t)gridDim.x * (NT / 32);
    for (int c = 0; c < nch; ++c) {
        const int f = c * 64 + li * 2;
        const bool fv = f < F;
        float we0[FE > 0 ? FE : 1], we1[FE > 0 ? FE : 1], dwe0[FE > 0 ? FE : 1], dwe1[FE > 0 ? FE : 1];
        float be0 = 1.f, be1 = 1.f, dbe0 = 0.f, dbe1 = 0.f;
        if (FE > 0) {
#pragma unroll
            for (int j = 0; j < FE; ++j) {
                we0[j] = fv ? __ldg(p.w_e + (size_t)f * FE + j) : 0.f;
                we1[j] = fv ? __ldg(p.w_e + (size_t)(f + 1) * FE + j) : 0.f;
                dwe0[j] = dwe1[j] = 0.f;
            }
            be0 = fv ? __ldg(p.b_e + f) : 0.f;
            be1 = fv ? __ldg(p.b_e + f + 1) : 0.f;
        }
        float sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f, mu0 = 0.f, mu1 = 0.f, rs0 = 1.f, rs1 = 1.f;
        if (fv && p.in_scale) {
            sc0 = __ldg(p.in_scale + f); sc1 = __ldg(p.in_scale + f + 1);
            sh0 = __ldg(p.in_shift + f); sh1 = __ldg(p.in_shift + f + 1);
        }
        if (fv && p.in_mean) {
            mu0 = __ldg(p.in_mean + f); mu1 = __ldg(p.in_mean + f + 1);
            rs0 = __ldg(p.in_rstd + f); rs1 = __ldg(p.in_rstd + f + 1);
        }
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
        for (int64_t s = warp_g * ipw + sub; s < p.n_src; s += n_warps * ipw) {
            int4 ob = __ldg(reinterpret_cast<const int4*>(p.onbr) + s);
            const int ov[4] = {ob.x, ob.y, ob.z, ob.w};
            float2 da[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                da[k] = make_float2(0.f, 0.f);
                if (ov[k] >= 0 && fv) da[k] = ldg2(p.d_agg + (size_t)ov[k] * F + f);
            }
            float2 xv = make_float2(0.f, 0.f);
            if (fv) xv = ldg2(p.x_in + (size_t)s * F + f);
            float y0 = fmaf(xv.x, sc0, sh0), y1 = fmaf(xv.y, sc1, sh1);
            float h0 = relu ? fmaxf(y0, 0.f) : y0, h1 = relu ? fmaxf(y1, 0.f) : y1;
            float dh0 = 0.f, dh1 = 0.f;
            if (s < p.n_tgt && fv && p.d_self) {
                float2 ds = ldg2(p.d_self + (size_t)s * F + f);
                dh0 = ds.x; dh1 = ds.y;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ov[k] < 0) continue;
                float ph0 = be0, ph1 = be1;
                float dp0 = h0 * da[k].x, dp1 = h1 * da[k].y;
                if (FE > 0) {
                    const float* er = p.ea_own + ((size_t)s * 4 + k) * FE;
#pragma unroll
                    for (int j = 0; j < FE; j += 4) {
                        float4 e = ldg4(er + j);
                        if (!p.only_dwe) {
                            ph0 = fmaf(we0[j], e.x, ph0); ph1 = fmaf(we1[j], e.x, ph1);
                            ph0 = fmaf(we0[j + 1], e.y, ph0); ph1 = fmaf(we1[j + 1], e.y, ph1);
                            ph0 = fmaf(we0[j + 2], e.z, ph0); ph1 = fmaf(we1[j + 2], e.z, ph1);
                            ph0 = fmaf(we0[j + 3], e.w, ph0); ph1 = fmaf(we1[j + 3], e.w, ph1);
                        }
                        dwe0[j] = fmaf(dp0, e.x, dwe0[j]); dwe1[j] = fmaf(dp1, e.x, dwe1[j]);
                        dwe0[j + 1] = fmaf(dp0, e.y, dwe0[j + 1]); dwe1[j + 1] = fmaf(dp1, e.y, dwe1[j + 1]);
                        dwe0[j + 2] = fmaf(dp0, e.z, dwe0[j + 2]); dwe1[j + 2] = fmaf(dp1, e.z, dwe1[j + 2]);
                        dwe0[j + 3] = fmaf(dp0, e.w, dwe0[j + 3]); dwe1[j + 3] = fmaf(dp1, e.w, dwe1[j + 3]);
                    }
                    dbe0 += dp0; dbe1 += dp1;
                }
                dh0 = fmaf(ph0, da[k].x, dh0);
                dh1 = fmaf(ph1, da[k].y, dh1);
            }
            if (p.dy_prev && fv) {
                if (relu && !(y0 > 0.f)) dh0 = 0.f;
                if (relu && !(y1 > 0.f)) dh1 = 0.f;
                *reinterpret_cast<float2*>(p.dy_prev + (size_t)s * F + f) = make_float2(dh0, dh1);
                s1a += dh0; s1b += dh1;
                s2a = fmaf(dh0, (xv.x - mu0) * rs0, s2a);
                s2b = fmaf(dh1, (xv.y - mu1) * rs1, s2b);
            }
        }
        if (fv) {
            if (FE > 0) {
#pragma unroll
                for (int j = 0; j < FE; ++j) {
                    atomicAdd(&sred[f * FE + j], (double)dwe0[j]);
                    atomicAdd(&sred[(f + 1) * FE + j], (double)dwe1[j]);
                }
                atomicAdd(&sred[F * FE + f], (double)dbe0);
                atomicAdd(&sred[F * FE + f + 1], (double)dbe1);
            }
            atomicAdd(&sred[F * (FE + 1) + f], (double)s1a);
            atomicAdd(&sred[F * (FE + 1) + f + 1], (double)s1b);
            atomicAdd(&sred[F * (FE + 1) + F + f], (double)s2a);
            atomicAdd(&sred[F * (FE + 1) + F + f + 1], (double)s2b);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plen; i += blockDim.x) p.partials[(size_t)blockIdx.x * plen + i] = sred[i];
}

}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_rowdot_bwd(const float* dlogits, const float* z_in, const float* in_scale, const float* in_shift,
                               const float* mean, const float* rstd, int relu_in, const float* w, int64_t n, int f,
                               int od, float* dy, double* partials, void* stream) {
    DGNN_REQUIRE(f % 4 == 0 && f <= 1024, "f must be a multiple of 4 and <= 1024");
    DGNN_REQUIRE(od >= 1 && od <= 4, "1 <= od <= 4");
    int grid = dgnn_small_grid();
    size_t smem = (size_t)(od * f + od + 2 * f) * sizeof(double);
    cudaStream_t st = as_stream(stream);
#define LAUNCH_RD(OD)                                                                                          \
    rowdot_bwd_kernel<OD><<<grid, 256, smem, st>>>(dlogits, z_in, in_scale, in_shift, mean, rstd, relu_in, w, n, f, \
                                                   dy, partials)
    switch (od) {
        case 1: LAUNCH_RD(1); break;
        case 2: LAUNCH_RD(2); break;
        case 3: LAUNCH_RD(3); break;
        default: LAUNCH_RD(4); break;
    }
#undef LAUNCH_RD
    return check_launch("dgnn_rowdot_bwd");
}

extern "C" int dgnn_dense_bwd(const float* dy, const float* z, const float* g, const float* a, const float* b,
                              const float* mean, const float* rstd, const float* w_cat, const int32_t* nbr,
                              int64_t n_tgt, int f_in, int f_out, int k_total, float* d_agg, float* d_self,
                              double* db_partials, void* stream) {
    DGNN_REQUIRE(f_in % 4 == 0 && f_out % 4 == 0, "widths must be multiples of 4");
    DGNN_REQUIRE(k_total == (nbr ? 2 * f_in : f_in), "k_total mismatch");
    DGNN_REQUIRE(dy && w_cat && d_self && (!nbr || d_agg), "null pointer");
    DenseBwdArgs p;
    p.dy = dy; p.z = z; p.nb = NormBwd{g, a, b, mean, rstd};
    p.w_cat = w_cat; p.nbr = nbr; p.n_tgt = n_tgt; p.f_in = f_in; p.f_out = f_out; p.k_total = k_total;
    p.lda = (f_out + TK - 1) / TK * TK + 4;
    p.d_agg = d_agg; p.d_self = d_self; p.db_partials = db_partials;
    size_t smem = ((size_t)TM * p.lda + 2 * TK * TN + ((f_out + 3) & ~3) + TM) * sizeof(float);
    DGNN_REQUIRE(smem <= 200 * 1024, "layer too wide for the generic FP32 path");
    if (int rc_ = ensure_dyn_smem((const void*)dense_bwd_kernel, 200 * 1024, "dgnn_dense_bwd")) return rc_;
    int grid = dgnn_layer_grid(f_in, f_out);
    dense_bwd_kernel<<<grid, NT, smem, as_stream(stream)>>>(p);
    return check_launch("dgnn_dense_bwd");
}

extern "C" int dgnn_dw_splits(int f_out, int k_total) {
    int chunks = ((f_out + DW_BM - 1) / DW_BM) * ((k_total + DW_BN - 1) / DW_BN);
    int s = (2 * sm_count()) / chunks;
    return s < 1 ? 1 : s;
}

extern "C" int dgnn_dw_bwd(const float* dy, const float* z, const float* g, const float* a, const float* b,
                           const float* mean, const float* rstd, const float* agg, const float* x_in,
                           const float* in_scale, const float* in_shift, int relu_in, int64_t n_tgt, int f_in,
                           int f_out, int k_total, float* partials, void* stream) {
    DGNN_REQUIRE(f_in % 4 == 0 && f_out % 4 == 0, "widths must be multiples of 4");
    DGNN_REQUIRE(k_total == (agg ? 2 * f_in : f_in), "k_total mismatch");
    DwArgs p;
    p.dy = dy; p.z = z; p.nb = NormBwd{g, a, b, mean, rstd};
    p.agg = agg; p.x_in = x_in; p.in_scale = in_scale; p.in_shift = in_shift; p.relu_in = relu_in;
    p.n_tgt = n_tgt; p.f_in = f_in; p.f_out = f_out; p.k_total = k_total;
    p.chunks_m = (f_out + DW_BM - 1) / DW_BM;
    p.chunks_n = (k_total + DW_BN - 1) / DW_BN;
    p.splits = dgnn_dw_splits(f_out, k_total);
    p.partials = partials;
    dw_bwd_kernel<<<p.chunks_m * p.chunks_n * p.splits, NT, 0, as_stream(stream)>>>(p);
    return check_launch("dgnn_dw_bwd");
}

extern "C" int dgnn_gather_bwd_grid(int f_in) {
    (void)f_in;
    return sm_count() * 2;
}

template <int FE>
static int launch_gather_bwd(const GatherBwdArgs& p, cudaStream_t st) {
    size_t smem = (size_t)(p.f_in * (FE + 1) + 2 * p.f_in) * sizeof(double);
    if (smem > 200 * 1024) return fail("dgnn_gather_bwd", "f_in too wide for the shared reduction");
    if (int rc_ = ensure_dyn_smem((const void*)gather_bwd_kernel<FE>, 200 * 1024, "dgnn_gather_bwd")) return rc_;
    gather_bwd_kernel<FE><<<dgnn_gather_bwd_grid(p.f_in), NT, smem, st>>>(p);
    return check_launch("dgnn_gather_bwd");
}

static int g_only_dwe = 0;

// dW_e / db_e only (same partial layout as dgnn_gather_bwd; its S1/S2 part is zero)
extern "C" int dgnn_edge_filter_bwd(const float* d_agg, const int32_t* onbr, const float* ea_own, int fe,
                                    const float* w_e, const float* b_e, const float* x_in, const float* in_scale,
                                    const float* in_shift, int relu_in, int64_t n_src, int64_t n_tgt, int f_in,
                                    double* partials, void* stream);

extern "C" int dgnn_gather_bwd(const float* d_agg, const float* d_self, const int32_t* onbr, const float* ea_own,
                               int fe, const float* w_e, const float* b_e, const float* x_in, const float* in_scale,
                               const float* in_shift, const float* in_mean, const float* in_rstd, int relu_in,
                               int64_t n_src, int64_t n_tgt, int f_in, float* dy_prev, double* partials, void* stream) {
    DGNN_REQUIRE(f_in % 4 == 0, "f_in must be a multiple of 4");
    if (w_e == nullptr) fe = 0;
    DGNN_REQUIRE(fe % 4 == 0 && fe <= 32, "edge feature width must be a multiple of 4 and <= 32");
    DGNN_REQUIRE(d_agg && onbr && x_in && partials, "null pointer");
    GatherBwdArgs p;
    p.d_agg = d_agg; p.d_self = d_self; p.onbr = onbr; p.ea_own = ea_own; p.w_e = w_e; p.b_e = b_e;
    p.x_in = x_in; p.in_scale = in_scale; p.in_shift = in_shift; p.in_mean = in_mean; p.in_rstd = in_rstd;
    p.relu_in = relu_in; p.n_src = n_src; p.n_tgt = n_tgt; p.f_in = f_in; p.dy_prev = dy_prev; p.partials = partials;
    p.only_dwe = (d_self == nullptr && dy_prev == nullptr && in_mean == nullptr && fe > 0) ? g_only_dwe : 0;
    cudaStream_t st = as_stream(stream);
    switch (fe) {
        case 0: return launch_gather_bwd<0>(p, st);
        case 4: return launch_gather_bwd<4>(p, st);
        case 8: return launch_gather_bwd<8>(p, st);
        case 12: return launch_gather_bwd<12>(p, st);
        case 16: return launch_gather_bwd<16>(p, st);
        case 20: return launch_gather_bwd<20>(p, st);
        case 24: return launch_gather_bwd<24>(p, st);
        case 28: return launch_gather_bwd<28>(p, st);
        case 32: return launch_gather_bwd<32>(p, st);
    }
    return fail("dgnn_gather_bwd", "unsupported edge feature width");
}

extern "C" int dgnn_edge_filter_bwd(const float* d_agg, const int32_t* onbr, const float* ea_own, int fe,
                                    const float* w_e, const float* b_e, const float* x_in, const float* in_scale,
                                    const float* in_shift, int relu_in, int64_t n_src, int64_t n_tgt, int f_in,
                                    double* partials, void* stream) {
    DGNN_REQUIRE(fe > 0 && w_e != nullptr, "no edge filter");
    g_only_dwe = 1;
    int rc = dgnn_gather_bwd(d_agg, nullptr, onbr, ea_own, fe, w_e, b_e, x_in, in_scale, in_shift, nullptr, nullptr,
                             relu_in, n_src, n_tgt, f_in, nullptr, partials, stream);
    g_only_dwe = 0;
    return rc;
}
