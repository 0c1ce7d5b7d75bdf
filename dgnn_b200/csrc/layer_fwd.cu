// Fused message-passing layer, forward (generic-width FP32 path).
//
// Replaces, per layer, the reference's lin_e GEMM, index_select, Hadamard product,
// scatter-mean, two node GEMMs, BatchNorm and ReLU
// (learning/surfaceNetStaticEdgeFilters.py:66-96,:217-219 and the PyG / torch_scatter kernels
// they reach): the E x F edge tensors are never materialised.
//
// One persistent CTA processes tiles of TM target cells:
//   1. gather: per (cell, 64-feature chunk) a group of lanes owns 2 features each, keeps the
//      matching rows of W_e in registers, loads the 4 neighbour rows (float2 per lane, whole
//      rows coalesced), applies the producer's norm affine + ReLU on load, evaluates the edge
//      filter phi = W_e . ea + b_e and accumulates agg = mean_k h(nbr_k) * phi_k into the shared
//      A tile [agg | h(self)].
//   2. dense: z = A . [W_j ; W_i]^T through tile_gemm (weights streamed from L2).
//   3. epilogue: bias / eval-norm affine + ReLU, coalesced float4 stores, per-channel
//      (sum, sum^2) partials for training BatchNorm.
#include "tile_gemm.cuh"

namespace dgnn {

struct LayerFwdArgs {
    const float* x_in;
    const float* in_scale;
    const float* in_shift;
    int relu_in;
    const int32_t* nbr;
    const float* ea;
    const float* w_e;
    const float* b_e;
    const float* wt_cat;
    const float* bias;
    const float* out_scale;
    const float* out_shift;
    int relu_out;
    int64_t n_tgt;
    int f_in, f_out, k_total, kp, lda;
    float* out;
    float* agg_save;
    double* stats;
};

template <int FE>
__device__ __forceinline__ void gather_tile(const LayerFwdArgs& p, float* a_s, int64_t tile0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = p.f_in;
    const int lpc = F > 32 ? 32 : (F > 16 ? 16 : (F > 8 ? 8 : 4));  // lanes per (cell, chunk) item
    const int ipw = 32 / lpc;                                       // items per warp pass
    const int sub = lane / lpc, li = lane % lpc;
    const int nch = (F + 63) >> 6;
    const bool relu = (p.relu_in & 1) != 0;
    const bool self_loop = (p.relu_in & 2) != 0;   // run.py:70-71,215-216: add_self_loops (edge_convs == 0 only)
    for (int c = 0; c < nch; ++c) {
        const int f = c * 64 + li * 2;
        const bool fv = f < F;
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
        for (int cell = warp * ipw + sub; cell < TM; cell += (NT / 32) * ipw) {
            const int64_t t = tile0 + cell;
            const bool tv = t < p.n_tgt;
            int4 nb = make_int4(-1, -1, -1, -1);
            if (tv) nb = __ldg(reinterpret_cast<const int4*>(p.nbr) + t);
            const int nbv[4] = {nb.x, nb.y, nb.z, nb.w};
            float2 xs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xs[k] = make_float2(0.f, 0.f);
                if (nbv[k] >= 0 && fv) xs[k] = ldg2(p.x_in + (size_t)nbv[k] * F + f);
            }
            float2 self = make_float2(0.f, 0.f);
            if (tv && fv) self = ldg2(p.x_in + (size_t)t * F + f);
            float a0 = 0.f, a1 = 0.f;
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (nbv[k] < 0) continue;  // uniform across the item's lanes
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
            if (fv) {
                float s0 = tv ? act(self.x, sc0, sh0, relu) : 0.f;
                float s1 = tv ? act(self.y, sc1, sh1, relu) : 0.f;
                if (self_loop && tv) {                 // the cell is its own fifth in-neighbour (no edge filter: phi = 1)
                    a0 += s0; a1 += s1; ++cnt;
                }
                float d = (float)(cnt > 0 ? cnt : 1);
                a0 = a0 / d;
                a1 = a1 / d;
                *reinterpret_cast<float2*>(a_s + cell * p.lda + f) = make_float2(a0, a1);
                *reinterpret_cast<float2*>(a_s + cell * p.lda + F + f) = make_float2(s0, s1);
                if (p.agg_save != nullptr && tv)
                    *reinterpret_cast<float2*>(p.agg_save + (size_t)t * F + f) = make_float2(a0, a1);
            }
        }
    }
}

// dense mode (no gather): A tile = h(x_in[tile rows])
__device__ __forceinline__ void load_tile_dense(const LayerFwdArgs& p, float* a_s, int64_t tile0) {
    const int F = p.f_in;
    const int f4 = F >> 2;
    const bool relu = (p.relu_in & 1) != 0;
    for (int idx = threadIdx.x; idx < TM * f4; idx += NT) {
        int r = idx / f4, c = (idx % f4) * 4;
        int64_t t = tile0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < p.n_tgt) {
            v = ldg4(p.x_in + (size_t)t * F + c);
            if (p.in_scale != nullptr) {
                float4 sc = ldg4(p.in_scale + c), sh = ldg4(p.in_shift + c);
                v.x = act(v.x, sc.x, sh.x, relu); v.y = act(v.y, sc.y, sh.y, relu);
                v.z = act(v.z, sc.z, sh.z, relu); v.w = act(v.w, sc.w, sh.w, relu);
            } else if (relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
        }
        *reinterpret_cast<float4*>(a_s + r * p.lda + c) = v;
    }
}

template <int FE>
__global__ void __launch_bounds__(NT, 2) layer_fwd_kernel(const LayerFwdArgs p) {
    extern __shared__ __align__(16) float smem[];
    float* a_s = smem;                              // TM * lda
    float* w_s = a_s + TM * p.lda;                  // 2 * TK * TN
    float* red = w_s + 2 * TK * TN;                 // 2 * TN
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    // zero the k-padding columns once (never overwritten afterwards)
    const int padw = p.lda - p.k_total;
    for (int idx = tid; idx < TM * padw; idx += NT)
        a_s[(idx / padw) * p.lda + p.k_total + idx % padw] = 0.f;
    double* my_stats = p.stats ? p.stats + (size_t)blockIdx.x * 2 * p.f_out : nullptr;
    if (my_stats)
        for (int c = tid; c < 2 * p.f_out; c += NT) my_stats[c] = 0.0;
    const int64_t n_tiles = (p.n_tgt + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile0 = tile * TM;
        __syncthreads();  // previous tile's readers of a_s / red are done
        if (p.nbr != nullptr) gather_tile<FE>(p, a_s, tile0);
        else load_tile_dense(p, a_s, tile0);
        __syncthreads();
        for (int n0 = 0; n0 < p.f_out; n0 += TN) {
            float acc[4][8];
            tile_gemm(acc, a_s, p.lda, p.wt_cat, p.k_total, p.f_out, n0, w_s);
            if (my_stats) {
                for (int c = tid; c < 2 * TN; c += NT) red[c] = 0.f;
                __syncthreads();
            }
            float cs[8], cq[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cs[j] = cq[j] = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + h * 64 + tx * 4;
                if (n >= p.f_out) continue;
                float4 bi = p.bias ? ldg4(p.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 os = make_float4(1.f, 1.f, 1.f, 1.f), oh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.out_scale) { os = ldg4(p.out_scale + n); oh = ldg4(p.out_shift + n); }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t t = tile0 + ty * 4 + i;
                    if (t >= p.n_tgt) continue;
                    float4 z = make_float4(acc[i][h * 4 + 0] + bi.x, acc[i][h * 4 + 1] + bi.y,
                                           acc[i][h * 4 + 2] + bi.z, acc[i][h * 4 + 3] + bi.w);
                    cs[h * 4 + 0] += z.x; cq[h * 4 + 0] = fmaf(z.x, z.x, cq[h * 4 + 0]);
                    cs[h * 4 + 1] += z.y; cq[h * 4 + 1] = fmaf(z.y, z.y, cq[h * 4 + 1]);
                    cs[h * 4 + 2] += z.z; cq[h * 4 + 2] = fmaf(z.z, z.z, cq[h * 4 + 2]);
                    cs[h * 4 + 3] += z.w; cq[h * 4 + 3] = fmaf(z.w, z.w, cq[h * 4 + 3]);
                    if (p.out_scale) {
                        z.x = fmaf(z.x, os.x, oh.x); z.y = fmaf(z.y, os.y, oh.y);
                        z.z = fmaf(z.z, os.z, oh.z); z.w = fmaf(z.w, os.w, oh.w);
                    }
                    if (p.relu_out) {
                        z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(p.out + (size_t)t * p.f_out + n) = z;
                }
            }
            if (my_stats) {
                // combine the two row groups of a warp, then one shared atomic per warp and column
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
                    cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 16);
                }
                if ((tid & 16) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int cl = (j >> 2) * 64 + tx * 4 + (j & 3);
                        atomicAdd(&red[cl], cs[j]);
                        atomicAdd(&red[TN + cl], cq[j]);
                    }
                }
                __syncthreads();
                if (tid < TN && n0 + tid < p.f_out) {
                    my_stats[n0 + tid] += (double)red[tid];
                    my_stats[p.f_out + n0 + tid] += (double)red[TN + tid];
                }
                __syncthreads();
            }
        }
    }
}

// agg[t] = (1/max(cnt,1)) * sum_k h(x[nbr[t,k]]) (*) phi[t,k,:] with the edge filter given as a matrix
// (the Updated-edge-filter variant carries the edge state from layer to layer,
// learning/surfaceNetUpdatedEdgeFilters.py:157-176,236-241).  One thread per (cell, 4 features).
__global__ void __launch_bounds__(256) gather_phi_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                                         const float* __restrict__ sh, int relu,
                                                         const int32_t* __restrict__ nbr, const float* __restrict__ phi,
                                                         long long n_tgt, int f, float* __restrict__ agg) {
    const int f4 = f >> 2;
    const long long total = n_tgt * f4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / f4;
        const int c = (int)(i % f4) * 4;
        const int4 nb = __ldg(reinterpret_cast<const int4*>(nbr) + t);
        const int nbv[4] = {nb.x, nb.y, nb.z, nb.w};
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sc) { s4 = ldg4(sc + c); h4 = ldg4(sh + c); }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (nbv[k] < 0) continue;
            ++cnt;
            float4 v = ldg4(x + (size_t)nbv[k] * f + c);
            float4 ph = ldg4(phi + ((size_t)t * 4 + k) * f + c);
            if (sc) {
                v.x = act(v.x, s4.x, h4.x, relu); v.y = act(v.y, s4.y, h4.y, relu);
                v.z = act(v.z, s4.z, h4.z, relu); v.w = act(v.w, s4.w, h4.w, relu);
            } else if (relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            a.x = fmaf(v.x, ph.x, a.x); a.y = fmaf(v.y, ph.y, a.y); a.z = fmaf(v.z, ph.z, a.z); a.w = fmaf(v.w, ph.w, a.w);
        }
        const float d = (float)(cnt > 0 ? cnt : 1);
        *reinterpret_cast<float4*>(agg + (size_t)t * f + c) = make_float4(a.x / d, a.y / d, a.z / d, a.w / d);
    }
}

}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_layer_grid(int f_in, int f_out) {
    (void)f_in; (void)f_out;
    int sms = sm_count();
    if (sms <= 0) return -1;
    return sms * 2;
}

template <int FE>
static int launch_layer_fwd(const LayerFwdArgs& p, size_t smem, int grid, cudaStream_t st) {
    if (int rc_ = ensure_dyn_smem((const void*)layer_fwd_kernel<FE>, 200 * 1024, "dgnn_layer_fwd")) return rc_;
    layer_fwd_kernel<FE><<<grid, NT, smem, st>>>(p);
    return check_launch("dgnn_layer_fwd");
}

extern "C" int dgnn_layer_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                              const int32_t* nbr, const float* ea, int fe,
                              const float* w_e, const float* b_e,
                              const float* wt_cat, const float* bias,
                              const float* out_scale, const float* out_shift, int relu_out,
                              int64_t n_tgt, int f_in, int f_out,
                              float* out, float* agg_save, double* stats, void* stream) {
    DGNN_REQUIRE(f_in > 0 && f_in % 4 == 0, "f_in must be a positive multiple of 4");
    DGNN_REQUIRE(f_out > 0 && f_out % 4 == 0, "f_out must be a positive multiple of 4");
    DGNN_REQUIRE(x_in && wt_cat && out, "null pointer");
    if (w_e == nullptr) fe = 0;
    DGNN_REQUIRE(fe % 4 == 0 && fe <= 32, "edge feature width must be a multiple of 4 and <= 32");
    DGNN_REQUIRE(nbr != nullptr || agg_save == nullptr, "agg_save needs a gather layer");
    DGNN_REQUIRE((relu_in & 2) == 0 || (fe == 0 && nbr != nullptr), "self loops (relu_in bit 1) need a gather layer without an edge filter");
    DGNN_REQUIRE(fe == 0 || (ea != nullptr && b_e != nullptr), "edge filter needs ea and b_e");
    LayerFwdArgs p;
    p.x_in = x_in; p.in_scale = in_scale; p.in_shift = in_shift; p.relu_in = relu_in;
    p.nbr = nbr; p.ea = ea; p.w_e = w_e; p.b_e = b_e; p.wt_cat = wt_cat; p.bias = bias;
    p.out_scale = out_scale; p.out_shift = out_shift; p.relu_out = relu_out;
    p.n_tgt = n_tgt; p.f_in = f_in; p.f_out = f_out;
    p.k_total = nbr ? 2 * f_in : f_in;
    p.kp = (p.k_total + TK - 1) / TK * TK;
    p.lda = p.kp + 4;
    p.out = out; p.agg_save = agg_save; p.stats = stats;
    size_t smem = ((size_t)TM * p.lda + 2 * TK * TN + 2 * TN) * sizeof(float);
    DGNN_REQUIRE(smem <= 200 * 1024, "layer too wide for the generic FP32 path (2*f_in <= 704)");
    int grid = dgnn_layer_grid(f_in, f_out);
    DGNN_REQUIRE(grid > 0, "no device");
    if (n_tgt <= 0 && stats == nullptr) return 0;
    cudaStream_t st = as_stream(stream);
    switch (fe) {
        case 0: return launch_layer_fwd<0>(p, smem, grid, st);
        case 4: return launch_layer_fwd<4>(p, smem, grid, st);
        case 8: return launch_layer_fwd<8>(p, smem, grid, st);
        case 12: return launch_layer_fwd<12>(p, smem, grid, st);
        case 16: return launch_layer_fwd<16>(p, smem, grid, st);
        case 20: return launch_layer_fwd<20>(p, smem, grid, st);
        case 24: return launch_layer_fwd<24>(p, smem, grid, st);
        case 28: return launch_layer_fwd<28>(p, smem, grid, st);
        case 32: return launch_layer_fwd<32>(p, smem, grid, st);
    }
    return fail("dgnn_layer_fwd", "unsupported edge feature width");
}

extern "C" int dgnn_gather_phi_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                                   const int32_t* nbr, const float* phi, int64_t n_tgt, int f, float* agg,
                                   void* stream) {
    DGNN_REQUIRE(f % 4 == 0 && f > 0, "f must be a positive multiple of 4");
    DGNN_REQUIRE(x_in && nbr && phi && agg, "null pointer");
    if (n_tgt <= 0) return 0;
    long long total = (long long)n_tgt * (f / 4);
    long long g = (total + 255) / 256, cap = (long long)sm_count() * 16;
    gather_phi_kernel<<<(int)(g < cap ? g : cap), 256, 0, as_stream(stream)>>>(x_in, in_scale, in_shift, relu_in, nbr, phi,
                                                                              n_tgt, f, agg);
    return check_launch("dgnn_gather_phi_fwd");
}
