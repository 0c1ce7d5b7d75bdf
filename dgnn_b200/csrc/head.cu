// Normalisation statistics, final linear, loss, optimiser, labels: the small kernels around the
// fused layers.  (learning/surfaceNetStaticEdgeFilters.py:116-123,180-187;
// learning/runModel.py:109-211,282,290; processing/generate_mesh.py:75,94-105.)
#include <math.h>

#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace dgnn {

thread_local char g_err[512] = {0};

static int g_sms[64] = {0};
static thread_local int g_sm_reserve = 0;   // SMs the persistent kernels leave free (dgnn_reserve_sms)
int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (g_sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        g_sms[dev] = n;
    }
    const int n = g_sms[dev] - g_sm_reserve;
    return n > 1 ? n : 1;
}

// (kernel, device) pairs whose dynamic shared-memory limit has been raised
static std::mutex g_cfg_mu;
static std::set<std::pair<const void*, int>> g_cfg;
int ensure_dyn_smem(const void* kernel, int bytes, const char* what) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(what, "cudaGetDevice failed");
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    if (g_cfg.count({kernel, dev})) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return fail(what, cudaGetErrorString(e));
    g_cfg.insert({kernel, dev});
    return 0;
}

// ---- norm finalize ------------------------------------------------------------------------
// one block; thread c per channel (loops if c > blockDim)
__global__ void norm_finalize_kernel(const double* __restrict__ stats, int n_partials, long long n_rows, int c,
                                     const float* __restrict__ weight, const float* __restrict__ bias, float eps,
                                     float momentum, int mode, float* running_mean, float* running_var,
                                     float* scale, float* shift, float* mean_o, float* rstd_o) {
    __shared__ double sh_s, sh_q;
    __shared__ double red_s[32], red_q[32];
    double tot_s = 0.0, tot_q = 0.0;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        // four independent accumulators per sum (loads in flight together), combined in a fixed order
        double s4[4] = {0.0, 0.0, 0.0, 0.0}, q4[4] = {0.0, 0.0, 0.0, 0.0};
        int p = 0;
        for (; p + 4 <= n_partials; p += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s4[u] += stats[(size_t)(p + u) * 2 * c + ch];
                q4[u] += stats[(size_t)(p + u) * 2 * c + c + ch];
            }
        }
        for (; p < n_partials; ++p) {
            s4[0] += stats[(size_t)p * 2 * c + ch];
            q4[0] += stats[(size_t)p * 2 * c + c + ch];
        }
        const double s = (s4[0] + s4[1]) + (s4[2] + s4[3]), q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        if (mode == 0) {
            double n = (double)n_rows;
            double m = s / n;
            double var = q / n - m * m;
            if (var < 0.0) var = 0.0;
            float rs = (float)(1.0 / sqrt(var + (double)eps));
            float w = weight ? weight[ch] : 1.f, b = bias ? bias[ch] : 0.f;
            float sc = w * rs;
            scale[ch] = sc;
            shift[ch] = b - (float)m * sc;
            mean_o[ch] = (float)m;
            rstd_o[ch] = rs;
            if (running_mean) {
                double unb = n > 1.0 ? var * n / (n - 1.0) : var;
                running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
                running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
            }
        } else {
            tot_s += s;
            tot_q += q;
        }
    }
    if (mode == 1) {
        tot_s = warp_sum(tot_s);
        tot_q = warp_sum(tot_q);
        int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) { red_s[w] = tot_s; red_q[w] = tot_q; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < (blockDim.x + 31) / 32; ++i) { a += red_s[i]; b += red_q[i]; }
            sh_s = a; sh_q = b;
        }
        __syncthreads();
        double n = (double)n_rows * (double)c;
        double m = sh_s / n;
        double var = sh_q / n - m * m;
        if (var < 0.0) var = 0.0;
        float rs = (float)(1.0 / (sqrt(var) + (double)eps));
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            float wv = weight ? weight[ch] : 1.f, bv = bias ? bias[ch] : 0.f;
            float sc = wv * rs;
            scale[ch] = sc;
            shift[ch] = bv - (float)m * sc;
            mean_o[ch] = (float)m;
            rstd_o[ch] = rs;
        }
    }
}

// BatchNorm (mode 0) finalize, parallel: a block handles 32 channels with 8 lanes each; lane q adds the partials
// q, q+8, ..., the eight lane sums are combined in lane order (fixed summation tree: reproducible)
__global__ void __launch_bounds__(256) norm_finalize_bn_kernel(const double* __restrict__ stats, int n_partials,
                                                               long long n_rows, int c, const float* __restrict__ weight,
                                                               const float* __restrict__ bias, float eps, float momentum,
                                                               float* running_mean, float* running_var, float* scale,
                                                               float* shift, float* mean_o, float* rstd_o) {
    __shared__ double red_s[8][32], red_q[8][32];
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + lane;
    double s = 0.0, sq = 0.0;
    if (ch < c)
        for (int p = q; p < n_partials; p += 8) {
            s += stats[(size_t)p * 2 * c + ch];
            sq += stats[(size_t)p * 2 * c + c + ch];
        }
    red_s[q][lane] = s; red_q[q][lane] = sq;
    __syncthreads();
    if (q == 0 && ch < c) {
        double ts = 0.0, tq = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ts += red_s[k][lane]; tq += red_q[k][lane]; }
        const double n = (double)n_rows;
        const double m = ts / n;
        double var = tq / n - m * m;
        if (var < 0.0) var = 0.0;
        const float rs = (float)(1.0 / sqrt(var + (double)eps));
        const float w = weight ? weight[ch] : 1.f, b = bias ? bias[ch] : 0.f;
        const float sc = w * rs;
        scale[ch] = sc;
        shift[ch] = b - (float)m * sc;
        mean_o[ch] = (float)m;
        rstd_o[ch] = rs;
        if (running_mean) {
            const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
            running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
            running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
        }
    }
}

__global__ void norm_eval_affine_kernel(const float* w, const float* b, const float* rm, const float* rv, float eps,
                                        int c, float* scale, float* shift) {
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    float rs = 1.f / sqrtf(rv[ch] + eps);
    float sc = (w ? w[ch] : 1.f) * rs;
    scale[ch] = sc;
    shift[ch] = (b ? b[ch] : 0.f) - rm[ch] * sc;
}

// ---- final linear: out[r,o] = sum_f h(r,f) w[o,f] + b[o] --------------------------------------
// LPR lanes per row, float4 per lane per step
template <int OD>
__global__ void __launch_bounds__(256) rowdot_fwd_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                                          const float* __restrict__ sh, int relu,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          long long n, int f, int lpr, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int sub = lane / lpr, li = lane % lpr;
    const int rows_per_warp = 32 / lpr;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r0 = warp_global * rows_per_warp; r0 < n; r0 += n_warps * rows_per_warp) {
        long long r = r0 + sub;
        float acc[OD];
#pragma unroll
        for (int o = 0; o < OD; ++o) acc[o] = 0.f;
        if (r < n) {
            for (int c = li * 4; c < f; c += lpr * 4) {
                float4 v = ldg4(x + (size_t)r * f + c);
                if (sc) {
                    float4 s4 = ldg4(sc + c), h4 = ldg4(sh + c);
                    v.x = act(v.x, s4.x, h4.x, relu); v.y = act(v.y, s4.y, h4.y, relu);
                    v.z = act(v.z, s4.z, h4.z, relu); v.w = act(v.w, s4.w, h4.w, relu);
                } else if (relu) {
                    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                }
#pragma unroll
                for (int o = 0; o < OD; ++o) {
                    float4 wv = ldg4(w + (size_t)o * f + c);
                    acc[o] = fmaf(v.x, wv.x, acc[o]); acc[o] = fmaf(v.y, wv.y, acc[o]);
                    acc[o] = fmaf(v.z, wv.z, acc[o]); acc[o] = fmaf(v.w, wv.w, acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < OD; ++o) {
            for (int d = lpr >> 1; d > 0; d >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], d);
        }
        if (r < n && li == 0) {
#pragma unroll
            for (int o = 0; o < OD; ++o) out[(size_t)r * OD + o] = acc[o] + (b ? b[o] : 0.f);
        }
    }
}

__global__ void affine_relu_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                   const float* __restrict__ sh, int relu, long long n4, int f, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)((i * 4) % f);
        float4 v = ldg4(x + i * 4);
        if (sc) {
            float4 s4 = ldg4(sc + c), h4 = ldg4(sh + c);
            v.x = act(v.x, s4.x, h4.x, relu); v.y = act(v.y, s4.y, h4.y, relu);
            v.z = act(v.z, s4.z, h4.z, relu); v.w = act(v.w, s4.w, h4.w, relu);
        } else if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        *reinterpret_cast<float4*>(out + i * 4) = v;
    }
}

// ---- loss -------------------------------------------------------------------------------------
__device__ __forceinline__ float weight_of(float w, int mode) {
    return mode == 0 ? w : (mode == 1 ? sqrtf(w) : (mode == 2 ? logf(1.f + w) : 1.f));
}

__device__ __forceinline__ void block_sum2(double a, double b, double* out2) {
    __shared__ double ra[32], rb[32];
    a = warp_sum(a); b = warp_sum(b);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { ra[w] = a; rb[w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) { x += ra[i]; y += rb[i]; }
        out2[0] = x; out2[1] = y;
    }
}

__global__ void __launch_bounds__(256) kl_loss_fwd_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                          int ys, const float* __restrict__ w, int ws, int mode,
                                                          long long n, double* __restrict__ partials) {
    double sl = 0.0, sw = 0.0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        float z0 = z[r * 2], z1 = z[r * 2 + 1];
        float m = fmaxf(z0, z1);
        float lse = m + logf(expf(z0 - m) + expf(z1 - m));
        float lp0 = z0 - lse, lp1 = z1 - lse;
        float y0 = y[r * ys], y1 = y[r * ys + 1];
        // F.kl_div(reduction='none'): y*(log y - logp), 0 where y == 0
        float l = 0.f;
        if (y0 > 0.f) l += y0 * (logf(y0) - lp0);
        if (y1 > 0.f) l += y1 * (logf(y1) - lp1);
        float wv = weight_of(w ? w[r * ws] : 1.f, w ? mode : 3);
        sl += (double)(l * wv);
        sw += (double)wv;
    }
    block_sum2(sl, sw, partials + 2 * blockIdx.x);
}

__global__ void kl_loss_finalize_kernel(const double* __restrict__ partials, int np, float* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < np; ++i) { a += partials[2 * i]; b += partials[2 * i + 1]; }
        out[0] = (float)(a / b);
        out[1] = (float)a;
        out[2] = (float)b;
    }
}

__global__ void __launch_bounds__(256) kl_loss_bwd_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                          int ys, const float* __restrict__ w, int ws, int mode,
                                                          long long n, const float* __restrict__ sums,
                                                          const float* __restrict__ gout, float* __restrict__ dz) {
    const float g = (gout ? gout[0] : 1.f) / sums[2];
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        float z0 = z[r * 2], z1 = z[r * 2 + 1];
        float m = fmaxf(z0, z1);
        float e0 = expf(z0 - m), e1 = expf(z1 - m);
        float inv = 1.f / (e0 + e1);
        float y0 = y[r * ys], y1 = y[r * ys + 1];
        float sy = (y0 > 0.f ? y0 : 0.f) + (y1 > 0.f ? y1 : 0.f);
        float wv = weight_of(w ? w[r * ws] : 1.f, w ? mode : 3) * g;
        dz[r * 2] = wv * (e0 * inv * sy - (y0 > 0.f ? y0 : 0.f));
        dz[r * 2 + 1] = wv * (e1 * inv * sy - (y1 > 0.f ? y1 : 0.f));
    }
}

// ---- bce / mse on one logit per cell (runModel.py:181-188) ----------------------------------------
//   kind 0 (bce): l_r = BCEWithLogits(z_r, y_r) = max(z,0) - z*y + log(1 + exp(-|z|)), weighted like kl
//   kind 1 (mse): F.mse_loss(sigmoid(z), y) is already the mean over cells; the reference multiplies that scalar by
//                 the weights and divides by their sum again, so the weights cancel: partials = (sum sq, count)
__global__ void __launch_bounds__(256) point_loss_fwd_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                             int ys, const float* __restrict__ w, int ws, int mode,
                                                             int kind, long long n, double* __restrict__ partials) {
    double sl = 0.0, sw = 0.0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const float zv = z[r], yv = y[r * ys];
        if (kind == 0) {
            const float l = fmaxf(zv, 0.f) - zv * yv + log1pf(expf(-fabsf(zv)));
            const float wv = weight_of(w ? w[r * ws] : 1.f, w ? mode : 3);
            sl += (double)(l * wv);
            sw += (double)wv;
        } else {
            const float d = 1.f / (1.f + expf(-zv)) - yv;
            sl += (double)(d * d);
            sw += 1.0;
        }
    }
    block_sum2(sl, sw, partials + 2 * blockIdx.x);
}

__global__ void __launch_bounds__(256) point_loss_bwd_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                             int ys, const float* __restrict__ w, int ws, int mode,
                                                             int kind, long long n, const float* __restrict__ sums,
                                                             const float* __restrict__ gout, float* __restrict__ dz) {
    const float g = (gout ? gout[0] : 1.f) / sums[2];
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const float zv = z[r], yv = y[r * ys];
        const float sg = 1.f / (1.f + expf(-zv));
        if (kind == 0) dz[r] = g * weight_of(w ? w[r * ws] : 1.f, w ? mode : 3) * (sg - yv);
        else dz[r] = g * 2.f * (sg - yv) * sg * (1.f - sg);
    }
}

__global__ void __launch_bounds__(256) edge_reg_kernel(const float* __restrict__ z, const long long* __restrict__ src,
                                                       const long long* __restrict__ tgt, long long ne,
                                                       double* __restrict__ partials) {
    double s = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        long long a = src[e], b = tgt[e];
        float pa = 1.f / (1.f + expf(z[a * 2 + 1] - z[a * 2]));
        float pb = 1.f / (1.f + expf(z[b * 2 + 1] - z[b * 2]));
        s += (double)fabsf(pa - pb);
    }
    block_sum2(s, 0.0, partials + 2 * blockIdx.x);
}

// backward of the regulariser: d|pa - pb| = sign(pa - pb) (d pa - d pb), p = sigmoid(z0 - z1).  The per-cell sum of the
// signs over its edges is an INTEGER, so it is accumulated with integer atomics (exact, order-independent) ...
__global__ void __launch_bounds__(256) edge_reg_sign_kernel(const float* __restrict__ z, const long long* __restrict__ src,
                                                            const long long* __restrict__ tgt, long long ne,
                                                            int* __restrict__ cnt) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        long long a = src[e], b = tgt[e];
        float pa = 1.f / (1.f + expf(z[a * 2 + 1] - z[a * 2]));
        float pb = 1.f / (1.f + expf(z[b * 2 + 1] - z[b * 2]));
        int sg = (pa > pb) - (pa < pb);
        if (sg != 0) { atomicAdd(&cnt[a], sg); atomicAdd(&cnt[b], -sg); }
    }
}
// ... and turned into the logit gradient per cell: dz0 = gout * scale * p (1 - p) * cnt, dz1 = -dz0
__global__ void __launch_bounds__(256) edge_reg_bwd_kernel(const float* __restrict__ z, const int* __restrict__ cnt,
                                                           long long n, float scale, const float* __restrict__ gout,
                                                           float* __restrict__ dz) {
    const float g0 = gout[0] * scale;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        float p = 1.f / (1.f + expf(z[c * 2 + 1] - z[c * 2]));
        float g = g0 * p * (1.f - p) * (float)cnt[c];
        reinterpret_cast<float2*>(dz)[c] = make_float2(g, -g);
    }
}

// ---- generic partial reductions ---------------------------------------------------------------
// out[j] = sum_i p[i, j] in double.  256 threads = 32 columns x 8 lanes; lane q adds the partials i = q, q+8, ..., the
// eight lane sums are added in lane order: a fixed summation tree, so the result is reproducible run to run.
template <typename T>
__global__ void __launch_bounds__(256) reduce_partials_kernel(const T* __restrict__ p, int np, long long len,
                                                              float* __restrict__ out) {
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const long long j = (long long)blockIdx.x * 32 + lane;
    double s = 0.0;
    if (j < len)
        for (int i = q; i < np; i += 8) s += (double)p[(size_t)i * len + j];
    red[q][lane] = s;
    __syncthreads();
    if (q == 0 && j < len) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][lane];
        out[j] = (float)t;
    }
}

__global__ void norm_bwd_coeffs_kernel(const float* s1, const float* s2, long long n_rows, int c, const float* weight,
                                       const float* rstd, int mode, float* g, float* a, float* b) {
    __shared__ float m1s, m2s;
    if (mode == 1) {
        // graph LayerNorm: scalar means over all rows x channels of (w*dy) and (w*dy*xhat)
        if (threadIdx.x == 0) {
            double t1 = 0.0, t2 = 0.0;
            for (int ch = 0; ch < c; ++ch) {
                float wv = weight ? weight[ch] : 1.f;
                t1 += (double)wv * s1[ch];
                t2 += (double)wv * s2[ch];
            }
            double n = (double)n_rows * c;
            m1s = (float)(t1 / n);
            m2s = (float)(t2 / n);
        }
        __syncthreads();
    }
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float wv = weight ? weight[ch] : 1.f;
        float gv = wv * rstd[ch];
        g[ch] = gv;
        if (mode == 0) {
            a[ch] = gv * s1[ch] / (float)n_rows;
            b[ch] = gv * s2[ch] / (float)n_rows;
        } else {
            a[ch] = rstd[ch] * m1s;
            b[ch] = rstd[ch] * m2s;
        }
    }
}

// ---- Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) ---------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i];
        float mi = m[i] + (gi - m[i]) * (1.f - b1);       // exp_avg.lerp_(grad, 1-beta1)
        float vi = v[i] * b2 + (1.f - b2) * gi * gi;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
    }
}

__global__ void adam_multi_kernel(const long long* __restrict__ table, float lr, float b1, float b2, float eps,
                                  float bc1, float bc2_sqrt) {
    const long long* row = table + (size_t)blockIdx.y * 5;
    float* p = reinterpret_cast<float*>(row[0]);
    const float* g = reinterpret_cast<const float*>(row[1]);
    float* m = reinterpret_cast<float*>(row[2]);
    float* v = reinterpret_cast<float*>(row[3]);
    const long long n = row[4];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i];
        float mi = m[i] + (gi - m[i]) * (1.f - b1);
        float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
    }
}

// The same update with the step count and the hyper-parameters read from device memory, so that a captured CUDA graph of
// the training step stays valid across replays: state[0] = step count (>= 1, advanced by the caller on the same stream
// before the launch), hyper = (lr, beta1, beta2, eps).
__global__ void adam_multi_dev_kernel(const long long* __restrict__ table, const float* __restrict__ hyper,
                                      const long long* __restrict__ state) {
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
    const double step = (double)state[0];
    const float bc1 = (float)(1.0 - pow((double)b1, step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
    const long long* row = table + (size_t)blockIdx.y * 5;
    float* p = reinterpret_cast<float*>(row[0]);
    const float* g = reinterpret_cast<const float*>(row[1]);
    float* m = reinterpret_cast<float*>(row[2]);
    float* v = reinterpret_cast<float*>(row[3]);
    const long long n = row[4];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i];
        float mi = m[i] + (gi - m[i]) * (1.f - b1);
        float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
    }
}

// dy = relu'(y) * dh, S1 = sum dy, S2 = sum dy*xhat; thread owns a 4-column group
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ z,
                                                      const float* __restrict__ sc, const float* __restrict__ sh,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                      int relu, long long n, int f, float* __restrict__ dy,
                                                      double* __restrict__ partials) {
    extern __shared__ double sred[];  // 2f
    for (int i = threadIdx.x; i < 2 * f; i += blockDim.x) sred[i] = 0.0;
    __syncthreads();
    const int f4 = f >> 2;
    const int rows_per_pass = blockDim.x / f4;
    const int c = (threadIdx.x % f4) * 4;
    const int rsub = threadIdx.x / f4;
    if (rsub < rows_per_pass) {
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (sc) { s4 = ldg4(sc + c); h4 = ldg4(sh + c); }
        if (mean) { m4 = ldg4(mean + c); r4 = ldg4(rstd + c); }
        float S1[4] = {0.f, 0.f, 0.f, 0.f}, S2[4] = {0.f, 0.f, 0.f, 0.f};
        for (long long r = (long long)blockIdx.x * rows_per_pass + rsub; r < n; r += (long long)gridDim.x * rows_per_pass) {
            float4 zv = ldg4(z + (size_t)r * f + c);
            float4 d = ldg4(dh + (size_t)r * f + c);
            float g[4] = {d.x, d.y, d.z, d.w};
            const float yv[4] = {fmaf(zv.x, s4.x, h4.x), fmaf(zv.y, s4.y, h4.y), fmaf(zv.z, s4.z, h4.z), fmaf(zv.w, s4.w, h4.w)};
            const float xh[4] = {(zv.x - m4.x) * r4.x, (zv.y - m4.y) * r4.y, (zv.z - m4.z) * r4.z, (zv.w - m4.w) * r4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (relu && !(yv[j] > 0.f)) g[j] = 0.f;
                S1[j] += g[j];
                S2[j] = fmaf(g[j], xh[j], S2[j]);
            }
            *reinterpret_cast<float4*>(dy + (size_t)r * f + c) = make_float4(g[0], g[1], g[2], g[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sred[c + j], (double)S1[j]);
            atomicAdd(&sred[f + c + j], (double)S2[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * f; i += blockDim.x) partials[(size_t)blockIdx.x * 2 * f + i] = sred[i];
}

// ---- labels / facets -------------------------------------------------------------------------------
__global__ void argmax_labels_kernel(const float* __restrict__ z, long long n, int od, uint8_t* __restrict__ lab) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        if (od == 2) lab[r] = z[r * 2 + 1] > z[r * 2] ? 1 : 0;  // argmax, ties -> 0 (generate_mesh.py:75)
        else lab[r] = z[r] > 0.f ? 1 : 0;
    }
}
// exportScore (processing/data.py:521-535): sigmoid and softmax of the logits, one pass
__global__ void scores_kernel(const float* __restrict__ z, long long n, int od, float* __restrict__ sig,
                              float* __restrict__ soft) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        if (od == 2) {
            const float a = z[r * 2], b = z[r * 2 + 1];
            sig[r * 2] = 1.f / (1.f + expf(-a));
            sig[r * 2 + 1] = 1.f / (1.f + expf(-b));
            const float m = fmaxf(a, b), ea = expf(a - m), eb = expf(b - m), inv = 1.f / (ea + eb);
            soft[r * 2] = ea * inv;
            soft[r * 2 + 1] = eb * inv;
        } else {
            sig[r] = 1.f / (1.f + expf(-z[r]));
            soft[r] = 1.f;                             // softmax over a single column
        }
    }
}
__global__ void interface_facets_kernel(const uint8_t* __restrict__ lab, long long nf, const int* __restrict__ nfac,
                                        long long n_facets, uint8_t* __restrict__ flag) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_facets; i += (long long)gridDim.x * blockDim.x) {
        int a = nfac[i * 2], b = nfac[i * 2 + 1];
        uint8_t la = (a < 0 || a >= nf) ? 1 : lab[a];  // infinite cell forced outside (:98-99)
        uint8_t lb = (b < 0 || b >= nf) ? 1 : lab[b];
        flag[i] = la != lb;
    }
}

}  // namespace dgnn

using namespace dgnn;

static inline int grid_for(long long n, int block, int cap) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

extern "C" int dgnn_version(void) { return DGNN_B200_VERSION; }
extern "C" const char* dgnn_last_error(void) { return g_err; }

extern "C" int dgnn_device_check(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail("dgnn_device_check", cudaGetErrorString(e));
    if (prop.major != 10) {
        char msg[128];
        snprintf(msg, sizeof(msg), "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                 prop.minor);
        return fail("dgnn_device_check", msg);
    }
    return 0;
}

extern "C" int dgnn_sm_count(void) { return sm_count(); }
// The persistent layer kernels fill every SM (one CTA each, all registers).  While a halo exchange runs next to a layer
// the calling thread reserves a few SMs so that the pack kernel and the NCCL kernels of the communication stream get
// scheduled immediately instead of behind the layer kernel.  Returns the previous value; thread-local.
extern "C" int dgnn_reserve_sms(int n) {
    const int prev = g_sm_reserve;
    g_sm_reserve = n < 0 ? 0 : n;
    return prev;
}
extern "C" int dgnn_small_grid(void) { return sm_count() * 4; }

extern "C" int dgnn_norm_finalize(const double* stats, int n_partials, int64_t n_rows, int c, const float* weight,
                                  const float* bias, float eps, float momentum, int mode, float* running_mean,
                                  float* running_var, float* scale, float* shift, float* mean, float* rstd,
                                  void* stream) {
    DGNN_REQUIRE(stats && scale && shift && mean && rstd, "null pointer");
    DGNN_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (batch) or 1 (graph layer norm)");
    if (mode == 0)
        norm_finalize_bn_kernel<<<(c + 31) / 32, 256, 0, as_stream(stream)>>>(stats, n_partials, n_rows, c, weight, bias, eps,
                                                                              momentum, running_mean, running_var, scale,
                                                                              shift, mean, rstd);
    else
        norm_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(stats, n_partials, n_rows, c, weight, bias, eps, momentum,
                                                               mode, running_mean, running_var, scale, shift, mean, rstd);
    return check_launch("dgnn_norm_finalize");
}

extern "C" int dgnn_norm_eval_affine(const float* weight, const float* bias, const float* running_mean,
                                     const float* running_var, float eps, int c, float* scale, float* shift,
                                     void* stream) {
    norm_eval_affine_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(weight, bias, running_mean, running_var,
                                                                            eps, c, scale, shift);
    return check_launch("dgnn_norm_eval_affine");
}

extern "C" int dgnn_rowdot_fwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                               const float* w, const float* b, int64_t n, int f, int od, float* out, void* stream) {
    DGNN_REQUIRE(f % 4 == 0 && od >= 1 && od <= 4, "f multiple of 4, 1 <= od <= 4");
    if (n <= 0) return 0;
    int lpr = 1;
    while (lpr * 4 < f && lpr < 32) lpr <<= 1;
    int grid = grid_for(n * lpr, 256, sm_count() * 8);
    cudaStream_t st = as_stream(stream);
    switch (od) {
        case 1: rowdot_fwd_kernel<1><<<grid, 256, 0, st>>>(x_in, in_scale, in_shift, relu_in, w, b, n, f, lpr, out); break;
        case 2: rowdot_fwd_kernel<2><<<grid, 256, 0, st>>>(x_in, in_scale, in_shift, relu_in, w, b, n, f, lpr, out); break;
        case 3: rowdot_fwd_kernel<3><<<grid, 256, 0, st>>>(x_in, in_scale, in_shift, relu_in, w, b, n, f, lpr, out); break;
        default: rowdot_fwd_kernel<4><<<grid, 256, 0, st>>>(x_in, in_scale, in_shift, relu_in, w, b, n, f, lpr, out); break;
    }
    return check_launch("dgnn_rowdot_fwd");
}

extern "C" int dgnn_affine_relu(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                                int64_t n, int f, float* out, void* stream) {
    DGNN_REQUIRE(f % 4 == 0, "f multiple of 4");
    long long n4 = n * f / 4;
    if (n4 <= 0) return 0;
    affine_relu_kernel<<<grid_for(n4, 256, sm_count() * 8), 256, 0, as_stream(stream)>>>(x_in, in_scale, in_shift,
                                                                                         relu_in, n4, f, out);
    return check_launch("dgnn_affine_relu");
}

extern "C" int dgnn_kl_loss_fwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                                int weight_mode, int64_t n, double* partials, void* stream) {
    kl_loss_fwd_kernel<<<dgnn_small_grid(), 256, 0, as_stream(stream)>>>(logits, y, y_stride, w, w_stride, weight_mode,
                                                                         n, partials);
    return check_launch("dgnn_kl_loss_fwd");
}
extern "C" int dgnn_kl_loss_finalize(const double* partials, int n_partials, float* out, void* stream) {
    kl_loss_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(partials, n_partials, out);
    return check_launch("dgnn_kl_loss_finalize");
}
extern "C" int dgnn_kl_loss_bwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                                int weight_mode, int64_t n, const float* sums, const float* grad_out, float* dlogits,
                                void* stream) {
    if (n <= 0) return 0;
    kl_loss_bwd_kernel<<<grid_for(n, 256, sm_count() * 8), 256, 0, as_stream(stream)>>>(
        logits, y, y_stride, w, w_stride, weight_mode, n, sums, grad_out, dlogits);
    return check_launch("dgnn_kl_loss_bwd");
}
extern "C" int dgnn_point_loss_fwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                                   int weight_mode, int kind, int64_t n, double* partials, void* stream) {
    DGNN_REQUIRE(logits && y && partials, "null pointer");
    DGNN_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (bce) or 1 (mse)");
    point_loss_fwd_kernel<<<dgnn_small_grid(), 256, 0, as_stream(stream)>>>(logits, y, y_stride, w, w_stride, weight_mode,
                                                                           kind, n, partials);
    return check_launch("dgnn_point_loss_fwd");
}
extern "C" int dgnn_point_loss_bwd(const float* logits, const float* y, int y_stride, const float* w, int w_stride,
                                   int weight_mode, int kind, int64_t n, const float* sums, const float* grad_out,
                                   float* dlogits, void* stream) {
    if (n <= 0) return 0;
    point_loss_bwd_kernel<<<grid_for(n, 256, sm_count() * 8), 256, 0, as_stream(stream)>>>(
        logits, y, y_stride, w, w_stride, weight_mode, kind, n, sums, grad_out, dlogits);
    return check_launch("dgnn_point_loss_bwd");
}
extern "C" int dgnn_edge_reg_fwd(const float* logits, const int64_t* src, const int64_t* tgt, int64_t n_edges,
                                 double* partials, void* stream) {
    edge_reg_kernel<<<dgnn_small_grid(), 256, 0, as_stream(stream)>>>(logits, (const long long*)src,
                                                                      (const long long*)tgt, n_edges, partials);
    return check_launch("dgnn_edge_reg_fwd");
}

extern "C" int dgnn_edge_reg_bwd(const float* logits, const int64_t* src, const int64_t* tgt, int64_t n_edges,
                                 int64_t n_rows, float scale, const float* grad_out, int32_t* sign_count, float* dlogits,
                                 void* stream) {
    DGNN_REQUIRE(logits && grad_out && sign_count && dlogits, "null pointer");
    cudaStream_t st = as_stream(stream);
    if (n_edges > 0) {
        DGNN_REQUIRE(src && tgt, "null pointer");
        edge_reg_sign_kernel<<<grid_for(n_edges, 256, sm_count() * 8), 256, 0, st>>>(logits, (const long long*)src,
                                                                                    (const long long*)tgt, n_edges, sign_count);
        if (check_launch("dgnn_edge_reg_bwd")) return 1;
    }
    if (n_rows > 0)
        edge_reg_bwd_kernel<<<grid_for(n_rows, 256, sm_count() * 8), 256, 0, st>>>(logits, sign_count, n_rows, scale, grad_out,
                                                                                  dlogits);
    return check_launch("dgnn_edge_reg_bwd");
}

extern "C" int dgnn_reduce_partials(const double* partials, int n_partials, int len, float* out, void* stream) {
    if (len <= 0) return 0;
    reduce_partials_kernel<double><<<(len + 31) / 32, 256, 0, as_stream(stream)>>>(partials, n_partials, len, out);
    return check_launch("dgnn_reduce_partials");
}
extern "C" int dgnn_reduce_partials_f32(const float* partials, int n_partials, int64_t len, float* out, void* stream) {
    if (len <= 0) return 0;
    reduce_partials_kernel<float><<<(int)((len + 31) / 32), 256, 0, as_stream(stream)>>>(partials, n_partials, len, out);
    return check_launch("dgnn_reduce_partials_f32");
}
extern "C" int dgnn_norm_bwd_coeffs(const float* s1, const float* s2, int64_t n_rows, int c, const float* weight,
                                    const float* rstd, int mode, float* g, float* a, float* b, void* stream) {
    norm_bwd_coeffs_kernel<<<1, 256, 0, as_stream(stream)>>>(s1, s2, n_rows, c, weight, rstd, mode, g, a, b);
    return check_launch("dgnn_norm_bwd_coeffs");
}

extern "C" int dgnn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, int step, void* stream) {
    if (n <= 0) return 0;
    DGNN_REQUIRE(step >= 1, "step counts from 1");
    float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    float bc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    adam_kernel<<<grid_for(n, 256, sm_count() * 4), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                                                 lr, beta1, beta2, eps, bc1, bc2);
    return check_launch("dgnn_adam_step");
}

extern "C" int dgnn_adam_multi(const int64_t* table, int n_tensors, int64_t max_n, float lr, float beta1, float beta2,
                               float eps, int step, void* stream) {
    if (n_tensors <= 0 || max_n <= 0) return 0;
    DGNN_REQUIRE(step >= 1, "step counts from 1");
    float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    float bc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    dim3 grid(grid_for(max_n, 256, 64), n_tensors);
    adam_multi_kernel<<<grid, 256, 0, as_stream(stream)>>>((const long long*)table, lr, beta1, beta2, eps, bc1, bc2);
    return check_launch("dgnn_adam_multi");
}

extern "C" int dgnn_adam_multi_dev(const int64_t* table, int n_tensors, int64_t max_n, const float* hyper,
                                   const int64_t* state, void* stream) {
    if (n_tensors <= 0 || max_n <= 0) return 0;
    DGNN_REQUIRE(table && hyper && state, "null pointer");
    dim3 grid(grid_for(max_n, 256, 64), n_tensors);
    adam_multi_dev_kernel<<<grid, 256, 0, as_stream(stream)>>>((const long long*)table, hyper, (const long long*)state);
    return check_launch("dgnn_adam_multi_dev");
}

extern "C" int dgnn_act_bwd(const float* dh, const float* z, const float* in_scale, const float* in_shift,
                            const float* mean, const float* rstd, int relu_in, int64_t n, int f, float* dy,
                            double* partials, void* stream) {
    DGNN_REQUIRE(f % 4 == 0 && f <= 1024, "f must be a multiple of 4 and <= 1024");
    act_bwd_kernel<<<dgnn_small_grid(), 256, 2 * f * sizeof(double), as_stream(stream)>>>(
        dh, z, in_scale, in_shift, mean, rstd, relu_in, n, f, dy, partials);
    return check_launch("dgnn_act_bwd");
}

extern "C" int dgnn_argmax_labels(const float* logits, int64_t n, int od, uint8_t* labels, void* stream) {
    DGNN_REQUIRE(od == 1 || od == 2, "od must be 1 or 2");
    if (n <= 0) return 0;
    argmax_labels_kernel<<<grid_for(n, 256, sm_count() * 8), 256, 0, as_stream(stream)>>>(logits, n, od, labels);
    return check_launch("dgnn_argmax_labels");
}
extern "C" int dgnn_scores(const float* logits, int64_t n, int od, float* sigmoid, float* softmax, void* stream) {
    DGNN_REQUIRE(od == 1 || od == 2, "od must be 1 or 2");
    DGNN_REQUIRE(logits && sigmoid && softmax, "null pointer");
    if (n <= 0) return 0;
    scores_kernel<<<grid_for(n, 256, sm_count() * 8), 256, 0, as_stream(stream)>>>(logits, n, od, sigmoid, softmax);
    return check_launch("dgnn_scores");
}
extern "C" int dgnn_interface_facets(const uint8_t* labels_finite, int64_t n_finite, const int32_t* nfacets,
                                     int64_t n_facets, uint8_t* flag, void* stream) {
    if (n_facets <= 0) return 0;
    interface_facets_kernel<<<grid_for(n_facets, 256, sm_count() * 8), 256, 0, as_stream(stream)>>>(
        labels_finite, n_finite, nfacets, n_facets, flag);
    return check_launch("dgnn_interface_facets");
}
