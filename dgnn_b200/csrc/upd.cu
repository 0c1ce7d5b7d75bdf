// Backward of the Updated-edge-filter conv (learning/surfaceNetUpdatedEdgeFilters.py:147-176, :236-241):
// the edge state e' is a materialised [4*n_tgt, f] matrix (row 4t+k = edge into target t through slot k),
// so both gradients are plain HBM-bound gathers; one thread per (row, 4 features), no atomics.
#include "common.cuh"

namespace dgnn {

// dphi[t,k,:] = h(x[nbr[t,k]]) (*) d_agg[t]  +  (phi[t,k,:] > 0) * de_next[eid[t,k],:]
//   first term : d(agg)/d(e') with d_agg already divided by max(cnt,1)
//   second term: gradient that reaches e' through relu(edge state) read by the NEXT layer (de_next indexed by
//                global edge id, NULL for the last layer)
__global__ void __launch_bounds__(256) upd_edge_bwd_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                                           const float* __restrict__ sh, int relu,
                                                           const int32_t* __restrict__ nbr,
                                                           const float* __restrict__ d_agg,
                                                           const float* __restrict__ phi,
                                                           const float* __restrict__ de_next,
                                                           const int32_t* __restrict__ eid, long long n_tgt, int f,
                                                           float* __restrict__ dphi) {
    const int f4 = f >> 2;
    const long long total = n_tgt * 4 * f4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / f4;
        const int c = (int)(i % f4) * 4;
        const long long t = row >> 2;
        const int s = __ldg(nbr + row);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s >= 0) {
            float4 v = ldg4(x + (size_t)s * f + c);
            if (sc) {                                  // norm affine of the producer layer applied on load
                const float4 a = ldg4(sc + c), b = ldg4(sh + c);
                v.x = fmaf(v.x, a.x, b.x); v.y = fmaf(v.y, a.y, b.y); v.z = fmaf(v.z, a.z, b.z); v.w = fmaf(v.w, a.w, b.w);
            }
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            const float4 g = ldg4(d_agg + (size_t)t * f + c);
            o = make_float4(v.x * g.x, v.y * g.y, v.z * g.z, v.w * g.w);
            if (de_next) {
                const int e = __ldg(eid + row);
                if (e >= 0) {
                    const float4 p = ldg4(phi + (size_t)row * f + c);
                    const float4 d = ldg4(de_next + (size_t)e * f + c);
                    o.x += p.x > 0.f ? d.x : 0.f; o.y += p.y > 0.f ? d.y : 0.f;
                    o.z += p.z > 0.f ? d.z : 0.f; o.w += p.w > 0.f ? d.w : 0.f;
                }
            }
        }
        *reinterpret_cast<float4*>(dphi + (size_t)row * f + c) = o;
    }
}

// dx[s] = relu'(x[s]) * ( d_self[s] (s < n_tgt) + sum_j phi[orow[s,j],:] (*) d_agg[onbr[s,j]] )
// onbr[s,j] = target of the j-th out-edge of s, orow[s,j] = its row 4t+k in phi; -1 = none.
__global__ void __launch_bounds__(256) gather_phi_bwd_kernel(const float* __restrict__ d_agg,
                                                             const float* __restrict__ d_self,
                                                             const int32_t* __restrict__ onbr,
                                                             const int32_t* __restrict__ orow,
                                                             const float* __restrict__ phi,
                                                             const float* __restrict__ x, int relu, long long n_src,
                                                             long long n_tgt, int f, float* __restrict__ dx) {
    const int f4 = f >> 2;
    const long long total = n_src * f4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / f4;
        const int c = (int)(i % f4) * 4;
        const int4 nb = __ldg(reinterpret_cast<const int4*>(onbr) + s);
        const int4 rw = __ldg(reinterpret_cast<const int4*>(orow) + s);
        const int nbv[4] = {nb.x, nb.y, nb.z, nb.w};
        const int rwv[4] = {rw.x, rw.y, rw.z, rw.w};
        float4 a = (s < n_tgt && d_self) ? ldg4(d_self + (size_t)s * f + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (nbv[j] < 0) continue;
            const float4 g = ldg4(d_agg + (size_t)nbv[j] * f + c);
            const float4 p = ldg4(phi + (size_t)rwv[j] * f + c);
            a.x = fmaf(p.x, g.x, a.x); a.y = fmaf(p.y, g.y, a.y); a.z = fmaf(p.z, g.z, a.z); a.w = fmaf(p.w, g.w, a.w);
        }
        if (relu) {
            const float4 v = ldg4(x + (size_t)s * f + c);
            a.x = v.x > 0.f ? a.x : 0.f; a.y = v.y > 0.f ? a.y : 0.f;
            a.z = v.z > 0.f ? a.z : 0.f; a.w = v.w > 0.f ? a.w : 0.f;
        }
        *reinterpret_cast<float4*>(dx + (size_t)s * f + c) = a;
    }
}

// dst[r,:] = (z[r,:] > 0) ? src[r,:] : 0   (ReLU backward of a materialised pre-activation)
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ src, const float* __restrict__ z,
                                                        long long n4, float* __restrict__ dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        const float4 p = __ldg(reinterpret_cast<const float4*>(z) + i);
        reinterpret_cast<float4*>(dst)[i] =
            make_float4(p.x > 0.f ? v.x : 0.f, p.y > 0.f ? v.y : 0.f, p.z > 0.f ? v.z : 0.f, p.w > 0.f ? v.w : 0.f);
    }
}

static inline int grid_for(long long total) {
    long long g = (total + 255) / 256, cap = (long long)sm_count() * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_upd_edge_bwd(const float* x_in, const float* in_scale, const float* in_shift, int relu_in,
                                 const int32_t* nbr, const float* d_agg,
                                 const float* phi, const float* de_next, const int32_t* eid, int64_t n_tgt, int f,
                                 float* dphi, void* stream) {
    DGNN_REQUIRE(f > 0 && f % 4 == 0, "f must be a positive multiple of 4");
    DGNN_REQUIRE(x_in && nbr && d_agg && dphi, "null pointer");
    DGNN_REQUIRE(!de_next || (phi && eid), "de_next needs phi and eid");
    if (n_tgt <= 0) return 0;
    upd_edge_bwd_kernel<<<grid_for((long long)n_tgt * f), 256, 0, as_stream(stream)>>>(x_in, in_scale, in_shift, relu_in, nbr, d_agg, phi,
                                                                                      de_next, eid, n_tgt, f, dphi);
    return check_launch("dgnn_upd_edge_bwd");
}

extern "C" int dgnn_gather_phi_bwd(const float* d_agg, const float* d_self, const int32_t* onbr, const int32_t* orow,
                                   const float* phi, const float* x_in, int relu_in, int64_t n_src, int64_t n_tgt,
                                   int f, float* dx, void* stream) {
    DGNN_REQUIRE(f > 0 && f % 4 == 0, "f must be a positive multiple of 4");
    DGNN_REQUIRE(d_agg && onbr && orow && phi && dx, "null pointer");
    DGNN_REQUIRE(!relu_in || x_in, "relu_in needs x_in");
    if (n_src <= 0) return 0;
    gather_phi_bwd_kernel<<<grid_for((long long)n_src * (f / 4)), 256, 0, as_stream(stream)>>>(
        d_agg, d_self, onbr, orow, phi, x_in, relu_in, n_src, n_tgt, f, dx);
    return check_launch("dgnn_gather_phi_bwd");
}

extern "C" int dgnn_relu_mask(const float* src, const float* z, int64_t n_floats, float* dst, void* stream) {
    DGNN_REQUIRE(n_floats % 4 == 0, "length must be a multiple of 4");
    DGNN_REQUIRE(src && z && dst, "null pointer");
    if (n_floats <= 0) return 0;
    relu_mask_kernel<<<grid_for(n_floats / 4), 256, 0, as_stream(stream)>>>(src, z, n_floats / 4, dst);
    return check_launch("dgnn_relu_mask");
}
