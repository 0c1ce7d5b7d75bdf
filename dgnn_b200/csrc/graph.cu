// Graph layout kernels: the facet adjacency of processing/data.py:434-439 re-laid into a
// fixed-width ELL-4 table with the reverse-facet slot, Morton codes, permutation application and
// the edge-attribute re-layout (incoming / own-slot order).  Integer work, bit-exact against
// oracle/graph.py.
#include "common.cuh"

namespace dgnn {

__global__ void ell_from_adj_kernel(const int32_t* __restrict__ adj, long long n, int32_t* __restrict__ nbr,
                                    uint8_t* __restrict__ rslot, int32_t* err) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int4 nb;
        int32_t own_ok = 1;
        int32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int2 row = __ldg(reinterpret_cast<const int2*>(adj) + i * 4 + k);
            own_ok &= (row.x == (int32_t)i);
            v[k] = row.y;
        }
        if (!own_ok) atomicMax(err, 1);
        nb = make_int4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<int4*>(nbr)[i] = nb;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int s = v[k];
            int found = -1;
            if (s >= 0 && s < n) {
#pragma unroll
                for (int kk = 3; kk >= 0; --kk) {  // descending so the FIRST match wins
                    int2 row = __ldg(reinterpret_cast<const int2*>(adj) + (long long)s * 4 + kk);
                    if (row.y == (int32_t)i) found = kk;
                }
            }
            if (found < 0) { atomicMax(err, 2); found = 0; }
            packed |= (uint32_t)found << (8 * k);
        }
        reinterpret_cast<uint32_t*>(rslot)[i] = packed;
    }
}

__global__ void ell_fill_kernel(const long long* __restrict__ src, const long long* __restrict__ tgt, long long ne,
                                long long n_rows, long long n_other, int32_t* __restrict__ nbr,
                                int32_t* __restrict__ eid, int32_t* __restrict__ cnt, int32_t* err) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        long long t = tgt[e];
        if (t < 0 || t >= n_rows) { atomicMax(err, 4); continue; }
        const long long sv = src[e];
        if (sv < 0 || sv >= n_other) { atomicMax(err, 4); continue; }   // a malformed edge_index must not become a gather index
        int slot = atomicAdd(&cnt[t], 1);
        if (slot >= 4) { atomicMax(err, 3); continue; }
        nbr[t * 4 + slot] = (int32_t)src[e];
        eid[t * 4 + slot] = (int32_t)e;
    }
}

// sort each row's <=4 entries by edge id (by_source: by (source, edge id), the order of PyG's
// SparseTensor(row=src, col=tgt).t() rows) so the layout is deterministic; pad with -1
__global__ void ell_sort_kernel(long long n_rows, int32_t* __restrict__ nbr, int32_t* __restrict__ eid,
                                int32_t* __restrict__ cnt, int by_source) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += (long long)gridDim.x * blockDim.x) {
        int c = cnt[t];
        if (c > 4) { c = 4; cnt[t] = 4; }
        int4 nb = reinterpret_cast<int4*>(nbr)[t], ei = reinterpret_cast<int4*>(eid)[t];
        int nv[4] = {nb.x, nb.y, nb.z, nb.w}, ev[4] = {ei.x, ei.y, ei.z, ei.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k >= c) { nv[k] = -1; ev[k] = 0x7fffffff; }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3 - a; ++b)
                if (by_source ? (nv[b + 1] >= 0 && (nv[b] < 0 || nv[b] > nv[b + 1] || (nv[b] == nv[b + 1] && ev[b] > ev[b + 1])))
                              : (ev[b] > ev[b + 1])) {
                    int t1 = ev[b]; ev[b] = ev[b + 1]; ev[b + 1] = t1;
                    int t2 = nv[b]; nv[b] = nv[b + 1]; nv[b + 1] = t2;
                }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k >= c) ev[k] = -1;
        reinterpret_cast<int4*>(nbr)[t] = make_int4(nv[0], nv[1], nv[2], nv[3]);
        reinterpret_cast<int4*>(eid)[t] = make_int4(ev[0], ev[1], ev[2], ev[3]);
    }
}

__global__ void morton_kernel(const float* __restrict__ pos, long long n, float lx, float ly, float lz, float ex,
                              float ey, float ez, int bits, unsigned long long* __restrict__ codes) {
    const float scale = (float)((1u << bits) - 1u);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float p[3] = {pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]};
        const float lo[3] = {lx, ly, lz}, ext[3] = {ex, ey, ez};
        unsigned q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            // (p - lo) / ext * scale with IEEE round-to-nearest at every step (matches NumPy float32)
            float v = __fmul_rn(__fdiv_rn(__fsub_rn(p[a], lo[a]), ext[a]), scale);
            v = fminf(fmaxf(v, 0.f), scale);
            q[a] = (unsigned)v;
        }
        unsigned long long code = 0;
        for (int b = 0; b < bits; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a)
                code |= (unsigned long long)((q[a] >> b) & 1u) << (3 * b + (2 - a));
        codes[i] = code;
    }
}

__global__ void perm_apply_ell_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm,
                                      const int32_t* __restrict__ inv, long long n, int32_t* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int4 nb = __ldg(reinterpret_cast<const int4*>(nbr) + perm[i]);
        nb.x = nb.x >= 0 ? inv[nb.x] : nb.x;
        nb.y = nb.y >= 0 ? inv[nb.y] : nb.y;
        nb.z = nb.z >= 0 ? inv[nb.z] : nb.z;
        nb.w = nb.w >= 0 ? inv[nb.w] : nb.w;
        reinterpret_cast<int4*>(out)[i] = nb;
    }
}

// one float4 per thread: dst[r, c4] = src[idx[r], c4]
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, long long n_rows,
                                   int row4, float* __restrict__ dst) {
    const long long total = n_rows * row4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / row4;
        int c = (int)(i % row4);
        int s = idx[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s >= 0) v = __ldg(reinterpret_cast<const float4*>(src) + (long long)s * row4 + c);
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, long long n_rows,
                                    int row4, float* __restrict__ dst) {
    const long long total = n_rows * row4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / row4;
        int c = (int)(i % row4);
        int d = idx[r];
        if (d >= 0) reinterpret_cast<float4*>(dst)[(long long)d * row4 + c] = __ldg(reinterpret_cast<const float4*>(src) + i);
    }
}

// e_id (nullable): row r of the edge list is row e_id[r] of `ea` (data.all.edge_attr[e_id] without materialising it)
// dst[idx[r],:] += src[r,:]; the indices of one call must be distinct (one peer's boundary list)
__global__ void add_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, long long n_rows,
                                int row4, float* __restrict__ dst) {
    const long long total = n_rows * row4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / row4;
        int c = (int)(i % row4);
        int d = idx[r];
        if (d < 0) continue;
        float4* o = reinterpret_cast<float4*>(dst) + (long long)d * row4 + c;
        const float4 a = *o, b = __ldg(reinterpret_cast<const float4*>(src) + i);
        *o = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

__global__ void edge_relayout_kernel(const float* __restrict__ ea, const long long* __restrict__ e_id,
                                     const int32_t* __restrict__ nbr, const uint8_t* __restrict__ rslot,
                                     const int32_t* __restrict__ perm, long long n, int fe4,
                                     float* __restrict__ ea_in, float* __restrict__ ea_own) {
    const long long total = n * 4 * fe4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long row = i / fe4;  // new*4 + k
        int c = (int)(i % fe4);
        long long nw = row >> 2;
        int k = (int)(row & 3);
        long long old = perm ? perm[nw] : nw;
        if (ea_own) {
            long long r = old * 4 + k;
            if (e_id) r = e_id[r];
            reinterpret_cast<float4*>(ea_own)[i] = __ldg(reinterpret_cast<const float4*>(ea) + r * fe4 + c);
        }
        if (ea_in) {
            long long s = nbr[old * 4 + k];
            int rs = rslot[old * 4 + k];
            long long r = s * 4 + rs;
            if (e_id) r = e_id[r];
            reinterpret_cast<float4*>(ea_in)[i] = __ldg(reinterpret_cast<const float4*>(ea) + r * fe4 + c);
        }
    }
}

}  // namespace dgnn

using namespace dgnn;

static inline int ggrid(long long n) {
    long long g = (n + 255) / 256;
    long long cap = (long long)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" int dgnn_ell_from_adjacency(const int32_t* adj, int64_t n, int32_t* nbr, uint8_t* rslot, int32_t* err_flag,
                                       void* stream) {
    DGNN_REQUIRE(adj && nbr && rslot && err_flag, "null pointer");
    if (n <= 0) return 0;
    ell_from_adj_kernel<<<ggrid(n), 256, 0, as_stream(stream)>>>(adj, n, nbr, rslot, err_flag);
    return check_launch("dgnn_ell_from_adjacency");
}

extern "C" int dgnn_ell_build(const int64_t* src, const int64_t* tgt, int64_t n_edges, int64_t n_rows, int64_t n_other,
                              int by_source, int32_t* nbr, int32_t* eid, int32_t* cnt, int32_t* err_flag, void* stream) {
    DGNN_REQUIRE(nbr && eid && cnt && err_flag, "null pointer");
    DGNN_REQUIRE(n_other < ((int64_t)1 << 31) && n_rows < ((int64_t)1 << 31), "more than 2^31 cells");
    cudaStream_t st = as_stream(stream);
    if (n_edges > 0) {
        ell_fill_kernel<<<ggrid(n_edges), 256, 0, st>>>((const long long*)src, (const long long*)tgt, n_edges, n_rows,
                                                        n_other, nbr, eid, cnt, err_flag);
        if (check_launch("dgnn_ell_build")) return 1;
    }
    if (n_rows > 0) ell_sort_kernel<<<ggrid(n_rows), 256, 0, st>>>(n_rows, nbr, eid, cnt, by_source);
    return check_launch("dgnn_ell_build");
}

extern "C" int dgnn_morton_codes(const float* pos, int64_t n, const float* lo_host, const float* hi_host, int bits,
                                 uint64_t* codes, void* stream) {
    DGNN_REQUIRE(bits >= 1 && bits <= 21, "bits in 1..21");
    if (n <= 0) return 0;
    float ext[3];
    for (int a = 0; a < 3; ++a) {
        ext[a] = hi_host[a] - lo_host[a];
        if (!(ext[a] > 1e-30f)) ext[a] = 1e-30f;
    }
    morton_kernel<<<ggrid(n), 256, 0, as_stream(stream)>>>(pos, n, lo_host[0], lo_host[1], lo_host[2], ext[0], ext[1],
                                                           ext[2], bits, (unsigned long long*)codes);
    return check_launch("dgnn_morton_codes");
}

extern "C" int dgnn_perm_apply_ell(const int32_t* nbr, const int32_t* perm, const int32_t* inv, int64_t n, int32_t* out,
                                   void* stream) {
    if (n <= 0) return 0;
    perm_apply_ell_kernel<<<ggrid(n), 256, 0, as_stream(stream)>>>(nbr, perm, inv, n, out);
    return check_launch("dgnn_perm_apply_ell");
}

extern "C" int dgnn_gather_rows(const float* src, const int32_t* idx, int64_t n_rows, int row_floats, float* dst,
                                void* stream) {
    DGNN_REQUIRE(row_floats % 4 == 0, "row length must be a multiple of 4 floats");
    if (n_rows <= 0) return 0;
    gather_rows_kernel<<<ggrid(n_rows * (row_floats / 4)), 256, 0, as_stream(stream)>>>(src, idx, n_rows, row_floats / 4, dst);
    return check_launch("dgnn_gather_rows");
}
extern "C" int dgnn_scatter_rows(const float* src, const int32_t* idx, int64_t n_rows, int row_floats, float* dst,
                                 void* stream) {
    DGNN_REQUIRE(row_floats % 4 == 0, "row length must be a multiple of 4 floats");
    if (n_rows <= 0) return 0;
    scatter_rows_kernel<<<ggrid(n_rows * (row_floats / 4)), 256, 0, as_stream(stream)>>>(src, idx, n_rows, row_floats / 4, dst);
    return check_launch("dgnn_scatter_rows");
}

extern "C" int dgnn_add_rows(const float* src, const int32_t* idx, int64_t n_rows, int row_floats, float* dst,
                             void* stream) {
    DGNN_REQUIRE(row_floats % 4 == 0, "row length must be a multiple of 4 floats");
    if (n_rows <= 0) return 0;
    add_rows_kernel<<<ggrid(n_rows * (row_floats / 4)), 256, 0, as_stream(stream)>>>(src, idx, n_rows, row_floats / 4, dst);
    return check_launch("dgnn_add_rows");
}

extern "C" int dgnn_edge_relayout(const float* ea, const int32_t* nbr, const uint8_t* rslot, const int32_t* perm,
                                  int64_t n, int fe, float* ea_in, float* ea_own, void* stream) {
    DGNN_REQUIRE(fe % 4 == 0, "edge feature width must be a multiple of 4");
    DGNN_REQUIRE(!ea_in || (nbr && rslot), "incoming order needs nbr and rslot");
    if (n <= 0) return 0;
    edge_relayout_kernel<<<ggrid(n * fe), 256, 0, as_stream(stream)>>>(ea, nullptr, nbr, rslot, perm, n, fe / 4, ea_in, ea_own);
    return check_launch("dgnn_edge_relayout");
}

extern "C" int dgnn_edge_relayout_idx(const float* ea, const int64_t* e_id, const int32_t* nbr, const uint8_t* rslot,
                                      const int32_t* perm, int64_t n, int fe, float* ea_in, float* ea_own,
                                      void* stream) {
    DGNN_REQUIRE(fe % 4 == 0, "edge feature width must be a multiple of 4");
    DGNN_REQUIRE(!ea_in || (nbr && rslot), "incoming order needs nbr and rslot");
    if (n <= 0) return 0;
    edge_relayout_kernel<<<ggrid(n * fe), 256, 0, as_stream(stream)>>>(ea, (const long long*)e_id, nbr, rslot, perm, n,
                                                                       fe / 4, ea_in, ea_own);
    return check_launch("dgnn_edge_relayout_idx");
}

// ---- per-graph feature standardisation (processing/data.py:467-506: sklearn StandardScaler per graph) --------------
// columns col0 .. col0+c-1 of x float32[n, ld]:  partials[block][0][j] = sum (x - m_j), partials[block][1][j] = sum (x - m_j)^2
// (m = NULL: zeros).  One thread per row chunk, double accumulation, per-block partials (deterministic).
namespace dgnn {
constexpr int STD_MAXC = 64;
template <typename T>
__global__ void __launch_bounds__(256) column_moments_kernel(const T* __restrict__ x, long long n, int ld, int col0,
                                                             int c, const double* __restrict__ m,
                                                             double* __restrict__ partials) {
    __shared__ double red[2][256];
    // thread = (row lane rsub, column): rows_per_pass = blockDim.x / c
    const int col = threadIdx.x % c, rsub = threadIdx.x / c, rpp = blockDim.x / c;
    double s1 = 0.0, s2 = 0.0;
    if (rsub < rpp) {
        const double mj = m ? m[col] : 0.0;
        for (long long r = (long long)blockIdx.x * rpp + rsub; r < n; r += (long long)gridDim.x * rpp) {
            const double d = (double)x[r * ld + col0 + col] - mj;
            s1 += d; s2 += d * d;
        }
    }
    red[0][threadIdx.x] = s1; red[1][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < c) {                              // fixed summation order over the row lanes
        double a = 0.0, b = 0.0;
        for (int q = 0; q < rpp; ++q) { a += red[0][q * c + threadIdx.x]; b += red[1][q * c + threadIdx.x]; }
        partials[(size_t)blockIdx.x * 2 * c + threadIdx.x] = a;
        partials[(size_t)blockIdx.x * 2 * c + c + threadIdx.x] = b;
    }
}

// out = (x - shift) * k (DIVIDE = false) or (x - shift) / k (DIVIDE = true, sklearn's `X -= mean; X /= scale`), evaluated
// in float64 and rounded to float32 once (processing/data.py:444-519: standardise in float64, then torch.float)
template <typename T, bool DIVIDE>
__global__ void __launch_bounds__(256) column_affine_kernel(const T* __restrict__ x, long long n, int ld, int col0,
                                                            int c, const double* __restrict__ shift,
                                                            const double* __restrict__ k, int ld_out, int col0_out,
                                                            float* __restrict__ out) {
    const long long total = n * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        const int j = (int)(i % c);
        const double d = (double)x[r * ld + col0 + j] - shift[j];
        out[r * ld_out + col0_out + j] = (float)(DIVIDE ? d / k[j] : d * k[j]);
    }
}
}  // namespace dgnn

extern "C" int dgnn_column_moments(const float* x, int64_t n, int ld, int col0, int c, const double* center,
                                   double* partials, int n_blocks, void* stream) {
    DGNN_REQUIRE(x && partials, "null pointer");
    DGNN_REQUIRE(c > 0 && c <= dgnn::STD_MAXC && col0 >= 0 && col0 + c <= ld, "column range");
    DGNN_REQUIRE(n_blocks > 0, "n_blocks");
    dgnn::column_moments_kernel<float><<<n_blocks, 256, 0, as_stream(stream)>>>(x, n, ld, col0, c, center, partials);
    return check_launch("dgnn_column_moments");
}

extern "C" int dgnn_column_moments_f64(const double* x, int64_t n, int ld, int col0, int c, const double* center,
                                       double* partials, int n_blocks, void* stream) {
    DGNN_REQUIRE(x && partials, "null pointer");
    DGNN_REQUIRE(c > 0 && c <= dgnn::STD_MAXC && col0 >= 0 && col0 + c <= ld, "column range");
    DGNN_REQUIRE(n_blocks > 0, "n_blocks");
    dgnn::column_moments_kernel<double><<<n_blocks, 256, 0, as_stream(stream)>>>(x, n, ld, col0, c, center, partials);
    return check_launch("dgnn_column_moments_f64");
}

extern "C" int dgnn_column_standardize_f64(const double* x, int64_t n, int ld, int col0, int c, const double* mean,
                                           const double* scale, int ld_out, int col0_out, float* out, void* stream) {
    DGNN_REQUIRE(x && mean && scale && out, "null pointer");
    DGNN_REQUIRE(c > 0 && col0 >= 0 && col0 + c <= ld && col0_out >= 0 && col0_out + c <= ld_out, "column range");
    if (n <= 0) return 0;
    dgnn::column_affine_kernel<double, true><<<ggrid(n * c), 256, 0, as_stream(stream)>>>(x, n, ld, col0, c, mean, scale,
                                                                                         ld_out, col0_out, out);
    return check_launch("dgnn_column_standardize_f64");
}

extern "C" int dgnn_column_affine(const float* x, int64_t n, int ld, int col0, int c, const double* shift,
                                  const double* inv_scale, int ld_out, float* out, void* stream) {
    DGNN_REQUIRE(x && shift && inv_scale && out, "null pointer");
    DGNN_REQUIRE(c > 0 && col0 >= 0 && col0 + c <= ld && col0 + c <= ld_out, "column range");
    if (n <= 0) return 0;
    dgnn::column_affine_kernel<float, false><<<ggrid(n * c), 256, 0, as_stream(stream)>>>(x, n, ld, col0, c, shift, inv_scale,
                                                                                         ld_out, col0, out);
    return check_launch("dgnn_column_affine");
}
