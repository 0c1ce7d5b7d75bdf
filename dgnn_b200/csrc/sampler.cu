// On-device neighbour-closure sampler over the ELL-4 in-edge table: the full-neighbourhood
// NeighborSampler of the reference's training loop (run.py:59-74, sizes = [-1] * hops) as k frontier expansions.
// One hop, given the current node list n_id (targets = all of it):
//   expand   : every target's in-edges in (target, ascending edge id) order -> (e_id, global source, local target)
//   mark     : first[s] = min position of an edge whose source s is not yet in n_id           (atomicMin: order-free)
//   flag     : position p introduces a new node iff first[src[p]] == p                          -> caller's prefix sum
//   assign   : new nodes appended to n_id in order of first appearance, loc[s] = their local id
//   localise : local source id of every edge
// All integer work; results are bit-exact against the oracle's restatement of the PyG sampler.
#include "common.cuh"

namespace dgnn {

__global__ void __launch_bounds__(256) sampler_expand_kernel(const long long* __restrict__ n_id, long long n_tgt,
                                                             const int32_t* __restrict__ in_src,
                                                             const int32_t* __restrict__ in_eid,
                                                             const long long* __restrict__ offs,
                                                             long long* __restrict__ e_id, long long* __restrict__ src_g,
                                                             long long* __restrict__ tgt_l) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_tgt; t += (long long)gridDim.x * blockDim.x) {
        const long long g = n_id[t];
        const int4 s4 = __ldg(reinterpret_cast<const int4*>(in_src) + g);
        const int4 e4 = __ldg(reinterpret_cast<const int4*>(in_eid) + g);
        const int sv[4] = {s4.x, s4.y, s4.z, s4.w}, ev[4] = {e4.x, e4.y, e4.z, e4.w};
        long long o = offs[t];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (sv[k] < 0) continue;                   // rows are filled from slot 0, ascending edge id
            e_id[o] = ev[k]; src_g[o] = sv[k]; tgt_l[o] = t;
            ++o;
        }
    }
}

__global__ void __launch_bounds__(256) sampler_degree_kernel(const long long* __restrict__ n_id, long long n_tgt,
                                                             const int32_t* __restrict__ in_src,
                                                             long long* __restrict__ deg) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_tgt; t += (long long)gridDim.x * blockDim.x) {
        const int4 s4 = __ldg(reinterpret_cast<const int4*>(in_src) + n_id[t]);
        deg[t] = (s4.x >= 0) + (s4.y >= 0) + (s4.z >= 0) + (s4.w >= 0);
    }
}

// loc[n_id[i]] = base + i
__global__ void __launch_bounds__(256) sampler_set_loc_kernel(const long long* __restrict__ ids, long long n, int base,
                                                              int32_t* __restrict__ loc) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        loc[ids[i]] = base < 0 ? -1 : base + (int)i;
}

__global__ void __launch_bounds__(256) sampler_mark_kernel(const long long* __restrict__ src_g, long long n_e,
                                                           const int32_t* __restrict__ loc, int32_t* __restrict__ first) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_e; p += (long long)gridDim.x * blockDim.x) {
        const long long s = src_g[p];
        if (loc[s] < 0) atomicMin(first + s, (int)p);
    }
}

__global__ void __launch_bounds__(256) sampler_flag_kernel(const long long* __restrict__ src_g, long long n_e,
                                                           const int32_t* __restrict__ loc,
                                                           const int32_t* __restrict__ first, long long* __restrict__ flag) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_e; p += (long long)gridDim.x * blockDim.x) {
        const long long s = src_g[p];
        flag[p] = (loc[s] < 0 && first[s] == (int)p) ? 1 : 0;
    }
}

// rank[p] = inclusive prefix sum of flag; new node of position p gets local id n_tgt + rank[p] - 1
__global__ void __launch_bounds__(256) sampler_assign_kernel(const long long* __restrict__ src_g, long long n_e,
                                                             const long long* __restrict__ flag,
                                                             const long long* __restrict__ rank, long long n_tgt,
                                                             int32_t* __restrict__ loc, int32_t* __restrict__ first,
                                                             long long* __restrict__ new_ids) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_e; p += (long long)gridDim.x * blockDim.x) {
        if (!flag[p]) continue;
        const long long s = src_g[p];
        const long long r = rank[p] - 1;
        new_ids[r] = s;
        loc[s] = (int)(n_tgt + r);
        first[s] = 0x7fffffff;                         // scratch back to "unseen" for the next hop / batch
    }
}

__global__ void __launch_bounds__(256) sampler_localise_kernel(const long long* __restrict__ src_g, long long n_e,
                                                               const int32_t* __restrict__ loc,
                                                               long long* __restrict__ src_l) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_e; p += (long long)gridDim.x * blockDim.x)
        src_l[p] = loc[src_g[p]];
}

static inline int sgrid(long long n) {
    long long g = (n + 255) / 256, cap = (long long)sm_count() * 8;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_sampler_degree(const int64_t* n_id, int64_t n_tgt, const int32_t* in_src, int64_t* deg, void* stream) {
    DGNN_REQUIRE(n_id && in_src && deg, "null pointer");
    if (n_tgt <= 0) return 0;
    sampler_degree_kernel<<<sgrid(n_tgt), 256, 0, as_stream(stream)>>>((const long long*)n_id, n_tgt, in_src, (long long*)deg);
    return check_launch("dgnn_sampler_degree");
}

extern "C" int dgnn_sampler_expand(const int64_t* n_id, int64_t n_tgt, const int32_t* in_src, const int32_t* in_eid,
                                   const int64_t* offsets, int64_t* e_id, int64_t* src_global, int64_t* tgt_local,
                                   void* stream) {
    DGNN_REQUIRE(n_id && in_src && in_eid && offsets && e_id && src_global && tgt_local, "null pointer");
    if (n_tgt <= 0) return 0;
    sampler_expand_kernel<<<sgrid(n_tgt), 256, 0, as_stream(stream)>>>((const long long*)n_id, n_tgt, in_src, in_eid,
                                                                       (const long long*)offsets, (long long*)e_id,
                                                                       (long long*)src_global, (long long*)tgt_local);
    return check_launch("dgnn_sampler_expand");
}

extern "C" int dgnn_sampler_set_loc(const int64_t* ids, int64_t n, int base, int32_t* loc, void* stream) {
    DGNN_REQUIRE(ids && loc, "null pointer");
    if (n <= 0) return 0;
    sampler_set_loc_kernel<<<sgrid(n), 256, 0, as_stream(stream)>>>((const long long*)ids, n, base, loc);
    return check_launch("dgnn_sampler_set_loc");
}

extern "C" int dgnn_sampler_mark(const int64_t* src_global, int64_t n_edges, const int32_t* loc, int32_t* first,
                                 int64_t* flag, void* stream) {
    DGNN_REQUIRE(src_global && loc && first && flag, "null pointer");
    if (n_edges <= 0) return 0;
    sampler_mark_kernel<<<sgrid(n_edges), 256, 0, as_stream(stream)>>>((const long long*)src_global, n_edges, loc, first);
    sampler_flag_kernel<<<sgrid(n_edges), 256, 0, as_stream(stream)>>>((const long long*)src_global, n_edges, loc, first,
                                                                      (long long*)flag);
    return check_launch("dgnn_sampler_mark");
}

extern "C" int dgnn_sampler_assign(const int64_t* src_global, int64_t n_edges, const int64_t* flag, const int64_t* rank,
                                   int64_t n_tgt, int32_t* loc, int32_t* first, int64_t* new_ids, int64_t* src_local,
                                   void* stream) {
    DGNN_REQUIRE(src_global && flag && rank && loc && first && src_local, "null pointer");
    if (n_edges <= 0) return 0;
    sampler_assign_kernel<<<sgrid(n_edges), 256, 0, as_stream(stream)>>>((const long long*)src_global, n_edges,
                                                                        (const long long*)flag, (const long long*)rank,
                                                                        n_tgt, loc, first, (long long*)new_ids);
    sampler_localise_kernel<<<sgrid(n_edges), 256, 0, as_stream(stream)>>>((const long long*)src_global, n_edges, loc,
                                                                          (long long*)src_local);
    return check_launch("dgnn_sampler_assign");
}
