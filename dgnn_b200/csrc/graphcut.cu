// Graph-cut regularisation of the cell labels on the device.
//
// processing/generate_mesh.py:15-58 builds a two-label energy over the finite cells,
//     E(l) = sum_c D(c, l_c) + w * #{facets (a,b) : l_a != l_b},
//     D(c,0) = round(z[c,1] * unary_weight),  D(c,1) = round(z[c,0] * unary_weight)   (the logit columns swapped, :25-26),
//     w = binary_weight (Potts, every finite-finite facet once, :31-39),
// and minimises it with gco's alpha-expansion.  For two labels and a Potts term the energy is submodular, a labelling no
// expansion move improves is a global minimum, and the global minimum is one s-t minimum cut:
//     cap(s -> c) = D(c,1),  cap(c -> t) = D(c,0)  (shifted to be non-negative),  cap(a <-> b) = w.
// Here the minimum cut is computed with a lock-free push-relabel (Hong & He): one thread per cell pushes its excess to
// the lowest residual neighbour or relabels itself, with integer atomics on the residual capacities and excesses, and a
// periodic global relabelling (backward breadth-first search from the sink over the residual graph).  The cell graph is
// the ELL-4 facet table; the terminal arcs are folded into a per-cell excess / sink capacity.
#include "common.cuh"

namespace dgnn {

// terminal arcs, pre-saturated: net = D(c,1) - D(c,0) > 0 leaves that much excess at c, < 0 a sink arc of -net
__global__ void __launch_bounds__(256) gc_terminals_kernel(const float* __restrict__ z, long long n, float uw,
                                                           long long* __restrict__ excess, long long* __restrict__ sink_cap,
                                                           long long* __restrict__ d0, long long* __restrict__ d1) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const long long c0 = llrintf(z[c * 2 + 1] * uw);      // cost of label 0 (inside): the OUTSIDE logit (:25-26)
        const long long c1 = llrintf(z[c * 2 + 0] * uw);
        if (d0) { d0[c] = c0; d1[c] = c1; }
        const long long net = c1 - c0;
        excess[c] = net > 0 ? net : 0;
        sink_cap[c] = net < 0 ? -net : 0;
    }
}

// `iters` asynchronous push / relabel attempts per cell
__global__ void __launch_bounds__(256) gc_push_relabel_kernel(long long n, const int32_t* __restrict__ nbr,
                                                              const uint8_t* __restrict__ rslot, int* cap,
                                                              unsigned long long* excess, long long* sink_cap, int* height,
                                                              int hmax, int iters) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int4 nb = reinterpret_cast<const int4*>(nbr)[c];
        const int nv[4] = {nb.x, nb.y, nb.z, nb.w};
        const uchar4 rs4 = reinterpret_cast<const uchar4*>(rslot)[c];
        const int rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
        for (int it = 0; it < iters; ++it) {
            long long e = (long long)atomicAdd(&excess[c], 0ull);
            int h = height[c];
            if (e <= 0 || h >= hmax) break;
            // the sink arc is private to the cell
            long long sc = sink_cap[c];
            if (sc > 0) {
                const long long d = e < sc ? e : sc;
                sink_cap[c] = sc - d;
                atomicAdd(&excess[c], (unsigned long long)(-d));
                continue;
            }
            int best = -1, hmin = 0x7fffffff;
            for (int k = 0; k < 4; ++k) {
                if (nv[k] < 0) continue;
                if (atomicAdd(&cap[c * 4 + k], 0) <= 0) continue;
                const int hv = atomicAdd(&height[nv[k]], 0);
                if (hv < hmin) { hmin = hv; best = k; }
            }
            if (best < 0) { height[c] = hmax; break; }           // no residual arc at all: the excess stays on the source side
            if (h > hmin) {
                const int v = nv[best];
                int cv = atomicAdd(&cap[c * 4 + best], 0);
                long long d = e < (long long)cv ? e : (long long)cv;
                if (d > 0) {
                    atomicSub(&cap[c * 4 + best], (int)d);        // only this cell lowers its own arc: never below zero
                    atomicAdd(&cap[(long long)v * 4 + rs[best]], (int)d);
                    atomicAdd(&excess[c], (unsigned long long)(-d));
                    atomicAdd(&excess[v], (unsigned long long)d);
                }
            } else {
                height[c] = hmin + 1 < hmax ? hmin + 1 : hmax;
            }
        }
    }
}

// global relabelling: exact distance to the sink in the residual graph (hmax = unreachable)
__global__ void __launch_bounds__(256) gc_bfs_init_kernel(long long n, const long long* __restrict__ sink_cap,
                                                          int* __restrict__ height, int hmax) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x)
        height[c] = sink_cap[c] > 0 ? 1 : hmax;
}
// cells at distance `level` pull in the neighbours u that own a residual arc u -> c
__global__ void __launch_bounds__(256) gc_bfs_step_kernel(long long n, const int32_t* __restrict__ nbr,
                                                          const uint8_t* __restrict__ rslot, const int* __restrict__ cap,
                                                          int* height, int level, int hmax, int* changed) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        if (height[c] != level) continue;
        const int4 nb = reinterpret_cast<const int4*>(nbr)[c];
        const int nv[4] = {nb.x, nb.y, nb.z, nb.w};
        const uchar4 rs4 = reinterpret_cast<const uchar4*>(rslot)[c];
        const int rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
        for (int k = 0; k < 4; ++k) {
            const int u = nv[k];
            if (u < 0) continue;
            if (cap[(long long)u * 4 + rs[k]] > 0 && height[u] == hmax) {   // benign race: every writer stores level + 1
                height[u] = level + 1;
                *changed = 1;
            }
        }
    }
}

__global__ void __launch_bounds__(256) gc_active_kernel(long long n, const long long* __restrict__ excess,
                                                        const int* __restrict__ height, int hmax, unsigned long long* count) {
    unsigned long long mine = 0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x)
        mine += (excess[c] > 0 && height[c] < hmax) ? 1ull : 0ull;
    mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, mine);
}

// sink side (can still reach the sink in the residual graph) = label 1 (outside), source side = label 0 (inside)
__global__ void __launch_bounds__(256) gc_labels_kernel(long long n, const int* __restrict__ height, int hmax,
                                                        uint8_t* __restrict__ labels) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x)
        labels[c] = height[c] < hmax ? 1 : 0;
}

// E(l) = sum_c D(c, l_c) + w * cut facets (each facet is stored at both of its cells: halve), per-block partials
__global__ void __launch_bounds__(256) gc_energy_kernel(long long n, const float* __restrict__ z, float uw,
                                                        const int32_t* __restrict__ nbr, int w,
                                                        const uint8_t* __restrict__ labels, long long* partial2) {
    long long data = 0, cut = 0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int l = labels[c];
        data += llrintf(z[c * 2 + (l ? 0 : 1)] * uw);
        const int4 nb = reinterpret_cast<const int4*>(nbr)[c];
        const int nv[4] = {nb.x, nb.y, nb.z, nb.w};
        for (int k = 0; k < 4; ++k)
            if (nv[k] >= 0 && labels[nv[k]] != l) cut += w;
    }
    __shared__ long long sd[256], sc[256];
    sd[threadIdx.x] = data; sc[threadIdx.x] = cut;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long a = 0, b = 0;
        for (int i = 0; i < 256; ++i) { a += sd[i]; b += sc[i]; }
        partial2[2 * blockIdx.x] = a;
        partial2[2 * blockIdx.x + 1] = b;
    }
}

}  // namespace dgnn

using namespace dgnn;

static inline int gc_grid(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (g > cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

extern "C" int dgnn_gc_terminals(const float* logits, int64_t n, float unary_weight, int64_t* excess, int64_t* sink_cap,
                                 int64_t* d0, int64_t* d1, void* stream) {
    DGNN_REQUIRE(logits && excess && sink_cap, "null pointer");
    if (n <= 0) return 0;
    gc_terminals_kernel<<<gc_grid(n), 256, 0, as_stream(stream)>>>(logits, n, unary_weight, (long long*)excess,
                                                                  (long long*)sink_cap, (long long*)d0, (long long*)d1);
    return check_launch("dgnn_gc_terminals");
}

extern "C" int dgnn_gc_push_relabel(int64_t n, const int32_t* nbr, const uint8_t* rslot, int32_t* cap, int64_t* excess,
                                    int64_t* sink_cap, int32_t* height, int hmax, int iters, void* stream) {
    DGNN_REQUIRE(nbr && rslot && cap && excess && sink_cap && height, "null pointer");
    if (n <= 0) return 0;
    gc_push_relabel_kernel<<<gc_grid(n), 256, 0, as_stream(stream)>>>(n, nbr, rslot, cap, (unsigned long long*)excess,
                                                                     (long long*)sink_cap, height, hmax, iters);
    return check_launch("dgnn_gc_push_relabel");
}

extern "C" int dgnn_gc_bfs_init(int64_t n, const int64_t* sink_cap, int32_t* height, int hmax, void* stream) {
    if (n <= 0) return 0;
    gc_bfs_init_kernel<<<gc_grid(n), 256, 0, as_stream(stream)>>>(n, (const long long*)sink_cap, height, hmax);
    return check_launch("dgnn_gc_bfs_init");
}

extern "C" int dgnn_gc_bfs_step(int64_t n, const int32_t* nbr, const uint8_t* rslot, const int32_t* cap, int32_t* height,
                                int level, int hmax, int32_t* changed, void* stream) {
    if (n <= 0) return 0;
    gc_bfs_step_kernel<<<gc_grid(n), 256, 0, as_stream(stream)>>>(n, nbr, rslot, cap, height, level, hmax, changed);
    return check_launch("dgnn_gc_bfs_step");
}

extern "C" int dgnn_gc_active(int64_t n, const int64_t* excess, const int32_t* height, int hmax, uint64_t* count,
                              void* stream) {
    if (n <= 0) return 0;
    gc_active_kernel<<<gc_grid(n), 256, 0, as_stream(stream)>>>(n, (const long long*)excess, height, hmax,
                                                               (unsigned long long*)count);
    return check_launch("dgnn_gc_active");
}

extern "C" int dgnn_gc_labels(int64_t n, const int32_t* height, int hmax, uint8_t* labels, void* stream) {
    if (n <= 0) return 0;
    gc_labels_kernel<<<gc_grid(n), 256, 0, as_stream(stream)>>>(n, height, hmax, labels);
    return check_launch("dgnn_gc_labels");
}

extern "C" int dgnn_gc_energy_grid(void) { return sm_count() * 4; }

extern "C" int dgnn_gc_energy(int64_t n, const float* logits, float unary_weight, const int32_t* nbr, int binary_weight,
                              const uint8_t* labels, int64_t* partials, void* stream) {
    DGNN_REQUIRE(logits && nbr && labels && partials, "null pointer");
    gc_energy_kernel<<<dgnn_gc_energy_grid(), 256, 0, as_stream(stream)>>>(n, logits, unary_weight, nbr, binary_weight, labels,
                                                                          (long long*)partials);
    return check_launch("dgnn_gc_energy");
}
