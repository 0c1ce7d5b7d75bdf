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
// that lands in TMEM (4 slots x F <= 512 columns).  Each thread owns one cell row and a strip of F/4 features;
// slot by slot it reads the neighbour row's strip (staged through a warp-private cp.async ring; the norm affine +
// ReLU of the producer layer applied on the way), the matching PHI strip from TMEM (tcgen05.ld) and accumulates
// h * phi in registers.
// dW_e is the product P^T . EA with P = dphi [edges x F]: both operands need the edge index contiguous,
// so each thread scatters its strip of P (rounded to TF32) and of EA (hi / lo) transposed into K-major
// shared-memory operands; one [F x 32] accumulator per row quarter lives in TMEM for the whole kernel.
//
// One persistent CTA per SM with 16 compute warps (4 per scheduler => 128 registers per thread).  There is no
// dedicated MMA warp: every warp counts itself in on a shared-memory counter after it has written its share of an
// operand stage, and the warp that completes the count issues the stage's tcgen05.mma + commits (elect-by-arrival).
// No CTA-wide barrier in the loop.
#include "umma.cuh"
#include "common.cuh"

namespace dgnn {

using namespace umma;

constexpr int G_NCW = 16;
constexpr int G_THREADS = G_NCW * 32;                // no dedicated MMA warp: the last warp to finish a stage issues its MMAs
constexpr int G_M = 128;
constexpr int G_EA_STAGES = 2;                       // EA operand ring (hi | lo per stage)
#ifndef DGNN_NST_NARROW
#define DGNN_NST_NARROW 4
#endif
constexpr int G_NST_NARROW = DGNN_NST_NARROW;          // x-ring stages of the gather / dW_e kernels at F <= 64
constexpr int G_EAW = 4;                             // gather_tc_kernel: EA warps (two per operand stage, 64 rows each)
constexpr int G_ATOM = G_M * 128;                    // 16 KB
constexpr int G_P_BYTES = G_ATOM + 2 * 32 * 128;     // P_hi [128 x 32 cells] + EA^T hi / lo [32 x 32 cells]
#ifndef DGNN_PF_AHEAD
#define DGNN_PF_AHEAD 2
#endif
constexpr int G_PF_AHEAD = DGNN_PF_AHEAD;              // L2 prefetch distance in tiles of this CTA
constexpr int DWE_ROUNDS = 1;                        // dW_e: 1 = P rounded to TF32 once, 2 = P split hi / lo (see dwe_tc_kernel)

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
    int f;                 // feature width of this launch's slice (multiple of 4, <= 128)
    int fp;                // 32, 64 or 128
    int ld;                // row stride of x / out / addend / z_prev (the layer's full width; pointers are pre-offset)
    int s_ld;              // full width: stride of the (S1 | S2) blocks and rows of the dW_e partials per CTA
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

// L2 prefetch of a contiguous byte range, one 128-byte line per thread and instruction (no register result, nothing to wait
// for): the tile after next is pulled from DRAM into L2 while this one is computed, so the register / cp.async loads that
// feed the pipeline one item ahead see L2 latency instead of DRAM latency.
__device__ __forceinline__ void l2_prefetch_range(const void* base, size_t bytes, int tid, int nthreads) {
    const char* b = reinterpret_cast<const char*>(base);
    for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)nthreads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
}

// ---------------------------------------------------------------------------------------------------
// MODE 0 / 1: phi on tensor cores, strips accumulated in registers.
//
// Everything is pipelined per ITEM = (tile, slot k):
//   x ring    the 32 rows x CPT features a warp consumes for one item are copied global -> shared with
//             cp.async (16 B per lane, CH lanes per row => full 32..128-byte row segments per request instead
//             of one sector per lane) into a warp-private, XOR-swizzled 2-stage ring; the copy of item n+1 is
//             in flight while item n is consumed, with no registers held.  MODE 1 appends two plain items per
//             tile: the rows' own d_self (k = 4) and z_prev (k = 5) strips.
//   EA ring   2 stages of (hi | lo) K-major operands; the global loads of item n+3 are issued, and the
//             registers of item n+2 stored, while item n is consumed.  The warp that completes a stage's arrival
//             count issues the 9 MMAs of PHI item n+2 (elect-by-arrival, see the file header).
//   PHI ring  4 TMEM buffers of FP columns (item n in buffer n & 3), ready two items before they are consumed;
//             no "free" barrier: the arrival count of item n already orders the overwrite of item n-4's buffer.
// NW compute warps: 16 (4 per scheduler, 128 registers) or 32 (8 per scheduler, 64 registers; twice the warps to hide the
// shared-memory / TMEM / mbarrier latencies with, half the strip per thread)
template <int CPT, int MODE, int NW, int NST>  // CPT: features per thread = fp / (NW / 4); NST: x-ring stages (power of 2)
__global__ void __launch_bounds__((NW + G_EAW) * 32, 1) gather_tc_kernel(const GatherTcArgs p) {
    constexpr int G_NCW = NW, G_THREADS = NW * 32;      // compute warps / threads (shadow the file-level constants of the dW_e kernel)
    constexpr int NT = (NW + G_EAW) * 32;               // all threads: + the EA warps
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t ea_empty[G_EA_STAGES], ea_ready[G_EA_STAGES];
    __shared__ uint64_t phi_full[4], phi_free[4];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int FP = CPT * (NW / 4);
    constexpr int CH = CPT / 4;                          // 16-byte chunks per strip row (2, 4 or 8)
    constexpr int RPI = 32 / CH;                         // rows copied by one cp.async warp instruction
    constexpr int PITCH = CPT * 4;                       // bytes per strip row
    constexpr int WSTAGE = 32 * PITCH;                   // bytes per warp and stage
    constexpr int IPT = MODE == 0 ? 4 : 6;               // x-ring items per tile
    uint8_t* we_hi = smem;                               // [FP rows x 128 B]
    uint8_t* we_lo = we_hi + FP * 128;
    uint8_t* ea_base = we_lo + FP * 128;                 // stages of (hi 16 KB | lo 16 KB)
    uint8_t* x_base = ea_base + (size_t)G_EA_STAGES * 2 * G_ATOM;   // [NST stages][G_NCW warps][32 rows x PITCH]
    float* aff_s = reinterpret_cast<float*>(x_base + (size_t)NST * G_NCW * WSTAGE);   // MODE 0: scale[FP] | shift[FP]
    static_assert(NST >= 2 && (NST & (NST - 1)) == 0 && NST - 1 <= IPT, "x ring: 2 or 4 stages");

    if (tid == 0) {
        for (int s = 0; s < G_EA_STAGES; ++s) { mbar_init(&ea_empty[s], 1); mbar_init(&ea_ready[s], 2); }
        for (int s = 0; s < 4; ++s) { mbar_init(&phi_full[s], 1); mbar_init(&phi_free[s], G_NCW); }
        fence_barrier_init();
    }
    // zero the operands (K padding stays zero), then WE[n][e] = w_e[n][e] (e < fe), WE[n][fe] = b_e[n]
    for (int i = tid; i < (2 * FP * 128 + G_EA_STAGES * 2 * G_ATOM) / 16; i += NT)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int i = tid; i < p.f * (p.fe + 1); i += NT) {
        const int n = i / (p.fe + 1), e = i % (p.fe + 1);
        float v = e < p.fe ? __ldg(p.w_e + (size_t)n * p.fe + e) : __ldg(p.b_e + n);
        float hi, lo;
        split_tf32(v, hi, lo);
        const uint32_t off = atom_off(n, e);
        *reinterpret_cast<float*>(we_hi + off) = hi;
        *reinterpret_cast<float*>(we_lo + off) = lo;
    }
    for (int i = tid; i < G_EA_STAGES * G_M; i += NT) {   // bias column of every EA stage
        const int s = i / G_M, r = i % G_M;
        *reinterpret_cast<float*>(ea_base + (size_t)s * 2 * G_ATOM + atom_off(r, p.fe)) = 1.0f;
    }
    if (MODE == 0) {                                             // producer norm affine, identity when absent / padded
        for (int i = tid; i < FP; i += NT) {
            const bool on = p.scale != nullptr && i < p.f;
            aff_s[i] = on ? __ldg(p.scale + i) : 1.0f;
            aff_s[FP + i] = on ? __ldg(p.shift + i) : 0.0f;
        }
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_tiles = (p.n_rows + G_M - 1) / G_M;
    // tiles of this CTA: blockIdx.x + i * gridDim.x, i < n_my;  PHI items: 4 per tile
    const uint32_t n_my = (int64_t)blockIdx.x < n_tiles ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    const uint32_t n_phi = n_my * 4u;

    if (warp >= NW) {
        // ---------------------------------------------------------------- EA warps: warp e owns operand stage e and the PHI
        // items n = e, e + 2, ..: EA rows of (tile n >> 2, slot n & 3) global -> registers (all of them in flight at once,
        // issued BEFORE the waits on the stage / PHI buffer), split hi / lo -> swizzled operand stage, then the same warp
        // issues PHI item n = EA . WE^T into TMEM buffer n & 3.  The compute warps only consume PHI: they never meet each
        // other at a barrier (elect-by-arrival made all 16 of them rendezvous once per item), and a slow warp delays the
        // others only through the 4-deep PHI ring.
        const int e = (warp - NW) & 1, half = (warp - NW) >> 1;   // operand stage, half of the tile's rows
        const uint32_t idesc = make_idesc_tf32(G_M, FP);
        const uint32_t wh = smem_u32(we_hi), wl = smem_u32(we_lo);
        const int ksteps = (p.fe + 1 + 7) >> 3;        // fe features + bias column, 8 per k-step (3 for fe = 20)
        const int fe4 = p.fe >> 2;
        const int n_f4 = (G_M / 2) * fe4;              // float4 pieces of this warp's 64 rows
        const uint32_t ah = smem_u32(ea_base) + (uint32_t)e * 2u * G_ATOM, al = ah + G_ATOM;
        constexpr int EV = 10;                         // pieces per lane and batch (one batch covers fe <= 20)
        // piece j of this lane in batch 0: global offset (floats from the item's first row) and operand offset (bytes) are
        // the same for every item - formed once (a runtime division per piece and item made these warps the slowest)
        uint32_t goff[EV], soff[EV];
#pragma unroll
        for (int j = 0; j < EV; ++j) {
            const int idx = lane + 32 * j;
            const int r = idx / fe4, c4 = idx - r * fe4;
            goff[j] = idx < n_f4 ? (uint32_t)(r * 4 * p.fe + c4 * 4) : 0xffffffffu;
            soff[j] = atom_off(half * (G_M / 2) + r, c4 * 4);                      // row = soff >> 7
        }
        for (uint32_t n = (uint32_t)e; n < n_phi; n += 2) {
            const uint32_t b = n & 3u;
            const int64_t t0 = ((int64_t)blockIdx.x + (int64_t)(n >> 2) * gridDim.x) * G_M;   // first row of the tile
            const int64_t r0 = t0 + half * (G_M / 2);
            const float* src = p.ea + ((size_t)r0 * 4 + (n & 3u)) * p.fe;
            const uint32_t row_lim = (uint32_t)(p.n_rows - t0 < G_M ? p.n_rows - t0 : G_M);   // rows of the tile that exist
#ifndef DGNN_NO_EA_PREFETCH
            // The four slots of a row share its 4 fe floats of EA, so only the first item of a tile misses L2; pull the
            // NEXT tile's rows (this warp's half of them, split with the other stage's warp) into L2 now, so that an
            // item costs this warp an L2 round trip, not a DRAM one (the EA warps set the pace of the kernel at F <= 64).
            if ((n & 3u) == (uint32_t)e && (n >> 2) + 1 < n_my) {
                const int64_t rn = r0 + (int64_t)gridDim.x * G_M;
                const int64_t left = p.n_rows - rn;
                const int64_t rows = left < G_M / 2 ? left : G_M / 2;
                const char* blk = reinterpret_cast<const char*>(p.ea + (size_t)rn * 4 * p.fe);
                const int64_t bytes = rows * 4 * p.fe * 4;
                for (int64_t o = ((int64_t)e * 32 + lane) * 128; o < bytes; o += 64 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + o));
            }
#endif
            float4 ev[EV];
#pragma unroll
            for (int j = 0; j < EV; ++j) {
                ev[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifndef DGNN_G_NOEALOAD   // timing experiment: no EA loads (wrong results)
                if (goff[j] != 0xffffffffu && (soff[j] >> 7) < row_lim) ev[j] = ldg4_pinned(src + goff[j]);
#endif
            }
            mbar_wait(&ea_empty[e], ((n >> 1) & 1u) ^ 1u);           // this stage's previous MMAs have read it
            mbar_wait(&phi_free[b], ((n >> 2) & 1u) ^ 1u);           // every compute warp is done with item n - 4
#pragma unroll
            for (int j = 0; j < EV; ++j) {
#ifdef DGNN_G_NOEASTORE   // timing experiment: no EA operand stores (wrong results)
                if (false)
#endif
                if (goff[j] != 0xffffffffu) {
                    float4 h, l;
                    split_tf32(ev[j].x, h.x, l.x); split_tf32(ev[j].y, h.y, l.y);
                    split_tf32(ev[j].z, h.z, l.z); split_tf32(ev[j].w, h.w, l.w);
                    sts128(ah + soff[j], h);
                    sts128(al + soff[j], l);
                }
            }
            for (int base = 32 * EV; base < n_f4; base += 32 * EV) {   // wide edge features (fe > 20): further batches
#pragma unroll
                for (int j = 0; j < EV; ++j) {
                    const int idx = base + lane + 32 * j;
                    const int r = idx / fe4, c4 = idx - r * fe4;
                    ev[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < n_f4 && r0 + r < p.n_rows) ev[j] = ldg4_pinned(src + (size_t)r * 4 * p.fe + c4 * 4);
                }
#pragma unroll
                for (int j = 0; j < EV; ++j) {
                    const int idx = base + lane + 32 * j;
                    if (idx < n_f4) {
                        const int r = idx / fe4, c4 = idx - r * fe4;
                        float4 h, l;
                        split_tf32(ev[j].x, h.x, l.x); split_tf32(ev[j].y, h.y, l.y);
                        split_tf32(ev[j].z, h.z, l.z); split_tf32(ev[j].w, h.w, l.w);
                        const uint32_t off = atom_off(half * (G_M / 2) + r, c4 * 4);
                        sts128(ah + off, h);
                        sts128(al + off, l);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ea_ready[e]);
                if (half == 1) {                       // the second-half warp issues the item once both halves are stored
                    mbar_wait(&ea_ready[e], (n >> 1) & 1u);
                    tc_fence_after_sync();
                    const uint32_t d = tmem_base + b * (uint32_t)FP;
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const uint32_t ko = kk * 32;
                        mma_tf32(d, make_desc(ah + ko), make_desc(wh + ko), idesc, kk > 0 ? 1u : 0u);
                        mma_tf32(d, make_desc(al + ko), make_desc(wh + ko), idesc, 1u);
                        mma_tf32(d, make_desc(ah + ko), make_desc(wl + ko), idesc, 1u);
                    }
                    mma_commit(&ea_empty[e]);
                    mma_commit(&phi_full[b]);
                }
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- compute warps
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        const int c0 = grp * CPT;                      // first feature of this thread's strip
        const bool relu = (p.relu & 1) != 0;
        const bool affine = p.scale != nullptr;
        const bool al8 = ((p.f | p.ld) & 7) == 0;
        uint8_t* xw = x_base + (size_t)warp * WSTAGE;                  // + stage * G_NCW * WSTAGE
        const uint32_t swz_l = ((uint32_t)lane / (8 / CH)) & (CH - 1); // chunk swizzle of this thread's row
        const int c_row = lane / CH, c_ch = lane % CH;                 // copy role: row within an instruction, chunk
        float s1d[CPT / 8], s2d[CPT / 8];              // running (S1, S2) of column c0 + 8*jj + (lane >> 2): per-CTA
#pragma unroll                                         // share of one column (<= 64 tiles x 32 rows) in fp32, doubles after
        for (int i = 0; i < CPT / 8; ++i) s1d[i] = s2d[i] = 0.f;

        auto tile_of = [&](uint32_t tc) { return (int64_t)blockIdx.x + (int64_t)tc * gridDim.x; };
        auto load_nbr = [&](uint32_t tc) {
            int4 nb = make_int4(-1, -1, -1, -1);
            if (tc < n_my) {
                const int64_t t = tile_of(tc) * G_M + row;
                if (t < p.n_rows) nb = ldg4i_pinned(p.nbr + (size_t)t * 4);
            }
            return nb;
        };
        // ---- x ring: cp.async of item (tile count tc, slot k) into stage `st`
        auto x_issue = [&](uint32_t tc, int k, const int (&nbv)[4], uint32_t st) {
            if (tc < n_my) {
                uint8_t* dst = xw + (size_t)st * G_NCW * WSTAGE;
                const int64_t t0 = tile_of(tc) * G_M + q * 32;
                const int fcol = c0 + c_ch * 4;
                const int mine = k < 4 ? nbv[k < 4 ? k : 0] : 0;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int r = i * RPI + c_row;     // row of the quadrant handled by this lane
                    int64_t sr = -1;
                    const float* base = p.x;
                    if (k < 4) {
                        sr = __shfl_sync(0xffffffffu, mine, r);
                    } else if (MODE == 1 && k == 4) {
                        base = p.addend;
                        if (base != nullptr && t0 + r < p.n_add_rows && t0 + r < p.n_rows) sr = t0 + r;
                    } else if (MODE == 1) {
                        base = p.z_prev;
                        if (base != nullptr && t0 + r < p.n_rows) sr = t0 + r;
                    }
                    if (sr >= 0 && fcol < p.f) {
                        const uint32_t sw = ((uint32_t)r / (8 / CH)) & (CH - 1);
                        cp_async16(dst + r * PITCH + (((uint32_t)c_ch ^ sw) << 4), base + (size_t)sr * p.ld + fcol);
                    }
                }
            }
            cp_async_commit();
        };

        // ---- output: a warp's [32 rows x CPT] block, staged row by row in its own x stage, leaves as full row segments
        // (CH lanes x 16 bytes per row and instruction).  One 32-byte sector per lane and instruction - every lane another
        // row - measured 2 us per tile (10.0 -> 8.0 us forward without any store): 2 048 separate requests per tile.
        auto flush_rows = [&](const uint8_t* stg, int64_t t0) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int r = i * RPI + c_row;
                const uint32_t sw = ((uint32_t)r / (8 / CH)) & (CH - 1);
                const float4 v = lds128(smem_u32(stg) + (uint32_t)r * PITCH + ((uint32_t)c_ch << 4));
                const int f = c0 + (int)(((uint32_t)c_ch ^ sw) << 2);
#ifndef DGNN_G_NOSTORE   // timing experiment: no output stores (no results)
                if (t0 + r < p.n_rows && f < p.f) *reinterpret_cast<float4*>(p.out + (size_t)(t0 + r) * p.ld + f) = v;
#endif
            }
        };

        // ---- prologue
        int4 nb4 = load_nbr(0);
        int nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
        uint32_t xi = 0;                               // x items consumed so far (stage = xi & (NST - 1))
#pragma unroll
        for (int j = 0; j < NST - 1; ++j) x_issue(0, j, nbv, (uint32_t)j);
        uint32_t pn = 0;                               // PHI items consumed so far
        for (uint32_t tc = 0; tc < n_my; ++tc) {
            const int64_t t = tile_of(tc) * G_M + row;
            const bool tv = t < p.n_rows;
            const int4 nbn4 = load_nbr(tc + 1);        // neighbours of the next tile (its first copy is issued in this one)
            const int nbn[4] = {nbn4.x, nbn4.y, nbn4.z, nbn4.w};
#ifdef DGNN_L2_PREFETCH        // measured on B200: slower (10.9 vs 10.4 us forward, 14.8 vs 12.2 us backward): the loads are not DRAM-latency bound
            if (tc + G_PF_AHEAD < n_my) {              // this CTA's tile G_PF_AHEAD tiles on: DRAM -> L2 now
                const int64_t r0 = tile_of(tc + G_PF_AHEAD) * G_M;
                const int64_t left = p.n_rows - r0;
                const size_t rows = (size_t)(left < G_M ? left : G_M);
                l2_prefetch_range(p.ea + (size_t)r0 * 4 * p.fe, rows * 4 * p.fe * 4, tid, G_THREADS);
                l2_prefetch_range(p.nbr + (size_t)r0 * 4, rows * 16, tid, G_THREADS);
                if (p.ld == p.f) {                     // full-width rows are contiguous
                    l2_prefetch_range(p.x + (size_t)r0 * p.ld, rows * p.ld * 4, tid, G_THREADS);
                    if (MODE == 1 && p.z_prev != nullptr) l2_prefetch_range(p.z_prev + (size_t)r0 * p.ld, rows * p.ld * 4, tid, G_THREADS);
                    if (MODE == 1 && p.addend != nullptr && r0 + (int64_t)rows <= p.n_add_rows)
                        l2_prefetch_range(p.addend + (size_t)r0 * p.ld, rows * p.ld * 4, tid, G_THREADS);
                }
            }
#endif
            const int cnt = (nbv[0] >= 0) + (nbv[1] >= 0) + (nbv[2] >= 0) + (nbv[3] >= 0);
            const float rcnt = cnt > 1 ? (cnt == 2 ? 0.5f : (cnt == 3 ? (1.0f / 3.0f) : 0.25f)) : 1.0f;   // 1 / max(cnt, 1)
            float acc[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i) acc[i] = 0.f;
#pragma unroll
            for (int k = 0; k < IPT; ++k, ++xi) {
                // the copy of the item NST - 1 ahead goes in flight (its stage was released by the __syncwarp at the end of the
                // last item).  Two stages = one item ahead is all that fits at F = 128; at F <= 64 the items are short (the
                // arithmetic of an item takes less than the round trip of its rows) and four stages keep three in flight.
                auto issue_next = [&]() {
                    constexpr int D = NST - 1;
                    if (k + D < IPT) x_issue(tc, k + D, nbv, (xi + D) & (NST - 1));
                    else x_issue(tc + 1, k + D - IPT, nbn, (xi + D) & (NST - 1));
                };
#ifdef DGNN_XISSUE_LATE      // variant: copy issued after the operand stores' fence.proxy.async (measured: no gain, MODE 1 slower)
                cp_async_wait<NST - 2>();
#else
                issue_next();
                cp_async_wait<NST - 1>();
#endif
                __syncwarp();
                const uint8_t* xs = xw + (size_t)(xi & (NST - 1)) * G_NCW * WSTAGE + lane * PITCH;
#ifdef DGNN_XISSUE_LATE
                if (k >= 4) issue_next();
#endif
                if (k < 4) {
                    const uint32_t b = pn & 3u, bu = pn >> 2;
                    mbar_wait(&phi_full[b], bu & 1);
                    tc_fence_after_sync();
#ifdef DGNN_XISSUE_LATE
                    issue_next();
#endif
                    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + b * (uint32_t)FP + (uint32_t)c0;
                    const bool valid = nbv[k] >= 0;
#pragma unroll
                    for (int j = 0; j < CPT; j += 8) {
                        const int f0 = c0 + j;
                        uint32_t ph[8];
                        tmem_ld8(trow + (uint32_t)j, ph);
                        float4 xa = lds128(smem_u32(xs) + ((((uint32_t)(j >> 2)) ^ swz_l) << 4));
                        float4 xb = lds128(smem_u32(xs) + ((((uint32_t)(j >> 2) + 1u) ^ swz_l) << 4));
                        if (MODE == 0 && affine) {               // warp-uniform addresses: shared-memory broadcasts
                            const uint32_t as = smem_u32(aff_s) + (uint32_t)f0 * 4u;
                            const float4 sa = lds128(as), sb = lds128(as + 16u);
                            const float4 ha = lds128(as + FP * 4u), hb = lds128(as + FP * 4u + 16u);
                            act8(xa, xb, sa, sb, ha, hb, true, relu);
                        } else if (MODE == 0) {
                            act8(xa, xb, xa, xa, xa, xa, false, relu);
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (valid) {
                            acc[j + 0] = fmaf(xa.x, __uint_as_float(ph[0]), acc[j + 0]);
                            acc[j + 1] = fmaf(xa.y, __uint_as_float(ph[1]), acc[j + 1]);
                            acc[j + 2] = fmaf(xa.z, __uint_as_float(ph[2]), acc[j + 2]);
                            acc[j + 3] = fmaf(xa.w, __uint_as_float(ph[3]), acc[j + 3]);
                            acc[j + 4] = fmaf(xb.x, __uint_as_float(ph[4]), acc[j + 4]);
                            acc[j + 5] = fmaf(xb.y, __uint_as_float(ph[5]), acc[j + 5]);
                            acc[j + 6] = fmaf(xb.z, __uint_as_float(ph[6]), acc[j + 6]);
                            acc[j + 7] = fmaf(xb.w, __uint_as_float(ph[7]), acc[j + 7]);
                        }
                    }
                    tc_fence_before_sync();             // orders these TMEM reads before the buffer is handed back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&phi_free[b]);
                    ++pn;
                    if (MODE == 0 && k == 3) {
                        // agg strip of this thread's row -> the warp's own (consumed) x stage, then out row-contiguous
#pragma unroll
                        for (int j = 0; j < CPT; j += 4)
                            sts128(smem_u32(xs) + ((((uint32_t)(j >> 2)) ^ swz_l) << 4),
                                   make_float4(acc[j] * rcnt, acc[j + 1] * rcnt, acc[j + 2] * rcnt, acc[j + 3] * rcnt));
                        flush_rows(xw + (size_t)(xi & (NST - 1)) * G_NCW * WSTAGE, tile_of(tc) * G_M + q * 32);
                    }
                } else if (MODE == 1 && k == 4) {
                    if (p.addend != nullptr && tv && t < p.n_add_rows) {
#pragma unroll
                        for (int j = 0; j < CPT; j += 4) {
                            if (c0 + j < p.f) {
                                const float4 a = lds128(smem_u32(xs) + ((((uint32_t)(j >> 2)) ^ swz_l) << 4));
                                acc[j] += a.x; acc[j + 1] += a.y; acc[j + 2] += a.z; acc[j + 3] += a.w;
                            }
                        }
                    }
                } else if (MODE == 1) {
#pragma unroll
                    for (int j = 0; j < CPT; j += 8) {
                        const int f0 = c0 + j;
                        const bool fvalid = f0 < p.f, fvalid_b = f0 + 4 < p.f;
                        float a8[8], dx[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { a8[i] = 0.f; dx[i] = 0.f; }
                        if (tv && fvalid) {
                            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) a8[i] = (i < 4 || fvalid_b) ? acc[j + i] : 0.f;
                            if (p.z_prev != nullptr) {
                                const float4 za = lds128(smem_u32(xs) + ((((uint32_t)(j >> 2)) ^ swz_l) << 4));
                                const float4 zb = fvalid_b ? lds128(smem_u32(xs) + ((((uint32_t)(j >> 2) + 1u) ^ swz_l) << 4)) : zero4;
                                const float zv[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
                                // ReLU mask of the producer layer; S2 is accumulated as sum(dh * z) and turned into
                                // sum(dh * xhat) = rstd * (sum(dh * z) - mean * S1) when the CTA partial is written
                                float sc[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, sh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                if (p.p_scale != nullptr) {
                                    float4 a = ldg4(p.p_scale + f0), bq = fvalid_b ? ldg4(p.p_scale + f0 + 4) : zero4;
                                    float4 c = ldg4(p.p_shift + f0), d = fvalid_b ? ldg4(p.p_shift + f0 + 4) : zero4;
                                    sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = bq.x; sc[5] = bq.y; sc[6] = bq.z; sc[7] = bq.w;
                                    sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
                                }
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    if (p.p_relu && !(fmaf(zv[i], sc[i], sh[i]) > 0.f)) a8[i] = 0.f;
                                    if (i >= 4 && !fvalid_b) a8[i] = 0.f;
                                    dx[i] = a8[i] * zv[i];
                                }
                            }
                        }
                        if (p.out != nullptr) {     // dy_prev chunk over the z chunk just read (same thread, same place)
                            sts128(smem_u32(xs) + ((((uint32_t)(j >> 2)) ^ swz_l) << 4), make_float4(a8[0], a8[1], a8[2], a8[3]));
                            sts128(smem_u32(xs) + ((((uint32_t)(j >> 2) + 1u) ^ swz_l) << 4), make_float4(a8[4], a8[5], a8[6], a8[7]));
                        }
                        if (p.s_partials != nullptr) {
                            s1d[j >> 3] += warp_colsum8(a8, lane);
                            s2d[j >> 3] += warp_colsum8(dx, lane);
                        }
                    }
                    if (p.out != nullptr) flush_rows(xw + (size_t)(xi & (NST - 1)) * G_NCW * WSTAGE, tile_of(tc) * G_M + q * 32);
                }
                __syncwarp();                          // every lane is done with this item's stage before it is refilled
            }
            nbv[0] = nbn[0]; nbv[1] = nbn[1]; nbv[2] = nbn[2]; nbv[3] = nbn[3];
        }
        cp_async_wait<0>();
        __syncwarp();
        // per-warp (S1, S2) partials into the warp's own (now idle) x-ring stage: doubles [2][CPT]
        if (MODE == 1 && p.s_partials != nullptr && (lane & 3) == 0) {
            double* mine = reinterpret_cast<double*>(xw);
#pragma unroll
            for (int jj = 0; jj < CPT / 8; ++jj) {
                mine[jj * 8 + (lane >> 2)] = (double)s1d[jj];
                mine[CPT + jj * 8 + (lane >> 2)] = (double)s2d[jj];
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (MODE == 1 && p.s_partials != nullptr) {
        double* my = p.s_partials + (size_t)blockIdx.x * 2 * p.s_ld;
        for (int c = tid; c < p.f; c += NT) {
            const int g = c / CPT, cl = c % CPT;
            double a = 0.0, b2 = 0.0;
            for (int qq = 0; qq < 4; ++qq) {           // the four row quadrants of feature group g, fixed order
                const double* w = reinterpret_cast<const double*>(x_base + (size_t)(g * 4 + qq) * WSTAGE);
                a += w[cl];
                b2 += w[CPT + cl];
            }
            my[c] = a;
            const double mu = p.p_mean != nullptr ? (double)__ldg(p.p_mean + c) : 0.0;
            const double rs = p.p_rstd != nullptr ? (double)__ldg(p.p_rstd + c) : 1.0;
            my[p.s_ld + c] = rs * (b2 - mu * a);          // sum(dh * xhat), xhat = (z - mean) * rstd
        }
    }
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// dW_e / db_e = P^T . EA over all edges, P[edge, f] = h(s)[f] * d_agg[onbr[s,k]][f], 3xTF32 (P and EA both split).
// (A row-block formulation like dw2_tc_kernel - bulk-copied z / EA stages, dedicated gather warps filling a ring of
// [32 x F] blocks, transposition by addressing - measured 16 - 23 us per 128 cells against 11.3 us here, independent of F:
// four operand hand-offs per 32-cell stage through the tensor core and back bound it, not the bytes.  Not kept.)
// Per row quarter q one operand stage (P^T [128 features x 32 cells], hi then lo | EA^T hi | EA^T lo); the strips a warp
// needs (its rows' z_prev strip, then the four gathered d_agg strips) arrive through the same warp-private
// cp.async ring as in gather_tc_kernel, one item ahead, so no load is waited for in registers.
template <int CPT, int NST>    // NST: x-ring stages (2 at F = 128, 4 below: see gather_tc_kernel)
__global__ void __launch_bounds__(G_THREADS, 1) dwe_tc_kernel(const GatherTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t p_empty[4];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t p_count[4];                      // counts the warps of the quarter that have stored their rows of the stage
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int CH = CPT / 4, RPI = 32 / CH, PITCH = CPT * 4, WSTAGE = 32 * PITCH;
    uint8_t* x_base = smem + (size_t)4 * G_P_BYTES;      // [2 stages][G_NCW warps][32 rows x PITCH]
    if (tid == 0) {
        for (int qq = 0; qq < 4; ++qq) { mbar_init(&p_empty[qq], 1); mbar_init(&p_count[qq], 4); }
        fence_barrier_init();
    }
    for (int i = tid; i < 4 * G_P_BYTES / 16; i += G_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    // row fe of every EA^T operand is all ones: column fe of the accumulator collects db_e = sum dphi
    for (int i = tid; i < 4 * 32; i += G_THREADS)
        *reinterpret_cast<float*>(smem + (size_t)(i >> 5) * G_P_BYTES + G_ATOM + atom_off(p.fe, i & 31)) = 1.0f;
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(&tmem_slot, 256);        // one [128 x 64] accumulator (P . EA_hi | P . EA_lo) per row quarter
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t n_tiles = (p.n_rows + G_M - 1) / G_M;
    const uint32_t n_my = (int64_t)blockIdx.x < n_tiles ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;

    {
        const uint32_t idesc = make_idesc_tf32(G_M, DWE_ROUNDS == 1 ? 64 : 32);
        // row quarter q = warp >> 2: the four warps of a quarter sit on four different schedulers, so a quarter that waits
        // for its MMAs does not idle a whole scheduler (TMEM is only touched in the read-out below)
        const int q = warp >> 2, grp = warp & 3;
        const int row = q * 32 + lane;
        const int c0 = grp * CPT;
        uint8_t* xw = x_base + (size_t)warp * WSTAGE;
        const uint32_t swz_l = ((uint32_t)lane / (8 / CH)) & (CH - 1);
        const int c_row = lane / CH, c_ch = lane % CH;
        uint8_t* pq = smem + (size_t)q * G_P_BYTES;
        uint8_t* eh = pq + G_ATOM;
        uint8_t* el = eh + 32 * 128;
        auto tile_of = [&](uint32_t tc) { return (int64_t)blockIdx.x + (int64_t)tc * gridDim.x; };
        auto load_nbr = [&](uint32_t tc) {
            int4 nb = make_int4(-1, -1, -1, -1);
            if (tc < n_my) {
                const int64_t t = tile_of(tc) * G_M + row;
                if (t < p.n_rows) nb = ldg4i_pinned(p.nbr + (size_t)t * 4);
            }
            return nb;
        };
        // item m of a tile: m = 0 the rows' own z_prev strip, m = 1..4 the d_agg strip of neighbour m-1
        auto x_issue = [&](uint32_t tc, int m, const int (&nbv)[4], uint32_t st) {
            if (tc < n_my) {
                uint8_t* dst = xw + (size_t)st * G_NCW * WSTAGE;
                const int64_t t0 = tile_of(tc) * G_M + q * 32;
                const int fcol = c0 + c_ch * 4;
                const int mine = m > 0 ? nbv[m > 0 ? m - 1 : 0] : 0;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int r = i * RPI + c_row;
                    int64_t sr = -1;
                    const float* base = p.x;
                    if (m > 0) {
                        sr = __shfl_sync(0xffffffffu, mine, r);
                    } else {
                        base = p.z_prev;
                        if (t0 + r < p.n_rows) sr = t0 + r;
                    }
                    if (sr >= 0 && fcol < p.f) {
                        const uint32_t sw = ((uint32_t)r / (8 / CH)) & (CH - 1);
                        cp_async16(dst + r * PITCH + (((uint32_t)c_ch ^ sw) << 4), base + (size_t)sr * p.ld + fcol);
                    }
                }
            }
            cp_async_commit();
        };
        int4 nb4 = load_nbr(0);
        int nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
        uint32_t xi = 0, it = 0;
        static_assert(NST >= 2 && (NST & (NST - 1)) == 0 && NST - 1 <= 5, "x ring: 2 or 4 stages");
#pragma unroll
        for (int j = 0; j < NST - 1; ++j) x_issue(0, j, nbv, (uint32_t)j);
        for (uint32_t tc = 0; tc < n_my; ++tc) {
            const int64_t t = tile_of(tc) * G_M + row;
            const bool tv = t < p.n_rows;
            const int4 nbn4 = load_nbr(tc + 1);
            const int nbn[4] = {nbn4.x, nbn4.y, nbn4.z, nbn4.w};
#ifdef DGNN_L2_PREFETCH        // measured on B200: slower (10.9 vs 10.4 us forward, 14.8 vs 12.2 us backward): the loads are not DRAM-latency bound
            if (tc + G_PF_AHEAD < n_my) {              // this CTA's tile G_PF_AHEAD tiles on: DRAM -> L2 now
                const int64_t r0 = tile_of(tc + G_PF_AHEAD) * G_M;
                const int64_t left = p.n_rows - r0;
                const size_t rows = (size_t)(left < G_M ? left : G_M);
                l2_prefetch_range(p.ea + (size_t)r0 * 4 * p.fe, rows * 4 * p.fe * 4, tid, G_THREADS);
                l2_prefetch_range(p.nbr + (size_t)r0 * 4, rows * 16, tid, G_THREADS);
                if (p.ld == p.f) {
                    l2_prefetch_range(p.x + (size_t)r0 * p.ld, rows * p.ld * 4, tid, G_THREADS);
                    l2_prefetch_range(p.z_prev + (size_t)r0 * p.ld, rows * p.ld * 4, tid, G_THREADS);
                }
            }
#endif
            float h[CPT];
            float4 ev[2];
            auto load_ev = [&](int k) {                // this warp's share of the EA row of slot k (global, one item ahead)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int e0 = grp * 8 + h2 * 4;
                    ev[h2] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e0 < p.fe && tv && nbv[k] >= 0) ev[h2] = ldg4_pinned(p.ea + ((size_t)t * 4 + k) * p.fe + e0);
                }
            };
#pragma unroll
            for (int m = 0; m < 5; ++m, ++xi) {
                {
                    constexpr int D = NST - 1;             // copies run D items ahead of their use
                    if (m + D < 5) x_issue(tc, m + D, nbv, (xi + D) & (NST - 1));
                    else x_issue(tc + 1, m + D - 5, nbn, (xi + D) & (NST - 1));
                }
                cp_async_wait<NST - 1>();
                __syncwarp();
                const uint8_t* xs = xw + (size_t)(xi & (NST - 1)) * G_NCW * WSTAGE + lane * PITCH;
                if (m == 0) {
                    load_ev(0);
#pragma unroll
                    for (int j = 0; j < CPT; j += 4) {
                        const int f0 = c0 + j;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (tv && f0 < p.f) {
                            v = lds128(smem_u32(xs) + ((((uint32_t)(j >> 2)) ^ swz_l) << 4));
                            if (p.p_scale != nullptr) {
                                float4 sc = ldg4(p.p_scale + f0), sh = ldg4(p.p_shift + f0);
                                v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                            }
                            if (p.p_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        }
                        h[j] = v.x; h[j + 1] = v.y; h[j + 2] = v.z; h[j + 3] = v.w;
                    }
                } else {
                    const int k = m - 1;
                    const bool valid = tv && nbv[k] >= 0;
                    mbar_wait(&p_empty[q], (it & 1) ^ 1);
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
                            sts32(smem_u32(eh) + off, hi);
                            sts32(smem_u32(el) + off, lo);
                        }
                    }
                    if (k < 3) load_ev(k + 1);
                    // Round 0 stores P_hi = rna_tf32(P) (+ EA^T above) and accumulates P_hi . (EA_hi + EA_lo).  DWE_ROUNDS = 2 adds a
                    // second round over the SAME operand buffer with P_lo = P - P_hi and P_lo . EA_hi (full 3xTF32).  Measured on
                    // B200 (tools/diag_grad_wide.py, eth widths): the gradient error of lin_e against the fp64 oracle is the same
                    // with and without it (median 9.7e-4 vs 9.9e-4 of the RMS: ReLU-mask flips dominate, the unbiased 2^-12
                    // rounding of P does not show), while the round costs 0.27 ms per 128-wide layer => one round.
#pragma unroll
                    for (int rnd = 0; rnd < DWE_ROUNDS; ++rnd) {
                        if (rnd == 1) mbar_wait(&p_empty[q], (it & 1) ^ 1);
#pragma unroll
                        for (int j = 0; j < CPT; j += 4) {
                            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid && c0 + j < p.f) d = lds128(smem_u32(xs) + ((((uint32_t)(j >> 2)) ^ swz_l) << 4));
                            const float dp[4] = {h[j] * d.x, h[j + 1] * d.y, h[j + 2] * d.z, h[j + 3] * d.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                // row = c0 + j + i, column = lane; c0 is a multiple of 8: row & 7 = (j & 4) + i
                                const uint32_t r8 = (uint32_t)((j & 4) + i);
                                const uint32_t off = (uint32_t)(c0 + j + i) * 128u + ((((uint32_t)lane >> 2) ^ r8) << 4) + (((uint32_t)lane & 3u) << 2);
                                const float hi = tf32_rna(dp[i]);
                                sts32(smem_u32(pq) + off, rnd == 0 ? hi : dp[i] - hi);
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (mbar_arrive_pending(&p_count[q]) == 1u) {   // last warp of the quarter: accumulate P^T . EA
                                mbar_wait(&p_count[q], it & 1);  // completed by this very arrival: acquires the other warps' stores
                                tc_fence_after_sync();
                                // every quarter accumulates into its own TMEM columns, its rounds in order: the sum over
                                // cells is evaluated in a fixed order whatever the timing of the warps
                                const uint32_t ph = smem_u32(pq), ehh = ph + G_ATOM, ell = ehh + 32 * 128;
                                const uint32_t dq = tmem_base + (uint32_t)(q * 64);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    const uint32_t ko = kk * 32;
                                    if (DWE_ROUNDS == 1) {
                                        // EA^T lo lies right behind EA^T hi: ONE MMA of N = 64 gives P . EA_hi | P . EA_lo in the
                                        // two column halves (4 instructions per stage instead of 8; the halves are added at the end)
                                        mma_tf32(dq, make_desc(ph + ko), make_desc(ehh + ko), idesc, (it > 0 || kk > 0) ? 1u : 0u);
                                    } else {
                                        mma_tf32(dq, make_desc(ph + ko), make_desc(ehh + ko), idesc, (it > 0 || kk > 0) ? 1u : 0u);
                                        if (rnd == 0) mma_tf32(dq, make_desc(ph + ko), make_desc(ell + ko), idesc, 1u);
                                    }
                                }
                                (void)ell;
                                mma_commit(&p_empty[q]);
                            }
                        }
                        __syncwarp();
                        ++it;
                    }
                }
                __syncwarp();
            }
            nbv[0] = nbn[0]; nbv[1] = nbn[1]; nbv[2] = nbn[2]; nbv[3] = nbn[3];
        }
        cp_async_wait<0>();
        // The quarters run independently; only after every warp has issued its last stage is each p_empty barrier at
        // most one phase behind, which makes the parity wait below unambiguous.
        __syncthreads();
        if (warp < 4) {                                  // warp w reads TMEM lanes (= feature rows) 32 w .. 32 w + 31
            const int row = warp * 32 + lane;
            float* outp = p.dwe_partials + (size_t)blockIdx.x * p.s_ld * 32;
            float v[32];
            if (it > 0) {
                // every stage's MMAs were issued and committed by some thread of its quarter; the last stage of each
                // quarter is the only one nobody has waited for yet
                for (int qq = 0; qq < 4; ++qq) mbar_wait(&p_empty[qq], (it - 1) & 1);
                tc_fence_after_sync();
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), v);       // quarter 0 (hi, then lo half), then 1, 2, 3 in order
#pragma unroll 1
                for (int hq = 1; hq < 8; ++hq) {
                    if (DWE_ROUNDS != 1 && (hq & 1)) continue;                 // two rounds: 32-column accumulators in the even slots
                    float w[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(hq * 32), w);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += w[i];
                }
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
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


}  // namespace dgnn

using namespace dgnn;

extern "C" int dgnn_gather_tc_supported(int f, int fe) {
    return (f % 4 == 0 && f >= 4 && fe % 4 == 0 && fe >= 4 && fe <= 28) ? 1 : 0;   // any width: 128-feature slices
}

constexpr int G_FSLICE = 128;   // features per launch: 4 PHI buffers x 128 columns = the 512 TMEM columns

static int fp_of(int f) { return f <= 32 ? 32 : (f <= 64 ? 64 : 128); }

template <int MODE>
static int launch_gather_tc(const GatherTcArgs& p, cudaStream_t st, const char* what) {
    const int cpt = p.fp / 4;
    const int nst = cpt == 32 ? 2 : G_NST_NARROW;  // x-ring stages: what fits at F = 128, deeper below
    size_t smem = (size_t)2 * p.fp * 128 + (size_t)G_EA_STAGES * 2 * G_ATOM + (size_t)nst * G_NCW * 32 * p.fp + (size_t)8 * p.fp + 1024;
#define LAUNCH_G(CPT, NW, NST)                                                                                 \
    do {                                                                                                       \
        if (int rc_ = ensure_dyn_smem((const void*)gather_tc_kernel<CPT, MODE, NW, NST>, 226 * 1024, what)) return rc_; \
        gather_tc_kernel<CPT, MODE, NW, NST><<<sm_count(), (NW + G_EAW) * 32, smem, st>>>(p);                  \
    } while (0)
    switch (cpt) {
        case 8: LAUNCH_G(8, 16, G_NST_NARROW); break;
        case 16: LAUNCH_G(16, 16, G_NST_NARROW); break;
        // (a 32-warp variant, 16 features per thread and 64 registers, measured 12.1 vs 10.5 us forward and 13.8 vs 12.1 us
        // backward at F = 128 on B200)
        case 32: LAUNCH_G(32, 16, 2); break;
        default: return fail(what, "unsupported feature width");
    }
#undef LAUNCH_G
    return check_launch(what);
}

static int launch_dwe_tc(const GatherTcArgs& p, cudaStream_t st, const char* what) {
    const int cpt = p.fp / 4;
    const int nst = cpt == 32 ? 2 : G_NST_NARROW;
    size_t smem = (size_t)4 * G_P_BYTES + (size_t)nst * G_NCW * 32 * p.fp + 1024;
#define LAUNCH_D(CPT, NST)                                                                                     \
    do {                                                                                                       \
        if (int rc_ = ensure_dyn_smem((const void*)dwe_tc_kernel<CPT, NST>, 226 * 1024, what)) return rc_;        \
        dwe_tc_kernel<CPT, NST><<<sm_count(), G_THREADS, smem, st>>>(p);                                       \
    } while (0)
    switch (cpt) {
        case 8: LAUNCH_D(8, G_NST_NARROW); break;
        case 16: LAUNCH_D(16, G_NST_NARROW); break;
        case 32: LAUNCH_D(32, 2); break;
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
    p.relu = relu_in; p.nbr = nbr; p.ea = ea; p.fe = fe;
    p.n_rows = n_tgt; p.ld = f_in; p.s_ld = f_in;
    for (int f0 = 0; f0 < f_in; f0 += G_FSLICE) {
        const int w = f_in - f0 < G_FSLICE ? f_in - f0 : G_FSLICE;
        p.f = w; p.fp = fp_of(w);
        p.x = x_in + f0; p.out = agg + f0;
        p.scale = in_scale ? in_scale + f0 : nullptr; p.shift = in_shift ? in_shift + f0 : nullptr;
        p.w_e = w_e + (size_t)f0 * fe; p.b_e = b_e + f0;
        if (int rc = launch_gather_tc<0>(p, as_stream(stream), "dgnn_gather_tc_fwd")) return rc;
    }
    return 0;
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
    p.nbr = onbr; p.ea = ea_own; p.fe = fe;
    p.n_rows = n_src; p.ld = f_in; p.s_ld = f_in;
    p.n_add_rows = n_tgt; p.p_relu = p_relu;
    cudaStream_t st = as_stream(stream);
    for (int f0 = 0; f0 < f_in; f0 += G_FSLICE) {
        const int w = f_in - f0 < G_FSLICE ? f_in - f0 : G_FSLICE;
        p.f = w; p.fp = fp_of(w);
        p.x = d_agg + f0; p.out = dy_prev ? dy_prev + f0 : nullptr;
        p.w_e = w_e + (size_t)f0 * fe; p.b_e = b_e + f0;
        p.addend = d_self ? d_self + f0 : nullptr;
        p.z_prev = z_prev ? z_prev + f0 : nullptr;
        p.p_scale = p_scale ? p_scale + f0 : nullptr; p.p_shift = p_shift ? p_shift + f0 : nullptr;
        p.p_mean = p_mean ? p_mean + f0 : nullptr; p.p_rstd = p_rstd ? p_rstd + f0 : nullptr;
        p.s_partials = s_partials ? s_partials + f0 : nullptr;
        p.dwe_partials = dwe_partials ? dwe_partials + (size_t)f0 * 32 : nullptr;
        if (dy_prev != nullptr || s_partials != nullptr)
            if (int rc = launch_gather_tc<1>(p, st, "dgnn_gather_tc_bwd")) return rc;
        if (dwe_partials != nullptr)
            if (int rc = launch_dwe_tc(p, st, "dgnn_gather_tc_bwd(dW_e)")) return rc;
    }
    return 0;
}
