// FP32 FMA tile GEMM shared by the forward / backward dense kernels.
//
//   C[TM x TN chunk] = A_s[TM x K] (shared memory, row-major, leading dim lda)
//                    x W[K x N]    (global, row-major, streamed in [TK x TN] chunks through a
//                                   cp.async double buffer)
//
// 256 threads; thread (ty = tid/16, tx = tid%16) owns rows ty*4..+3 and columns
// n0 + tx*4..+3 and n0 + 64 + tx*4..+3.  This is the generic-width path; the tensor-core
// (tcgen05) path replaces it where the widths allow.
#pragma once
#include "common.cuh"

namespace dgnn {

constexpr int TM = 64;    // rows (cells) per tile
constexpr int NT = 256;   // threads per CTA
constexpr int TK = 16;    // k rows per staged weight chunk
constexpr int TN = 128;   // columns per pass

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}

// stage W[k0..k0+TK) x [n0..n0+TN) into w_s (TK*TN floats), zero-filling out-of-range parts
__device__ __forceinline__ void stage_w_chunk(float* w_s, const float* __restrict__ W, int K, int N,
                                              int k0, int n0) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < (TK * TN / 4) / NT; ++it) {
        int idx = tid + it * NT;
        int r = idx / (TN / 4);
        int c4 = idx % (TN / 4);
        int k = k0 + r, n = n0 + c4 * 4;
        bool ok = (k < K) && (n < N);
        const float* src = ok ? (W + (size_t)k * N + n) : W;
        cp_async16_zfill(w_s + r * TN + c4 * 4, src, ok);
    }
}

// acc[4][8] = A_s x W[:, n0..n0+TN).  kp = K rounded up to TK (A_s columns K..kp are zero).
// w_s: 2*TK*TN floats.  All threads must call; ends with the weight buffers free.
__device__ __forceinline__ void tile_gemm(float (&acc)[4][8], const float* __restrict__ a_s, int lda,
                                          const float* __restrict__ W, int K, int N, int n0, float* w_s) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int nchunks = (K + TK - 1) / TK;
    stage_w_chunk(w_s, W, K, N, 0, n0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        float* cur = w_s + (c & 1) * (TK * TN);
        if (c + 1 < nchunks) {
            stage_w_chunk(w_s + ((c + 1) & 1) * (TK * TN), W, K, N, (c + 1) * TK, n0);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* ap = a_s + (ty * 4) * lda + c * TK;
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(ap + i * lda + kk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 b0 = *reinterpret_cast<const float4*>(cur + (kk + j) * TN + tx * 4);
                float4 b1 = *reinterpret_cast<const float4*>(cur + (kk + j) * TN + 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float av = j == 0 ? a[i].x : (j == 1 ? a[i].y : (j == 2 ? a[i].z : a[i].w));
                    acc[i][0] = fmaf(av, b0.x, acc[i][0]);
                    acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                    acc[i][2] = fmaf(av, b0.z, acc[i][2]);
                    acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                    acc[i][4] = fmaf(av, b1.x, acc[i][4]);
                    acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                    acc[i][6] = fmaf(av, b1.z, acc[i][6]);
                    acc[i][7] = fmaf(av, b1.w, acc[i][7]);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace dgnn
