// tile_mlp.cuh — fp32 (CUDA-core FFMA) fused row-tile MLP engine.
//
// A CTA of 256 threads owns a tile of TM rows.  Activations live in shared memory, row-major
// [TM][K+4]; every Linear is a register-tiled GEMM (4 rows x 8 columns per thread) whose weight
// operand W_t[K][H] is streamed from L2 in 8-row chunks with double-buffered cp.async; the layer
// output goes back to the same shared buffer (bias + SELU fused), the last layer stays in
// registers for LayerNorm (row statistics by warp shuffles) and the caller's epilogue.
#pragma once
#include "common.cuh"

namespace g4c {

template <int H>
struct Cfg {
    static_assert(H >= 16 && H <= 256 && (H & (H - 1)) == 0, "hidden width must be 16..256, power of two");
    static constexpr int NT = 256;
    static constexpr int TX = H / 8;          // threads across the H output columns
    static constexpr int TY = NT / TX;        // threads across rows
    static constexpr int TM = 4 * TY;         // rows per tile
    static constexpr int LD = H + 4;          // padded row stride of an H-wide smem buffer
    static constexpr int KC = 8;              // weight rows per cp.async chunk
    static constexpr int WST = 2 * KC * H;    // floats of weight staging (2 stages)
    static constexpr int SMALL_LD = 12;       // row stride of the narrow-segment buffer (K <= 8)
};

template <int H>
__device__ __forceinline__ int frag_col(int tx, int c) {
    return (c < 4) ? tx * 4 + c : H / 2 + tx * 4 + (c - 4);
}

template <int H>
__device__ __forceinline__ void load_w_chunk(float* dst, const float* __restrict__ Wt, int k0, int kc, int tid) {
    const float4* src = reinterpret_cast<const float4*>(Wt + (size_t)k0 * H);
    const int n4 = kc * H / 4;
    for (int i = tid; i < n4; i += Cfg<H>::NT) cp_async16(dst + i * 4, src + i);
}

// acc[4][8] += A[TM][K] (smem, row stride lda) x Wt[K][H] (global).  Ends with a __syncthreads().
template <int H>
__device__ __forceinline__ void gemm_seg(float (&acc)[4][8], const float* A, int lda, int K,
                                         const float* __restrict__ Wt, float* wst, int tid) {
    using C = Cfg<H>;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const float* a_row = A + (size_t)(ty * 4) * lda;
    const int nch = (K + C::KC - 1) / C::KC;
    load_w_chunk<H>(wst, Wt, 0, min(C::KC, K), tid);
    cp_async_commit();
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) {
            load_w_chunk<H>(wst + ((c + 1) & 1) * C::KC * H, Wt, (c + 1) * C::KC, min(C::KC, K - (c + 1) * C::KC), tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* w = wst + (c & 1) * C::KC * H + tx * 4;
        const int k0 = c * C::KC;
        const int kc = min(C::KC, K - k0);
        if (kc == C::KC) {
#pragma unroll
            for (int kk = 0; kk < C::KC; kk += 4) {
                float4 a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a_row + i * lda + k0 + kk);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w0 = *reinterpret_cast<const float4*>(w + (kk + q) * H);
                    const float4 w1 = *reinterpret_cast<const float4*>(w + (kk + q) * H + H / 2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
                        acc[i][0] = fmaf(av, w0.x, acc[i][0]);
                        acc[i][1] = fmaf(av, w0.y, acc[i][1]);
                        acc[i][2] = fmaf(av, w0.z, acc[i][2]);
                        acc[i][3] = fmaf(av, w0.w, acc[i][3]);
                        acc[i][4] = fmaf(av, w1.x, acc[i][4]);
                        acc[i][5] = fmaf(av, w1.y, acc[i][5]);
                        acc[i][6] = fmaf(av, w1.z, acc[i][6]);
                        acc[i][7] = fmaf(av, w1.w, acc[i][7]);
                    }
                }
            }
        } else {
            for (int kk = 0; kk < kc; ++kk) {
                const float4 w0 = *reinterpret_cast<const float4*>(w + kk * H);
                const float4 w1 = *reinterpret_cast<const float4*>(w + kk * H + H / 2);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float av = a_row[i * lda + k0 + kk];
                    acc[i][0] = fmaf(av, w0.x, acc[i][0]);
                    acc[i][1] = fmaf(av, w0.y, acc[i][1]);
                    acc[i][2] = fmaf(av, w0.z, acc[i][2]);
                    acc[i][3] = fmaf(av, w0.w, acc[i][3]);
                    acc[i][4] = fmaf(av, w1.x, acc[i][4]);
                    acc[i][5] = fmaf(av, w1.y, acc[i][5]);
                    acc[i][6] = fmaf(av, w1.z, acc[i][6]);
                    acc[i][7] = fmaf(av, w1.w, acc[i][7]);
                }
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
}

// X[m][col] = selu(acc + bias) for this thread's fragment; then __syncthreads().
template <int H>
__device__ __forceinline__ void store_hidden(const float (&acc)[4][8], const float* __restrict__ bias,
                                             float* X, int tid) {
    using C = Cfg<H>;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const float4 b0 = *reinterpret_cast<const float4*>(bias + tx * 4);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + H / 2 + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* row = X + (size_t)(ty * 4 + i) * C::LD;
        float4 o0 = make_float4(selu(acc[i][0] + b0.x), selu(acc[i][1] + b0.y), selu(acc[i][2] + b0.z), selu(acc[i][3] + b0.w));
        float4 o1 = make_float4(selu(acc[i][4] + b1.x), selu(acc[i][5] + b1.y), selu(acc[i][6] + b1.z), selu(acc[i][7] + b1.w));
        *reinterpret_cast<float4*>(row + tx * 4) = o0;
        *reinterpret_cast<float4*>(row + H / 2 + tx * 4) = o1;
    }
    __syncthreads();
}

// acc <- LayerNorm(acc + bias) (or just acc + bias when gamma == nullptr), rows spread over TX lanes.
template <int H>
__device__ __forceinline__ void finalize_rows(float (&acc)[4][8], const float* __restrict__ bias,
                                              const float* __restrict__ gamma, const float* __restrict__ beta, int tid) {
    using C = Cfg<H>;
    const int tx = tid % C::TX;
    const float4 b0 = *reinterpret_cast<const float4*>(bias + tx * 4);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + H / 2 + tx * 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] += bb[c];
    if (gamma == nullptr) return;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + tx * 4);
    const float4 g1 = *reinterpret_cast<const float4*>(gamma + H / 2 + tx * 4);
    const float4 e0 = *reinterpret_cast<const float4*>(beta + tx * 4);
    const float4 e1 = *reinterpret_cast<const float4*>(beta + H / 2 + tx * 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) s += acc[i][c];
#pragma unroll
        for (int off = C::TX / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float mean = s * (1.f / H);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            acc[i][c] -= mean;
            q = fmaf(acc[i][c], acc[i][c], q);
        }
#pragma unroll
        for (int off = C::TX / 2; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
        const float rstd = 1.f / sqrtf(q * (1.f / H) + kLnEps);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(acc[i][c] * rstd, gg[c], ee[c]);
    }
}

// Run linear_2..linear_L of an MLP whose linear_1 pre-activation is already in `acc`.
// X is the [TM][LD] chain buffer.  On return acc holds the (LayerNorm-ed) output of the last layer,
// valid only when mlp.out_width == H.  When out_width < 16 the last hidden activation is left in X.
template <int H>
__device__ __forceinline__ void chain_tail(float (&acc)[4][8], const G4cMlp& mlp, float* X, float* wst, int tid) {
    using C = Cfg<H>;
    const bool narrow = mlp.out_width != H;
    const int n_wide = narrow ? mlp.n_layers - 1 : mlp.n_layers;   // layers computed as H-wide GEMMs
    for (int l = 1; l < n_wide; ++l) {
        store_hidden<H>(acc, mlp.b[l - 1], X, tid);
        zero_acc(acc);
        gemm_seg<H>(acc, X, C::LD, H, mlp.W_t[l], wst, tid);
    }
    if (narrow) {
        store_hidden<H>(acc, mlp.b[n_wide - 1], X, tid);
    } else {
        finalize_rows<H>(acc, mlp.b[n_wide - 1], mlp.ln_gamma, mlp.ln_beta, tid);
    }
}

// Cooperative tile load: rows[m] (>= 0) selects the source row of `src` (row stride `stride`),
// rows[m] < 0 zero-fills.  width % 4 == 0 path is vectorised.
template <int H>
__device__ __forceinline__ void load_tile_wide(float* X, const float* __restrict__ src, int stride,
                                               const int* rows, float scale, int tid) {
    using C = Cfg<H>;
    constexpr int V = H / 4;
    for (int idx = tid; idx < C::TM * V; idx += C::NT) {
        const int m = idx / V, c4 = idx % V;
        const int r = rows[m];
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= 0) {
            x = ldg_stream(src + (size_t)r * stride + c4 * 4);
            if (scale != 1.f) { x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale; }
        }
        *reinterpret_cast<float4*>(X + (size_t)m * C::LD + c4 * 4) = x;
    }
}

}  // namespace g4c
