// pair_common.cuh — pieces shared by the CTA-pair tensor-core kernels (mp_edge_pair.cu, mp_row_pair.cu):
// thread-role constants, register re-partitioning, the hidden-layer epilogue, 256-bit global stores.
#pragma once
#include "tc2_core.cuh"
#include "mp_pair.h"

namespace g4c {
namespace pairk {

using namespace tc2;

constexpr int H = 128;
constexpr int IMG = 128 * 128;          // bytes of a 128-row image
constexpr int HIMG = 64 * 128;          // bytes of a 64-row image (one CTA's half of a weight K-block)
constexpr int NT = 896;                 // warps 0-15 epilogue, 16-23 loaders, 24 MMA issuer, 25-27 idle
constexpr int N_EPI_WARPS = 16, W_LOAD0 = 16, N_LOAD_WARPS = 8, W_MMA = 24;
constexpr int NEPI = 512;
// registers per thread after setmaxnreg: the CTA's pool is what it was launched with, 896 * 72 = 64512
// = 512*80 (epilogue) + 256*64 (loaders) + 128*56 (MMA issuer + idle warps)
constexpr int kRegsEpi = 80, kRegsLoad = 64, kRegsMisc = 56;
constexpr int kMaxSmem = 232448;        // 227 KiB opt-in limit

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// 256-bit row store that does not allocate in L1: the (small, 227 KiB of it being shared memory) L1 is left to the
// index loads and the few spilled registers; with allocating stores the edge kernel ran 7 % slower
__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// ---------------------------------------------------------------------------------------------- epilogues
// hidden layer: x = selu(acc*inv (+ bias)); written as fp16 (hi, lo) A operand columns of this thread's row
__device__ __forceinline__ void epilogue_hidden(uint32_t d_addr, uint32_t ah_addr, uint32_t al_addr, float inv,
                                                const float* __restrict__ bias) {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tmem_ld32f(d_addr + c0, v);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + c0 + i));
            const float x0 = selu_fast(fmaf(v[i], inv, b.x)), x1 = selu_fast(fmaf(v[i + 1], inv, b.y));
            const float x2 = selu_fast(fmaf(v[i + 2], inv, b.z)), x3 = selu_fast(fmaf(v[i + 3], inv, b.w));
            split2(x0, x1, hi[i / 2], lo[i / 2]);
            split2(x2, x3, hi[i / 2 + 1], lo[i / 2 + 1]);
        }
        tmem_st16(ah_addr + c0 / 2, hi);
        tmem_st16(al_addr + c0 / 2, lo);
    }
    tmem_wait_st();
}


}  // namespace pairk
}  // namespace g4c
