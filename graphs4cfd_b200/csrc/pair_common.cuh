// pair_common.cuh — pieces shared by the CTA-pair tensor-core kernels (mp_edge_pair.cu, mp_row_pair.cu):
// thread-role constants, register re-partitioning, 256-bit global stores.
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
// registers per thread after setmaxnreg: the CTA's pool is what it was launched with, 896 * 72 = 64512
// = 512*80 (epilogue) + 256*64 (loaders) + 128*56 (MMA issuer + idle warps)
constexpr int kRegsEpi = 80, kRegsLoad = 64, kRegsMisc = 56;
constexpr int kMaxSmem = 232448;        // 227 KiB opt-in limit

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


}  // namespace pairk
}  // namespace g4c
