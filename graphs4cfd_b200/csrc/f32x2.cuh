// f32x2.cuh — packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2, one issue slot for two fp32 lanes) and the
// helpers of the tensor-core epilogues built on them.  The fused edge / row kernels are bound by instruction issue,
// not by the fp32 pipe (profiles/r2a_edge_v3_instruction_mix.txt), so every elementwise stage works on register pairs.
// All packed operations round to nearest-even exactly like their scalar forms.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace g4c {
namespace p2 {

__device__ __forceinline__ uint64_t pk(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t bc(float a) { return pk(a, a); }
__device__ __forceinline__ void upk(uint64_t r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ void upk_u(uint64_t r, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0,%1}, %2;" : "=r"(a), "=r"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2(float t) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return e;
}

// (a, b) = hi + lo with hi = fp16(x), lo = fp16(x - hi): the operand split of tc_core.cuh's split2 with the
// subtraction done on the pair (F2FP, 2 x HADD2.F32, FADD2, F2FP)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    float r0, r1;
    upk(sub2(pk(a, b), pk(hf.x, hf.y)), r0, r1);
    const __half2 l = __floats2half2_rn(r0, r1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace p2
}  // namespace g4c
