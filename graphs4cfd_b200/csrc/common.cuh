// common.cuh — shared device helpers for libg4c (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/g4c.h"

namespace g4c {

constexpr float kSeluAlpha = 1.6732632423543772848170429916717f;
constexpr float kSeluScale = 1.0507009873554804934193349852946f;
constexpr float kLnEps = 1e-5f;

// torch.nn.functional.selu / torch.tanh on fp32 (ATen: scale*(x>0 ? x : alpha*(exp(x)-1)))
__device__ __forceinline__ float selu(float x) {
    return x > 0.f ? kSeluScale * x : (kSeluScale * kSeluAlpha) * (expf(x) - 1.f);
}
__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == G4C_ACT_SELU) return selu(x);
    if (act == G4C_ACT_TANH) return tanhf(x);
    return x;
}

// branch-free SELU on the SFU: exp(x) = ex2.approx(x*log2(e)) (2 ulp); absolute error of the negative branch
// ~2e-7, the same order as the cancellation in the reference's fp32 `exp(x) - 1`.
__device__ __forceinline__ float selu_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    const float neg = fmaf(kSeluScale * kSeluAlpha, e, -(kSeluScale * kSeluAlpha));
    return x > 0.f ? kSeluScale * x : neg;
}
__device__ __forceinline__ float apply_act_fast(float x, int act) {
    if (act == G4C_ACT_SELU) return selu_fast(x);
    if (act == G4C_ACT_TANH) return tanhf(x);
    return x;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// streaming 16B load that does not pollute L1 (row data is consumed once per CTA)
__device__ __forceinline__ float4 ldg_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

void set_error(const char* fmt, ...);
void count_launch(int n = 1, bool tensor_core = false);      // tensor_core: a tcgen05 kernel (g4c_tc_launch_count)
int check_launch(const char* what);

// ---- per-DEVICE launch configuration.  cudaFuncSetAttribute and the SM count belong to the current device, and a process may
// drive several (a model on cuda:1 while cuda:0 is current elsewhere): remember what was configured per device ordinal.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int device_sms() {
    static int n_sm[kMaxDevices] = {0};
    const int dev = current_device();
    if (n_sm[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        n_sm[dev] = n > 0 ? n : 148;
    }
    return n_sm[dev];
}
// opt in to `smem` bytes of dynamic shared memory for `kernel` on the current device (once per device and size); `state` is the
// caller's static per-kernel table.  Returns false when the runtime refuses.
template <typename K>
inline bool ensure_dynamic_smem(K kernel, int smem, int (&state)[kMaxDevices]) {
    const int dev = current_device();
    if (smem > state[dev]) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
        state[dev] = smem;
    }
    return true;
}

}  // namespace g4c
