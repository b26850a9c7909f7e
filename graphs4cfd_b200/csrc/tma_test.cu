// tma_test.cu — hardware self tests of the bulk-tensor (TMA) assumptions mp_edge_v5.cu is built on, one warp each:
//   test 0  3-D tile load, box 16 x 1 x 32 of the [rows, k, 128] view, SWIZZLE_64B: where the bytes land in shared memory
//           (read back with the loader's lane = row formula)
//   test 1  3-D tile store, box 8 x 1 x 32, SWIZZLE_32B, from a tile staged with the epilogue's formula
//   test 3  test 0 with the tile hanging over the end of the tensor: out-of-bounds rows are zero-filled and still count
//           towards the mbarrier's transaction bytes
// (tests/test_gpu_tma_primitives.py)
#include <cuda.h>
#include "tc2_core.cuh"
#include "mp_pair.h"

namespace g4c {
namespace ep5 {
bool encode_rows(CUtensorMap* m, const float* base, int64_t rows, int k, int box_cols, int swizzle_bytes, int box_rows);
}
namespace tmat {

using namespace tc2;

struct Maps { CUtensorMap in3, out3; };

__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (!ok && ++spins > (1u << 16)) __trap();
    } while (!ok);
}

// c0, j, n0: tile origin (column, in-edge slot, first target row)
__global__ void __launch_bounds__(32, 1) tma_test_kernel(int test, const __grid_constant__ Maps tm, const float* src,
                                                          float* out, int c0, int j, int n0) {
    __shared__ __align__(1024) uint8_t tile[2048];
    __shared__ uint64_t bar;
    const int lane = threadIdx.x;
    const uint32_t t = smem_u32(tile), b = smem_u32(&bar);
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    if (test == 0 || test == 3) {
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(2048) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(t), "l"(&tm.in3), "r"(c0), "r"(j), "r"(n0), "r"(b) : "memory");
        }
        wait_bar(b, 0);
    }
    if (test == 0 || test == 3) {
        // the loader's read-back: lane = row, 64-byte pitch, 16-byte chunk c at c ^ ((lane >> 1) & 3)
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 v = lds_f4(t + lane * 64 + ((c ^ swz) << 4));
            reinterpret_cast<float4*>(out + lane * 16)[c] = v;
        }
    } else if (test == 1) {
        // the epilogue's staging: lane = row, 32-byte pitch, chunk c at c ^ ((lane >> 2) & 1); values = src[lane * 8 + 0..7]
        const uint32_t o0 = t + (uint32_t)lane * 32u + ((uint32_t)((lane >> 2) & 1) << 4);
        const float4 v0 = reinterpret_cast<const float4*>(src + lane * 8)[0], v1 = reinterpret_cast<const float4*>(src + lane * 8)[1];
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(o0), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(o0 ^ 16u), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(&tm.out3), "r"(c0), "r"(j), "r"(n0), "r"(t) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    }
}

}  // namespace tmat

// src: [rows * k, 128]; out: [32, 16] (tests 0, 3) or [rows * k, 128] (test 1)
int tma_test_launch(int test, const float* src, int64_t rows, int k, float* out, int c0, int j, int n0, cudaStream_t st) {
    tmat::Maps tm;
    const bool ok = ep5::encode_rows(&tm.in3, src, rows, k, 16, 64, 32) && ep5::encode_rows(&tm.out3, test == 1 ? out : src, rows, k, 8, 32, 32);
    if (!ok) { set_error("g4c_debug_tma: cuTensorMapEncodeTiled failed"); return G4C_ECUDA; }
    tmat::tma_test_kernel<<<1, 32, 0, st>>>(test, tm, src, out, c0, j, n0);
    count_launch();
    return check_launch("tma_test_kernel");
}

}  // namespace g4c
