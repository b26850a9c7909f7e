// tc2_test.cu — hardware self tests of the tc2_core.cuh primitives (g4c_debug_tc2):
//   test 1: D = A W^T, cta_group::1, A operand written to TMEM by tcgen05.st (registers -> TMEM)
//   test 2: D = P + A W^T, cta_group::1, A and the accumulator's initial value P delivered by
//           tcgen05.cp from swizzled shared-memory images (flags&1: no MMA, D = P round trip)
//   test 3: D[256,128] = A W^T on a CTA pair (cluster of 2, cta_group::2, M = 256, B rows split between
//           the two CTAs, multicast commit, remote mbarrier arrive); flags&1: A delivered by tcgen05.cp
// All use the 3-term fp16 split (hi*hi + lo*hi + hi*lo) exactly like the production kernels.
#include "tc2_core.cuh"

namespace g4c {
namespace tc2 {

constexpr int IMG = 128 * 128;          // bytes of a 128-row image
constexpr int HIMG = 64 * 128;          // bytes of a 64-row image (half of B on a CTA pair)
constexpr int NCOMPUTE = 256;
constexpr int NTHREADS = NCOMPUTE + 32;

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory"); }

// warp-cooperative: stage fp32 row `src_row` (128 floats) as (hi, lo) fp16 into images
// img[0]=hi k0..63, img[1]=hi k64..127, img[2]=lo k0..63, img[3]=lo k64..127 at image row r
__device__ __forceinline__ void stage_row_split(uint8_t* img, const float* __restrict__ src_row, int r, int lane) {
    const float4 x = *reinterpret_cast<const float4*>(src_row + lane * 4);
    uint2 h, l;
    split2(x.x, x.y, h.x, l.x);
    split2(x.z, x.w, h.y, l.y);
    const int kb = lane >> 4, chunk = (lane & 15) >> 1, sub = (lane & 1) * 8;
    const uint32_t off = (uint32_t)kb * IMG + img_off(r, chunk) + sub;
    *reinterpret_cast<uint2*>(img + off) = h;
    *reinterpret_cast<uint2*>(img + 2 * IMG + off) = l;
}
// warp-cooperative: stage fp32 row (128 floats, scaled) into four 32-column fp32 images at image row r
__device__ __forceinline__ void stage_row_f32(uint8_t* img, const float* __restrict__ src_row, float scale, int r, int lane) {
    float4 x = *reinterpret_cast<const float4*>(src_row + lane * 4);
    x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
    *reinterpret_cast<float4*>(img + (uint32_t)(lane >> 3) * IMG + img_off(r, lane & 7)) = x;
}

// thread (row, half): split 64 fp32 values into TMEM A columns (hi at a_hi + 32*half, lo at a_lo + 32*half)
__device__ __forceinline__ void st_split64(uint32_t a_hi, uint32_t a_lo, const float* __restrict__ x64) {
    uint32_t h[32], l[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) split2(x64[2 * i], x64[2 * i + 1], h[i], l[i]);
    tmem_st32(a_hi, h);
    tmem_st32(a_lo, l);
}

struct TestSmem1 {
    uint8_t w[2][2 * IMG];               // per K-block: hi image | lo image (128 rows each)
    uint8_t img_a[4][IMG];
    uint8_t img_p[4][IMG];
    uint64_t w_full, d_full;
    uint32_t tmem_base;
};

// tests 1 and 2
__global__ void __launch_bounds__(NTHREADS, 1)
ts_test_kernel(const float* A, const uint8_t* Wpack, float inv_scale, const float* P, float* D, int test, int flags) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    TestSmem1& s = *reinterpret_cast<TestSmem1*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&s.w_full, 1);
        mbar_init(&s.d_full, 1);
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc<1>(&s.tmem_base, 256); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const uint32_t D_COL = 0, AH_COL = 128, AL_COL = 192;

    if (warp == 8) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&s.w_full, 4 * IMG);
            bulk_g2s(s.w[0], Wpack, 4 * IMG, &s.w_full);
        }
    } else {
        const int row = (warp & 3) * 32 + lane, half = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        if (test == 1) {
            float x[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) x[i] = A[(size_t)row * 128 + half * 64 + i];
            st_split64(tmem + lane_base + AH_COL + 32 * half, tmem + lane_base + AL_COL + 32 * half, x);
            tmem_wait_st();
            tc_fence_before();
        } else {
            for (int i = 0; i < 16; ++i) {
                const int r = warp * 16 + i;
                stage_row_split(s.img_a[0], A + (size_t)r * 128, r, lane);
                stage_row_f32(s.img_p[0], P + (size_t)r * 128, 1.f / inv_scale, r, lane);
            }
            fence_proxy_async();
        }
        compute_sync();
        if (tid == 0) {
            tc_fence_after();
            mbar_wait(&s.w_full, 0);
            bool init = false;
            if (test == 2) {
                for (int q = 0; q < 4; ++q)
                    for (int j = 0; j < 4; ++j)
                        tmem_cp_128x256b<1>(tmem + D_COL + 32 * q + 8 * j, make_desc_sw128(smem_u32(s.img_p[q]) + 32 * j));
                init = true;
                for (int kb = 0; kb < 2; ++kb)
                    for (int j = 0; j < 4; ++j) {
                        tmem_cp_128x256b<1>(tmem + AH_COL + 32 * kb + 8 * j, make_desc_sw128(smem_u32(s.img_a[kb]) + 32 * j));
                        tmem_cp_128x256b<1>(tmem + AL_COL + 32 * kb + 8 * j, make_desc_sw128(smem_u32(s.img_a[2 + kb]) + 32 * j));
                    }
            }
            if (!(test == 2 && (flags & 1))) {
                const uint32_t idesc = idesc_f16(128, 128);
                for (int kb = 0; kb < 2; ++kb)
                    for (int j = 0; j < 4; ++j) {
                        const int ks = kb * 4 + j;
                        const uint64_t wh = make_desc_sw128(smem_u32(s.w[kb]) + 32 * j);
                        const uint64_t wl = make_desc_sw128(smem_u32(s.w[kb]) + IMG + 32 * j);
                        umma_ts<1>(tmem + D_COL, tmem + AH_COL + 8 * ks, wh, idesc, (init || ks > 0) ? 1u : 0u);
                        umma_ts<1>(tmem + D_COL, tmem + AL_COL + 8 * ks, wh, idesc, 1u);
                        umma_ts<1>(tmem + D_COL, tmem + AH_COL + 8 * ks, wl, idesc, 1u);
                    }
            }
            umma_commit<1>(&s.d_full);
        }
        mbar_wait(&s.d_full, 0);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + lane_base + D_COL + (uint32_t)(half * 64 + c0), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) D[(size_t)row * 128 + half * 64 + c0 + i] = v[i] * inv_scale;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc<1>(tmem, 256);
}

struct TestSmem3 {
    uint8_t w[2][2 * HIMG];              // per K-block: hi image | lo image, 64 rows each (this CTA's half of W's rows)
    uint8_t img_a[4][IMG];
    uint64_t w_full, a_ready, d_full, bench_done;
    uint32_t tmem_base;
};

// test 3: CTA pair
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
pair_test_kernel(const float* A, const uint8_t* Wpair, float inv_scale, float* D, int flags) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    TestSmem3& s = *reinterpret_cast<TestSmem3*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    if (tid == 0) {
        mbar_init(&s.w_full, 1);
        mbar_init(&s.a_ready, 16);       // 8 compute warps of each CTA
        mbar_init(&s.d_full, 1);
        mbar_init(&s.bench_done, 1);
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc<2>(&s.tmem_base, 512); tmem_relinquish<2>(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const uint32_t D_COL = 0, AH_COL = 128, AL_COL = 192;
    const bool use_cp = (flags & 1) != 0;

    if (warp == 8) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&s.w_full, 4 * HIMG);
            bulk_g2s(s.w[0], Wpair + (size_t)rank * 4 * HIMG, 4 * HIMG, &s.w_full);
            if (rank == 0) {
                mbar_wait<true>(&s.a_ready, 0);
                tc_fence_after();
                if (use_cp) {
                    for (int kb = 0; kb < 2; ++kb)
                        for (int j = 0; j < 4; ++j) {
                            tmem_cp_128x256b<2>(tmem + AH_COL + 32 * kb + 8 * j, make_desc_sw128(smem_u32(s.img_a[kb]) + 32 * j));
                            tmem_cp_128x256b<2>(tmem + AL_COL + 32 * kb + 8 * j, make_desc_sw128(smem_u32(s.img_a[2 + kb]) + 32 * j));
                        }
                }
                const uint32_t idesc = idesc_f16(256, 128);
                for (int kb = 0; kb < 2; ++kb)
                    for (int j = 0; j < 4; ++j) {
                        const int ks = kb * 4 + j;
                        const uint64_t wh = make_desc_sw128(smem_u32(s.w[kb]) + 32 * j);
                        const uint64_t wl = make_desc_sw128(smem_u32(s.w[kb]) + HIMG + 32 * j);
                        umma_ts<2>(tmem + D_COL, tmem + AH_COL + 8 * ks, wh, idesc, ks > 0 ? 1u : 0u);
                        umma_ts<2>(tmem + D_COL, tmem + AL_COL + 8 * ks, wh, idesc, 1u);
                        umma_ts<2>(tmem + D_COL, tmem + AH_COL + 8 * ks, wl, idesc, 1u);
                    }
                umma_commit<2>(&s.d_full, 3);
            }
        }
        __syncwarp();
    } else {
        const int row = (warp & 3) * 32 + lane, half = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float* Ablk = A + (size_t)rank * 128 * 128;
        if (!use_cp) {
            float x[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) x[i] = Ablk[(size_t)row * 128 + half * 64 + i];
            st_split64(tmem + lane_base + AH_COL + 32 * half, tmem + lane_base + AL_COL + 32 * half, x);
            tmem_wait_st();
            tc_fence_before();
        } else {
            for (int i = 0; i < 16; ++i) {
                const int r = warp * 16 + i;
                stage_row_split(s.img_a[0], Ablk + (size_t)r * 128, r, lane);
            }
            fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) {
            mbar_wait(&s.w_full, 0);                       // this CTA's half of B has landed
            mbar_arrive_cluster(mapa(smem_u32(&s.a_ready), 0));
        }
        mbar_wait(&s.d_full, 0);
        tc_fence_after();
        float* Dblk = D + (size_t)rank * 128 * 128;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + lane_base + D_COL + (uint32_t)(half * 64 + c0), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) Dblk[(size_t)row * 128 + half * 64 + c0 + i] = v[i] * inv_scale;
        }
        tc_fence_before();
    }
    tc_fence_before();
    cluster_sync_all();
    // ---- flags bit 2: MMA issue-rate microbenchmark.  After the functional test, the issuer thread repeats the
    // 24-MMA layer (flags >> 8) times and reports clock cycles per MMA in D[0] (bit 3: the MMAs alternate between two
    // accumulators instead of chaining on one: separates the dependent-accumulate latency from the pipe rate)
    // bits 4 / 5 / 6: meanwhile the eight compute warps of both CTAs hammer TMEM with tcgen05.ld / tcgen05.st (columns
    // 384.., not used by the MMAs) or shared memory with 128-bit loads: what the epilogue and loader warps of the real
    // kernels do next to the tensor pipe
    if ((flags & 4) && warp < 8 && (flags & (16 | 32 | 64))) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int iters = (flags >> 8) * 24 * 64 / 40;         // roughly as long as the MMA loop
        float acc = 0.f;
        uint32_t z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0;
        for (int it = 0; it < iters; ++it) {
            if (flags & 16) {
                float v[32];
                tmem_ld32(tmem + lane_base + 384u + (uint32_t)((warp >> 2) * 32), v);
                acc += v[it & 31];
            }
            if (flags & 32) {
                tmem_st32(tmem + lane_base + 384u + (uint32_t)((warp >> 2) * 32), z);
                tmem_wait_st();
            }
            if (flags & 64) {
                const float4 q = *reinterpret_cast<const float4*>(s.img_a[it & 3] + ((lane * 144 + it * 16) & (IMG - 16)));
                acc += q.x + q.w;
            }
        }
        if (acc == 123.456f) D[1] = acc;
    }
    if ((flags & 4) && warp == 8 && lane == 0 && rank == 0) {
        const int reps = flags >> 8;
        const bool two_acc = (flags & 8) != 0;
        const uint32_t idesc = idesc_f16(256, 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int kb = 0; kb < 2; ++kb)
                for (int j = 0; j < 4; ++j) {
                    const int ks = kb * 4 + j;
                    const uint64_t wh = make_desc_sw128(smem_u32(s.w[kb]) + 32 * j);
                    const uint64_t wl = make_desc_sw128(smem_u32(s.w[kb]) + HIMG + 32 * j);
                    umma_ts<2>(tmem + D_COL, tmem + AH_COL + 8 * ks, wh, idesc, 1u);
                    umma_ts<2>(tmem + (two_acc ? 256u : D_COL), tmem + AL_COL + 8 * ks, wh, idesc, 1u);
                    umma_ts<2>(tmem + D_COL, tmem + AH_COL + 8 * ks, wl, idesc, 1u);
                }
        umma_commit<2>(&s.bench_done, 1);
        mbar_wait(&s.bench_done, 0);
        const long long t1 = clock64();
        D[0] = (float)(t1 - t0) / (float)(24 * reps);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 8) tmem_dealloc<2>(tmem, 512);
}

}  // namespace tc2

int tc2_test_launch(int test, const float* A, const void* Wpack, float inv_scale, const float* P, float* D, int flags, cudaStream_t st) {
    if (test == 1 || test == 2) {
        const int smem = (int)sizeof(tc2::TestSmem1);
        if (cudaFuncSetAttribute(tc2::ts_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return check_launch("ts_test_kernel attribute");
        tc2::ts_test_kernel<<<1, tc2::NTHREADS, smem, st>>>(A, static_cast<const uint8_t*>(Wpack), inv_scale, P, D, test, flags);
        count_launch();
        return check_launch("ts_test_kernel");
    }
    if (test == 3) {
        const int smem = (int)sizeof(tc2::TestSmem3);
        if (cudaFuncSetAttribute(tc2::pair_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return check_launch("pair_test_kernel attribute");
        tc2::pair_test_kernel<<<2, tc2::NTHREADS, smem, st>>>(A, static_cast<const uint8_t*>(Wpack), inv_scale, D, flags);
        count_launch();
        return check_launch("pair_test_kernel");
    }
    set_error("g4c_debug_tc2: unknown test %d", test);
    return G4C_EINVAL;
}

}  // namespace g4c
