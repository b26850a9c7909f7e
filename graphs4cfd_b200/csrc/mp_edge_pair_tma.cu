// mp_edge_pair_tma.cu — edge_pair_kernel (mp_edge_pair.cu, "v3") with bulk-tensor (TMA) data paths.  "v4".
//
// STATUS: EXPERIMENTAL and OPT-IN (G4C_EDGE_MODE=1..4 or g4c_debug_set_edge_mode): written after the round's GPU budget
// was spent; it compiles for sm_100a but has NOT run on hardware yet.  The default path is the v3 kernel, untouched.
//
// Why (profiles/r1d_edge_pair_v3_phases.txt): v3 is limited by the SM's load/store pipe, not by HBM
// or the tensor pipe.  Per 128-edge slot the pipe handles 384 LDGSTS (8 row pieces each), 384 lane = row LDS.128 and
// 64 STG.256 that each touch 32 different 128-byte lines (one line per pass), about 7.5k cycles against the 6.7k-cycle
// HBM bound (tools/edge_kernel_model.py), and the epilogue's stores queue behind the loaders' copies.  This version takes
// the regular streams off that pipe:
//   mode 1  e' leaves through shared memory and the TMA engine: each epilogue warp stages 32 rows x 8 columns (1 KiB,
//           SWIZZLE_32B so that lane = row STS.128 are bank-conflict free) and one elected lane issues
//           cp.async.bulk.tensor.3d.global.shared::cta (box 8 x 1 x 32 of the [N, k, 128] view of e').  The 24 KiB
//           this needs come from storing the loaders' row pieces unpadded with a manual XOR swizzle (64-byte pitch,
//           16-byte chunk c of row r at c ^ ((r >> 1) & 3) — the SWIZZLE_64B pattern) instead of an 80-byte pitch.
//   mode 2  additionally the e and P_c[tgt] row pieces of a loader stage arrive as two TMA tiles (box 16 x 1 x 32 of
//           the [N, k, 128] view of e; box 16 x 32 of P_c; SWIZZLE_64B = the ring's layout) signalled on a per-warp,
//           per-stage mbarrier; only the gathered P_r[src] pieces stay on cp.async (4 of the 12 LDGSTS per stage, and
//           none of the e / P_c address arithmetic).
//   mode 3  additionally the gathered P_r[src] pieces arrive through TMA (tile::gather4: four source rows per copy, eight
//           copies per stage issued by lanes 0-7 with their own row coordinates): no LDGSTS at all in the loaders.
//   mode 4  mode 3 with 96 / 48 registers per epilogue / loader thread instead of 88 / 64 (the loaders of mode 3 hold no
//           address arrays any more).
// Restrictions (checked by the launcher, which falls back to v3): fixed in-degree (fixed_k > 0), edges stored in
// aggregation order (no edge_perm / tgt_perm).  That covers the level-1 kNN launches and every REMuS angle level,
// i.e. the launches that dominate the step.  Arithmetic, TMEM layout, MMA issue, hidden epilogues and the LayerNorm
// are those of v3; results are expected to be bitwise identical to v3.
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cuda.h>
#include "tc2_core.cuh"
#include "mp_pair.h"

namespace g4c {
namespace ep4 {

using namespace tc2;

constexpr int H = 128;
constexpr int HIMG = 64 * 128;
constexpr int NT = 896;                  // warps 0-15 epilogue, 16-23 loaders, 24 MMA issuer, 25-27 idle
constexpr int N_EPI_WARPS = 16;
constexpr int N_LOAD_WARPS = 8;
constexpr int W_LOAD0 = 16, W_MMA = 24;
constexpr int kRegsMisc = 24;       // epilogue / loader registers are template parameters of the kernel (88 / 64 as in v3, or 96 / 48)

constexpr int SCOLS = 16;                // columns per loader stage
constexpr int NCS = H / SCOLS;
constexpr int NCS_W = NCS / 2;           // stages per slot per loader warp
constexpr int ARR = 32 * 64;             // one array's 32 row pieces of a stage: 64-byte pitch, XOR swizzled
constexpr int STG = 3 * ARR;             // e | P_r | P_c
constexpr int NSTG = 2;
constexpr int OPIECE = 32 * 32;          // e' staging of one epilogue warp: 32 rows x 8 columns, SWIZZLE_32B

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

struct Maps {
    CUtensorMap e_in;                    // [N, k, 128] fp32, box 16 x 1 x 32, SWIZZLE_64B
    CUtensorMap p_c;                     // [N, 128] fp32, box 16 x 32, SWIZZLE_64B
    CUtensorMap e_out;                   // [N, k, 128] fp32, box 8 x 1 x 32, SWIZZLE_32B
    CUtensorMap p_r;                     // [*, 128] fp32, box 16 x 1 (tile::gather4: four rows per copy), SWIZZLE_64B
};

struct Smem {
    uint8_t w[3][4 * HIMG];                          // 96 KiB, 1024-byte aligned (UMMA SWIZZLE_128B images)
    uint8_t ring[N_LOAD_WARPS][NSTG][STG];           // 96 KiB, every array 2048-byte aligned (SWIZZLE_64B repeats every 512 B)
    uint8_t ostage[N_EPI_WARPS][OPIECE];             // 16 KiB, 1024-byte aligned (SWIZZLE_32B repeats every 256 B)
    float cst[5][H];
    float part[2][2][4][H];
    uint64_t w_full;
    uint64_t in_ready[2];
    uint64_t a_ready[2];
    uint64_t d_free[2];
    uint64_t d_full[2];
    uint64_t ld_full[N_LOAD_WARPS][NSTG];            // mode 2: the two TMA tiles of a loader stage have landed
    uint32_t tmem_base;
};
static_assert(sizeof(Smem) <= 232448, "shared memory exceeds the 227 KiB opt-in limit");
static_assert(offsetof(Smem, ring) % 2048 == 0 && offsetof(Smem, ostage) % 1024 == 0, "TMA tile alignment");

// ---- optional in-kernel phase profile (make EXTRA=-DG4C_PROFILE), same laps as mp_edge_pair.cu; read back with g4c_debug_profile()
#ifdef G4C_PROFILE
__device__ unsigned long long g_prof[64];
#define PROF_DECL unsigned int prof_t0 = 0; unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_START() prof_t0 = clock()
#define PROF_LAP(i) do { const unsigned int t1 = clock(); prof_acc[i] += (unsigned int)(t1 - prof_t0); prof_t0 = t1; } while (0)
#define PROF_FLUSH(base, cond) do { if (blockIdx.x == 0 && lane == 0 && (cond)) for (int i = 0; i < 8; ++i) atomicAdd(&g_prof[(base) + i], prof_acc[i]); } while (0)
#else
#define PROF_DECL
#define PROF_START()
#define PROF_LAP(i)
#define PROF_FLUSH(base, cond)
#endif

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void epi_sync_all() { asm volatile("bar.sync 5, 512;" ::: "memory"); }
__device__ __forceinline__ void quarter_sync(int lq) { asm volatile("bar.sync %0, 128;" ::"r"(1 + lq) : "memory"); }

__device__ __forceinline__ float selu_over_lambda_l2(float t) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    const float neg = fmaf(kSeluAlpha, e, -kSeluAlpha);
    return t > 0.f ? t * kLn2 : neg;
}

// Sleeping mbarrier wait with a tighter watchdog than tc2_core.cuh's (2^24 polls x 20 us): a protocol error in this
// not-yet-run kernel traps after about five seconds instead of minutes.  Hides tc2::mbar_wait_sleep_a inside this namespace.
__device__ __forceinline__ void mbar_wait_sleep_a(uint32_t addr, uint32_t parity) {
    uint32_t spins = 0, ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (!ok && ++spins > (1u << 18)) __trap();
    } while (!ok);
}

// ------------------------------------------------------------------ bulk-tensor copies (one thread issues)
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// four rows r0..r3 of a 2-D tensor, box_cols columns from c0 each, to four consecutive row pieces at dst
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* m, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(m), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, int c0, int c1, int c2, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every bulk group this thread committed has finished READING its shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent one
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// kLoad: 0 = every row piece by cp.async (mode 1), 1 = e / P_c tiles by TMA (mode 2), 2 = P_r[src] by TMA gather4 as well (mode 3)
template <int kLoad, int kRegsEpi = 88, int kRegsLoad = 64>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) edge_tma_kernel(const EdgeArgs a, const __grid_constant__ Maps tm) {
    constexpr bool kTmaLoad = kLoad >= 1, kGather = kLoad == 2;
    static_assert(512 * kRegsEpi + 256 * kRegsLoad + 128 * kRegsMisc <= 896 * 72, "register budget");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int tid = threadIdx.x, warp = tid >> 5;
    int lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int nl = a.n_layers;
    const int k = a.fixed_k;
    const int n_units = (int)((a.n_targets + 127) / 128);
    const int n_up = (n_units + 1) / 2;
    const int up0 = blockIdx.x >> 1, up_stride = gridDim.x >> 1;

    if (tid == 0) {
        mbar_init(&s.w_full, 1);
        for (int c = 0; c < 2; ++c) {
            mbar_init(&s.in_ready[c], 2 * N_LOAD_WARPS);
            mbar_init(&s.a_ready[c], 2 * N_EPI_WARPS);
            mbar_init(&s.d_free[c], N_EPI_WARPS);
            mbar_init(&s.d_full[c], 1);
        }
        for (int w = 0; w < N_LOAD_WARPS; ++w)
            for (int g = 0; g < NSTG; ++g) mbar_init(&s.ld_full[w][g], 1);
        fence_barrier_init();
    }
    if (warp == W_MMA) { tmem_alloc<2>(&s.tmem_base, 512); tmem_relinquish<2>(); }
    if (tid == 0) {
        mbar_arrive_expect_tx(&s.w_full, (uint32_t)nl * 4 * HIMG);
        for (int l = 0; l < nl; ++l) bulk_g2s(s.w[l], a.W[l] + (size_t)rank * 4 * HIMG, 4 * HIMG, &s.w_full);
    }
    if (tid < H) {
        for (int l = 0; l < 3; ++l) {
            float b = 0.f;
            if (l > 0 && l < nl) b = a.bias[l][tid] * (l < nl - 1 ? kLog2e : 1.f);
            s.cst[l][tid] = b;
        }
        s.cst[3][tid] = a.gamma ? a.gamma[tid] : 1.f;
        s.cst[4][tid] = a.beta ? a.beta[tid] : 0.f;
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const uint32_t sb = smem_u32(smem_raw);
    const uint32_t a_cst = sb + (uint32_t)offsetof(Smem, cst), a_part = sb + (uint32_t)offsetof(Smem, part);
    const uint32_t a_in_ready = sb + (uint32_t)offsetof(Smem, in_ready), a_a_ready = sb + (uint32_t)offsetof(Smem, a_ready);
    const uint32_t a_d_free = sb + (uint32_t)offsetof(Smem, d_free), a_d_full = sb + (uint32_t)offsetof(Smem, d_full);

    if (warp < N_EPI_WARPS) {
        // ====================================================================== epilogue warps
        setmaxnreg_inc<kRegsEpi>();
        const int etid = __shfl_sync(0xffffffffu, tid, tid & 31);
        lane = etid & 31;
        const int lq = (etid >> 5) & 3, cq = etid >> 7;
        const int row = lq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
        const uint32_t leader_a_ready0 = mapa(a_a_ready, 0);
        const uint32_t my_cst = a_cst + 128u * cq, my_part = a_part + 4u * row;
        // e' staging of this warp: row = lane at a 32-byte pitch, 16-byte chunk c at c ^ ((lane >> 2) & 1)  (SWIZZLE_32B)
        // (tile base = ost0 & ~1023, the other chunk = ost0 ^ 16: one live register)
        const uint32_t ost0 = sb + (uint32_t)offsetof(Smem, ostage) + (uint32_t)(etid >> 5) * OPIECE + (uint32_t)lane * 32u +
                              ((uint32_t)((lane >> 2) & 1) << 4);
        // lane 0 issues, commits and waits for every bulk store of this warp (bulk groups are per thread).
        // Two-layer MLPs (REMuS angle / edge models) leave the third layer's 32 KiB of weight space unused: there the staging is
        // double buffered (pieces alternate between two 1 KiB tiles per warp, the wait lets one store stay in flight).
        const bool two_buf = nl == 2;
        const uint32_t ost_alt = sb + (uint32_t)offsetof(Smem, w) + 2u * 4u * HIMG + (uint32_t)(etid >> 5) * (2u * OPIECE) +
                                 (uint32_t)lane * 32u + ((uint32_t)((lane >> 2) & 1) << 4);
        uint32_t n_dfull[2] = {0, 0};
        const bool has_ln = a.gamma != nullptr;
        int pbuf = 0;
        PROF_DECL
        PROF_START();

        for (int up = up0; up < n_up; up += up_stride) {
            const int n_unit0 = (up * 2 + (int)rank) * 128;         // n_targets < 2^31 (checked by g4c_edge_aggr_fwd)
            const bool live = n_unit0 + row < a.n_targets;
            float agg[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) agg[i] = 0.f;

            for (int j0 = 0; j0 < k; j0 += 2) {
                const int nch = min(2, k - j0);
                for (int l = 0; l < nl; ++l) {
                    const float cl = a.inv_scale[l] * (l > 0 ? kSeluScale : 1.f);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c >= nch) continue;
                        const uint32_t d_addr = tmem + lane_base + 256u * c + 32u * cq;
                        if (warp == 0) mbar_wait_sleep_a(a_d_full + 8u * c, n_dfull[c] & 1);
                        ++n_dfull[c];
                        epi_sync_all();
                        tc_fence_after();
                        PROF_LAP(l < nl - 1 ? 0 : 1);        // waiting for the MMAs (hidden / last layer)
                        if (l < nl - 1) {
                            const float c2 = cl * kLog2e;
#pragma unroll
                            for (int h16 = 0; h16 < 2; ++h16) {
                                float v[16];
                                tmem_ld16f(d_addr + 16u * h16, v);
                                uint32_t hi[8], lo[8];
                                const uint32_t bs = my_cst + 512u * l + 64u * h16;
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    const float4 b = lds_f4(bs + 4u * i);
                                    const float x0 = selu_over_lambda_l2(fmaf(v[i], c2, b.x));
                                    const float x1 = selu_over_lambda_l2(fmaf(v[i + 1], c2, b.y));
                                    const float x2 = selu_over_lambda_l2(fmaf(v[i + 2], c2, b.z));
                                    const float x3 = selu_over_lambda_l2(fmaf(v[i + 3], c2, b.w));
                                    split2(x0, x1, hi[i / 2], lo[i / 2]);
                                    split2(x2, x3, hi[i / 2 + 1], lo[i / 2 + 1]);
                                }
                                tmem_st8(tmem + lane_base + 256u * c + 128u + 16u * cq + 8u * h16, hi);
                                tmem_st8(tmem + lane_base + 256u * c + 192u + 16u * cq + 8u * h16, lo);
                            }
                            tmem_wait_st();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_remote(leader_a_ready0 + 8u * c);
                            PROF_LAP(2);                     // hidden epilogue
                        } else {
                            const uint32_t bs = my_cst + 512u * l;
                            float y[32];
                            tmem_ld16_nowait(d_addr, y);
                            tmem_ld16_nowait(d_addr + 16u, y + 16);
                            tmem_wait_ld();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_a(a_d_free + 8u * c);
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 b4 = lds_f4(bs + 4u * i);
                                y[i] = fmaf(y[i], cl, b4.x);
                                y[i + 1] = fmaf(y[i + 1], cl, b4.y);
                                y[i + 2] = fmaf(y[i + 2], cl, b4.z);
                                y[i + 3] = fmaf(y[i + 3], cl, b4.w);
                            }
                            float mean = 0.f, rstd = 1.f;
                            if (has_ln) {
                                float mh[2], M2h[2];
#pragma unroll
                                for (int h16 = 0; h16 < 2; ++h16) {
                                    float sum = 0.f;
#pragma unroll
                                    for (int i = 0; i < 16; i += 4)
                                        sum += (y[16 * h16 + i] + y[16 * h16 + i + 1]) + (y[16 * h16 + i + 2] + y[16 * h16 + i + 3]);
                                    mh[h16] = sum * (1.f / 16.f);
                                    float sq = 0.f;
#pragma unroll
                                    for (int i = 0; i < 16; ++i) {
                                        const float dlt = y[16 * h16 + i] - mh[h16];
                                        sq = fmaf(dlt, dlt, sq);
                                    }
                                    M2h[h16] = sq;
                                }
                                const float dm = mh[0] - mh[1];
                                const uint32_t pa = my_part + 4096u * pbuf;
                                sts_f1(pa + 512u * cq, 0.5f * (mh[0] + mh[1]));
                                sts_f1(pa + 2048u + 512u * cq, (M2h[0] + M2h[1]) + 8.f * dm * dm);
                                PROF_LAP(3);                 // last layer: read + statistics
                                quarter_sync(lq);
                                PROF_LAP(4);                 // last layer: barrier
                                const float m0 = lds_f1(pa), m1 = lds_f1(pa + 512u), m2 = lds_f1(pa + 1024u), m3 = lds_f1(pa + 1536u);
                                mean = 0.25f * ((m0 + m1) + (m2 + m3));
                                const float d0 = m0 - mean, d1 = m1 - mean, d2 = m2 - mean, d3 = m3 - mean;
                                const float M2 = ((lds_f1(pa + 2048u) + lds_f1(pa + 2560u)) + (lds_f1(pa + 3072u) + lds_f1(pa + 3584u))) +
                                                 32.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
                                rstd = 1.f / sqrtf(M2 * (1.f / H) + kLnEps);
                                pbuf ^= 1;
                            }
                            const int j = j0 + c;
#pragma unroll
                            for (int i8 = 0; i8 < 32; i8 += 8) {
                                float* o = y + i8;
                                if (has_ln) {
#pragma unroll
                                    for (int u = 0; u < 8; u += 4) {
                                        const float4 g = lds_f4(my_cst + 1536u + 4u * (i8 + u));
                                        const float4 be = lds_f4(my_cst + 2048u + 4u * (i8 + u));
                                        o[u] = fmaf((o[u] - mean) * rstd, g.x, be.x);
                                        o[u + 1] = fmaf((o[u + 1] - mean) * rstd, g.y, be.y);
                                        o[u + 2] = fmaf((o[u + 2] - mean) * rstd, g.z, be.z);
                                        o[u + 3] = fmaf((o[u + 3] - mean) * rstd, g.w, be.w);
                                    }
                                }
                                if (live) {
#pragma unroll
                                    for (int u = 0; u < 8; ++u) agg[i8 + u] += o[u];
                                }
                                if (a.e_out != nullptr) {       // warp-uniform
                                    if (a.act_e_out == G4C_ACT_SELU) {
#pragma unroll
                                        for (int u = 0; u < 8; ++u) o[u] = selu_fast(o[u]);
                                    }
                                    // the staging tile is free once the previous piece's bulk store has read it; everything above
                                    // (normalise, aggregate, SELU of this piece) ran while the TMA engine was reading
                                    const uint32_t ost = two_buf ? ost_alt + (uint32_t)((i8 >> 3) & 1) * OPIECE : ost0;
                                    if (lane == 0) {
                                        if (two_buf) bulk_wait_read1(); else bulk_wait_read0();
                                    }
                                    __syncwarp();
                                    sts_f4(ost, o[0], o[1], o[2], o[3]);
                                    sts_f4(ost ^ 16u, o[4], o[5], o[6], o[7]);
                                    fence_proxy_async();        // generic-proxy writes -> visible to the TMA engine
                                    __syncwarp();
                                    // rows past the last target hold values of no edge; the tensor map's bounds clip them
                                    if (lane == 0) {
                                        tma_store_3d(&tm.e_out, cq * 32 + i8, j, n_unit0 + lq * 32, ost & ~1023u);
                                        bulk_commit();
                                    }
                                }
                            }
                            PROF_LAP(5);                     // last layer: normalise, aggregate, stage + bulk store
                        }
                    }
                }
            }
            if (live) {
                const float rc = (a.aggr == G4C_AGGR_MEAN) ? 1.f / (float)max(k, 1) : 1.f;
                float* dst = a.agg_out + (size_t)(n_unit0 + row) * H + cq * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] = agg[i + u] * rc;
                    stg256(dst + i, o);
                }
            }
            PROF_LAP(6);
        }
        PROF_FLUSH(0, warp == 0);
    } else if (warp < W_LOAD0 + N_LOAD_WARPS) {
        // ====================================================================== loader warps
        setmaxnreg_dec<kRegsLoad>();
        const int lw = (warp - W_LOAD0) & 3, hf = (warp - W_LOAD0) >> 2;
        const float ps = a.p_scale;
        const uint32_t lane_base = (uint32_t)(lw * 32) << 16;
        const uint32_t ring0 = smem_u32(s.ring[warp - W_LOAD0][0]);
        const uint32_t bar0 = smem_u32(&s.ld_full[warp - W_LOAD0][0]);
        const uint32_t leader_in_ready0 = mapa(a_in_ready, 0);
        const int row_in_pair = (int)rank * 128 + lw * 32 + lane;
        const int sub = lane >> 2, piece = lane & 3;      // cp.async: 8 rows per instruction, 4 x 16 B per row piece
        // destination of this lane's 16-byte piece inside an array: rows 8 i + sub, i = 0..3 (+512 i bytes); the swizzle
        // term (row >> 1) & 3 does not depend on i
        const uint32_t dst_off = (uint32_t)sub * 64u + ((uint32_t)(piece ^ ((sub >> 1) & 3)) << 4);
        const bool leader = elect_one();
        mbar_wait(&s.w_full, 0);

        int i_up = up0;
        int i_j = 0, i_cs = 0;
        int64_t i_n = -1;                 // this lane's target in the unit pair being issued, -1: none
        int nx_srow = -1;
        uint32_t oe[4], os[4], ot[4], vmask = 0;
        int gr[4] = {0, 0, 0, 0};         // kGather: lanes 0-7 hold the source rows of tile rows 4 lane .. 4 lane + 3
        bool i_live = false;
        auto load_src = [&](int j) -> int {               // source row of the j-th in-edge of this lane's target
            return (i_n >= 0 && j < k) ? __ldg(a.src + i_n * k + j) : -1;
        };
        auto spread_slot = [&](int j, int srow) {
            vmask = 0;
            if (kGather) {
                // a tile row without an edge reads row 0 of P_r: its accumulator row is never stored (rows are independent
                // through the MMAs and the LayerNorm; agg is skipped and the e' store is clipped for it)
#pragma unroll
                for (int i = 0; i < 4; ++i) gr[i] = max(__shfl_sync(0xffffffffu, srow, (4 * lane + i) & 31), 0);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int sr = __shfl_sync(0xffffffffu, srow, 8 * i + sub);
                    const bool ok = sr >= 0;
                    vmask |= ok ? (1u << i) : 0u;
                    os[i] = ok ? (uint32_t)sr * 32u + piece : 0u;
                    if (!kTmaLoad) {
                        const long long er = __shfl_sync(0xffffffffu, (long long)(i_n >= 0 ? i_n * k + j : -1), 8 * i + sub);
                        oe[i] = ok ? (uint32_t)er * 32u + piece : 0u;
                    }
                }
            }
        };
        auto seek_unit = [&]() {
            i_live = i_up < n_up;         // fixed in-degree: every unit pair has k slots
            if (i_live) {
                const int64_t n = (int64_t)i_up * 256 + row_in_pair;
                i_n = n < a.n_targets ? n : -1;
                i_j = 0; i_cs = 0;
                if (!kTmaLoad) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long long tr = __shfl_sync(0xffffffffu, (long long)i_n, 8 * i + sub);
                        ot[i] = tr >= 0 ? (uint32_t)tr * 32u + piece : 0u;
                    }
                }
                spread_slot(0, load_src(0));
            }
        };
        auto issue_stage = [&](uint32_t stage_addr, uint32_t bar_addr) {
            if (i_live) {
                const uint32_t dst0 = stage_addr + dst_off;
                const int col0 = (2 * i_cs + hf) * SCOLS;                   // first column of this stage
                const char* br = reinterpret_cast<const char*>(a.P_r) + col0 * 4;
                const float* pr[4];
                uint32_t sz[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sz[i] = (vmask >> i) & 1u ? 16u : 0u;
                    pr[i] = reinterpret_cast<const float*>(br + (size_t)os[i] * 16);
                }
                if (kGather) {
                    const int n0w = (i_up * 2 + (int)rank) * 128 + lw * 32;
                    if (leader) {
                        mbar_arrive_expect_tx_a(bar_addr, 3 * ARR);
                        tma_load_3d(stage_addr, &tm.e_in, col0, i_j, n0w, bar_addr);
                        tma_load_2d(stage_addr + 2 * ARR, &tm.p_c, col0, n0w, bar_addr);
                    }
                    __syncwarp();
                    // eight copies of four gathered rows each; divergent operands: ptxas serialises the lanes (R2UR + BRA.U.ANY)
                    if (lane < 8) tma_gather4(stage_addr + ARR + 256u * lane, &tm.p_r, col0, gr[0], gr[1], gr[2], gr[3], bar_addr);
                } else if (kTmaLoad) {
                    asm volatile(
                        "cp.async.cg.shared.global [%0 + 2048], [%1], 16, %5;\n\t"
                        "cp.async.cg.shared.global [%0 + 2560], [%2], 16, %6;\n\t"
                        "cp.async.cg.shared.global [%0 + 3072], [%3], 16, %7;\n\t"
                        "cp.async.cg.shared.global [%0 + 3584], [%4], 16, %8;\n"
                        ::"r"(dst0), "l"(pr[0]), "l"(pr[1]), "l"(pr[2]), "l"(pr[3]), "r"(sz[0]), "r"(sz[1]), "r"(sz[2]), "r"(sz[3])
                        : "memory");
                    if (leader) {
                        // rows past the last target are zero-filled by the tensor maps' bounds (and count towards the bytes)
                        const int n0w = (i_up * 2 + (int)rank) * 128 + lw * 32;
                        mbar_arrive_expect_tx_a(bar_addr, 2 * ARR);
                        tma_load_3d(stage_addr, &tm.e_in, col0, i_j, n0w, bar_addr);
                        tma_load_2d(stage_addr + 2 * ARR, &tm.p_c, col0, n0w, bar_addr);
                    }
                } else {
                    const char* be = reinterpret_cast<const char*>(a.e_in) + col0 * 4;
                    const char* bc = reinterpret_cast<const char*>(a.P_c) + col0 * 4;
                    const float* pe[4];
                    const float* pc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        pe[i] = reinterpret_cast<const float*>(be + (size_t)oe[i] * 16);
                        pc[i] = reinterpret_cast<const float*>(bc + (size_t)ot[i] * 16);
                    }
                    asm volatile(
                        "cp.async.cg.shared.global [%0], [%1], 16, %13;\n\t"
                        "cp.async.cg.shared.global [%0 + 2048], [%2], 16, %13;\n\t"
                        "cp.async.cg.shared.global [%0 + 4096], [%3], 16, %13;\n\t"
                        "cp.async.cg.shared.global [%0 + 512], [%4], 16, %14;\n\t"
                        "cp.async.cg.shared.global [%0 + 2560], [%5], 16, %14;\n\t"
                        "cp.async.cg.shared.global [%0 + 4608], [%6], 16, %14;\n\t"
                        "cp.async.cg.shared.global [%0 + 1024], [%7], 16, %15;\n\t"
                        "cp.async.cg.shared.global [%0 + 3072], [%8], 16, %15;\n\t"
                        "cp.async.cg.shared.global [%0 + 5120], [%9], 16, %15;\n\t"
                        "cp.async.cg.shared.global [%0 + 1536], [%10], 16, %16;\n\t"
                        "cp.async.cg.shared.global [%0 + 3584], [%11], 16, %16;\n\t"
                        "cp.async.cg.shared.global [%0 + 5632], [%12], 16, %16;\n"
                        ::"r"(dst0), "l"(pe[0]), "l"(pr[0]), "l"(pc[0]), "l"(pe[1]), "l"(pr[1]), "l"(pc[1]), "l"(pe[2]), "l"(pr[2]),
                        "l"(pc[2]), "l"(pe[3]), "l"(pr[3]), "l"(pc[3]), "r"(sz[0]), "r"(sz[1]), "r"(sz[2]), "r"(sz[3])
                        : "memory");
                }
                static_assert(ARR == 2048 && STG == 6144, "offsets in the cp.async blocks above");
                if (++i_cs == 1) nx_srow = load_src(i_j + 1);               // source ids of the next slot: three stages of slack
                if (i_cs == NCS_W) {
                    i_cs = 0;
                    if (++i_j < k) spread_slot(i_j, nx_srow);
                    else { i_up += up_stride; seek_unit(); }
                }
            }
            cp_async_commit();
        };

        seek_unit();
#pragma unroll 1
        for (int p = 0; p < NSTG - 1; ++p) issue_stage(ring0 + p * STG, bar0 + 8u * p);
        PROF_DECL
        PROF_START();
        uint32_t q = 0, n_slot0 = 0, n_slot1 = 0;
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);
        for (int up = up0; up < n_up; up += up_stride) {
            for (int j = 0; j < k; ++j) {
                const int cc = j & 1;
                const uint32_t d_col = tmem + lane_base + 256u * cc;
#pragma unroll 1
                for (int cw = 0; cw < NCS_W; ++cw, ++q) {
                    const int cs = 2 * cw + hf;
                    cp_async_wait<NSTG - 2>();                   // the cp.async part of stage q has landed
                    if (kTmaLoad) mbar_wait_sleep_a(bar0 + 8u * (q % NSTG), (q / NSTG) & 1);      // ... and its two TMA tiles
                    __syncwarp();
                    PROF_LAP(0);                                 // waiting for the staged rows
                    issue_stage(ring0 + ((q + NSTG - 1) % NSTG) * STG, bar0 + 8u * ((q + NSTG - 1) % NSTG));
                    if (cw == 0) {
                        mbar_wait_sleep_a(a_d_free + 8u * cc, ((cc ? n_slot1 : n_slot0) + 1) & 1);
                        tc_fence_after();
                        PROF_LAP(1);                             // waiting for the accumulator to be released
                    }
                    const uint32_t st = ring0 + (q % NSTG) * STG + lane * 64;
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        float4 xe[2], xr[2], xc[2];
#pragma unroll
                        for (int v4 = 0; v4 < 2; ++v4) {
                            const uint32_t ch = ((uint32_t)(2 * h8 + v4) ^ swz) << 4;
                            xe[v4] = lds_f4(st + ch);
                            xr[v4] = lds_f4(st + ARR + ch);
                            xc[v4] = lds_f4(st + 2 * ARR + ch);
                        }
                        uint32_t eh[4], el[4], pp[8];
#pragma unroll
                        for (int v4 = 0; v4 < 2; ++v4) {
                            split2(xe[v4].x, xe[v4].y, eh[2 * v4], el[2 * v4]);
                            split2(xe[v4].z, xe[v4].w, eh[2 * v4 + 1], el[2 * v4 + 1]);
                            pp[4 * v4] = __float_as_uint((xr[v4].x + xc[v4].x) * ps);
                            pp[4 * v4 + 1] = __float_as_uint((xr[v4].y + xc[v4].y) * ps);
                            pp[4 * v4 + 2] = __float_as_uint((xr[v4].z + xc[v4].z) * ps);
                            pp[4 * v4 + 3] = __float_as_uint((xr[v4].w + xc[v4].w) * ps);
                        }
                        tmem_st4(d_col + 128u + 8u * cs + 4u * h8, eh);
                        tmem_st4(d_col + 192u + 8u * cs + 4u * h8, el);
                        tmem_st8(d_col + 16u * cs + 8u * h8, pp);
                    }
                    if (cw == NCS_W - 1) {
                        tmem_wait_st();
                        tc_fence_before();
                    }
                    __syncwarp();                                // stage buffer may be refilled
                    if (cw == NCS_W - 1) {
                        if (lane == 0) mbar_arrive_remote(leader_in_ready0 + 8u * cc);
                        if (cc) ++n_slot1; else ++n_slot0;
                    }
                    PROF_LAP(2);                                 // read back, split / add, TMEM writes, next prefetch
                }
            }
        }
        cp_async_wait<0>();
        PROF_FLUSH(8, warp == W_LOAD0);
    } else {
        setmaxnreg_dec<kRegsMisc>();
        if (warp == W_MMA && rank == 0) {
            // ================================================================== MMA issuer (leader CTA), as in v3
            const uint32_t idesc = idesc_f16(256, 128);
            const uint64_t w_desc = make_desc_sw128(smem_u32(s.w[0]));
            uint32_t n_chain[2] = {0, 0}, n_ar[2] = {0, 0};
            for (int up = up0; up < n_up; up += up_stride) {
                for (int j0 = 0; j0 < k; j0 += 2) {
                    const int nch = min(2, k - j0);
                    for (int l = 0; l < nl; ++l) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            if (c >= nch) continue;
                            const uint32_t d_col = tmem + 256u * c, ah = d_col + 128u, al = d_col + 192u;
                            if (l == 0) {
                                if (lane == 0) {
                                    mbar_wait_sleep_a(a_in_ready + 8u * c, n_chain[c] & 1);
                                    tc_fence_after();
                                }
                                ++n_chain[c];
                            } else {
                                if (lane == 0) {
                                    mbar_wait_sleep_a(a_a_ready + 8u * c, n_ar[c] & 1);
                                    tc_fence_after();
                                }
                                ++n_ar[c];
                            }
                            if (elect_one()) {
                                const uint64_t wb = w_desc + (uint64_t)((l * 4 * HIMG) >> 4);
#pragma unroll 1
                                for (int ks = 0; ks < 8; ++ks) {
                                    const uint64_t wh = wb + (uint64_t)(((ks >> 2) * 2 * HIMG + (ks & 3) * 32) >> 4);
                                    const uint64_t wl = wh + (uint64_t)(HIMG >> 4);
                                    umma_ts<2>(d_col, ah + 8 * ks, wh, idesc, (l == 0 || ks > 0) ? 1u : 0u);
                                    umma_ts<2>(d_col, al + 8 * ks, wh, idesc, 1u);
                                    umma_ts<2>(d_col, ah + 8 * ks, wl, idesc, 1u);
                                }
                                umma_commit_a<2>(a_d_full + 8u * c, 3);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    }

    // shared memory must outlive the last bulk store's read.  (Placed here and not at the end of the epilogue role: code
    // after that role's unit loop makes ptxas spill 16 of the agg[] registers.)
    if (warp < N_EPI_WARPS && (tid & 31) == 0) bulk_wait_read0();
    tc_fence_before();
    cluster_sync_all();
    if (warp == W_MMA) tmem_dealloc<2>(tmem, 512);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// fp32 view [rows, (k,) 128] of a row-major feature matrix; box = box_cols x (1 x) box_rows
static bool encode(CUtensorMap* m, const float* base, int64_t rows, int k, int box_cols, CUtensorMapSwizzle swz, int box_rows = 32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint32_t ones[3] = {1, 1, 1};
    CUresult r;
    if (k > 0) {
        const cuuint64_t dims[3] = {128, (cuuint64_t)k, (cuuint64_t)rows};
        const cuuint64_t strides[2] = {512, (cuuint64_t)512 * k};
        const cuuint32_t box[3] = {(cuuint32_t)box_cols, 1, (cuuint32_t)box_rows};
        r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[2] = {128, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {512};
        const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
        r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS;
}

// the same, callable from tma_test.cu (swizzle given in bytes: 32 / 64 / 128)
bool encode_rows(CUtensorMap* m, const float* base, int64_t rows, int k, int box_cols, int swizzle_bytes, int box_rows) {
    const CUtensorMapSwizzle swz = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                                      : CU_TENSOR_MAP_SWIZZLE_128B;
    return encode(m, base, rows, k, box_cols, swz, box_rows);
}

}  // namespace ep4

int edge_pair_tma_profile(unsigned long long* out64) {
#ifdef G4C_PROFILE
    if (cudaMemcpyFromSymbol(out64, ep4::g_prof, sizeof(unsigned long long) * 64) != cudaSuccess) return check_launch("profile read");
    unsigned long long zero[64] = {0};
    cudaMemcpyToSymbol(ep4::g_prof, zero, sizeof(zero));
    return G4C_OK;
#else
    (void)out64;
    set_error("libg4c was built without -DG4C_PROFILE");
    return G4C_EUNSUPPORTED;
#endif
}

static int g_edge_mode = -1;      // -1: read G4C_EDGE_MODE on first use

void edge_pair_set_mode(int mode) { g_edge_mode = mode; }

int edge_pair_mode() {
    if (g_edge_mode < 0) {
        const char* e = std::getenv("G4C_EDGE_MODE");
        g_edge_mode = e ? std::atoi(e) : 0;
        if (g_edge_mode < 0 || g_edge_mode > 4) g_edge_mode = 0;
    }
    return g_edge_mode;
}

bool edge_pair_tma_supported(const EdgeArgs& a) {
    return a.fixed_k > 0 && a.edge_perm == nullptr && a.tgt_perm == nullptr && a.n_targets > 0;
}

int edge_pair_tma_launch(const EdgeArgs& a, int mode, cudaStream_t st) {
    if (!edge_pair_tma_supported(a)) { set_error("edge_pair_tma_launch: needs fixed_k > 0 and no permutations"); return G4C_EUNSUPPORTED; }
    if (a.act_e_out != G4C_ACT_NONE && a.act_e_out != G4C_ACT_SELU) { set_error("g4c_edge_aggr_fwd: act_e_out must be none or selu"); return G4C_EUNSUPPORTED; }
    ep4::Maps tm;
    bool ok = ep4::encode(&tm.e_in, a.e_in, a.n_targets, a.fixed_k, 16, CU_TENSOR_MAP_SWIZZLE_64B) &&
              ep4::encode(&tm.p_c, a.P_c, a.n_targets, 0, 16, CU_TENSOR_MAP_SWIZZLE_64B) &&
              ep4::encode(&tm.e_out, a.e_out ? a.e_out : a.e_in, a.n_targets, a.fixed_k, 8, CU_TENSOR_MAP_SWIZZLE_32B) &&
              // tile::gather4 map: one-row box, four row coordinates per copy.  The descriptor does not carry the number of source
              // rows (G4cEdgeDesc has no such field); the bound only matters for out-of-range coordinates, which are never issued.
              ep4::encode(&tm.p_r, a.P_r, (int64_t)1 << 28, 0, 16, CU_TENSOR_MAP_SWIZZLE_64B, 1);
    if (!ok) { set_error("edge_pair_tma_launch: cuTensorMapEncodeTiled failed"); return G4C_ECUDA; }
    static bool configured = false;
    const int smem = (int)sizeof(ep4::Smem);
    if (!configured) {
        if (cudaFuncSetAttribute(ep4::edge_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaFuncSetAttribute(ep4::edge_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaFuncSetAttribute(ep4::edge_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaFuncSetAttribute(ep4::edge_tma_kernel<2, 96, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return check_launch("edge_tma_kernel attribute");
        configured = true;
    }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    const int64_t n_units = (a.n_targets + 127) / 128, n_up = (n_units + 1) / 2;
    const int pairs = (int)std::min<int64_t>(n_up, n_sm / 2);
    if (mode >= 4) ep4::edge_tma_kernel<2, 96, 48><<<2 * pairs, ep4::NT, smem, st>>>(a, tm);
    else if (mode == 3) ep4::edge_tma_kernel<2><<<2 * pairs, ep4::NT, smem, st>>>(a, tm);
    else if (mode == 2) ep4::edge_tma_kernel<1><<<2 * pairs, ep4::NT, smem, st>>>(a, tm);
    else ep4::edge_tma_kernel<0><<<2 * pairs, ep4::NT, smem, st>>>(a, tm);
    count_launch();
    return check_launch("edge_tma_kernel");
}

}  // namespace g4c
