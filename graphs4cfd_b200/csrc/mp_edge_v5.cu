// mp_edge_v5.cu — fused edge-MLP + LayerNorm + aggregation kernel for launches with a fixed in-degree (kNN levels and
// every REMuS angle level: the launches that dominate a rollout step).  Fifth generation of the kernel behind
// g4c_edge_aggr_fwd; launches with a CSR or a permutation keep the cp.async kernel of mp_edge_pair.cu ("v3").
//
// Reference arithmetic (graphs4cfd/nn/blocks.py:181-183, 328-330, 376-378):
//     e' = LN(MLP(cat(e, S[src], T[tgt]))) ;  agg[t] = mean/sum over the in-edges of t of e'
// with the first Linear split exactly, W1 [e, S[src], T[tgt]] + b1 = W1e e + P_r[src] + P_c[tgt] (see mp_edge_pair.cu).
//
// Work decomposition, TMEM layout, MMA issue and mbarrier protocol are v3's: a CTA pair (cta_group::2, M = 256) owns two
// consecutive units of 128 targets; slot j of a unit = the j-th in-edge of each of its targets, so tile row m always
// belongs to target m; two chains (even / odd slots) alternate in TMEM: columns [256c, +128) accumulator, [+128, +64) A
// hi, [+192, +64) A lo; the weights of all layers stay resident in shared memory (96 KiB per CTA).
//
// What changed against v3, and why (profiles/r2a_edge_v3_instruction_mix.txt: v3 executes 36.1k warp instructions per
// 128-edge slot = 9.0k per scheduler against an HBM-bound slot time of 6.7k cycles — it is bound by instruction ISSUE):
//   * regular streams ride the TMA engine (measured: 2.74 -> 2.46 ms at 1M targets, k = 6, 3 layers): e and P_c[tgt]
//     arrive as tensor tiles (box 16 x 1 x 32 of the [N, k, 128] view of e, box 16 x 32 of P_c, SWIZZLE_64B) on a
//     per-warp, per-stage mbarrier, only the gathered P_r[src] pieces stay on cp.async; e' leaves through a SWIZZLE_32B
//     staging tile and cp.async.bulk.tensor stores (no STG.256 that touches 32 lines per instruction);
//   * every elementwise stage works on packed fp32 pairs (FFMA2 / FADD2 / FMUL2, f32x2.cuh);
//   * hidden activations are kept as SELU(x) / (lambda ln 2): the positive branch is the log2-domain pre-activation
//     itself (no multiply), the constant rides on the next layer's 1/s;
//   * LayerNorm output = FFMA2(FFMA2(y, rstd, -mean rstd), gamma', beta') with gamma' = gamma log2(e), beta' = beta
//     log2(e) when the SELU of the model (nn/mus_gnn.py:321) follows, so the result is already the exponent's argument;
//     the aggregation accumulates that scaled value and is un-scaled once per unit;
//   * the loaders wait for "accumulator released" on a hardware named barrier (bar.arrive by the 16 epilogue warps,
//     bar.sync by the 8 loader warps) instead of polling an mbarrier (v3: 3.5k polling instructions per slot);
//   * P_r / P_c may arrive pre-multiplied by the layer-1 scale (p_scale == 1: no multiply in the loaders).
#include <algorithm>
#include <cstddef>
#include <cuda.h>
#include "tc2_core.cuh"
#include "f32x2.cuh"
#include "mp_pair.h"

namespace g4c {
namespace ep5 {

using namespace tc2;
using namespace p2;

constexpr int H = 128;
constexpr int HIMG = 64 * 128;
constexpr int NT = 896;                  // warps 0-15 epilogue, 16-23 loaders, 24 MMA issuer, 25-27 idle
constexpr int N_EPI_WARPS = 16;
constexpr int N_LOAD_WARPS = 8;
constexpr int W_LOAD0 = 16, W_MMA = 24;
constexpr int kRegsEpi = 88, kRegsLoad = 64, kRegsMisc = 24;
static_assert(512 * kRegsEpi + 256 * kRegsLoad + 128 * kRegsMisc <= 896 * 72, "register budget");

constexpr int SCOLS = 16;                // columns per loader stage
constexpr int NCS = H / SCOLS;
constexpr int NCS_W = NCS / 2;           // stages per slot per loader warp (the two warps of a lane quarter alternate)
constexpr int ARR = 32 * 64;             // one array's 32 row pieces of a stage: 64-byte pitch, SWIZZLE_64B
constexpr int STG = 3 * ARR;             // e | P_r | P_c
constexpr int NSTG = 2;
constexpr int OPIECE = 32 * 32;          // e' staging of one epilogue warp: 32 rows x 8 columns, SWIZZLE_32B
constexpr int BAR_DFREE0 = 6;                   // named barriers: 1-4 lane quarters, 5 all epilogue warps, 6 / 7 accumulator released

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
constexpr float kA = kSeluAlpha * kLog2e;            // SELU(x) / (lambda ln 2) = t > 0 ? t : kA 2^t - kA,  t = x log2(e)

struct Maps {
    CUtensorMap e_in;                    // [N, k, 128] fp32, box 16 x 1 x 32, SWIZZLE_64B
    CUtensorMap p_c;                     // [N, 128] fp32, box 16 x 32, SWIZZLE_64B
    CUtensorMap e_out;                   // [N, k, 128] fp32, box 8 x 1 x 32, SWIZZLE_32B
};

struct Smem {
    uint8_t w[3][4 * HIMG];                          // 96 KiB, 1024-byte aligned (UMMA SWIZZLE_128B images)
    uint8_t ring[N_LOAD_WARPS][NSTG][STG];           // 96 KiB, every array 2048-byte aligned
    uint8_t ostage[N_EPI_WARPS][OPIECE];             // 16 KiB, 1024-byte aligned
    float cst[5][H];                                 // [l] bias of layer l (hidden: times log2 e), [3] gamma', [4] beta'
    float part[2][2][4][H];                          // LayerNorm partials [buffer][mean | M2][column quarter][row]
    uint64_t w_full;
    uint64_t in_ready[2];                            // leader: A operand + initial accumulator of chain c written
    uint64_t a_ready[2];                             // leader: next layer's A operand written
    uint64_t d_full[2];                              // local, multicast commit
    uint64_t ld_full[N_LOAD_WARPS][NSTG];            // the two TMA tiles of a loader stage have landed
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
// "accumulator of chain c released": 16 epilogue warps arrive, 8 loader warps wait (no polling, no issue slots while waiting)
__device__ __forceinline__ void dfree_arrive(int c) { asm volatile("bar.arrive %0, 768;" ::"r"(BAR_DFREE0 + c) : "memory"); }
__device__ __forceinline__ void dfree_wait(int c) { asm volatile("bar.sync %0, 768;" ::"r"(BAR_DFREE0 + c) : "memory"); }

// sleeping mbarrier wait with a short watchdog (a protocol error traps after seconds instead of hanging the GPU)
__device__ __forceinline__ void mbar_wait_sleep_w(uint32_t addr, uint32_t parity) {
    uint32_t spins = 0, ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (!ok && ++spins > (1u << 20)) __trap();
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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, int c0, int c1, int c2, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) edge_v5_kernel(const EdgeArgs a, const __grid_constant__ Maps tm) {
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
    // the SELU that the models apply to a block's edge output: fold log2(e) into the LayerNorm affine
    const bool selu_out = a.e_out != nullptr && a.act_e_out == G4C_ACT_SELU;
    const float fold = selu_out ? kLog2e : 1.f;

    if (tid == 0) {
        mbar_init(&s.w_full, 1);
        for (int c = 0; c < 2; ++c) {
            mbar_init(&s.in_ready[c], 2 * N_LOAD_WARPS);
            mbar_init(&s.a_ready[c], 2 * N_EPI_WARPS);
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
        s.cst[3][tid] = (a.gamma ? a.gamma[tid] : 1.f) * fold;
        s.cst[4][tid] = (a.beta ? a.beta[tid] : 0.f) * fold;
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const uint32_t sb = smem_u32(smem_raw);
    const uint32_t a_cst = sb + (uint32_t)offsetof(Smem, cst), a_part = sb + (uint32_t)offsetof(Smem, part);
    const uint32_t a_in_ready = sb + (uint32_t)offsetof(Smem, in_ready), a_a_ready = sb + (uint32_t)offsetof(Smem, a_ready);
    const uint32_t a_d_full = sb + (uint32_t)offsetof(Smem, d_full);

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
        const uint32_t ost0 = sb + (uint32_t)offsetof(Smem, ostage) + (uint32_t)(etid >> 5) * OPIECE + (uint32_t)lane * 32u +
                              ((uint32_t)((lane >> 2) & 1) << 4);
        // two-layer MLPs (REMuS) leave the third layer's 32 KiB of weight space unused: staging is double buffered there
        const bool two_buf = nl == 2;
        const uint32_t ost_alt = sb + (uint32_t)offsetof(Smem, w) + 2u * 4u * HIMG + (uint32_t)(etid >> 5) * (2u * OPIECE) +
                                 (uint32_t)lane * 32u + ((uint32_t)((lane >> 2) & 1) << 4);
        uint32_t n_dfull[2] = {0, 0};
        const bool has_ln = a.gamma != nullptr;
        const bool want_e = a.e_out != nullptr;
        int pbuf = 0;
        PROF_DECL
        PROF_START();

        for (int up = up0; up < n_up; up += up_stride) {
            const int n_unit0 = (up * 2 + (int)rank) * 128;         // n_targets < 2^31 (checked by g4c_edge_aggr_fwd)
            const bool live = n_unit0 + row < a.n_targets;
            uint64_t agg[16];                                       // 32 columns as pairs, in units of `fold`
#pragma unroll
            for (int i = 0; i < 16; ++i) agg[i] = 0ull;

            for (int j0 = 0; j0 < k; j0 += 2) {
                const int nch = min(2, k - j0);
                for (int l = 0; l < nl; ++l) {
                    // scale of this layer's accumulator: 1/s, times lambda ln 2 when its input was a deferred-constant SELU
                    const float cl = a.inv_scale[l] * (l > 0 ? kSeluScale * kLn2 : 1.f);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c >= nch) continue;
                        const uint32_t d_addr = tmem + lane_base + 256u * c + 32u * cq;
                        if (warp == 0) mbar_wait_sleep_w(a_d_full + 8u * c, n_dfull[c] & 1);
                        ++n_dfull[c];
                        epi_sync_all();
                        tc_fence_after();
                        PROF_LAP(l < nl - 1 ? 0 : 1);        // waiting for the MMAs (hidden / last layer)
                        if (l < nl - 1) {
                            // ---- hidden layer: x'' = SELU(acc cl + b) / (lambda ln 2) as fp16 (hi, lo) A operand columns
                            const uint64_t c2 = bc(cl * kLog2e), pA = bc(kA), nA = bc(-kA);
#pragma unroll
                            for (int h16 = 0; h16 < 2; ++h16) {
                                float v[16];
                                tmem_ld16f(d_addr + 16u * h16, v);
                                uint32_t hi[8], lo[8];
                                const uint32_t bs = my_cst + 512u * l + 64u * h16;
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    const float4 b = lds_f4(bs + 4u * i);
                                    float t0, t1, t2, t3, n0, n1, n2, n3;
                                    upk(fma2(pk(v[i], v[i + 1]), c2, pk(b.x, b.y)), t0, t1);
                                    upk(fma2(pk(v[i + 2], v[i + 3]), c2, pk(b.z, b.w)), t2, t3);
                                    upk(fma2(pk(ex2(t0), ex2(t1)), pA, nA), n0, n1);
                                    upk(fma2(pk(ex2(t2), ex2(t3)), pA, nA), n2, n3);
                                    split_pair(t0 > 0.f ? t0 : n0, t1 > 0.f ? t1 : n1, hi[i / 2], lo[i / 2]);
                                    split_pair(t2 > 0.f ? t2 : n2, t3 > 0.f ? t3 : n3, hi[i / 2 + 1], lo[i / 2 + 1]);
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
                            // ---- last layer: LayerNorm, aggregation, store.  The accumulator is read once and released at once.
                            const uint32_t bs = my_cst + 512u * l;
                            float y[32];
                            tmem_ld16_nowait(d_addr, y);
                            tmem_ld16_nowait(d_addr + 16u, y + 16);
                            tmem_wait_ld();
                            tc_fence_before();
                            dfree_arrive(c);
                            uint64_t yp[16];
                            const uint64_t clp = bc(cl);
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 b4 = lds_f4(bs + 4u * i);
                                yp[i / 2] = fma2(pk(y[i], y[i + 1]), clp, pk(b4.x, b4.y));
                                yp[i / 2 + 1] = fma2(pk(y[i + 2], y[i + 3]), clp, pk(b4.z, b4.w));
                            }
                            float mean = 0.f, rstd = 1.f;
                            if (has_ln) {
                                float mh[2], M2h[2];
#pragma unroll
                                for (int h16 = 0; h16 < 2; ++h16) {
                                    const uint64_t* p = yp + 8 * h16;
                                    float s0, s1;
                                    upk(add2(add2(add2(p[0], p[1]), add2(p[2], p[3])), add2(add2(p[4], p[5]), add2(p[6], p[7]))), s0, s1);
                                    mh[h16] = (s0 + s1) * (1.f / 16.f);
                                    const uint64_t mp = bc(mh[h16]);
                                    uint64_t q0 = 0ull, q1 = 0ull;
#pragma unroll
                                    for (int i = 0; i < 8; i += 2) {
                                        const uint64_t d0 = sub2(p[i], mp), d1 = sub2(p[i + 1], mp);
                                        q0 = fma2(d0, d0, q0);
                                        q1 = fma2(d1, d1, q1);
                                    }
                                    upk(add2(q0, q1), s0, s1);
                                    M2h[h16] = s0 + s1;
                                }
                                const float dm = mh[0] - mh[1];
                                const uint32_t pa = my_part + 4096u * pbuf;
                                sts_f1(pa + 512u * cq, 0.5f * (mh[0] + mh[1]));                       // mean of this thread's 32 columns
                                sts_f1(pa + 2048u + 512u * cq, (M2h[0] + M2h[1]) + 8.f * dm * dm);     // their M2 (Chan)
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
                            const uint64_t rs = bc(rstd), nm = bc(-mean * rstd);
                            const uint64_t lp = bc(kSeluScale * kLn2), sa = bc(kSeluScale * kSeluAlpha), nsa = bc(-kSeluScale * kSeluAlpha);
                            const int j = j0 + c;
#pragma unroll
                            for (int i8 = 0; i8 < 32; i8 += 8) {
                                // o = fold * LN(y): normalise, affine (gamma', beta' carry `fold`), aggregate
                                uint64_t o[4];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const float4 g = lds_f4(my_cst + 1536u + 4u * (i8 + 4 * u));
                                    const float4 be = lds_f4(my_cst + 2048u + 4u * (i8 + 4 * u));
                                    o[2 * u] = fma2(fma2(yp[i8 / 2 + 2 * u], rs, nm), pk(g.x, g.y), pk(be.x, be.y));
                                    o[2 * u + 1] = fma2(fma2(yp[i8 / 2 + 2 * u + 1], rs, nm), pk(g.z, g.w), pk(be.z, be.w));
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) agg[i8 / 2 + u] = add2(agg[i8 / 2 + u], o[u]);
                                if (want_e) {                   // warp-uniform
                                    float r[8];
                                    if (selu_out) {
                                        // lambda SELU(x) from t = x log2 e: t > 0 ? t lambda ln 2 : lambda alpha 2^t - lambda alpha
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            float t0, t1, p0, p1, n0, n1;
                                            upk(o[u], t0, t1);
                                            upk(mul2(o[u], lp), p0, p1);
                                            upk(fma2(pk(ex2(t0), ex2(t1)), sa, nsa), n0, n1);
                                            r[2 * u] = t0 > 0.f ? p0 : n0;
                                            r[2 * u + 1] = t1 > 0.f ? p1 : n1;
                                        }
                                    } else {
#pragma unroll
                                        for (int u = 0; u < 4; ++u) upk(o[u], r[2 * u], r[2 * u + 1]);
                                    }
                                    // the staging tile is free once the previous piece's bulk store has read it
                                    const uint32_t ost = two_buf ? ost_alt + (uint32_t)((i8 >> 3) & 1) * OPIECE : ost0;
                                    if (lane == 0) {
                                        if (two_buf) bulk_wait_read1(); else bulk_wait_read0();
                                    }
                                    __syncwarp();
                                    sts_f4(ost, r[0], r[1], r[2], r[3]);
                                    sts_f4(ost ^ 16u, r[4], r[5], r[6], r[7]);
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
                const float rc = ((a.aggr == G4C_AGGR_MEAN) ? 1.f / (float)max(k, 1) : 1.f) / fold;
                float* dst = a.agg_out + (size_t)(n_unit0 + row) * H + cq * 32;
                const uint64_t rcp = bc(rc);
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 4; ++u) upk(mul2(agg[i / 2 + u], rcp), o[2 * u], o[2 * u + 1]);
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
        const bool scale_p = ps != 1.f;
        const uint64_t psp = bc(ps);
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
        mbar_wait(&s.w_full, 0);           // in_ready is only signalled once this CTA's weights have landed

        // ---- issue cursor: runs NSTG-1 stages ahead of the processing cursor
        int i_up = up0;
        int i_j = 0, i_cs = 0;
        int64_t i_n = -1;                 // this lane's target in the unit pair being issued, -1: none
        int nx_srow = -1;
        uint32_t os[4], vmask = 0;        // offsets (16-byte units) of this lane's piece in the four source rows it copies
        bool i_live = false;
        auto load_src = [&](int j) -> int {               // source row of the j-th in-edge of this lane's target
            return (i_n >= 0 && j < k) ? __ldg(a.src + i_n * k + j) : -1;
        };
        auto spread_slot = [&](int srow) {
            vmask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int sr = __shfl_sync(0xffffffffu, srow, 8 * i + sub);
                const bool ok = sr >= 0;
                vmask |= ok ? (1u << i) : 0u;
                os[i] = ok ? (uint32_t)sr * 32u + piece : 0u;
            }
        };
        auto seek_unit = [&]() {
            i_live = i_up < n_up;         // fixed in-degree: every unit pair has k slots
            if (i_live) {
                const int64_t n = (int64_t)i_up * 256 + row_in_pair;
                i_n = n < a.n_targets ? n : -1;
                i_j = 0; i_cs = 0;
                spread_slot(load_src(0));
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
                    sz[i] = (vmask >> i) & 1u ? 16u : 0u;        // 0: nothing is read, the destination is zero-filled
                    pr[i] = reinterpret_cast<const float*>(br + (size_t)os[i] * 16);
                }
                asm volatile(
                    "cp.async.cg.shared.global [%0 + 2048], [%1], 16, %5;\n\t"
                    "cp.async.cg.shared.global [%0 + 2560], [%2], 16, %6;\n\t"
                    "cp.async.cg.shared.global [%0 + 3072], [%3], 16, %7;\n\t"
                    "cp.async.cg.shared.global [%0 + 3584], [%4], 16, %8;\n"
                    ::"r"(dst0), "l"(pr[0]), "l"(pr[1]), "l"(pr[2]), "l"(pr[3]), "r"(sz[0]), "r"(sz[1]), "r"(sz[2]), "r"(sz[3])
                    : "memory");
                static_assert(ARR == 2048 && STG == 6144, "offsets in the cp.async block above");
                if (leader) {
                    // rows past the last target are zero-filled by the tensor maps' bounds (and count towards the bytes)
                    const int n0w = (i_up * 2 + (int)rank) * 128 + lw * 32;
                    mbar_arrive_expect_tx_a(bar_addr, 2 * ARR);
                    tma_load_3d(stage_addr, &tm.e_in, col0, i_j, n0w, bar_addr);
                    tma_load_2d(stage_addr + 2 * ARR, &tm.p_c, col0, n0w, bar_addr);
                }
                if (++i_cs == 1) nx_srow = load_src(i_j + 1);               // source ids of the next slot: three stages of slack
                if (i_cs == NCS_W) {
                    i_cs = 0;
                    if (++i_j < k) spread_slot(nx_srow);
                    else { i_up += up_stride; seek_unit(); }
                }
            }
            cp_async_commit();              // one (possibly empty) group per stage keeps the wait depth constant
        };

        seek_unit();
#pragma unroll 1
        for (int p = 0; p < NSTG - 1; ++p) issue_stage(ring0 + p * STG, bar0 + 8u * p);
        PROF_DECL
        PROF_START();
        uint32_t q = 0, n_slot0 = 0, n_slot1 = 0;      // slots filled per chain
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);
        for (int up = up0; up < n_up; up += up_stride) {
            for (int j = 0; j < k; ++j) {
                const int cc = j & 1;
                const uint32_t d_col = tmem + lane_base + 256u * cc;
#pragma unroll 1
                for (int cw = 0; cw < NCS_W; ++cw, ++q) {
                    const int cs = 2 * cw + hf;
                    cp_async_wait<NSTG - 2>();                   // the cp.async part of stage q has landed
                    mbar_wait_sleep_w(bar0 + 8u * (q % NSTG), (q / NSTG) & 1);      // ... and its two TMA tiles
                    __syncwarp();
                    PROF_LAP(0);                                 // waiting for the staged rows
                    issue_stage(ring0 + ((q + NSTG - 1) % NSTG) * STG, bar0 + 8u * ((q + NSTG - 1) % NSTG));
                    if (cw == 0) {
                        // last-layer epilogue of the previous slot on this chain has read the accumulator
                        if ((cc ? n_slot1 : n_slot0) > 0) dfree_wait(cc);
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
                            split_pair(xe[v4].x, xe[v4].y, eh[2 * v4], el[2 * v4]);
                            split_pair(xe[v4].z, xe[v4].w, eh[2 * v4 + 1], el[2 * v4 + 1]);
                            uint64_t p01 = add2(pk(xr[v4].x, xr[v4].y), pk(xc[v4].x, xc[v4].y));
                            uint64_t p23 = add2(pk(xr[v4].z, xr[v4].w), pk(xc[v4].z, xc[v4].w));
                            if (scale_p) { p01 = mul2(p01, psp); p23 = mul2(p23, psp); }
                            upk_u(p01, pp[4 * v4], pp[4 * v4 + 1]);
                            upk_u(p23, pp[4 * v4 + 2], pp[4 * v4 + 3]);
                        }
                        tmem_st4(d_col + 128u + 8u * cs + 4u * h8, eh);       // A hi: k = 16 cs + 8 h8 .. +7
                        tmem_st4(d_col + 192u + 8u * cs + 4u * h8, el);       // A lo
                        tmem_st8(d_col + 16u * cs + 8u * h8, pp);             // accumulator columns 16 cs + 8 h8 .. +7
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
        // drain: the epilogue arrives once per slot, the loaders waited once per slot but the first of each chain
        if (n_slot0 > 0) dfree_wait(0);
        if (n_slot1 > 0) dfree_wait(1);
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
                                    mbar_wait_sleep_w(a_in_ready + 8u * c, n_chain[c] & 1);
                                    tc_fence_after();
                                }
                                ++n_chain[c];
                            } else {
                                if (lane == 0) {
                                    mbar_wait_sleep_w(a_a_ready + 8u * c, n_ar[c] & 1);
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
    // after that role's unit loop makes ptxas spill accumulator registers.)
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

// fp32 view [rows, (k,) 128] of a row-major feature matrix; box = box_cols x (1 x) box_rows; swizzle in bytes (32 / 64 / 128)
bool encode_rows(CUtensorMap* m, const float* base, int64_t rows, int k, int box_cols, int swizzle_bytes, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const CUtensorMapSwizzle swz = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                                      : CU_TENSOR_MAP_SWIZZLE_128B;
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

}  // namespace ep5

int edge_v5_profile(unsigned long long* out64) {
#ifdef G4C_PROFILE
    if (cudaMemcpyFromSymbol(out64, ep5::g_prof, sizeof(unsigned long long) * 64) != cudaSuccess) return check_launch("profile read");
    unsigned long long zero[64] = {0};
    cudaMemcpyToSymbol(ep5::g_prof, zero, sizeof(zero));
    return G4C_OK;
#else
    (void)out64;
    set_error("libg4c was built without -DG4C_PROFILE");
    return G4C_EUNSUPPORTED;
#endif
}

bool edge_v5_supported(const EdgeArgs& a) {
    return a.fixed_k > 0 && a.edge_perm == nullptr && a.tgt_perm == nullptr && a.n_targets > 0;
}

int edge_v5_launch(const EdgeArgs& a, cudaStream_t st) {
    if (!edge_v5_supported(a)) { set_error("edge_v5_launch: needs fixed_k > 0 and no permutations"); return G4C_EUNSUPPORTED; }
    if (a.act_e_out != G4C_ACT_NONE && a.act_e_out != G4C_ACT_SELU) { set_error("g4c_edge_aggr_fwd: act_e_out must be none or selu"); return G4C_EUNSUPPORTED; }
    ep5::Maps tm;
    const bool ok = ep5::encode_rows(&tm.e_in, a.e_in, a.n_targets, a.fixed_k, 16, 64, 32) &&
                    ep5::encode_rows(&tm.p_c, a.P_c, a.n_targets, 0, 16, 64, 32) &&
                    ep5::encode_rows(&tm.e_out, a.e_out ? a.e_out : a.e_in, a.n_targets, a.fixed_k, 8, 32, 32);
    if (!ok) { set_error("edge_v5_launch: cuTensorMapEncodeTiled failed"); return G4C_ECUDA; }
    static int configured[kMaxDevices] = {0};
    const int smem = (int)sizeof(ep5::Smem);
    if (!ensure_dynamic_smem(ep5::edge_v5_kernel, smem, configured)) return check_launch("edge_v5_kernel attribute");
    const int n_sm = device_sms();
    const int64_t n_units = (a.n_targets + 127) / 128, n_up = (n_units + 1) / 2;
    const int pairs = (int)std::min<int64_t>(n_up, n_sm / 2);
    ep5::edge_v5_kernel<<<2 * pairs, ep5::NT, smem, st>>>(a, tm);
    count_launch(1, true);
    return check_launch("edge_v5_kernel");
}

}  // namespace g4c
