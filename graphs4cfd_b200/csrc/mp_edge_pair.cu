// mp_edge_pair.cu — fused edge-MLP + aggregation kernel of the message-passing block, hidden = 128,
// on CTA pairs (cta_group::2) with every operand that is re-used kept on chip.  (third generation)
//
// Reference arithmetic (graphs4cfd/nn/blocks.py:181-183, 328-330, 376-378):
//     e' = LN(MLP(cat(e, S[src], T[tgt]))) ;  agg[t] = mean/sum over the in-edges of t of e'
// The first Linear is split exactly:  W1 [e, S[src], T[tgt]] + b1 = W1e e + P_r[src] + P_c[tgt]  with
// P_r = S W1s^T and P_c = T W1t^T + b1 computed once per NODE by the row kernel (mp_row_pair.cu), so the
// per-edge GEMMs are all K = 128 and the [E,3H] concatenation never exists.
//
// Work decomposition: a CTA pair owns two consecutive units of 128 targets (one per CTA).  Slot j of a
// unit is the tile made of the j-th in-edge of each of its 128 targets, so tile row m always belongs to
// target m: TMEM lane m / accumulator row m, and the aggregation is a register accumulation in the
// epilogue threads that own the row (fixed order, no atomics).
//
// Per CTA (896 threads):
//   warps 0-15  epilogue: thread (row, column quarter) owns 32 columns of its row (warp w: TMEM lane quarter
//               w & 3, column quarter w >> 2).  TMEM -> registers, bias, SELU, fp16 (hi, lo) split written back
//               to TMEM as the next layer's A operand (tcgen05.st); last layer: LayerNorm (row statistics
//               combined over the four column quarters through shared memory and a 128-thread named barrier
//               per lane quarter), aggregation, activation, 256-bit stores of e'.
//   warps 16-23 loaders, two per TMEM lane quarter (32 tile rows each, alternate column stages).  Rows of e, P_r[src], P_c[tgt] are
//               fetched with cp.async (16 B per lane, 64-byte row pieces, no registers held across the HBM/L2
//               latency) into a private two-stage ring of 16-column stages; row pieces are stored at an
//               80-byte pitch so that "lane = row" 16-byte reads are bank-conflict free.  The warp then reads
//               its rows back with lane = row, splits e into fp16 (hi, lo) and writes it with tcgen05.st as the
//               layer-1 A operand, and writes (P_r + P_c) * s as the INITIAL VALUE of the accumulator.
//   warp 24     (leader CTA) issues every tcgen05.mma of the pair; M = 256, N = 128, the B operand (weights) is
//               resident in shared memory, each CTA holding 64 of the 128 output rows of all layers as
//               pre-split, pre-swizzled fp16 (hi, lo) images (96 KiB).
// Two chains (even / odd slots) alternate so that the MMAs of one overlap the epilogue of the other.
// TMEM: chain c uses columns [256c, 256c+128) accumulator, [256c+128, +64) A hi, [256c+192, +64) A lo.
//
// SELU bookkeeping: hidden activations are kept as x' = SELU(x) / lambda (the A operand of the next layer); the
// factor lambda is applied with the next layer's 1/s on its accumulator.  The pre-activation is formed directly
// in the log2 domain, t = acc * (c * log2 e) + b * log2 e, so that x' = t > 0 ? t * ln 2 : alpha * 2^t - alpha
// costs FFMA + MUFU.EX2 + FFMA + FSETP + predicated FMUL per element.
#include <algorithm>
#include <cstddef>
#include "tc2_core.cuh"
#include "mp_pair.h"

namespace g4c {
namespace ep {

using namespace tc2;

constexpr int H = 128;
constexpr int HIMG = 64 * 128;           // bytes of a 64-row operand image (one CTA's half of a weight K-block)
constexpr int NT = 896;                  // warps 0-15 epilogue, 16-23 loaders, 24 MMA issuer, 25-27 idle
constexpr int N_EPI_WARPS = 16;
constexpr int N_LOAD_WARPS = 8;
constexpr int W_LOAD0 = 16, W_MMA = 24;
// registers per thread after setmaxnreg: the CTA's pool is what it was launched with, 896 * 72 = 64512
// = 512*88 (epilogue) + 256*64 (loaders) + 128*24 (MMA issuer + idle warps); 96/48/24 measured 4 % slower
#ifndef G4C_EP_REGS_EPI
#define G4C_EP_REGS_EPI 88
#define G4C_EP_REGS_LOAD 64
#define G4C_EP_REGS_MISC 24
#endif
constexpr int kRegsEpi = G4C_EP_REGS_EPI, kRegsLoad = G4C_EP_REGS_LOAD, kRegsMisc = G4C_EP_REGS_MISC;
static_assert(512 * kRegsEpi + 256 * kRegsLoad + 128 * kRegsMisc <= 896 * 72, "register budget");

constexpr int SCOLS = 16;                // columns per loader stage
constexpr int NCS = H / SCOLS;           // stages per slot
constexpr int NCS_W = NCS / 2;           // ... per loader warp: the two warps of a lane quarter take alternate stages
constexpr int PITCH = 80;                // bytes between staged 64-byte row pieces (lanes r..r+7 hit 8 distinct 16 B bank groups)
constexpr int ARR = 32 * PITCH;          // one array's 32 row pieces of a stage
constexpr int STG = 3 * ARR;             // stage = e | P_r | P_c pieces of the warp's 32 rows, 16 columns
constexpr int NSTG = 2;

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

struct Smem {
    uint8_t w[3][4 * HIMG];          // layer l: K-block 0 hi | lo, K-block 1 hi | lo (64-row images)
    uint8_t ring[N_LOAD_WARPS][NSTG][STG];      // per loader warp
    float cst[5][H];                 // [l] bias of layer l (hidden layers: times log2 e), [3] gamma, [4] beta
    float part[2][2][4][H];          // LayerNorm partials [buffer][mean | M2][column quarter][row]
    uint64_t w_full;
    uint64_t in_ready[2];            // leader: A operand + initial accumulator of chain c written (16 loader warps of the pair)
    uint64_t a_ready[2];             // leader: next layer's A operand written (32 epilogue warps of the pair)
    uint64_t d_free[2];              // local: accumulator of chain c has been read by the last-layer epilogue (16 warps)
    uint64_t d_full[2];              // local, multicast commit
    uint32_t tmem_base;
};
// ---- optional in-kernel phase profile (make EXTRA=-DG4C_PROFILE): cycles per role and phase in CTA 0, accumulated by
// lane 0 of one warp per role (epilogue warp 0, loader warp 16, the MMA warp); read back with g4c_debug_profile().
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

static_assert(sizeof(Smem) <= 232448, "edge kernel shared memory exceeds the 227 KiB opt-in limit");

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
__device__ __forceinline__ void epi_sync_all() { asm volatile("bar.sync 5, 512;" ::: "memory"); }
// the four warps that share a TMEM lane quarter (one per column quarter)
__device__ __forceinline__ void quarter_sync(int lq) { asm volatile("bar.sync %0, 128;" ::"r"(1 + lq) : "memory"); }

// SELU(x) / lambda from t = x * log2(e)
__device__ __forceinline__ float selu_over_lambda_l2(float t) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    const float neg = fmaf(kSeluAlpha, e, -kSeluAlpha);
    return t > 0.f ? t * kLn2 : neg;
}

// largest in-degree over the (up to) 256 targets of unit pair `up`; executed by a full warp
__device__ __forceinline__ int pair_maxdeg(const EdgeArgs& a, int up, int lane) {
    if (a.fixed_k > 0) return a.fixed_k;
    int m = 0;
    const int64_t n0 = (int64_t)up * 256;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t n = n0 + i * 32 + lane;
        if (n < a.n_targets) m = max(m, a.rowptr[n + 1] - a.rowptr[n]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    return m;
}

struct RowMeta {
    int deg, base, trow;             // trow < 0: no such target
};
__device__ __forceinline__ RowMeta row_meta(const EdgeArgs& a, int64_t n) {
    RowMeta r{0, 0, -1};
    if (n < a.n_targets) {
        if (a.fixed_k > 0) { r.base = (int)(n * a.fixed_k); r.deg = a.fixed_k; }
        else { r.base = a.rowptr[n]; r.deg = a.rowptr[n + 1] - r.base; }
        r.trow = a.tgt_perm ? a.tgt_perm[n] : (int)n;
    }
    return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) edge_pair_kernel(const EdgeArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int tid = threadIdx.x, warp = tid >> 5;
    int lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int nl = a.n_layers;
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
        fence_barrier_init();
    }
    if (warp == W_MMA) { tmem_alloc<2>(&s.tmem_base, 512); tmem_relinquish<2>(); }
    if (tid == 0) {
        mbar_arrive_expect_tx(&s.w_full, (uint32_t)nl * 4 * HIMG);
        for (int l = 0; l < nl; ++l) bulk_g2s(s.w[l], a.W[l] + (size_t)rank * 4 * HIMG, 4 * HIMG, &s.w_full);
    }
    if (tid < H) {
        // per-column constants: layer 0's bias lives in P_c; hidden-layer biases are used in the log2 domain
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
    // 32-bit shared-space addresses of everything the inner loops touch (see lds_f4 in tc2_core.cuh)
    const uint32_t sb = smem_u32(smem_raw);
    const uint32_t a_cst = sb + (uint32_t)offsetof(Smem, cst), a_part = sb + (uint32_t)offsetof(Smem, part);
    const uint32_t a_in_ready = sb + (uint32_t)offsetof(Smem, in_ready), a_a_ready = sb + (uint32_t)offsetof(Smem, a_ready);
    const uint32_t a_d_free = sb + (uint32_t)offsetof(Smem, d_free), a_d_full = sb + (uint32_t)offsetof(Smem, d_full);

    if (warp < N_EPI_WARPS) {
        // ====================================================================== epilogue warps
        setmaxnreg_inc<kRegsEpi>();
        // the thread index goes through an (identity) shuffle here: ptxas otherwise rematerialises it with an S2R (a
        // ~30-cycle round trip) wherever a thread-index-derived address is needed in the inner loops
        const int etid = __shfl_sync(0xffffffffu, tid, tid & 31);
        lane = etid & 31;
        const int lq = (etid >> 5) & 3, cq = etid >> 7;
        const int row = lq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
        const uint32_t leader_a_ready0 = mapa(a_a_ready, 0);          // chain c: + 8 c
        const uint32_t my_cst = a_cst + 128u * cq, my_part = a_part + 4u * row;
        uint32_t n_dfull[2] = {0, 0};
        const bool has_ln = a.gamma != nullptr;
        int pbuf = 0;
        PROF_DECL
        PROF_START();

        for (int up = up0; up < n_up; up += up_stride) {
            const int maxdeg = pair_maxdeg(a, up, lane);
            const RowMeta rm = row_meta(a, ((int64_t)up * 2 + rank) * 128 + row);
            float agg[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) agg[i] = 0.f;

            for (int j0 = 0; j0 < maxdeg; j0 += 2) {
                const int nch = min(2, maxdeg - j0);
                for (int l = 0; l < nl; ++l) {
                    // scale of this layer's accumulator: 1/s, times lambda when its input was a deferred-lambda SELU
                    const float cl = a.inv_scale[l] * (l > 0 ? kSeluScale : 1.f);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c >= nch) continue;
                        const uint32_t d_addr = tmem + lane_base + 256u * c + 32u * cq;
                        // one warp watches the mbarrier; the other fifteen block on a hardware barrier (no issue slots)
                        if (warp == 0) mbar_wait_sleep_a(a_d_full + 8u * c, n_dfull[c] & 1);
                        ++n_dfull[c];
                        epi_sync_all();
                        tc_fence_after();
                        PROF_LAP(l < nl - 1 ? 0 : 1);        // waiting for the MMAs (hidden / last layer)
                        if (l < nl - 1) {
                            // ---- hidden layer: x' = SELU(acc * cl + b) / lambda as fp16 (hi, lo) A operand columns
                            const float c2 = cl * kLog2e;
#pragma unroll
                            for (int h16 = 0; h16 < 2; ++h16) {          // 16 columns at a time (register pressure: agg[] stays live)
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
                            // ---- last layer: LayerNorm, aggregation, store.  The accumulator is read once and released
                            // at once (the loaders may refill the chain while the rest of this epilogue runs); the row
                            // statistics (mean / M2 of each 16-column piece, combined with Chan's formula) cross the column
                            // quarters through shared memory and one 128-thread barrier.
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
                                const uint32_t pa = my_part + 4096u * pbuf;           // part[pbuf][0][0][row]
                                sts_f1(pa + 512u * cq, 0.5f * (mh[0] + mh[1]));                       // mean of this thread's 32 columns
                                sts_f1(pa + 2048u + 512u * cq, (M2h[0] + M2h[1]) + 8.f * dm * dm);     // their M2
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
                            const bool live = rm.trow >= 0 && j < rm.deg;
                            float* dst = nullptr;
                            if (live && a.e_out) {
                                const int slot = rm.base + j;
                                const int erow = a.edge_perm ? a.edge_perm[slot] : slot;
                                dst = a.e_out + (size_t)erow * H + cq * 32;
                            }
                            const bool selu_out = a.act_e_out == G4C_ACT_SELU;
#pragma unroll
                            for (int i8 = 0; i8 < 32; i8 += 8) {              // 8 columns at a time, in place: every chunk's store
                                float* o = y + i8;                            // reads its own registers (no write-after-read stall
                                if (has_ln) {                                 // behind the previous chunk's pending store)
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
                                    if (dst) {
                                        if (selu_out) {
#pragma unroll
                                            for (int u = 0; u < 8; ++u) o[u] = selu_fast(o[u]);
                                        }
                                        stg256(dst + i8, o);
                                    }
                                }
                            }
                            PROF_LAP(5);                     // last layer: normalise, aggregate, store
                        }
                    }
                }
            }
            // ---- aggregated messages of this unit
            if (rm.trow >= 0) {
                const float rc = (a.aggr == G4C_AGGR_MEAN) ? 1.f / (float)max(rm.deg, 1) : 1.f;
                float* dst = a.agg_out + (size_t)rm.trow * H + cq * 32;
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
        const int lw = (warp - W_LOAD0) & 3, hf = (warp - W_LOAD0) >> 2;     // lane quarter; which half of the column stages
        const float ps = a.p_scale;
        const uint32_t lane_base = (uint32_t)(lw * 32) << 16;
        const uint32_t ring0 = smem_u32(s.ring[warp - W_LOAD0][0]);
        const uint32_t leader_in_ready0 = mapa(a_in_ready, 0);        // chain c: + 8 c
        const int row_in_pair = (int)rank * 128 + lw * 32 + lane;
        const int sub = lane >> 2, piece = lane & 3;      // cp.async: 8 rows per instruction, 4 x 16 B per row piece
        mbar_wait(&s.w_full, 0);           // in_ready is only signalled once this CTA's weights have landed

        // ---- issue cursor: runs NSTG-1 stages ahead of the processing cursor
        int i_up = up0;
        int i_j = 0, i_cs = 0, i_maxdeg = 0, nx_erow = -1, nx_srow = -1;
        // what this lane copies in the current slot, for tile rows 8 i + sub: offsets of its 16-byte piece in the edge /
        // source / target rows, in units of 16 bytes (row * 32 + piece; one IMAD.WIDE per address), bit i of vmask = row exists
        uint32_t oe[4], os[4], ot[4], vmask = 0;
        RowMeta i_rm{0, 0, -1};
        bool i_live = false;
        auto load_idx = [&](int j, int& erow, int& srow) {      // storage row of the j-th in-edge of this lane's target, its source
            erow = -1; srow = -1;
            if (i_rm.trow >= 0 && j < i_rm.deg) {
                const int slot = i_rm.base + j;
                erow = a.edge_perm ? __ldg(a.edge_perm + slot) : slot;
                srow = __ldg(a.src + slot);
            }
        };
        auto spread_slot = [&](int erow, int srow) {
            vmask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int er = __shfl_sync(0xffffffffu, erow, 8 * i + sub);
                const int sr = __shfl_sync(0xffffffffu, srow, 8 * i + sub);
                const bool ok = er >= 0;
                vmask |= ok ? (1u << i) : 0u;
                oe[i] = ok ? (uint32_t)er * 32u + piece : 0u;
                os[i] = ok ? (uint32_t)sr * 32u + piece : 0u;
            }
        };
        auto seek_unit = [&]() {              // position on the first slot of the first non-empty unit pair at or after i_up
            i_live = false;
            while (i_up < n_up) {
                i_maxdeg = pair_maxdeg(a, i_up, lane);
                if (i_maxdeg > 0) { i_live = true; break; }
                i_up += up_stride;
            }
            if (i_live) {
                i_rm = row_meta(a, (int64_t)i_up * 256 + row_in_pair);
                i_j = 0; i_cs = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int tr = __shfl_sync(0xffffffffu, i_rm.trow, 8 * i + sub);
                    ot[i] = tr >= 0 ? (uint32_t)tr * 32u + piece : 0u;
                }
                int er, sr;
                load_idx(0, er, sr);
                spread_slot(er, sr);
            }
        };
        auto issue_stage = [&](uint32_t stage_addr) {     // cp.async the 16-column stage (i_up, i_j, i_cs) of this warp's rows
            if (i_live) {
                const uint32_t dst0 = stage_addr + (uint32_t)sub * PITCH + piece * 16;
                const int colb = (2 * i_cs + hf) * (SCOLS * 4);             // byte offset of this stage's columns in a row
                const char* be = reinterpret_cast<const char*>(a.e_in) + colb;
                const char* br = reinterpret_cast<const char*>(a.P_r) + colb;
                const char* bc = reinterpret_cast<const char*>(a.P_c) + colb;
                const float* pe[4];
                const float* pr[4];
                const float* pc[4];
                uint32_t sz[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sz[i] = (vmask >> i) & 1u ? 16u : 0u;        // 0: nothing is read, the destination is zero-filled
                    pe[i] = reinterpret_cast<const float*>(be + (size_t)oe[i] * 16);
                    pr[i] = reinterpret_cast<const float*>(br + (size_t)os[i] * 16);
                    pc[i] = reinterpret_cast<const float*>(bc + (size_t)ot[i] * 16);
                }
                // one asm statement: the twelve addresses are live in distinct registers, so the copies issue back to
                // back (address registers recycled between consecutive LDGSTS stall on their release by the LSU)
                asm volatile(
                    "cp.async.cg.shared.global [%0], [%1], 16, %13;\n\t"
                    "cp.async.cg.shared.global [%0 + 2560], [%2], 16, %13;\n\t"
                    "cp.async.cg.shared.global [%0 + 5120], [%3], 16, %13;\n\t"
                    "cp.async.cg.shared.global [%0 + 640], [%4], 16, %14;\n\t"
                    "cp.async.cg.shared.global [%0 + 3200], [%5], 16, %14;\n\t"
                    "cp.async.cg.shared.global [%0 + 5760], [%6], 16, %14;\n\t"
                    "cp.async.cg.shared.global [%0 + 1280], [%7], 16, %15;\n\t"
                    "cp.async.cg.shared.global [%0 + 3840], [%8], 16, %15;\n\t"
                    "cp.async.cg.shared.global [%0 + 6400], [%9], 16, %15;\n\t"
                    "cp.async.cg.shared.global [%0 + 1920], [%10], 16, %16;\n\t"
                    "cp.async.cg.shared.global [%0 + 4480], [%11], 16, %16;\n\t"
                    "cp.async.cg.shared.global [%0 + 7040], [%12], 16, %16;\n"
                    ::"r"(dst0), "l"(pe[0]), "l"(pr[0]), "l"(pc[0]), "l"(pe[1]), "l"(pr[1]), "l"(pc[1]), "l"(pe[2]), "l"(pr[2]),
                    "l"(pc[2]), "l"(pe[3]), "l"(pr[3]), "l"(pc[3]), "r"(sz[0]), "r"(sz[1]), "r"(sz[2]), "r"(sz[3])
                    : "memory");
                static_assert(ARR == 2560 && 8 * PITCH == 640, "offsets in the cp.async block above");
                // advance to the next stage in execution order
                if (++i_cs == 1) load_idx(i_j + 1, nx_erow, nx_srow);       // indices of the next slot: three stages of slack
                if (i_cs == NCS_W) {
                    i_cs = 0;
                    if (++i_j < i_maxdeg) spread_slot(nx_erow, nx_srow);
                    else { i_up += up_stride; seek_unit(); }
                }
            }
            cp_async_commit();              // one (possibly empty) group per stage keeps the wait depth constant
        };

        seek_unit();
#pragma unroll 1
        for (int p = 0; p < NSTG - 1; ++p) issue_stage(ring0 + p * STG);
        PROF_DECL
        PROF_START();
        uint32_t q = 0, n_slot0 = 0, n_slot1 = 0;      // slots completed per chain (scalars: the chain index is dynamic here)
        for (int up = up0; up < n_up; up += up_stride) {
            const int maxdeg = pair_maxdeg(a, up, lane);
            for (int j = 0; j < maxdeg; ++j) {
                const int cc = j & 1;
                const uint32_t d_col = tmem + lane_base + 256u * cc;
#pragma unroll 1
                for (int cw = 0; cw < NCS_W; ++cw, ++q) {
                    const int cs = 2 * cw + hf;
                    cp_async_wait<NSTG - 2>();                   // stage q has landed
                    __syncwarp();
                    PROF_LAP(0);                                 // waiting for the staged rows
                    // prefetch this warp's next stage into the other buffer (released by the __syncwarp that ended the
                    // previous iteration): it stays in flight for the whole processing of stage q
                    issue_stage(ring0 + ((q + NSTG - 1) % NSTG) * STG);
                    if (cw == 0) {
                        mbar_wait_sleep_a(a_d_free + 8u * cc, ((cc ? n_slot1 : n_slot0) + 1) & 1);      // last-layer epilogue of the previous slot on this chain
                        tc_fence_after();
                        PROF_LAP(1);                             // waiting for the accumulator to be released
                    }
                    const uint32_t st = ring0 + (q % NSTG) * STG + lane * PITCH;
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {             // 8 columns at a time
                        float4 xe[2], xr[2], xc[2];
#pragma unroll
                        for (int v4 = 0; v4 < 2; ++v4) {
                            xe[v4] = lds_f4(st + h8 * 32 + v4 * 16);
                            xr[v4] = lds_f4(st + ARR + h8 * 32 + v4 * 16);
                            xc[v4] = lds_f4(st + 2 * ARR + h8 * 32 + v4 * 16);
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
        PROF_FLUSH(8, warp == W_LOAD0);
    } else {
        setmaxnreg_dec<kRegsMisc>();
        if (warp == W_MMA && rank == 0) {
            // ================================================================== MMA issuer (leader CTA)
            // both CTAs' weights are in place once in_ready completes (every loader warp waited on its w_full)
            const uint32_t idesc = idesc_f16(256, 128);
            // descriptors differ only in their 14-bit start-address field (byte address >> 4): add offsets to a base
            const uint64_t w_desc = make_desc_sw128(smem_u32(s.w[0]));
            uint32_t n_chain[2] = {0, 0}, n_ar[2] = {0, 0};
            PROF_DECL
            PROF_START();
            for (int up = up0; up < n_up; up += up_stride) {
                const int maxdeg = pair_maxdeg(a, up, lane);
                for (int j0 = 0; j0 < maxdeg; j0 += 2) {
                    const int nch = min(2, maxdeg - j0);
                    for (int l = 0; l < nl; ++l) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            if (c >= nch) continue;
                            const uint32_t d_col = tmem + 256u * c, ah = d_col + 128u, al = d_col + 192u;
                            if (l == 0) {
                                if (lane == 0) {
                                    mbar_wait_sleep_a(a_in_ready + 8u * c, n_chain[c] & 1);
                                    tc_fence_after();
                                    PROF_LAP(0);             // waiting for the loaders
                                }
                                ++n_chain[c];
                            } else {
                                if (lane == 0) {
                                    mbar_wait_sleep_a(a_a_ready + 8u * c, n_ar[c] & 1);
                                    tc_fence_after();
                                    PROF_LAP(1);             // waiting for the epilogue
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
                                PROF_LAP(2);                 // issuing
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            PROF_FLUSH(16, true);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == W_MMA) tmem_dealloc<2>(tmem, 512);
}

}  // namespace ep

int edge_pair_profile(unsigned long long* out64) {
#ifdef G4C_PROFILE
    if (cudaMemcpyFromSymbol(out64, ep::g_prof, sizeof(unsigned long long) * 64) != cudaSuccess) return check_launch("profile read");
    unsigned long long zero[64] = {0};
    cudaMemcpyToSymbol(ep::g_prof, zero, sizeof(zero));
    return G4C_OK;
#else
    (void)out64;
    set_error("libg4c was built without -DG4C_PROFILE");
    return G4C_EUNSUPPORTED;
#endif
}

int edge_pair_launch(const EdgeArgs& a, cudaStream_t st) {
    // variant: 0 = pick (v5 for fixed in-degree launches without permutations, this kernel otherwise), 1 = this kernel, 2 = v5
    if (a.variant == G4C_EDGE_V5 || (a.variant == G4C_EDGE_AUTO && edge_v5_supported(a))) return edge_v5_launch(a, st);
    if (a.variant != G4C_EDGE_AUTO && a.variant != G4C_EDGE_V3) { set_error("g4c_edge_aggr_fwd: variant=%d (0..2)", a.variant); return G4C_EINVAL; }
    if (a.act_e_out != G4C_ACT_NONE && a.act_e_out != G4C_ACT_SELU) {
        // the models only ever apply F.selu to a block's edge / angle output (nn/mus_gnn.py:321, nn/remus_gnn.py:143)
        set_error("g4c_edge_aggr_fwd: act_e_out must be none or selu");
        return G4C_EUNSUPPORTED;
    }
    static int configured[kMaxDevices] = {0};
    const int smem = (int)sizeof(ep::Smem);
    if (!ensure_dynamic_smem(ep::edge_pair_kernel, smem, configured)) return check_launch("edge_pair_kernel attribute");
    const int n_sm = device_sms();
    const int64_t n_units = (a.n_targets + 127) / 128, n_up = (n_units + 1) / 2;
    const int pairs = (int)std::min<int64_t>(n_up, n_sm / 2);
    ep::edge_pair_kernel<<<2 * pairs, ep::NT, smem, st>>>(a);
    count_launch(1, true);
    return check_launch("edge_pair_kernel");
}

}  // namespace g4c
