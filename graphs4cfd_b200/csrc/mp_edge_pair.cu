// mp_edge_pair.cu — fused edge-MLP + aggregation kernel of the message-passing block, hidden = 128,
// on CTA pairs (cta_group::2) with every operand that is re-used kept on chip.
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
// epilogue thread that owns the row (fixed order, no atomics).
//
// Per CTA (512 threads):
//   warps 0-7   epilogue: thread (row, half) owns 64 columns of its row.  TMEM -> registers, bias, SELU,
//               fp16 (hi, lo) split written back to TMEM as the next layer's A operand (tcgen05.st); last
//               layer: LayerNorm, aggregation, activation, 256-bit stores of e'.
//   warps 8-11  loaders, one per TMEM lane quarter (32 tile rows each).  Rows of e, P_r[src], P_c[tgt] are
//               fetched with cp.async (16 B per lane, whole 128-byte row pieces, no registers held across the
//               HBM/L2 latency) into a private two-stage ring of 32-column stages; row pieces are stored at a
//               144-byte pitch so that "lane = row" 16-byte reads are bank-conflict free.  The warp then reads
//               its rows back with lane = row, splits e into fp16 (hi, lo) and writes it with tcgen05.st as the
//               layer-1 A operand, and writes (P_r + P_c) * s as the INITIAL VALUE of the accumulator.
//   warp 12     (leader CTA) issues every tcgen05.mma of the pair; M = 256, N = 128, the B
//               operand (weights) is resident in shared memory, each CTA holding 64 of the 128 output rows
//               of all layers as pre-split, pre-swizzled fp16 (hi, lo) images (96 KiB).
// Two chains (even / odd slots) alternate so that the MMAs of one overlap the epilogue of the other.
// TMEM: chain c uses columns [256c, 256c+128) accumulator, [256c+128, +64) A hi, [256c+192, +64) A lo.
#include <algorithm>
#include "pair_common.cuh"

namespace g4c {
namespace ep {

using namespace tc2;
using namespace pairk;

constexpr int PITCH = 144;               // bytes between staged 128-byte row pieces (36 words: lanes r..r+7 hit 8 distinct 16 B bank groups)
constexpr int ARR = 32 * PITCH;          // one array's 32 row pieces of a stage
constexpr int STG = 3 * ARR;             // stage = e | P_r | P_c pieces of the warp's 32 rows, 32 columns
constexpr int NSTG = 2;

struct Smem {
    uint8_t w[3][4 * HIMG];          // layer l: K-block 0 hi | lo, K-block 1 hi | lo (64-row images)
    uint8_t ring[4][NSTG][STG];      // per loader warp
    float part[2][2][128];           // LayerNorm partial sums [pass][half][row]
    uint64_t w_full;
    uint64_t in_ready[2];            // leader: A operand + initial accumulator of chain c written (8 loader warps of the pair)
    uint64_t a_ready[2];             // leader: next layer's A operand written (16 epilogue warps of the pair)
    uint64_t d_free[2];              // local: accumulator of chain c has been read by the last-layer epilogue (8 warps)
    uint64_t d_full[2];              // local, multicast commit
    uint32_t tmem_base;
};
// ---- optional in-kernel phase profile (make EXTRA=-DG4C_PROFILE): cycles spent per role and phase by CTA 0,
// accumulated by lane 0 of each warp; read back with g4c_debug_profile().
#ifdef G4C_PROFILE
__device__ unsigned long long g_prof[64];
#define PROF_DECL unsigned int prof_t0 = 0; unsigned long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
#define PROF_START() prof_t0 = clock()
#define PROF_LAP(i) do { const unsigned int t1 = clock(); prof_acc[i] += (unsigned int)(t1 - prof_t0); prof_t0 = t1; } while (0)
#define PROF_FLUSH(base) do { if (blockIdx.x == 0 && lane == 0) for (int i = 0; i < 6; ++i) atomicAdd(&g_prof[(base) + i], prof_acc[i]); } while (0)
#else
#define PROF_DECL
#define PROF_START()
#define PROF_LAP(i)
#define PROF_FLUSH(base)
#endif

static_assert(sizeof(Smem) <= 232448, "edge kernel shared memory exceeds the 227 KiB opt-in limit");

// largest in-degree over the (up to) 256 targets of unit pair `up`; executed by a full warp
__device__ __forceinline__ int pair_maxdeg(const EdgeArgs& a, int64_t up, int lane) {
    if (a.fixed_k > 0) return a.fixed_k;
    int m = 0;
    const int64_t n0 = up * 256;
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
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int nl = a.n_layers;
    const int64_t n_units = (a.n_targets + 127) / 128;
    const int64_t n_up = (n_units + 1) / 2;
    const int64_t up0 = blockIdx.x >> 1, up_stride = gridDim.x >> 1;

    if (tid == 0) {
        mbar_init(&s.w_full, 1);
        for (int c = 0; c < 2; ++c) {
            mbar_init(&s.in_ready[c], 8);
            mbar_init(&s.a_ready[c], 16);
            mbar_init(&s.d_free[c], 8);
            mbar_init(&s.d_full[c], 1);
        }
        fence_barrier_init();
    }
    if (warp == 12) { tmem_alloc<2>(&s.tmem_base, 512); tmem_relinquish<2>(); }
    if (tid == 0) {
        mbar_arrive_expect_tx(&s.w_full, (uint32_t)nl * 4 * HIMG);
        for (int l = 0; l < nl; ++l) bulk_g2s(s.w[l], a.W[l] + (size_t)rank * 4 * HIMG, 4 * HIMG, &s.w_full);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;

    if (warp < 8) {
        // ====================================================================== epilogue warps
        setmaxnreg_inc<kRegsEpi>();
        const int row = (warp & 3) * 32 + lane, half = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t leader_a_ready[2] = {mapa(smem_u32(&s.a_ready[0]), 0), mapa(smem_u32(&s.a_ready[1]), 0)};
        uint32_t n_dfull[2] = {0, 0};
        PROF_DECL
        PROF_START();
        const float* gamma = a.gamma ? a.gamma + half * 64 : nullptr;
        const float* beta = a.beta ? a.beta + half * 64 : nullptr;

        for (int64_t up = up0; up < n_up; up += up_stride) {
            const int maxdeg = pair_maxdeg(a, up, lane);
            const RowMeta rm = row_meta(a, (up * 2 + rank) * 128 + row);
            float agg[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) agg[i] = 0.f;

            for (int j0 = 0; j0 < maxdeg; j0 += 2) {
                const int nch = min(2, maxdeg - j0);
                for (int l = 0; l < nl; ++l) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c >= nch) continue;
                        const uint32_t d_addr = tmem + lane_base + 256u * c + 64u * half;
                        mbar_wait(&s.d_full[c], n_dfull[c] & 1);
                        ++n_dfull[c];
                        tc_fence_after();
                        PROF_LAP(0);                     // waiting for MMA completion
                        if (l < nl - 1) {
                            epilogue_hidden(d_addr, tmem + lane_base + 256u * c + 128u + 32u * half,
                                            tmem + lane_base + 256u * c + 192u + 32u * half, a.inv_scale[l],
                                            l == 0 ? nullptr : a.bias[l] + half * 64);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(leader_a_ready[c]);
                            PROF_LAP(1);                 // hidden epilogue
                        } else {
                            // ---- last layer: LayerNorm, aggregation, store
                            float y[64];
                            const float inv = a.inv_scale[l];
                            const float* bias = (l == 0) ? nullptr : a.bias[l] + half * 64;
                            tmem_ld32f(d_addr, y);
                            tmem_ld32f(d_addr + 32, y + 32);
#pragma unroll
                            for (int i = 0; i < 64; i += 4) {
                                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + i));
                                y[i] = fmaf(y[i], inv, b.x);
                                y[i + 1] = fmaf(y[i + 1], inv, b.y);
                                y[i + 2] = fmaf(y[i + 2], inv, b.z);
                                y[i + 3] = fmaf(y[i + 3], inv, b.w);
                            }
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&s.d_free[c]);
                            PROF_LAP(2);                 // last epilogue up to the release of the accumulator
                            if (gamma) {
                                float sum = 0.f;
#pragma unroll
                                for (int i = 0; i < 64; ++i) sum += y[i];
                                s.part[0][half][row] = sum;
                                epi_sync();
                                const float mean = (s.part[0][0][row] + s.part[0][1][row]) * (1.f / H);
                                float sq = 0.f;
#pragma unroll
                                for (int i = 0; i < 64; ++i) {
                                    y[i] -= mean;
                                    sq = fmaf(y[i], y[i], sq);
                                }
                                s.part[1][half][row] = sq;
                                epi_sync();
                                const float rstd = 1.f / sqrtf((s.part[1][0][row] + s.part[1][1][row]) * (1.f / H) + kLnEps);
#pragma unroll
                                for (int i = 0; i < 64; i += 4) {
                                    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + i));
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + i));
                                    y[i] = fmaf(y[i] * rstd, g.x, b.x);
                                    y[i + 1] = fmaf(y[i + 1] * rstd, g.y, b.y);
                                    y[i + 2] = fmaf(y[i + 2] * rstd, g.z, b.z);
                                    y[i + 3] = fmaf(y[i + 3] * rstd, g.w, b.w);
                                }
                            }
                            const int j = j0 + c;
                            if (rm.trow >= 0 && j < rm.deg) {
#pragma unroll
                                for (int i = 0; i < 64; ++i) agg[i] += y[i];
                                if (a.e_out) {
                                    const int slot = rm.base + j;
                                    const int erow = a.edge_perm ? a.edge_perm[slot] : slot;
                                    float* dst = a.e_out + (size_t)erow * H + half * 64;
#pragma unroll
                                    for (int i = 0; i < 64; i += 8) {
                                        float o[8];
#pragma unroll
                                        for (int u = 0; u < 8; ++u) o[u] = apply_act_fast(y[i + u], a.act_e_out);
                                        stg256(dst + i, o);
                                    }
                                }
                            }
                            PROF_LAP(3);                 // LayerNorm, aggregation, stores
                        }
                    }
                }
            }
            // ---- aggregated messages of this unit
            if (rm.trow >= 0) {
                const float cnt = (a.aggr == G4C_AGGR_MEAN) ? (float)max(rm.deg, 1) : 1.f;
                float* dst = a.agg_out + (size_t)rm.trow * H + half * 64;
#pragma unroll
                for (int i = 0; i < 64; i += 8) {
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] = agg[i + u] / cnt;
                    stg256(dst + i, o);
                }
            }
            PROF_LAP(4);
        }
        PROF_FLUSH(warp < 4 ? 0 : 8);
    } else if (warp < 12) {
        // ====================================================================== loader warps
        setmaxnreg_dec<kRegsLoad>();
        const int lw = warp - 8;
        const float ps = a.p_scale;
        const uint32_t lane_base = (uint32_t)(lw * 32) << 16;
        const uint32_t ring0 = smem_u32(s.ring[lw][0]);
        const uint32_t leader_in_ready[2] = {mapa(smem_u32(&s.in_ready[0]), 0), mapa(smem_u32(&s.in_ready[1]), 0)};
        const int64_t row_in_pair = (int64_t)rank * 128 + lw * 32 + lane;
        mbar_wait(&s.w_full, 0);           // in_ready is only signalled once this CTA's weights have landed

        // issue cursor (runs one stage ahead of the processing cursor); lane = tile row
        int64_t i_up = up0;
        int i_j = 0, i_cs = 0, i_maxdeg = 0, i_erow = -1, i_srow = -1, nx_erow = -1, nx_srow = -1;
        RowMeta i_rm{0, 0, -1};
        bool i_live = false;
        auto issue_seek_unit = [&]() {        // position on the first slot of the first non-empty unit at or after i_up
            i_live = false;
            while (i_up < n_up) {
                i_maxdeg = pair_maxdeg(a, i_up, lane);
                if (i_maxdeg > 0) { i_live = true; break; }
                i_up += up_stride;
            }
            if (i_live) { i_rm = row_meta(a, i_up * 256 + row_in_pair); i_j = 0; i_cs = 0; }
        };
        auto load_idx = [&](int j, int& erow, int& srow) {      // storage row of the j-th in-edge of this lane's target, its source
            erow = -1; srow = -1;
            if (i_rm.trow >= 0 && j < i_rm.deg) {
                const int slot = i_rm.base + j;
                erow = a.edge_perm ? __ldg(a.edge_perm + slot) : slot;
                srow = __ldg(a.src + slot);
            }
        };
        auto issue_load_idx = [&]() { load_idx(i_j, i_erow, i_srow); };
        auto issue_stage = [&](uint32_t stage_addr) {     // cp.async the 32-column stage (i_up, i_j, i_cs) of this warp's rows
            const int sub = lane >> 3, piece = lane & 7;   // 4 rows per instruction, 8 x 16 B per row piece
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + sub;
                const int er = __shfl_sync(0xffffffffu, i_erow, rr);
                const int sr = __shfl_sync(0xffffffffu, i_srow, rr);
                const int tr = __shfl_sync(0xffffffffu, i_rm.trow, rr);
                const bool ok = er >= 0;
                const uint32_t dst = stage_addr + (uint32_t)rr * PITCH + piece * 16;
                const size_t col = (size_t)i_cs * 32 + piece * 4;
                cp_async16_zfill(dst, a.e_in + (ok ? (size_t)er * H : 0) + col, ok);
                cp_async16_zfill(dst + ARR, a.P_r + (ok ? (size_t)sr * H : 0) + col, ok);
                cp_async16_zfill(dst + 2 * ARR, a.P_c + (ok ? (size_t)tr * H : 0) + col, ok);
            }
            cp_async_commit();
        };
        auto issue_advance = [&]() {          // next stage in execution order
            if (++i_cs < 4) {
                if (i_cs == 1) load_idx(i_j + 1, nx_erow, nx_srow);     // indices of the next slot: three stages of slack
                return;
            }
            i_cs = 0;
            if (++i_j < i_maxdeg) { i_erow = nx_erow; i_srow = nx_srow; return; }
            i_up += up_stride;
            issue_seek_unit();
            if (i_live) issue_load_idx();
        };

        issue_seek_unit();
        if (i_live) { issue_load_idx(); issue_stage(ring0); issue_advance(); }
        uint32_t q = 0, n_slot[2] = {0, 0};
        PROF_DECL
        PROF_START();
        for (int64_t up = up0; up < n_up; up += up_stride) {
            const int maxdeg = pair_maxdeg(a, up, lane);
            for (int j = 0; j < maxdeg; ++j) {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    if ((j & 1) != cc) continue;
                    const uint32_t d_col = tmem + lane_base + 256u * cc;
#pragma unroll 1
                    for (int cs = 0; cs < 4; ++cs, ++q) {
                        // prefetch the next stage, then wait for this one
                        if (i_live) { issue_stage(ring0 + ((q + 1) & 1) * STG); issue_advance(); PROF_LAP(0); cp_async_wait<1>(); }
                        else cp_async_wait<0>();
                        __syncwarp();
                        PROF_LAP(1);                     // waiting for the staged rows
                        if (cs == 0) {
                            mbar_wait(&s.d_free[cc], (n_slot[cc] + 1) & 1);      // last-layer epilogue of the previous slot on this chain
                            tc_fence_after();
                            PROF_LAP(2);                 // waiting for the accumulator to be released
                        }
                        const uint8_t* st = s.ring[lw][q & 1] + lane * PITCH;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {                          // 16 columns at a time
                            uint32_t eh[8], el[8], pp[16];
#pragma unroll
                            for (int v4 = 0; v4 < 4; ++v4) {
                                const float4 xe = *reinterpret_cast<const float4*>(st + hh * 64 + v4 * 16);
                                const float4 xr = *reinterpret_cast<const float4*>(st + ARR + hh * 64 + v4 * 16);
                                const float4 xc = *reinterpret_cast<const float4*>(st + 2 * ARR + hh * 64 + v4 * 16);
                                split2(xe.x, xe.y, eh[2 * v4], el[2 * v4]);
                                split2(xe.z, xe.w, eh[2 * v4 + 1], el[2 * v4 + 1]);
                                pp[4 * v4] = __float_as_uint((xr.x + xc.x) * ps);
                                pp[4 * v4 + 1] = __float_as_uint((xr.y + xc.y) * ps);
                                pp[4 * v4 + 2] = __float_as_uint((xr.z + xc.z) * ps);
                                pp[4 * v4 + 3] = __float_as_uint((xr.w + xc.w) * ps);
                            }
                            tmem_st8(d_col + 128u + 16u * cs + 8u * hh, eh);       // A hi: k = 32 cs + 16 hh .. +15
                            tmem_st8(d_col + 192u + 16u * cs + 8u * hh, el);       // A lo
                            tmem_st16(d_col + 32u * cs + 16u * hh, pp);            // accumulator columns 32 cs + 16 hh .. +15
                        }
                        if (cs == 3) {
                            tmem_wait_st();
                            tc_fence_before();
                        }
                        __syncwarp();                                             // stage buffer may be refilled
                        if (cs == 3) {
                            if (lane == 0) mbar_arrive_cluster(leader_in_ready[cc]);
                            ++n_slot[cc];
                        }
                        PROF_LAP(3);                     // split / add / TMEM writes
                    }
                }
            }
        }
        PROF_FLUSH(16);
    } else {
        setmaxnreg_dec<kRegsMisc>();
        if (warp == 12 && rank == 0) {
            // ================================================================== MMA / copy issuer (leader CTA)
            // both CTAs' weights are in place once in_ready completes (every loader warp waited on its w_full)
            const uint32_t idesc = idesc_f16(256, 128);
            // descriptors differ only in their 14-bit start-address field (byte address >> 4): add offsets to a base
            const uint64_t w_desc = make_desc_sw128(smem_u32(s.w[0]));
            uint32_t n_chain[2] = {0, 0}, n_ar[2] = {0, 0};
            PROF_DECL
            PROF_START();
            for (int64_t up = up0; up < n_up; up += up_stride) {
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
                                    mbar_wait<true>(&s.in_ready[c], n_chain[c] & 1);
                                    tc_fence_after();
                                    PROF_LAP(0);         // waiting for the loaders
                                }
                                ++n_chain[c];
                            } else {
                                if (lane == 0) {
                                    mbar_wait<true>(&s.a_ready[c], n_ar[c] & 1);
                                    tc_fence_after();
                                    PROF_LAP(1);         // waiting for the epilogue
                                }
                                ++n_ar[c];
                            }
                            if (lane == 0) {
                                const uint64_t wb = w_desc + (uint64_t)((l * 4 * HIMG) >> 4);
#pragma unroll 1
                                for (int ks = 0; ks < 8; ++ks) {
                                    const uint64_t wh = wb + (uint64_t)(((ks >> 2) * 2 * HIMG + (ks & 3) * 32) >> 4);
                                    const uint64_t wl = wh + (uint64_t)(HIMG >> 4);
                                    umma_ts<2>(d_col, ah + 8 * ks, wh, idesc, (l == 0 || ks > 0) ? 1u : 0u);
                                    umma_ts<2>(d_col, al + 8 * ks, wh, idesc, 1u);
                                    umma_ts<2>(d_col, ah + 8 * ks, wl, idesc, 1u);
                                }
                                umma_commit<2>(&s.d_full[c], 3);
                                PROF_LAP(2);             // issuing
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            PROF_FLUSH(24);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 12) tmem_dealloc<2>(tmem, 512);
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
    static bool configured = false;
    const int smem = (int)sizeof(ep::Smem);
    if (!configured) {
        if (cudaFuncSetAttribute(ep::edge_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return check_launch("edge_pair_kernel attribute");
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
    ep::edge_pair_kernel<<<2 * pairs, ep::NT, smem, st>>>(a);
    count_launch();
    return check_launch("edge_pair_kernel");
}

}  // namespace g4c
