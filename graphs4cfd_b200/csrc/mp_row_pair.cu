// mp_row_pair.cu — row-tile MLP on CTA pairs (cta_group::2), hidden = 128, fp16x3: the node model of the
// message-passing block (blocks.py:185), encoders / decoders, the DownMP / UpMP MLPs, and the bare Linear that
// makes the per-node products P_r, P_c of the split edge model (g4c_rowmlp_tc_fwd).
//
// Same machinery as mp_edge_pair.cu — weights resident in shared memory as half images per CTA, A operands and
// accumulators in TMEM, loader warps (cp.async rings, lane = row read-back, tcgen05.st), two chains alternating so the
// MMAs of one tile overlap the epilogue of the other — with these differences:
//   * a "slot" is a pair-tile (256 consecutive rows, 128 per CTA); the tiles a pair owns alternate chains;
//   * layer 1 consumes the concatenated input K-block by K-block (64 columns each, or one K = 16 step for a
//     narrow segment): K-block b is written by the loader warps into half b & 1 of the chain's A operand columns
//     (rows fetched with cp.async into a private ring of 16-column stages, read back with lane = row, split into
//     fp16 (hi, lo), tcgen05.st), so any number of segments streams through the same 64 TMEM columns of A;
//   * the last layer is either 128 wide (optional LayerNorm, activation, 256-bit row stores) or narrower than
//     16 (an N = 16 MMA; bias, optional residual, scalar stores: the decoder, nn/mus_gnn.py:369-373).
#include <algorithm>
#include "pair_common.cuh"

namespace g4c {
namespace rp {

using namespace tc2;
using namespace pairk;

constexpr int MAX_KB = 6;
constexpr int SCOLS = 16;               // columns per loader stage
constexpr int PITCH = 80;               // bytes between staged 64-byte row pieces (conflict-free lane = row 16-byte reads)
constexpr int STG = 32 * PITCH;         // one stage: 32 row pieces
constexpr int NSTG = 4;                 // stages per loader warp
constexpr int RING_BYTES = N_LOAD_WARPS * NSTG * STG;

struct KBlock {
    const float* ptr;
    const int32_t* gather;
    int32_t stride, col0, width;        // width 64 (half of a wide segment) or 1..16 (narrow segment)
    float scale;
};

struct Args {
    G4cRowTcDesc d;
    KBlock kb[MAX_KB];
    int32_t n_kb, _pad;
    uint32_t w_off[3];                  // byte offset of each layer's images in the weight region
    uint32_t w_bytes, ring_off, tail_off;
    int64_t n_pt;                       // pair-tiles
};

struct Tail {
    float part[4][128];                 // LayerNorm partials [column quarter][row] (sums, then centred squares)
    uint64_t w_full;
    uint64_t full[4], empty[4];         // per A-operand half (2 * chain + half): full = leader, the 8 loader warps of the pair
                                        // that wrote it; empty = local, multicast commit of the MMAs that read it
    uint64_t a_ready[2], d_free[2];     // leader, 32 epilogue warps of the pair
    uint64_t d_full[2];                 // local, multicast commit
    uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) row_pair_kernel(const Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    Tail& s = *reinterpret_cast<Tail*>(smem + a.tail_off);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const G4cRowTcDesc& d = a.d;
    const int nl = d.n_layers, NKB = a.n_kb;
    const bool narrow_out = d.out_width != H;
    const int64_t pt0 = blockIdx.x >> 1, pt_stride = gridDim.x >> 1;

    if (tid == 0) {
        mbar_init(&s.w_full, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&s.full[i], 8); mbar_init(&s.empty[i], 1); }
        for (int c = 0; c < 2; ++c) {
            mbar_init(&s.a_ready[c], 2 * N_EPI_WARPS);
            mbar_init(&s.d_free[c], 2 * N_EPI_WARPS);
            mbar_init(&s.d_full[c], 1);
        }
        fence_barrier_init();
    }
    if (warp == W_MMA) { tmem_alloc<2>(&s.tmem_base, 512); tmem_relinquish<2>(); }
    if (tid == 0) {
        // this CTA's half of every layer: layer l is stored [rank][...] in global memory
        mbar_arrive_expect_tx(&s.w_full, a.w_bytes);
        for (int l = 0; l < nl; ++l) {
            const uint32_t bytes = (l + 1 < nl ? a.w_off[l + 1] : a.w_bytes) - a.w_off[l];
            bulk_g2s(smem + a.w_off[l], d.W[l] + (size_t)rank * bytes, bytes, &s.w_full);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;

    if (warp < N_EPI_WARPS) {
        // ====================================================================== epilogue warps: thread (row, column quarter)
        setmaxnreg_inc<kRegsEpi>();
        const int lq = warp & 3, cq = warp >> 2;
        const int row = lq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
        const uint32_t leader_a_ready[2] = {mapa(smem_u32(&s.a_ready[0]), 0), mapa(smem_u32(&s.a_ready[1]), 0)};
        const uint32_t leader_d_free[2] = {mapa(smem_u32(&s.d_free[0]), 0), mapa(smem_u32(&s.d_free[1]), 0)};
        uint32_t n_dfull[2] = {0, 0};
        const float* gamma = d.gamma ? d.gamma + cq * 32 : nullptr;
        const float* beta = d.beta ? d.beta + cq * 32 : nullptr;

        for (int64_t ptb = pt0; ptb < a.n_pt; ptb += 2 * pt_stride) {
            const int nch = (ptb + pt_stride < a.n_pt) ? 2 : 1;
            for (int l = 0; l < nl; ++l) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c >= nch) continue;
                    const int64_t R = ((ptb + c * pt_stride) * 2 + rank) * 128 + row;     // this thread's global row
                    const uint32_t d_addr = tmem + lane_base + 256u * c + 32u * cq;
                    mbar_wait_sleep(&s.d_full[c], n_dfull[c] & 1);
                    ++n_dfull[c];
                    tc_fence_after();
                    const float inv = d.inv_scale[l];
                    const float* bias = d.bias[l] + cq * 32;
                    if (l < nl - 1 && d.dual) {
                        // ---- first of two bare Linears of the same input: store it, keep the A operand, hand the accumulator back
                        float y[32];
                        tmem_ld32f(d_addr, y);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(leader_a_ready[c]);
                        if (R < d.rows) {
                            float* dst = d.out + (size_t)R * d.out_stride + cq * 32;
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + i));
                                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + i + 4));
                                float o[8] = {fmaf(y[i], inv, b0.x), fmaf(y[i + 1], inv, b0.y), fmaf(y[i + 2], inv, b0.z), fmaf(y[i + 3], inv, b0.w),
                                              fmaf(y[i + 4], inv, b1.x), fmaf(y[i + 5], inv, b1.y), fmaf(y[i + 6], inv, b1.z), fmaf(y[i + 7], inv, b1.w)};
                                stg256(dst + i, o);
                            }
                        }
                    } else if (l < nl - 1) {
                        // ---- hidden layer: x = selu(acc * inv + bias) as fp16 (hi, lo) A operand columns, 16 columns at a time
#pragma unroll
                        for (int h16 = 0; h16 < 2; ++h16) {
                            float v[16];
                            tmem_ld16f(d_addr + 16u * h16, v);
                            uint32_t hi[8], lo[8];
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + 16 * h16 + i));
                                const float x0 = selu_fast(fmaf(v[i], inv, b.x)), x1 = selu_fast(fmaf(v[i + 1], inv, b.y));
                                const float x2 = selu_fast(fmaf(v[i + 2], inv, b.z)), x3 = selu_fast(fmaf(v[i + 3], inv, b.w));
                                split2(x0, x1, hi[i / 2], lo[i / 2]);
                                split2(x2, x3, hi[i / 2 + 1], lo[i / 2 + 1]);
                            }
                            tmem_st8(tmem + lane_base + 256u * c + 128u + 16u * cq + 8u * h16, hi);
                            tmem_st8(tmem + lane_base + 256u * c + 192u + 16u * cq + 8u * h16, lo);
                        }
                        tmem_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(leader_a_ready[c]);
                    } else if (narrow_out) {
                        // ---- last layer narrower than 16: columns 0..out_width-1 of the N = 32 accumulator (column quarter 0)
                        float y[16];
                        if (cq == 0) tmem_ld16f(tmem + lane_base + 256u * c, y);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(leader_d_free[c]);
                        if (cq == 0 && R < d.rows) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                if (i < d.out_width) {
                                    float v = fmaf(y[i], inv, __ldg(d.bias[l] + i));
                                    v = apply_act_fast(v, d.act_out);
                                    if (d.residual) v += __ldg(d.residual + (size_t)R * d.res_stride + i);
                                    d.out[(size_t)R * d.out_stride + i] = v;
                                }
                            }
                        }
                    } else {
                        // ---- last layer 128 wide: LayerNorm (row statistics combined over the four column quarters with Chan's
                        // formula through shared memory and one 128-thread barrier per row quarter), activation, store
                        float y[32];
                        tmem_ld32f(d_addr, y);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(leader_d_free[c]);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + i));
                            y[i] = fmaf(y[i], inv, b.x);
                            y[i + 1] = fmaf(y[i + 1], inv, b.y);
                            y[i + 2] = fmaf(y[i + 2], inv, b.z);
                            y[i + 3] = fmaf(y[i + 3], inv, b.w);
                        }
                        float mean = 0.f, rstd = 1.f;
                        if (gamma) {
                            // two-pass statistics through ONE 2 KiB array of per-quarter partials (a third ring stage matters
                            // more to this kernel than a barrier): sums, barrier, read, barrier, centred squares, barrier, read
                            float sum = 0.f;
#pragma unroll
                            for (int i = 0; i < 32; ++i) sum += y[i];
                            s.part[cq][row] = sum;
                            asm volatile("bar.sync %0, 128;" ::"r"(1 + lq) : "memory");
                            mean = ((s.part[0][row] + s.part[1][row]) + (s.part[2][row] + s.part[3][row])) * (1.f / H);
                            asm volatile("bar.sync %0, 128;" ::"r"(1 + lq) : "memory");
                            float sq = 0.f;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float dlt = y[i] - mean;
                                sq = fmaf(dlt, dlt, sq);
                            }
                            s.part[cq][row] = sq;
                            asm volatile("bar.sync %0, 128;" ::"r"(1 + lq) : "memory");
                            const float var = ((s.part[0][row] + s.part[1][row]) + (s.part[2][row] + s.part[3][row])) * (1.f / H);
                            rstd = 1.f / sqrtf(var + kLnEps);
                        }
                        if (R < d.rows) {
                            float* dst = (d.dual ? d.out2 : d.out) + (size_t)R * d.out_stride + cq * 32;
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {          // normalise, activate and store 8 columns at a time
                                float* o = y + i;
                                if (gamma) {
#pragma unroll
                                    for (int u = 0; u < 8; u += 4) {
                                        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + i + u));
                                        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + i + u));
                                        o[u] = fmaf((o[u] - mean) * rstd, g.x, b.x);
                                        o[u + 1] = fmaf((o[u + 1] - mean) * rstd, g.y, b.y);
                                        o[u + 2] = fmaf((o[u + 2] - mean) * rstd, g.z, b.z);
                                        o[u + 3] = fmaf((o[u + 3] - mean) * rstd, g.w, b.w);
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 8; ++u) o[u] = apply_act_fast(o[u], d.act_out);
                                stg256(dst + i, o);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp < W_LOAD0 + N_LOAD_WARPS) {
        // ====================================================================== loader warps: two per row quarter, taking
        // alternate K-blocks of the global (tile, K-block) sequence
        setmaxnreg_dec<kRegsLoad>();
        const int lw = (warp - W_LOAD0) & 3, hf = (warp - W_LOAD0) >> 2;
        const uint32_t lane_base = (uint32_t)(lw * 32) << 16;
        const uint32_t leader_full0 = mapa(smem_u32(&s.full[0]), 0);      // full[] is contiguous: + 8 bytes per buffer
        const uint32_t ring0 = smem_u32(smem + a.ring_off) + (uint32_t)(warp - W_LOAD0) * (NSTG * STG);
        const int sub = lane >> 2, piece = lane & 3;                       // cp.async: 8 rows per instruction, 4 x 16 B per piece
        mbar_wait(&s.w_full, 0);              // full[] is only signalled once this CTA's weights have landed

        // ---- issue cursor over this warp's WIDE K-blocks (narrow ones are loaded directly), NSTG - 1 stages ahead
        int64_t i_pt = pt0;
        int i_b = -1, i_cs = 0;
        uint32_t i_g = 0xffffffffu, off[4], vmask = 0;     // 16-byte-unit offsets of this lane's pieces of tile rows 8 i + sub
        bool i_live = true;
        auto seek_kblock = [&]() {            // next wide K-block of this warp at or after the cursor
            while (true) {
                ++i_b; ++i_g;
                if (i_b == NKB) { i_b = 0; i_pt += pt_stride; }
                if (i_pt >= a.n_pt) { i_live = false; return; }
                if ((int)(i_g & 1u) == hf && a.kb[i_b].width == 64) break;
            }
            const KBlock& kb = a.kb[i_b];
            vmask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t R = (i_pt * 2 + rank) * 128 + lw * 32 + 8 * i + sub;
                const bool ok = R < d.rows;
                const int64_t sr = ok ? (kb.gather ? (int64_t)__ldg(kb.gather + R) : R) : 0;
                vmask |= ok ? (1u << i) : 0u;
                off[i] = (uint32_t)((sr * kb.stride + kb.col0) >> 2) + piece;
            }
            i_cs = 0;
        };
        auto issue_stage = [&](uint32_t stage_addr) {
            if (i_live) {
                const char* base = reinterpret_cast<const char*>(a.kb[i_b].ptr) + i_cs * (SCOLS * 4);
                const uint32_t dst0 = stage_addr + (uint32_t)sub * PITCH + piece * 16;
                const float* p0 = reinterpret_cast<const float*>(base + (size_t)off[0] * 16);
                const float* p1 = reinterpret_cast<const float*>(base + (size_t)off[1] * 16);
                const float* p2 = reinterpret_cast<const float*>(base + (size_t)off[2] * 16);
                const float* p3 = reinterpret_cast<const float*>(base + (size_t)off[3] * 16);
                asm volatile(
                    "cp.async.cg.shared.global [%0], [%1], 16, %5;\n\t"
                    "cp.async.cg.shared.global [%0 + 640], [%2], 16, %6;\n\t"
                    "cp.async.cg.shared.global [%0 + 1280], [%3], 16, %7;\n\t"
                    "cp.async.cg.shared.global [%0 + 1920], [%4], 16, %8;\n"
                    ::"r"(dst0), "l"(p0), "l"(p1), "l"(p2), "l"(p3), "r"((vmask & 1u) ? 16u : 0u), "r"((vmask & 2u) ? 16u : 0u),
                    "r"((vmask & 4u) ? 16u : 0u), "r"((vmask & 8u) ? 16u : 0u)
                    : "memory");
                static_assert(8 * PITCH == 640, "offsets in the cp.async block above");
                if (++i_cs == 4) seek_kblock();
            }
            cp_async_commit();              // one (possibly empty) group per stage keeps the wait depth constant
        };
        seek_kblock();
#pragma unroll 1
        for (int p = 0; p < NSTG - 1; ++p) issue_stage(ring0 + p * STG);

        uint32_t g = 0, q = 0, ubits = 0;     // K-blocks seen, ring stages consumed, use parity of the four A-operand halves
        int tile_ord = 0;
        for (int64_t pt = pt0; pt < a.n_pt; pt += pt_stride, ++tile_ord) {
            const int c = tile_ord & 1;
            for (int b = 0; b < NKB; ++b, ++g) {
                const int buf = 2 * c + (b & 1);
                const uint32_t upar = (ubits >> buf) & 1u;
                ubits ^= 1u << buf;
                if ((int)(g & 1u) != hf) continue;
                const KBlock& kb = a.kb[b];
                const uint32_t ah = tmem + lane_base + 256u * c + 128u + 32u * (b & 1), al = ah + 64u;
                // the MMAs of the previous use of this half (or, for the last K-blocks of a tile, of its last layer) are done
                mbar_wait_sleep(&s.empty[buf], upar ^ 1u);
                tc_fence_after();
                if (kb.width == 64) {
#pragma unroll 1
                    for (int cs = 0; cs < 4; ++cs, ++q) {
                        cp_async_wait<NSTG - 2>();
                        __syncwarp();
                        const uint32_t st = ring0 + (q % NSTG) * STG + lane * PITCH;
                        float4 x[4];
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) x[v4] = lds_f4(st + v4 * 16);
                        __syncwarp();                                    // every lane has read stage q: refill the buffer of stage q - 1 ... q + 3
                        issue_stage(ring0 + ((q + NSTG - 1) % NSTG) * STG);
                        uint32_t h[8], lo[8];
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            split2(x[v4].x * kb.scale, x[v4].y * kb.scale, h[2 * v4], lo[2 * v4]);
                            split2(x[v4].z * kb.scale, x[v4].w * kb.scale, h[2 * v4 + 1], lo[2 * v4 + 1]);
                        }
                        tmem_st8(ah + 8u * cs, h);
                        tmem_st8(al + 8u * cs, lo);
                    }
                } else {
                    // narrow segment: lane = row, one K = 16 step, scalar loads (rows need not be 16-byte aligned)
                    const int64_t R = (pt * 2 + rank) * 128 + lw * 32 + lane;
                    int64_t src_row = -1;
                    if (R < d.rows) src_row = kb.gather ? (int64_t)kb.gather[R] : R;
                    uint32_t h[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float x0 = 0.f, x1 = 0.f;
                        if (src_row >= 0) {
                            if (2 * i < kb.width) x0 = __ldg(kb.ptr + (size_t)src_row * kb.stride + kb.col0 + 2 * i) * kb.scale;
                            if (2 * i + 1 < kb.width) x1 = __ldg(kb.ptr + (size_t)src_row * kb.stride + kb.col0 + 2 * i + 1) * kb.scale;
                        }
                        split2(x0, x1, h[i], lo[i]);
                    }
                    tmem_st8(ah, h);
                    tmem_st8(al, lo);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_full0 + 8u * (uint32_t)buf);
            }
        }
        cp_async_wait<0>();
    } else {
        setmaxnreg_dec<kRegsMisc>();
        if (warp == W_MMA && rank == 0) {
            // ================================================================== MMA issuer (leader CTA)
            const uint32_t idesc = idesc_f16(256, 128), idesc_narrow = idesc_f16(256, 32);
            const uint64_t w_desc = make_desc_sw128(smem_u32(smem));
            uint32_t ubits = 0, n_chain[2] = {0, 0}, n_ar[2] = {0, 0};
            for (int64_t ptb = pt0; ptb < a.n_pt; ptb += 2 * pt_stride) {
                const int nch = (ptb + pt_stride < a.n_pt) ? 2 : 1;
                for (int l = 0; l < nl; ++l) {
                    const bool narrow_layer = narrow_out && l == nl - 1;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c >= nch) continue;
                        const uint32_t d_col = tmem + 256u * c, ah = d_col + 128u, al = d_col + 192u;
                        if (l == 0) {
                            if (lane == 0) mbar_wait_sleep(&s.d_free[c], (n_chain[c] + 1) & 1);
                            ++n_chain[c];
                            for (int b = 0; b < NKB; ++b) {
                                const int buf = 2 * c + (b & 1);
                                const uint32_t upar = (ubits >> buf) & 1u;
                                ubits ^= 1u << buf;
                                if (lane == 0) {
                                    mbar_wait_sleep(&s.full[buf], upar);
                                    tc_fence_after();
                                }
                                if (elect_one()) {           // uniform-datapath issue loop (see elect_one in tc2_core.cuh)
                                    const int nks = a.kb[b].width == 64 ? 4 : 1;
                                    const uint32_t hb = 32u * (b & 1);
                                    const uint64_t wb = w_desc + (uint64_t)((a.w_off[0] + b * 2 * HIMG) >> 4);
#pragma unroll 4
                                    for (int j = 0; j < nks; ++j) {
                                        const uint64_t wh = wb + (uint64_t)((32 * j) >> 4), wl = wh + (uint64_t)(HIMG >> 4);
                                        umma_ts<2>(d_col, ah + hb + 8 * j, wh, idesc, (b > 0 || j > 0) ? 1u : 0u);
                                        umma_ts<2>(d_col, al + hb + 8 * j, wh, idesc, 1u);
                                        umma_ts<2>(d_col, ah + hb + 8 * j, wl, idesc, 1u);
                                    }
                                    // this half may be refilled once these MMAs are done -- unless it is its last use in the tile and
                                    // later layers keep their A operand in the same columns: then it is released after the last layer
                                    if (nl == 1 || b + 2 < NKB) umma_commit<2>(&s.empty[buf], 3);
                                }
                            }
                            if (lane == 0) umma_commit<2>(&s.d_full[c], 3);
                        } else {
                            if (lane == 0) {
                                mbar_wait_sleep(&s.a_ready[c], n_ar[c] & 1);
                                tc_fence_after();
                            }
                            if (elect_one()) {
                                // K = 128 from the A operand the epilogue wrote; 64-row images, or 16-row images (N = 32)
                                const uint32_t kb_bytes = narrow_layer ? 4096u : 2u * HIMG, lo_off = narrow_layer ? 2048u : (uint32_t)HIMG;
                                const uint64_t wb = w_desc + (uint64_t)(a.w_off[l] >> 4);
                                const uint32_t id = narrow_layer ? idesc_narrow : idesc;
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks) {     // straight-line issue: the rolled loop costs the single issuing thread ~110 cycles per MMA (64 at the pipe's rate)
                                    const uint64_t wh = wb + (uint64_t)(((ks >> 2) * kb_bytes + (ks & 3) * 32) >> 4);
                                    const uint64_t wl = wh + (uint64_t)(lo_off >> 4);
                                    umma_ts<2>(d_col, ah + 8 * ks, wh, id, ks > 0 ? 1u : 0u);
                                    umma_ts<2>(d_col, al + 8 * ks, wh, id, 1u);
                                    umma_ts<2>(d_col, ah + 8 * ks, wl, id, 1u);
                                }
                                umma_commit<2>(&s.d_full[c], 3);
                                if (l == nl - 1) {         // the chain's A operand columns are free for the next tile's K-blocks
                                    umma_commit<2>(&s.empty[2 * c], 3);
                                    if (NKB >= 2) umma_commit<2>(&s.empty[2 * c + 1], 3);
                                }
                            }
                            ++n_ar[c];
                        }
                        __syncwarp();
                    }
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == W_MMA) tmem_dealloc<2>(tmem, 512);
}

}  // namespace rp

int row_pair_launch(const G4cRowTcDesc& d, cudaStream_t st) {
    rp::Args a;
    a.d = d;
    a.n_kb = 0;
    for (int sgi = 0; sgi < d.n_segs; ++sgi) {
        const G4cSeg& sg = d.seg[sgi];
        const int nblk = sg.width == 128 ? 2 : 1;
        for (int h = 0; h < nblk; ++h) {
            if (a.n_kb >= rp::MAX_KB) { set_error("g4c_rowmlp_tc_fwd: more than %d K-blocks", rp::MAX_KB); return G4C_EUNSUPPORTED; }
            rp::KBlock& kb = a.kb[a.n_kb++];
            kb.ptr = sg.ptr; kb.gather = sg.gather; kb.stride = sg.stride; kb.scale = sg.scale;
            kb.col0 = 64 * h;
            kb.width = sg.width == 128 ? 64 : sg.width;
        }
    }
    const bool narrow = d.out_width != 128;
    if (narrow && d.n_layers == 1) { set_error("g4c_rowmlp_tc_fwd: a single narrow Linear is unsupported"); return G4C_EUNSUPPORTED; }
    uint32_t off = 0;
    for (int l = 0; l < d.n_layers; ++l) {
        a.w_off[l] = off;
        const bool narrow_layer = narrow && l == d.n_layers - 1;
        const uint32_t per_kb = narrow_layer ? 4096u : 2u * pairk::HIMG;
        off += per_kb * (l == 0 ? (uint32_t)a.n_kb : 2u);
    }
    a.w_bytes = off;
    a.ring_off = (off + 1023u) & ~1023u;
    const uint32_t tail = (uint32_t)sizeof(rp::Tail);
    a.tail_off = (a.ring_off + (uint32_t)rp::RING_BYTES + 15u) & ~15u;
    const int smem = (int)(a.tail_off + tail);
    if (smem > pairk::kMaxSmem) { set_error("g4c_rowmlp_tc_fwd: weights (%u bytes per CTA) leave no room for the input rings", off); return G4C_EUNSUPPORTED; }
    const int64_t n_pt = (d.rows + 255) / 256;
    a.n_pt = n_pt;
    static int configured[kMaxDevices] = {0};
    if (!ensure_dynamic_smem(rp::row_pair_kernel, smem, configured)) return check_launch("row_pair_kernel attribute");
    const int n_sm = device_sms();
    const int pairs = (int)std::min<int64_t>(n_pt, n_sm / 2);
    rp::row_pair_kernel<<<2 * pairs, pairk::NT, smem, st>>>(a);
    count_launch(1, true);
    return check_launch("row_pair_kernel");
}

}  // namespace g4c
