// mp_tc.cu — fused message-passing block on the 5th-gen tensor cores (tcgen05 + TMEM), hidden = 128.
//
// Same unit decomposition as the fp32 kernel (one CTA = 128 consecutive targets; slot j stages the
// j-th in-edge of every target so TMEM lane m always belongs to target m), but every Linear is a
// 128x128xK tcgen05.mma with fp32 accumulation in TMEM.  fp32 parity is kept by splitting both
// operands into fp16 (hi, lo) pairs and issuing three MMAs per K-step (hi*hi + lo*hi + hi*lo): 22
// significant bits per operand, error of the dropped lo*lo term ~2^-22.
//
//   warps 0-7 : stage fp32 rows from HBM -> (hi,lo) fp16 SWIZZLE_128B operand images in shared memory,
//               run the epilogues (TMEM -> registers: bias, SELU, re-split as the next layer's A operand;
//               last layer: LayerNorm, edge store, per-target aggregation in registers)
//   thread 0  : issues every tcgen05.mma and the tcgen05.commit that publish completion on mbarriers
//   warp 8    : weight producer — streams the pre-split, pre-swizzled weight images (32 KiB per 64-wide
//               K-block) from L2 into a 3-stage shared-memory ring with the TMA engine's 1-D bulk copy
#include <algorithm>
#include "tc_core.cuh"

namespace g4c {
namespace tc {

constexpr int H = 128;
constexpr int TMU = 128;                         // targets per unit == MMA M
constexpr int NSTAGE = 3;                        // weight ring stages
constexpr int STAGE_BYTES = 2 * kBlockBytes;     // hi | lo image of one K-block
constexpr int ABUF_BYTES = 4 * kBlockBytes;      // hi kb0, hi kb1, lo kb0, lo kb1 of a 128-wide operand
constexpr int NCOMPUTE = 256;
constexpr int NTHREADS = NCOMPUTE + 32;
constexpr uint32_t TMEM_COLS = 128;

struct Smem {
    uint8_t a[2][ABUF_BYTES];
    uint8_t w[NSTAGE][STAGE_BYTES];
    float part[2][2][TMU];                       // LayerNorm partial sums: [pass][column half][row]
    uint64_t w_full[NSTAGE];
    uint64_t w_empty[NSTAGE];
    uint64_t a0_free;
    uint64_t d_full;
    uint32_t tmem_base;
    int maxdeg;
};

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory"); }

struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
};

// Stage 16 rows per warp of an fp32 [*,128] matrix as the (hi,lo) operand image `img`.
// `my_row` is this lane's row index for the row it owns (row = 32*(warp%4) + lane), < 0 = zero row.
__device__ __forceinline__ void stage_rows(uint8_t* img, const float* __restrict__ src, int my_row, int warp, int lane) {
    const int q = warp & 3, half = warp >> 2;
    const int kb = lane >> 4;                                    // columns 4*lane .. 4*lane+3
    const int chunk = (lane & 15) >> 1, sub = (lane & 1) * 8;
#pragma unroll
    for (int i0 = 0; i0 < 16; i0 += 4) {
        float4 x[4];
        int rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int owner = half * 16 + i0 + u;
            rr[u] = __shfl_sync(0xffffffffu, my_row, owner);
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rr[u] >= 0) x[u] = ldg_stream(src + (size_t)rr[u] * H + lane * 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = q * 32 + half * 16 + i0 + u;
            uint2 h, l;
            split2(x[u].x, x[u].y, h.x, l.x);
            split2(x[u].z, x[u].w, h.y, l.y);
            const uint32_t off = (uint32_t)kb * kBlockBytes + (uint32_t)r * 128 + (uint32_t)((chunk ^ (r & 7)) << 4) + sub;
            *reinterpret_cast<uint2*>(img + off) = h;
            *reinterpret_cast<uint2*>(img + 2 * kBlockBytes + off) = l;
        }
    }
}

// thread 0: one Linear over `nkb` K-blocks whose A images start at a_img (hi at +0, lo at +2 blocks)
__device__ __forceinline__ void issue_layer(Smem& s, Ring& ring, uint32_t d_tmem, const uint8_t* a_img, int nkb, bool first) {
    const uint32_t a_base = smem_u32(a_img);
    for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&s.w_full[ring.stage], ring.phase);
        tc_fence_after();
        const uint32_t w_base = smem_u32(s.w[ring.stage]);
        issue_kblock_x3(d_tmem, a_base + kb * kBlockBytes, a_base + (2 + kb) * kBlockBytes, w_base, w_base + kBlockBytes,
                        first && kb == 0);
        umma_commit(&s.w_empty[ring.stage]);
        ring.advance();
    }
}

// epilogue of a hidden layer: TMEM -> bias + SELU -> (hi,lo) image of the next layer's A operand
__device__ __forceinline__ void epilogue_hidden(uint32_t d_tmem, int row, int half, float inv_scale,
                                                const float* __restrict__ bias, uint8_t* a_img) {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tmem_ld32(d_tmem + (uint32_t)(half * 64 + c0), v);
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = selu(fmaf(v[g + i], inv_scale, __ldg(bias + half * 64 + c0 + g + i)));
            // this thread's 64 columns are exactly K-block `half` of the next operand
            const int k = c0 + g;
            uint4 h, l;
            split2(x[0], x[1], h.x, l.x);
            split2(x[2], x[3], h.y, l.y);
            split2(x[4], x[5], h.z, l.z);
            split2(x[6], x[7], h.w, l.w);
            const uint32_t off = (uint32_t)half * kBlockBytes + (uint32_t)row * 128 + (uint32_t)(((k >> 3) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(a_img + off) = h;
            *reinterpret_cast<uint4*>(a_img + 2 * kBlockBytes + off) = l;
        }
    }
}

// last layer: y[64] = LayerNorm(acc*inv_scale + bias) for this thread's 64 columns of its row
__device__ __forceinline__ void epilogue_final(Smem& s, uint32_t d_tmem, int row, int half, float inv_scale,
                                               const float* __restrict__ bias, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float (&y)[64]) {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tmem_ld32(d_tmem + (uint32_t)(half * 64 + c0), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) y[c0 + i] = fmaf(v[i], inv_scale, __ldg(bias + half * 64 + c0 + i));
    }
    if (gamma == nullptr) return;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) sum += y[i];
    s.part[0][half][row] = sum;
    compute_sync();
    const float mean = (s.part[0][0][row] + s.part[0][1][row]) * (1.f / H);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        y[i] -= mean;
        sq = fmaf(y[i], y[i], sq);
    }
    s.part[1][half][row] = sq;
    compute_sync();
    const float rstd = 1.f / sqrtf((s.part[1][0][row] + s.part[1][1][row]) * (1.f / H) + kLnEps);
#pragma unroll
    for (int i = 0; i < 64; ++i) y[i] = fmaf(y[i] * rstd, __ldg(gamma + half * 64 + i), __ldg(beta + half * 64 + i));
}

__device__ __forceinline__ void store_row64(float* __restrict__ dst, const float (&y)[64], int act) {
#pragma unroll
    for (int i = 0; i < 64; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(apply_act(y[i], act), apply_act(y[i + 1], act),
                                                          apply_act(y[i + 2], act), apply_act(y[i + 3], act));
}

__global__ void __launch_bounds__(NTHREADS, 1) mp_tc_kernel(const G4cMpDesc d) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();      // SWIZZLE_128B operand images need 1024-byte alignment
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_units = (d.n_targets + TMU - 1) / TMU;
    const int nl_e = d.edge_mlp.n_layers, nl_n = d.node_mlp.n_layers;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
        mbar_init(&s.a0_free, 1);
        mbar_init(&s.d_full, 1);
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(&s.tmem_base, TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;

    Ring ring;                       // weight ring position (producer and MMA issuer walk the same schedule)
    uint32_t ph_a0 = 0, ph_d = 0;    // phases of a0_free / d_full

    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        // ---- per-row metadata, held in registers by the two threads (column halves) that own the row
        const int row = (warp & 3) * 32 + lane;             // valid for compute warps
        const int half = (warp >> 2) & 1;
        int deg = 0, base = 0, trow = -1;
        if (warp < 8) {
            const int64_t n = unit * TMU + row;
            if (n < d.n_targets) {
                if (d.fixed_k > 0) { base = (int)(n * d.fixed_k); deg = d.fixed_k; }
                else { base = d.rowptr[n]; deg = d.rowptr[n + 1] - base; }
                trow = d.tgt_perm ? d.tgt_perm[n] : (int)n;
            }
        }
        if (tid == 0) s.maxdeg = 0;
        __syncthreads();
        if (warp < 8 && deg > 0) atomicMax(&s.maxdeg, deg);
        __syncthreads();
        const int maxdeg = s.maxdeg;

        if (warp == 8) {
            // ================= weight producer =================
            if (lane == 0) {
                auto feed = [&](const G4cMlp& m, int nl, int nkb_first) {
                    for (int l = 0; l < nl; ++l) {
                        const int nkb = (l == 0) ? nkb_first : 2;
                        const uint8_t* src = static_cast<const uint8_t*>(m.W_pack[l]);
                        for (int kb = 0; kb < nkb; ++kb) {
                            mbar_wait(&s.w_empty[ring.stage], ring.phase ^ 1);
                            mbar_arrive_expect_tx(&s.w_full[ring.stage], STAGE_BYTES);
                            bulk_g2s(s.w[ring.stage], src + (size_t)kb * STAGE_BYTES, STAGE_BYTES, &s.w_full[ring.stage]);
                            ring.advance();
                        }
                    }
                };
                for (int j = 0; j < maxdeg; ++j) feed(d.edge_mlp, nl_e, 6);
                feed(d.node_mlp, nl_n, 4);
            }
            __syncwarp();
        } else {
            // ================= compute warps =================
            float agg[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) agg[i] = 0.f;
            float y[64];

            for (int j = 0; j < maxdeg; ++j) {
                int erow = -1, srow = -1;
                if (j < deg) {
                    const int slot = base + j;
                    erow = d.edge_perm ? d.edge_perm[slot] : slot;
                    srow = d.src[slot];
                }
                // ---- layer 1 = three K-segments: e -> A0, S[src] -> A1, T[tgt] -> A0 (after a0_free)
                stage_rows(s.a[0], d.e_in, erow, warp, lane);
                fence_proxy_async();
                compute_sync();
                if (tid == 0) {
                    tc_fence_after();
                    issue_layer(s, ring, tmem, s.a[0], 2, true);
                    umma_commit(&s.a0_free);
                }
                stage_rows(s.a[1], d.src_feat, srow, warp, lane);
                fence_proxy_async();
                compute_sync();
                if (tid == 0) {
                    tc_fence_after();
                    issue_layer(s, ring, tmem, s.a[1], 2, false);
                }
                mbar_wait(&s.a0_free, ph_a0);
                ph_a0 ^= 1;
                stage_rows(s.a[0], d.tgt_feat, trow, warp, lane);
                fence_proxy_async();
                compute_sync();
                if (tid == 0) {
                    tc_fence_after();
                    issue_layer(s, ring, tmem, s.a[0], 2, false);
                    umma_commit(&s.d_full);
                }
                // ---- hidden layers
                int cur = 1;                                   // epilogue of layer l writes A[cur]
                for (int l = 1; l < nl_e; ++l) {
                    mbar_wait(&s.d_full, ph_d);
                    ph_d ^= 1;
                    tc_fence_after();
                    epilogue_hidden(tmem + ((uint32_t)((warp & 3) * 32) << 16), row, half, d.edge_mlp.w_inv_scale[l - 1],
                                    d.edge_mlp.b[l - 1], s.a[cur]);
                    tc_fence_before();
                    fence_proxy_async();
                    compute_sync();
                    if (tid == 0) {
                        tc_fence_after();
                        issue_layer(s, ring, tmem, s.a[cur], 2, true);
                        umma_commit(&s.d_full);
                    }
                    cur ^= 1;
                }
                // ---- last layer: LayerNorm, edge store, aggregation
                mbar_wait(&s.d_full, ph_d);
                ph_d ^= 1;
                tc_fence_after();
                epilogue_final(s, tmem + ((uint32_t)((warp & 3) * 32) << 16), row, half, d.edge_mlp.w_inv_scale[nl_e - 1],
                               d.edge_mlp.b[nl_e - 1], d.edge_mlp.ln_gamma, d.edge_mlp.ln_beta, y);
                tc_fence_before();
                if (erow >= 0) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) agg[i] += y[i];
                    if (d.e_out) store_row64(d.e_out + (size_t)erow * H + half * 64, y, d.act_e_out);
                }
                compute_sync();      // every TMEM read of this slot is done before the next slot's first MMA
            }

            // ---- node MLP: cat(agg, T)
            {
                const float cnt = (d.aggr == G4C_AGGR_MEAN) ? (float)max(deg, 1) : 1.f;
#pragma unroll
                for (int g = 0; g < 64; g += 8) {
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = agg[g + i] / cnt;
                    uint4 h, l;
                    split2(x[0], x[1], h.x, l.x);
                    split2(x[2], x[3], h.y, l.y);
                    split2(x[4], x[5], h.z, l.z);
                    split2(x[6], x[7], h.w, l.w);
                    const uint32_t off = (uint32_t)half * kBlockBytes + (uint32_t)row * 128 + (uint32_t)(((g >> 3) ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(s.a[1] + off) = h;
                    *reinterpret_cast<uint4*>(s.a[1] + 2 * kBlockBytes + off) = l;
                }
                stage_rows(s.a[0], d.tgt_feat, trow, warp, lane);
                fence_proxy_async();
                compute_sync();
                if (tid == 0) {
                    tc_fence_after();
                    issue_layer(s, ring, tmem, s.a[1], 2, true);
                    issue_layer(s, ring, tmem, s.a[0], 2, false);
                    umma_commit(&s.d_full);
                }
                int cur = 1;
                for (int l = 1; l < nl_n; ++l) {
                    mbar_wait(&s.d_full, ph_d);
                    ph_d ^= 1;
                    tc_fence_after();
                    epilogue_hidden(tmem + ((uint32_t)((warp & 3) * 32) << 16), row, half, d.node_mlp.w_inv_scale[l - 1],
                                    d.node_mlp.b[l - 1], s.a[cur]);
                    tc_fence_before();
                    fence_proxy_async();
                    compute_sync();
                    if (tid == 0) {
                        tc_fence_after();
                        issue_layer(s, ring, tmem, s.a[cur], 2, true);
                        umma_commit(&s.d_full);
                    }
                    cur ^= 1;
                }
                mbar_wait(&s.d_full, ph_d);
                ph_d ^= 1;
                tc_fence_after();
                epilogue_final(s, tmem + ((uint32_t)((warp & 3) * 32) << 16), row, half, d.node_mlp.w_inv_scale[nl_n - 1],
                               d.node_mlp.b[nl_n - 1], d.node_mlp.ln_gamma, d.node_mlp.ln_beta, y);
                tc_fence_before();
                if (trow >= 0) store_row64(d.t_out + (size_t)trow * H + half * 64, y, d.act_t_out);
                compute_sync();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------ GEMM-core self test
// D[128,128] = A[128,K] * W^T through exactly the staging / descriptor / MMA / TMEM-load code above.
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_test_kernel(const float* A, const uint8_t* W_pack, float inv_scale, int K, float* D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&s.w_full[i], 1); mbar_init(&s.w_empty[i], 1); }
        mbar_init(&s.a0_free, 1);
        mbar_init(&s.d_full, 1);
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(&s.tmem_base, TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const int nkb = K / 64;
    Ring ring;
    if (warp == 8) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&s.w_empty[ring.stage], ring.phase ^ 1);
                mbar_arrive_expect_tx(&s.w_full[ring.stage], STAGE_BYTES);
                bulk_g2s(s.w[ring.stage], W_pack + (size_t)kb * STAGE_BYTES, STAGE_BYTES, &s.w_full[ring.stage]);
                ring.advance();
            }
        }
    } else {
        const int row = (warp & 3) * 32 + lane, half = (warp >> 2) & 1;
        // A has row stride K here: stage K-block by K-block through the generic helper (row stride 128 only when K == 128)
        for (int idx = tid; idx < 128 * (K / 8); idx += NCOMPUTE) {
            const int r = idx / (K / 8), k0 = (idx % (K / 8)) * 8;
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = A[(size_t)r * K + k0 + i];
            store_split8(s.a[0], s.a[0] + 2 * kBlockBytes, r, k0, x);
        }
        fence_proxy_async();
        compute_sync();
        if (tid == 0) {
            tc_fence_after();
            issue_layer(s, ring, tmem, s.a[0], nkb, true);
            umma_commit(&s.d_full);
        }
        mbar_wait(&s.d_full, 0);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 64 + c0), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) D[(size_t)row * 128 + half * 64 + c0 + i] = v[i] * inv_scale;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tc

static int tc_smem_bytes() { return (int)sizeof(tc::Smem); }

static int num_sms_tc() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int mp_tc_dispatch(const G4cMpDesc& d, cudaStream_t st) {
    if (d.precision != G4C_PREC_FP16X3) { set_error("g4c_mp_fwd: precision=%d is not implemented (fp32, fp16x3)", d.precision); return G4C_EUNSUPPORTED; }
    if (d.hidden != tc::H) { set_error("g4c_mp_fwd: the tensor-core path is built for hidden=128 (got %d); use precision fp32", d.hidden); return G4C_EUNSUPPORTED; }
    for (int l = 0; l < d.edge_mlp.n_layers; ++l) if (!d.edge_mlp.W_pack[l]) { set_error("g4c_mp_fwd: edge_mlp.W_pack[%d] is NULL", l); return G4C_EINVAL; }
    for (int l = 0; l < d.node_mlp.n_layers; ++l) if (!d.node_mlp.W_pack[l]) { set_error("g4c_mp_fwd: node_mlp.W_pack[%d] is NULL", l); return G4C_EINVAL; }
    static bool configured = false;
    const int smem = tc_smem_bytes();
    if (!configured) {
        if (cudaFuncSetAttribute(tc::mp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return check_launch("mp_tc_kernel attribute");
        configured = true;
    }
    const int64_t n_units = (d.n_targets + tc::TMU - 1) / tc::TMU;
    const int grid = (int)std::min<int64_t>(n_units, num_sms_tc());
    tc::mp_tc_kernel<<<grid, tc::NTHREADS, smem, st>>>(d);
    count_launch();
    return check_launch("mp_tc_kernel");
}

int tc_gemm_test_launch(const float* A, const void* W_pack, float inv_scale, int K, float* D, cudaStream_t st) {
    if (K != 64 && K != 128) { set_error("g4c_debug_tc_gemm: K must be 64 or 128"); return G4C_EINVAL; }
    static bool configured = false;
    const int smem = tc_smem_bytes();
    if (!configured) {
        if (cudaFuncSetAttribute(tc::tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return check_launch("tc_gemm_test_kernel attribute");
        configured = true;
    }
    tc::tc_gemm_test_kernel<<<1, tc::NTHREADS, smem, st>>>(A, static_cast<const uint8_t*>(W_pack), inv_scale, K, D);
    count_launch();
    return check_launch("tc_gemm_test_kernel");
}

}  // namespace g4c
