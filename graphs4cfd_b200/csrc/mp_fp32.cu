// mp_fp32.cu — fp32 fused message-passing block and fused row-MLP (CUDA-core path).
//
// mp_kernel: one CTA owns TM consecutive targets.  For slot j = 0..maxdeg-1 it stages the j-th
// in-edge of every target (ELL view of the sorted-by-target edge list): e row, gathered source
// row, and the target tile (loaded once) are the three K-segments of the edge MLP's first layer,
// so the [E,3H] concatenation of the reference never exists.  Row m of every slot tile belongs to
// target m, which makes the aggregation a register accumulation (no atomics, fixed order) and
// lets the node MLP run on the same tile without leaving the SM.
#include <algorithm>
#include "tile_mlp.cuh"

namespace g4c {

template <int H>
__global__ void __launch_bounds__(256, (H <= 128 ? 2 : 1)) mp_kernel(const G4cMpDesc d) {
    using C = Cfg<H>;
    extern __shared__ __align__(16) float smem[];
    float* T = smem;                       // target tile  [TM][LD]
    float* X = T + C::TM * C::LD;          // chain buffer [TM][LD]
    float* S = X + C::TM * C::LD;          // source tile  [TM][LD]
    float* wst = S + C::TM * C::LD;        // weight staging
    int* s_deg = reinterpret_cast<int*>(wst + C::WST);
    int* s_base = s_deg + C::TM;
    int* s_trow = s_base + C::TM;
    int* s_erow = s_trow + C::TM;
    int* s_srow = s_erow + C::TM;
    __shared__ int s_maxdeg;

    const int tid = threadIdx.x;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const int64_t n_units = (d.n_targets + C::TM - 1) / C::TM;

    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t n0 = unit * C::TM;
        if (tid == 0) s_maxdeg = 0;
        __syncthreads();
        for (int m = tid; m < C::TM; m += C::NT) {
            const int64_t n = n0 + m;
            int deg = 0, base = 0, trow = -1;
            if (n < d.n_targets) {
                if (d.fixed_k > 0) { base = (int)(n * d.fixed_k); deg = d.fixed_k; }
                else { base = d.rowptr[n]; deg = d.rowptr[n + 1] - base; }
                trow = d.tgt_perm ? d.tgt_perm[n] : (int)n;
            }
            s_deg[m] = deg; s_base[m] = base; s_trow[m] = trow;
            if (deg > 0) atomicMax(&s_maxdeg, deg);
        }
        __syncthreads();
        const int maxdeg = s_maxdeg;
        load_tile_wide<H>(T, d.tgt_feat, H, s_trow, 1.f, tid);

        float agg[4][8];
        zero_acc(agg);
        float acc[4][8];

        for (int j = 0; j < maxdeg; ++j) {
            for (int m = tid; m < C::TM; m += C::NT) {
                int erow = -1, srow = -1;
                if (j < s_deg[m]) {
                    const int slot = s_base[m] + j;
                    erow = d.edge_perm ? d.edge_perm[slot] : slot;
                    srow = d.src[slot];
                }
                s_erow[m] = erow; s_srow[m] = srow;
            }
            __syncthreads();
            load_tile_wide<H>(X, d.e_in, H, s_erow, 1.f, tid);
            load_tile_wide<H>(S, d.src_feat, H, s_srow, 1.f, tid);
            __syncthreads();
            zero_acc(acc);
            gemm_seg<H>(acc, X, C::LD, H, d.edge_mlp.W_t[0], wst, tid);
            gemm_seg<H>(acc, S, C::LD, H, d.edge_mlp.W_t[0] + (size_t)H * H, wst, tid);
            gemm_seg<H>(acc, T, C::LD, H, d.edge_mlp.W_t[0] + (size_t)2 * H * H, wst, tid);
            chain_tail<H>(acc, d.edge_mlp, X, wst, tid);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int erow = s_erow[ty * 4 + i];
                if (erow >= 0) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) agg[i][c] += acc[i][c];
                    if (d.e_out) {
                        float* o = d.e_out + (size_t)erow * H;
                        *reinterpret_cast<float4*>(o + tx * 4) = make_float4(
                            apply_act(acc[i][0], d.act_e_out), apply_act(acc[i][1], d.act_e_out),
                            apply_act(acc[i][2], d.act_e_out), apply_act(acc[i][3], d.act_e_out));
                        *reinterpret_cast<float4*>(o + H / 2 + tx * 4) = make_float4(
                            apply_act(acc[i][4], d.act_e_out), apply_act(acc[i][5], d.act_e_out),
                            apply_act(acc[i][6], d.act_e_out), apply_act(acc[i][7], d.act_e_out));
                    }
                }
            }
            __syncthreads();   // s_erow / X are rewritten by the next slot
        }

        // aggregated messages become the first K-segment of the node MLP
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int deg = s_deg[ty * 4 + i];
            // mean = sum / clamp(count, 1), a true division as in torch_geometric.utils.scatter
            const float cnt = (d.aggr == G4C_AGGR_MEAN) ? (float)max(deg, 1) : 1.f;
            float* row = X + (size_t)(ty * 4 + i) * C::LD;
            *reinterpret_cast<float4*>(row + tx * 4) =
                make_float4(agg[i][0] / cnt, agg[i][1] / cnt, agg[i][2] / cnt, agg[i][3] / cnt);
            *reinterpret_cast<float4*>(row + H / 2 + tx * 4) =
                make_float4(agg[i][4] / cnt, agg[i][5] / cnt, agg[i][6] / cnt, agg[i][7] / cnt);
        }
        __syncthreads();
        zero_acc(acc);
        gemm_seg<H>(acc, X, C::LD, H, d.node_mlp.W_t[0], wst, tid);
        gemm_seg<H>(acc, T, C::LD, H, d.node_mlp.W_t[0] + (size_t)H * H, wst, tid);
        chain_tail<H>(acc, d.node_mlp, X, wst, tid);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int trow = s_trow[ty * 4 + i];
            if (trow >= 0) {
                float* o = d.t_out + (size_t)trow * H;
                *reinterpret_cast<float4*>(o + tx * 4) = make_float4(
                    apply_act(acc[i][0], d.act_t_out), apply_act(acc[i][1], d.act_t_out),
                    apply_act(acc[i][2], d.act_t_out), apply_act(acc[i][3], d.act_t_out));
                *reinterpret_cast<float4*>(o + H / 2 + tx * 4) = make_float4(
                    apply_act(acc[i][4], d.act_t_out), apply_act(acc[i][5], d.act_t_out),
                    apply_act(acc[i][6], d.act_t_out), apply_act(acc[i][7], d.act_t_out));
            }
        }
        __syncthreads();   // tiles and meta are rewritten by the next unit
    }
}

// ---------------------------------------------------------------------------------- row MLP
template <int H>
__global__ void __launch_bounds__(256, (H <= 128 ? 2 : 1)) rowmlp_kernel(const G4cRowMlpDesc d) {
    using C = Cfg<H>;
    extern __shared__ __align__(16) float smem[];
    float* X = smem;                       // chain buffer / first wide segment
    float* Y = X + C::TM * C::LD;          // second wide segment
    float* Z = Y + C::TM * C::LD;          // narrow segment [TM][SMALL_LD]
    float* wst = Z + C::TM * C::SMALL_LD;
    int* s_rows = reinterpret_cast<int*>(wst + C::WST);

    const int tid = threadIdx.x;
    const int tx = tid % C::TX, ty = tid / C::TX;
    const int64_t n_tiles = (d.rows + C::TM - 1) / C::TM;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = tile * C::TM;
        float acc[4][8];
        zero_acc(acc);
        int koff = 0, n_wide = 0;
        // stage all segments, then run linear_1 segment by segment
        const float* seg_buf[G4C_MAX_SEGS];
        int seg_ld[G4C_MAX_SEGS];
        for (int s = 0; s < d.n_segs; ++s) {
            const G4cSeg& sg = d.seg[s];
            __syncthreads();
            for (int m = tid; m < C::TM; m += C::NT) {
                const int64_t r = r0 + m;
                s_rows[m] = (r < d.rows) ? (sg.gather ? sg.gather[r] : (int)r) : -1;
            }
            __syncthreads();
            if (sg.width == H) {
                float* buf = (n_wide == 0) ? X : Y;
                ++n_wide;
                load_tile_wide<H>(buf, sg.ptr, sg.stride, s_rows, sg.scale, tid);
                seg_buf[s] = buf; seg_ld[s] = C::LD;
            } else {
                for (int idx = tid; idx < C::TM * sg.width; idx += C::NT) {
                    const int m = idx / sg.width, c = idx % sg.width;
                    const int r = s_rows[m];
                    Z[m * C::SMALL_LD + c] = (r >= 0) ? sg.scale * sg.ptr[(size_t)r * sg.stride + c] : 0.f;
                }
                seg_buf[s] = Z; seg_ld[s] = C::SMALL_LD;
            }
        }
        __syncthreads();
        for (int s = 0; s < d.n_segs; ++s) {
            gemm_seg<H>(acc, seg_buf[s], seg_ld[s], d.seg[s].width, d.mlp.W_t[0] + (size_t)koff * H, wst, tid);
            koff += d.seg[s].width;
        }
        chain_tail<H>(acc, d.mlp, X, wst, tid);

        if (d.mlp.out_width == H) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t r = r0 + ty * 4 + i;
                if (r < d.rows) {
                    float* o = d.out + (size_t)r * d.out_stride;
                    float v[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        v[c] = acc[i][c];
                        if (d.residual) v[c] += d.residual[(size_t)r * d.res_stride + frag_col<H>(tx, c)];
                        v[c] = apply_act(v[c], d.act_out);
                    }
                    *reinterpret_cast<float4*>(o + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(o + H / 2 + tx * 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        } else {
            // narrow last layer (decoder): one dot product of length H per (row, output)
            const int nout = d.mlp.out_width;
            const float* W = d.mlp.W_t[d.mlp.n_layers - 1];   // torch layout [nout][H]
            const float* b = d.mlp.b[d.mlp.n_layers - 1];
            for (int idx = tid; idx < C::TM * nout; idx += C::NT) {
                const int m = idx / nout, c = idx % nout;
                const int64_t r = r0 + m;
                if (r >= d.rows) continue;
                const float* x = X + (size_t)m * C::LD;
                const float* w = W + (size_t)c * H;
                float s = 0.f;
#pragma unroll 8
                for (int k = 0; k < H; ++k) s = fmaf(x[k], __ldg(w + k), s);
                s += b[c];
                if (d.residual) s += d.residual[(size_t)r * d.res_stride + c];
                d.out[(size_t)r * d.out_stride + c] = apply_act(s, d.act_out);
            }
        }
        __syncthreads();
    }
}

template <int H>
static size_t mp_smem_bytes() {
    using C = Cfg<H>;
    return (size_t)(3 * C::TM * C::LD + C::WST) * sizeof(float) + 5 * C::TM * sizeof(int);
}
template <int H>
static size_t rowmlp_smem_bytes() {
    using C = Cfg<H>;
    return (size_t)(2 * C::TM * C::LD + C::TM * C::SMALL_LD + C::WST) * sizeof(float) + C::TM * sizeof(int);
}

static int num_sms() { return device_sms(); }

template <int H>
static int launch_mp(const G4cMpDesc& d, cudaStream_t st) {
    using C = Cfg<H>;
    static int configured[kMaxDevices] = {0};
    const size_t smem = mp_smem_bytes<H>();
    if (!ensure_dynamic_smem(mp_kernel<H>, (int)smem, configured)) return check_launch("mp_kernel attribute");
    const int64_t n_units = (d.n_targets + C::TM - 1) / C::TM;
    const int per_sm = (H <= 128) ? 2 : 1;
    const int grid = (int)std::min<int64_t>(n_units, (int64_t)num_sms() * per_sm);
    mp_kernel<H><<<grid, C::NT, smem, st>>>(d);
    count_launch();
    return check_launch("mp_kernel");
}

template <int H>
static int launch_rowmlp(const G4cRowMlpDesc& d, cudaStream_t st) {
    using C = Cfg<H>;
    static int configured[kMaxDevices] = {0};
    const size_t smem = rowmlp_smem_bytes<H>();
    if (!ensure_dynamic_smem(rowmlp_kernel<H>, (int)smem, configured)) return check_launch("rowmlp_kernel attribute");
    const int64_t n_tiles = (d.rows + C::TM - 1) / C::TM;
    const int per_sm = (H <= 128) ? 2 : 1;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)num_sms() * per_sm);
    rowmlp_kernel<H><<<grid, C::NT, smem, st>>>(d);
    count_launch();
    return check_launch("rowmlp_kernel");
}

int mp_fp32_dispatch(const G4cMpDesc& d, cudaStream_t st) {
    switch (d.hidden) {
        case 16: return launch_mp<16>(d, st);
        case 32: return launch_mp<32>(d, st);
        case 64: return launch_mp<64>(d, st);
        case 128: return launch_mp<128>(d, st);
        case 256: return launch_mp<256>(d, st);
    }
    set_error("g4c_mp_fwd: hidden=%d unsupported (16,32,64,128,256)", d.hidden);
    return G4C_EUNSUPPORTED;
}

int rowmlp_fp32_dispatch(const G4cRowMlpDesc& d, cudaStream_t st) {
    switch (d.mlp.hidden) {
        case 16: return launch_rowmlp<16>(d, st);
        case 32: return launch_rowmlp<32>(d, st);
        case 64: return launch_rowmlp<64>(d, st);
        case 128: return launch_rowmlp<128>(d, st);
        case 256: return launch_rowmlp<256>(d, st);
    }
    set_error("g4c_rowmlp_fwd: hidden=%d unsupported (16,32,64,128,256)", d.mlp.hidden);
    return G4C_EUNSUPPORTED;
}

}  // namespace g4c
