// misc.cu — bandwidth-bound helpers of the hot path: segmented reduce (pooling), REMuS geometry
// (projection, edge->node least squares, kNN interpolation), rollout state update, halo staging.
// All are single-pass, float4-vectorised where the row width allows, one warp (or sub-warp) per row.
#include <algorithm>
#include "common.cuh"

namespace g4c {

static inline int grid_for(int64_t work_items, int per_block) {
    int64_t g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > (int64_t)148 * 32) g = (int64_t)148 * 32;
    return (int)g;
}

// out[g, :] = act(reduce_{i in [ptr[g], ptr[g+1])} x[idx[i], :]); one thread per (group, float4)
__global__ void seg_reduce_kernel(const G4cSegReduceDesc d) {
    const int V = d.width / 4;
    const int64_t total = d.n_groups * V;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = t / V;
        const int c4 = (int)(t % V);
        const int lo = d.ptr[g], hi = d.ptr[g + 1];
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = lo; i < hi; ++i) {
            const int r = d.idx ? d.idx[i] : i;
            const float4 x = ldg_stream(d.x + (size_t)r * d.width + c4 * 4);
            s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
        }
        if (d.aggr == G4C_AGGR_MEAN) {
            const float cnt = (float)max(hi - lo, 1);
            s.x /= cnt; s.y /= cnt; s.z /= cnt; s.w /= cnt;
        }
        s.x = apply_act(s.x, d.act_out); s.y = apply_act(s.y, d.act_out);
        s.z = apply_act(s.z, d.act_out); s.w = apply_act(s.w, d.act_out);
        *reinterpret_cast<float4*>(d.out + (size_t)g * d.width + c4 * 4) = s;
    }
}

// out[j, f] = V[col[j], 2f]*U[j,0] + V[col[j], 2f+1]*U[j,1]; out[j, F+x] = extra[x][col[j]]
__global__ void project_kernel(const G4cProjectDesc d) {
    const int W = d.n_feat + d.n_extra;
    const int64_t total = d.n_edges * W;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / W;
        const int f = (int)(t % W);
        const int n = d.col[j];
        float o;
        if (f < d.n_feat) {
            const float2 v = *reinterpret_cast<const float2*>(d.V + (size_t)n * 2 * d.n_feat + 2 * f);
            const float2 u = *reinterpret_cast<const float2*>(d.U + (size_t)j * 2);
            // same association as (v * u).sum(-1): v.x*u.x + v.y*u.y without fused contraction
            o = __fadd_rn(__fmul_rn(v.x, u.x), __fmul_rn(v.y, u.y));
        } else {
            o = d.extra[f - d.n_feat][n];
        }
        d.out[t] = o;
    }
}

// V[n, 2f+c] = (residual) + sum_m Uinv[n, c, m] * e[n*k+m, f]
__global__ void edge_to_node_kernel(const G4cEdgeToNodeDesc d) {
    const int F = d.n_feat;
    const int64_t total = d.n_nodes * F;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = t / F;
        const int f = (int)(t % F);
        const float* ui = d.Uinv + (size_t)n * 2 * d.k;
        float sx = 0.f, sy = 0.f;
        for (int m = 0; m < d.k; ++m) {
            const float e = d.e[((size_t)n * d.k + m) * F + f];
            sx = fmaf(ui[m], e, sx);
            sy = fmaf(ui[d.k + m], e, sy);
        }
        float* o = d.V + (size_t)n * d.out_stride + 2 * f;
        if (d.residual) {
            sx += d.residual[(size_t)n * d.res_stride + 2 * f];
            sy += d.residual[(size_t)n * d.res_stride + 2 * f + 1];
        }
        *reinterpret_cast<float2*>(o) = make_float2(sx, sy);
    }
}

// y[y_row[i], :] = sum_m w[i*k+m] * x[x_idx[i*k+m], :] / sum_m w[i*k+m]
__global__ void interp_kernel(const G4cInterpDesc d) {
    const int V = d.width / 4;
    const int64_t total = d.n_out * V;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / V;
        const int c4 = (int)(t % V);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        float ws = 0.f;
        for (int m = 0; m < d.k; ++m) {
            const float w = d.w[i * d.k + m];
            const float4 x = *reinterpret_cast<const float4*>(d.x + (size_t)d.x_idx[i * d.k + m] * d.width + c4 * 4);
            s.x += x.x * w; s.y += x.y * w; s.z += x.z * w; s.w += x.w * w;
            ws += w;
        }
        const int64_t r = d.y_row ? d.y_row[i] : i;
        *reinterpret_cast<float4*>(d.y + (size_t)r * d.width + c4 * 4) = make_float4(s.x / ws, s.y / ws, s.z / ws, s.w / ws);
    }
}

// outputs[:, t*nf + c] = pred[:, c];  field <- cat(field[:, nf:], pred)
__global__ void step_update_kernel(const G4cStepUpdateDesc d) {
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < d.n_nodes; n += (int64_t)gridDim.x * blockDim.x) {
        float* f = d.node_in + (size_t)n * d.in_stride;
        const float* p = d.pred + (size_t)n * d.nf;
        for (int c = 0; c + d.nf < d.field_width; ++c) f[c] = f[c + d.nf];
        for (int c = 0; c < d.nf; ++c) {
            const float v = p[c];
            f[d.field_width - d.nf + c] = v;
            d.outputs[(size_t)n * d.out_stride + (size_t)d.t * d.nf + c] = v;
        }
    }
}

__global__ void halo_pack_kernel(const G4cHaloDesc d) {
    const int V = d.width / 4;
    const int64_t total = d.n_rows * V;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / V;
        const int c4 = (int)(t % V);
        *reinterpret_cast<float4*>(d.dst + (size_t)i * d.width + c4 * 4) =
            *reinterpret_cast<const float4*>(d.src + (size_t)d.idx[i] * d.width + c4 * 4);
    }
}
__global__ void halo_unpack_kernel(const G4cHaloDesc d) {
    const int V = d.width / 4;
    const int64_t total = d.n_rows * V;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / V;
        const int c4 = (int)(t % V);
        *reinterpret_cast<float4*>(d.dst + (size_t)d.idx[i] * d.width + c4 * 4) =
            *reinterpret_cast<const float4*>(d.src + (size_t)i * d.width + c4 * 4);
    }
}

// ---- halo exchange over peer memory: pack + put + signal + wait + unpack in one kernel (see G4cHaloPutDesc)
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__global__ void __launch_bounds__(256) halo_put_kernel(const G4cHaloPutDesc d) {
    __shared__ int s_last;
    // state[0] = exchanges completed so far on this rank (the same number on every rank: the exchange sequence is collective);
    // it is rewritten only after EVERY block has passed the second counter below, so every block reads the same value here
    const uint64_t seq = *reinterpret_cast<volatile uint64_t*>(d.state);
    const uint64_t epoch = seq + 1;
    const size_t half = (seq & 1) ? (size_t)d.mail_stride : 0;       // the mailbox half this exchange uses (see g4c.h)
    const int V = d.width / 4;
    const int64_t total = d.n_rows * V;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t i = t / V;
        const int c4 = (int)(t % V);
        int p = 0;
        while (p + 1 < d.n_peers && i >= d.seg_start[p + 1]) ++p;
        const float4 v = *reinterpret_cast<const float4*>(d.src + (size_t)d.send_idx[i] * d.width + c4 * 4);
        *reinterpret_cast<float4*>(d.dst[p] + half + (size_t)(i - d.seg_start[p]) * d.width + c4 * 4) = v;       // NVLink store
    }
    __threadfence_system();                       // this thread's peer stores are performed system-wide
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(reinterpret_cast<unsigned long long*>(d.state + 1), 1ull) == (unsigned long long)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        // the last block to finish its stores: every block's stores are fenced -> tell the neighbours
        __threadfence_system();
        if ((int)threadIdx.x < d.n_peers) st_release_sys(d.peer_flag[threadIdx.x], epoch);
    }
    // every block waits for the neighbours' rows (the flags are in THIS rank's memory: the polling stays on this GPU)
    if ((int)threadIdx.x < d.n_peers) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(d.my_flag[threadIdx.x]) < epoch) {
            // ranks may legitimately be seconds apart (plan building, graph capture, a host-side check on one rank); only a
            // neighbour that stays silent for a minute is treated as a protocol error, which must not hang the GPU for good
            if (global_timer_ns() - t0 > 60000000000ull) __trap();
        }
    }
    __syncthreads();
    // mailbox half -> the ghost rows of the feature array (L2 loads: the rows were written by other GPUs)
    const int64_t total_in = d.n_recv * V;
    const float4* mail = reinterpret_cast<const float4*>(d.mail + half);
    float4* ghost = reinterpret_cast<float4*>(d.ghost);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total_in; t += stride) ghost[t] = __ldcg(mail + t);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(reinterpret_cast<unsigned long long*>(d.state + 2), 1ull) == (unsigned long long)gridDim.x - 1) {
            d.state[1] = 0;
            d.state[2] = 0;
            __threadfence();
            *reinterpret_cast<volatile uint64_t*>(d.state) = epoch;
        }
    }
}

int halo_put_launch(const G4cHaloPutDesc& d, cudaStream_t st) {
    const int64_t work = std::max(d.n_rows, d.n_recv) * (d.width / 4);
    // every block spins until the neighbours answer, so the grid must be co-resident: at most one block per SM
    const int grid = (int)std::min<int64_t>(std::max<int64_t>(grid_for(work, 256), 1), device_sms());
    halo_put_kernel<<<grid, 256, 0, st>>>(d);
    count_launch();
    return check_launch("halo_put_kernel");
}

int seg_reduce_launch(const G4cSegReduceDesc& d, cudaStream_t st) {
    if (d.n_groups == 0) return G4C_OK;
    seg_reduce_kernel<<<grid_for(d.n_groups * (d.width / 4), 256), 256, 0, st>>>(d);
    count_launch();
    return check_launch("seg_reduce_kernel");
}
int project_launch(const G4cProjectDesc& d, cudaStream_t st) {
    if (d.n_edges == 0) return G4C_OK;
    project_kernel<<<grid_for(d.n_edges * (d.n_feat + d.n_extra), 256), 256, 0, st>>>(d);
    count_launch();
    return check_launch("project_kernel");
}
int edge_to_node_launch(const G4cEdgeToNodeDesc& d, cudaStream_t st) {
    if (d.n_nodes == 0) return G4C_OK;
    edge_to_node_kernel<<<grid_for(d.n_nodes * d.n_feat, 256), 256, 0, st>>>(d);
    count_launch();
    return check_launch("edge_to_node_kernel");
}
int interp_launch(const G4cInterpDesc& d, cudaStream_t st) {
    if (d.n_out == 0) return G4C_OK;
    interp_kernel<<<grid_for(d.n_out * (d.width / 4), 256), 256, 0, st>>>(d);
    count_launch();
    return check_launch("interp_kernel");
}
int step_update_launch(const G4cStepUpdateDesc& d, cudaStream_t st) {
    if (d.n_nodes == 0) return G4C_OK;
    step_update_kernel<<<grid_for(d.n_nodes, 256), 256, 0, st>>>(d);
    count_launch();
    return check_launch("step_update_kernel");
}
int halo_launch(const G4cHaloDesc& d, cudaStream_t st, bool pack) {
    if (d.n_rows == 0) return G4C_OK;
    const int grid = grid_for(d.n_rows * (d.width / 4), 256);
    if (pack) halo_pack_kernel<<<grid, 256, 0, st>>>(d);
    else halo_unpack_kernel<<<grid, 256, 0, st>>>(d);
    count_launch();
    return check_launch(pack ? "halo_pack_kernel" : "halo_unpack_kernel");
}

}  // namespace g4c
