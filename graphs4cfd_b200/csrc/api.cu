// api.cu — extern "C" surface of libg4c.so (see include/g4c.h): argument validation, dispatch,
// thread-local error text, launch counter, host-side plan helper.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "mp_pair.h"

namespace g4c {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0}, g_tc_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n, bool tensor_core) {
    g_launches.fetch_add(n, std::memory_order_relaxed);
    if (tensor_core) g_tc_launches.fetch_add(n, std::memory_order_relaxed);
}
int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return G4C_ECUDA;
    }
    return G4C_OK;
}

int mp_fp32_dispatch(const G4cMpDesc& d, cudaStream_t st);
int rowmlp_fp32_dispatch(const G4cRowMlpDesc& d, cudaStream_t st);
int seg_reduce_launch(const G4cSegReduceDesc& d, cudaStream_t st);
int project_launch(const G4cProjectDesc& d, cudaStream_t st);
int edge_to_node_launch(const G4cEdgeToNodeDesc& d, cudaStream_t st);
int interp_launch(const G4cInterpDesc& d, cudaStream_t st);
int step_update_launch(const G4cStepUpdateDesc& d, cudaStream_t st);
int halo_launch(const G4cHaloDesc& d, cudaStream_t st, bool pack);
int knn_launch(const G4cKnnDesc& d, cudaStream_t st);
int halo_put_launch(const G4cHaloPutDesc& d, cudaStream_t st);
int tc2_test_launch(int test, const float* A, const void* Wpack, float inv_scale, const float* P, float* D, int flags, cudaStream_t st);

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int check_mlp(const char* who, const G4cMlp& m, int expect_in) {
    if (m.n_layers < 2 || m.n_layers > G4C_MAX_LAYERS) { set_error("%s: n_layers=%d (2..3)", who, m.n_layers); return G4C_EINVAL; }
    if (expect_in >= 0 && m.in_width != expect_in) { set_error("%s: in_width=%d, expected %d", who, m.in_width, expect_in); return G4C_EINVAL; }
    if (m.out_width != m.hidden && (m.out_width < 1 || m.out_width >= 16)) {
        set_error("%s: out_width=%d must equal hidden=%d or be < 16", who, m.out_width, m.hidden); return G4C_EUNSUPPORTED; }
    if (m.out_width != m.hidden && m.ln_gamma) { set_error("%s: layer_norm on a narrow output is unsupported", who); return G4C_EUNSUPPORTED; }
    for (int l = 0; l < m.n_layers; ++l) {
        if (!m.W_t[l] || !m.b[l]) { set_error("%s: NULL weight/bias at layer %d", who, l + 1); return G4C_EINVAL; }
        if (!aligned16(m.W_t[l]) || !aligned16(m.b[l])) { set_error("%s: weights must be 16-byte aligned", who); return G4C_EINVAL; }
    }
    if ((m.ln_gamma == nullptr) != (m.ln_beta == nullptr)) { set_error("%s: ln_gamma/ln_beta must both be set or NULL", who); return G4C_EINVAL; }
    return G4C_OK;
}

}  // namespace g4c

using namespace g4c;

extern "C" {

int g4c_version(void) { return G4C_VERSION; }
const char* g4c_last_error(void) { return g_err; }
int64_t g4c_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int64_t g4c_tc_launch_count(void) { return g_tc_launches.load(std::memory_order_relaxed); }

int g4c_rowmlp_fwd(const G4cRowMlpDesc* d, void* stream) {
    if (!d) { set_error("g4c_rowmlp_fwd: NULL descriptor"); return G4C_EINVAL; }
    if (d->rows < 0 || d->n_segs < 1 || d->n_segs > G4C_MAX_SEGS) { set_error("g4c_rowmlp_fwd: rows=%lld n_segs=%d", (long long)d->rows, d->n_segs); return G4C_EINVAL; }
    int kin = 0, n_wide = 0, n_narrow = 0;
    for (int s = 0; s < d->n_segs; ++s) {
        const G4cSeg& sg = d->seg[s];
        if (!sg.ptr || sg.width < 1 || sg.stride < sg.width) { set_error("g4c_rowmlp_fwd: bad segment %d", s); return G4C_EINVAL; }
        if (sg.width == d->mlp.hidden) {
            ++n_wide;
            if (!aligned16(sg.ptr) || (sg.stride & 3)) { set_error("g4c_rowmlp_fwd: wide segment %d must be 16-byte aligned", s); return G4C_EINVAL; }
        } else if (sg.width <= 8) ++n_narrow;
        else { set_error("g4c_rowmlp_fwd: segment width %d must be <= 8 or == hidden", sg.width); return G4C_EUNSUPPORTED; }
        kin += sg.width;
    }
    if (n_wide > 2 || n_narrow > 1) { set_error("g4c_rowmlp_fwd: at most 2 wide + 1 narrow segments"); return G4C_EUNSUPPORTED; }
    int rc = check_mlp("g4c_rowmlp_fwd", d->mlp, kin);
    if (rc) return rc;
    if (!d->out || d->out_stride < d->mlp.out_width) { set_error("g4c_rowmlp_fwd: bad out"); return G4C_EINVAL; }
    if (d->mlp.out_width == d->mlp.hidden && (!aligned16(d->out) || (d->out_stride & 3))) { set_error("g4c_rowmlp_fwd: out must be 16-byte aligned"); return G4C_EINVAL; }
    if (d->rows == 0) return G4C_OK;
    return rowmlp_fp32_dispatch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_mp_fwd(const G4cMpDesc* d, void* stream) {
    if (!d) { set_error("g4c_mp_fwd: NULL descriptor"); return G4C_EINVAL; }
    const int H = d->hidden;
    if (d->n_targets < 0 || d->n_edges < 0) { set_error("g4c_mp_fwd: negative sizes"); return G4C_EINVAL; }
    if (d->n_edges > 0x7fffffffLL || d->n_targets > 0x7fffffffLL) { set_error("g4c_mp_fwd: int32 index range exceeded"); return G4C_EUNSUPPORTED; }
    if (d->fixed_k < 0 || (d->fixed_k == 0 && !d->rowptr)) { set_error("g4c_mp_fwd: need fixed_k > 0 or rowptr"); return G4C_EINVAL; }
    if (d->fixed_k > 0 && d->n_edges != d->n_targets * d->fixed_k) { set_error("g4c_mp_fwd: n_edges != n_targets*fixed_k"); return G4C_EINVAL; }
    if (!d->tgt_feat || !d->t_out || (d->n_edges > 0 && (!d->src || !d->e_in || !d->src_feat))) { set_error("g4c_mp_fwd: NULL tensor"); return G4C_EINVAL; }
    if (d->e_out == d->e_in && d->e_out) { set_error("g4c_mp_fwd: e_out must not alias e_in"); return G4C_EINVAL; }
    if (d->t_out == d->tgt_feat || d->t_out == d->src_feat) { set_error("g4c_mp_fwd: t_out must not alias its inputs"); return G4C_EINVAL; }
    if (!aligned16(d->e_in) || !aligned16(d->src_feat) || !aligned16(d->tgt_feat) || !aligned16(d->e_out) || !aligned16(d->t_out)) {
        set_error("g4c_mp_fwd: feature matrices must be 16-byte aligned"); return G4C_EINVAL; }
    int rc = check_mlp("g4c_mp_fwd(edge_mlp)", d->edge_mlp, 3 * H);
    if (rc) return rc;
    rc = check_mlp("g4c_mp_fwd(node_mlp)", d->node_mlp, 2 * H);
    if (rc) return rc;
    if (d->edge_mlp.hidden != H || d->node_mlp.hidden != H || d->edge_mlp.out_width != H || d->node_mlp.out_width != H) {
        set_error("g4c_mp_fwd: every layer width must equal hidden=%d", H); return G4C_EUNSUPPORTED; }
    if (d->aggr != G4C_AGGR_MEAN && d->aggr != G4C_AGGR_SUM) { set_error("g4c_mp_fwd: aggr=%d", d->aggr); return G4C_EINVAL; }
    if (d->n_targets == 0) return G4C_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->precision != G4C_PREC_FP32) {
        // the tensor-core block is three calls: g4c_rowmlp_tc_fwd (dual: P_r, P_c), g4c_edge_aggr_fwd, g4c_rowmlp_tc_fwd (node model)
        set_error("g4c_mp_fwd: precision=%d; this entry point is the fused CUDA-core block (G4C_PREC_FP32)", d->precision); return G4C_EUNSUPPORTED; }
    return mp_fp32_dispatch(*d, st);
}

int g4c_rowmlp_tc_fwd(const G4cRowTcDesc* d, void* stream) {
    if (!d) { set_error("g4c_rowmlp_tc_fwd: NULL descriptor"); return G4C_EINVAL; }
    if (d->rows < 0 || d->n_segs < 1 || d->n_segs > G4C_MAX_SEGS || d->n_layers < 1 || d->n_layers > 3) {
        set_error("g4c_rowmlp_tc_fwd: rows=%lld n_segs=%d n_layers=%d", (long long)d->rows, d->n_segs, d->n_layers); return G4C_EINVAL; }
    for (int s = 0; s < d->n_segs; ++s) {
        const G4cSeg& sg = d->seg[s];
        if (!sg.ptr || sg.stride < sg.width) { set_error("g4c_rowmlp_tc_fwd: bad segment %d", s); return G4C_EINVAL; }
        if (sg.width == 128) {
            if (!aligned16(sg.ptr) || (sg.stride & 3)) { set_error("g4c_rowmlp_tc_fwd: wide segment %d must be 16-byte aligned", s); return G4C_EINVAL; }
        } else if (sg.width < 1 || sg.width > 16) { set_error("g4c_rowmlp_tc_fwd: segment width %d must be 128 or 1..16", sg.width); return G4C_EUNSUPPORTED; }
    }
    if (d->out_width != 128 && (d->out_width < 1 || d->out_width > 15)) { set_error("g4c_rowmlp_tc_fwd: out_width=%d must be 128 or 1..15", d->out_width); return G4C_EUNSUPPORTED; }
    if (d->out_width != 128 && d->gamma) { set_error("g4c_rowmlp_tc_fwd: layer_norm on a narrow output is unsupported"); return G4C_EUNSUPPORTED; }
    if (d->out_width == 128 && d->residual) { set_error("g4c_rowmlp_tc_fwd: residual is only supported on a narrow output"); return G4C_EUNSUPPORTED; }
    if (!d->out || d->out_stride < d->out_width) { set_error("g4c_rowmlp_tc_fwd: bad out"); return G4C_EINVAL; }
    if (d->out_width == 128 && ((reinterpret_cast<uintptr_t>(d->out) & 31) || (d->out_stride & 7))) { set_error("g4c_rowmlp_tc_fwd: out must be 32-byte aligned"); return G4C_EINVAL; }
    for (int l = 0; l < d->n_layers; ++l)
        if (!d->W[l] || !aligned16(d->W[l]) || !d->bias[l]) { set_error("g4c_rowmlp_tc_fwd: bad weights at layer %d", l + 1); return G4C_EINVAL; }
    if ((d->gamma == nullptr) != (d->beta == nullptr)) { set_error("g4c_rowmlp_tc_fwd: gamma/beta must both be set or NULL"); return G4C_EINVAL; }
    if (d->dual) {
        if (d->n_layers != 2 || d->n_segs != 1 || d->seg[0].width != 128 || d->out_width != 128 || d->gamma || d->act_out != G4C_ACT_NONE) {
            set_error("g4c_rowmlp_tc_fwd: dual needs n_layers=2, one 128-wide segment, 128-wide outputs, no LayerNorm / activation"); return G4C_EUNSUPPORTED; }
        if (!d->out2 || (reinterpret_cast<uintptr_t>(d->out2) & 31)) { set_error("g4c_rowmlp_tc_fwd: dual needs a 32-byte aligned out2"); return G4C_EINVAL; }
    }
    if (d->rows == 0) return G4C_OK;
    return row_pair_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_edge_aggr_fwd(const G4cEdgeDesc* d, void* stream) {
    if (!d) { set_error("g4c_edge_aggr_fwd: NULL descriptor"); return G4C_EINVAL; }
    if (d->n_targets < 0 || d->n_edges < 0 || d->n_edges > 0x7fffffffLL || d->n_targets > 0x7fffffffLL) { set_error("g4c_edge_aggr_fwd: bad sizes"); return G4C_EINVAL; }
    if (d->n_layers < 2 || d->n_layers > 3) { set_error("g4c_edge_aggr_fwd: n_layers=%d (2..3)", d->n_layers); return G4C_EUNSUPPORTED; }
    if (d->fixed_k < 0 || (d->fixed_k == 0 && !d->rowptr)) { set_error("g4c_edge_aggr_fwd: need fixed_k > 0 or rowptr"); return G4C_EINVAL; }
    if (d->fixed_k > 0 && d->n_edges != d->n_targets * d->fixed_k) { set_error("g4c_edge_aggr_fwd: n_edges != n_targets*fixed_k"); return G4C_EINVAL; }
    if (!d->agg_out || !d->P_c || (d->n_edges > 0 && (!d->src || !d->e_in || !d->P_r))) { set_error("g4c_edge_aggr_fwd: NULL tensor"); return G4C_EINVAL; }
    if (!aligned16(d->e_in) || !aligned16(d->P_r) || !aligned16(d->P_c) || ((reinterpret_cast<uintptr_t>(d->e_out) | reinterpret_cast<uintptr_t>(d->agg_out)) & 31)) {
        set_error("g4c_edge_aggr_fwd: inputs must be 16-byte, outputs 32-byte aligned"); return G4C_EINVAL; }
    if (d->e_out && d->e_out == d->e_in) { set_error("g4c_edge_aggr_fwd: e_out must not alias e_in"); return G4C_EINVAL; }
    for (int l = 0; l < d->n_layers; ++l) {
        if (!d->W[l] || !aligned16(d->W[l]) || (l > 0 && !d->bias[l])) { set_error("g4c_edge_aggr_fwd: bad weights at layer %d", l + 1); return G4C_EINVAL; }
    }
    if ((d->gamma == nullptr) != (d->beta == nullptr)) { set_error("g4c_edge_aggr_fwd: gamma/beta must both be set or NULL"); return G4C_EINVAL; }
    if (d->aggr != G4C_AGGR_MEAN && d->aggr != G4C_AGGR_SUM) { set_error("g4c_edge_aggr_fwd: aggr=%d", d->aggr); return G4C_EINVAL; }
    if (d->n_targets == 0) return G4C_OK;
    return edge_pair_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_seg_reduce_fwd(const G4cSegReduceDesc* d, void* stream) {
    if (!d || !d->ptr || !d->x || !d->out || d->n_groups < 0) { set_error("g4c_seg_reduce_fwd: bad descriptor"); return G4C_EINVAL; }
    if ((d->width & 3) || !aligned16(d->x) || !aligned16(d->out)) { set_error("g4c_seg_reduce_fwd: width %% 4 and 16-byte alignment required"); return G4C_EINVAL; }
    return seg_reduce_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_project_fwd(const G4cProjectDesc* d, void* stream) {
    if (!d || !d->col || !d->V || !d->U || !d->out || d->n_edges < 0 || d->n_feat < 1 || d->n_extra < 0 || d->n_extra > 2) {
        set_error("g4c_project_fwd: bad descriptor"); return G4C_EINVAL; }
    for (int x = 0; x < d->n_extra; ++x) if (!d->extra[x]) { set_error("g4c_project_fwd: NULL extra"); return G4C_EINVAL; }
    return project_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_edge_to_node_fwd(const G4cEdgeToNodeDesc* d, void* stream) {
    if (!d || !d->Uinv || !d->e || !d->V || d->n_nodes < 0 || d->k < 1 || d->n_feat < 1 || d->out_stride < 2 * d->n_feat || (d->out_stride & 1)) {
        set_error("g4c_edge_to_node_fwd: bad descriptor"); return G4C_EINVAL; }
    return edge_to_node_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_interp_fwd(const G4cInterpDesc* d, void* stream) {
    if (!d || !d->x_idx || !d->w || !d->x || !d->y || d->n_out < 0 || d->k < 1 || (d->width & 3) || !aligned16(d->x) || !aligned16(d->y)) {
        set_error("g4c_interp_fwd: bad descriptor"); return G4C_EINVAL; }
    return interp_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_step_update(const G4cStepUpdateDesc* d, void* stream) {
    if (!d || !d->pred || !d->node_in || !d->outputs || d->nf < 1 || d->field_width < d->nf || d->in_stride < d->field_width) {
        set_error("g4c_step_update: bad descriptor"); return G4C_EINVAL; }
    return step_update_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_halo_pack(const G4cHaloDesc* d, void* stream) {
    if (!d || !d->idx || !d->src || !d->dst || (d->width & 3)) { set_error("g4c_halo_pack: bad descriptor"); return G4C_EINVAL; }
    return halo_launch(*d, static_cast<cudaStream_t>(stream), true);
}
int g4c_halo_unpack(const G4cHaloDesc* d, void* stream) {
    if (!d || !d->idx || !d->src || !d->dst || (d->width & 3)) { set_error("g4c_halo_unpack: bad descriptor"); return G4C_EINVAL; }
    return halo_launch(*d, static_cast<cudaStream_t>(stream), false);
}

int g4c_halo_put(const G4cHaloPutDesc* d, void* stream) {
    if (!d || !d->state || d->n_rows < 0 || d->n_peers < 0 || d->n_peers > G4C_MAX_PEERS || (d->width & 3) || d->width < 4) {
        set_error("g4c_halo_put: bad descriptor"); return G4C_EINVAL; }
    if (d->n_rows > 0 && (!d->src || !d->send_idx)) { set_error("g4c_halo_put: NULL source"); return G4C_EINVAL; }
    if (d->n_recv < 0 || d->mail_stride < 0 || (d->mail_stride & 3) || (d->n_recv > 0 && (!d->mail || !d->ghost || !aligned16(d->mail) || !aligned16(d->ghost)))) {
        set_error("g4c_halo_put: bad mailbox"); return G4C_EINVAL; }
    if (d->n_recv * d->width > d->mail_stride) { set_error("g4c_halo_put: n_recv rows do not fit a mailbox half"); return G4C_EINVAL; }
    if (d->seg_start[0] != 0 || d->seg_start[d->n_peers] != d->n_rows) { set_error("g4c_halo_put: seg_start does not cover n_rows"); return G4C_EINVAL; }
    for (int p = 0; p < d->n_peers; ++p) {
        if (!d->peer_flag[p] || !d->my_flag[p]) { set_error("g4c_halo_put: NULL flag for neighbour %d", p); return G4C_EINVAL; }
        if (d->seg_start[p + 1] < d->seg_start[p] || (d->seg_start[p + 1] > d->seg_start[p] && !d->dst[p])) {
            set_error("g4c_halo_put: bad segment for neighbour %d", p); return G4C_EINVAL; }
    }
    return halo_put_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_plan_knn(const G4cKnnDesc* d, void* stream) {
    if (!d || !d->pos || !d->query || !d->cell_start || !d->sorted_idx || !d->nbr || d->n_points < 0 || d->n_queries < 0) {
        set_error("g4c_plan_knn: bad descriptor"); return G4C_EINVAL; }
    if (d->k < 1 || d->k > 16) { set_error("g4c_plan_knn: k=%d (1..16)", d->k); return G4C_EUNSUPPORTED; }
    if (d->gx < 1 || d->gy < 1 || !(d->cell > 0.f)) { set_error("g4c_plan_knn: bad grid"); return G4C_EINVAL; }
    if (d->n_points > 0x7fffffffLL || d->n_queries > 0x7fffffffLL) { set_error("g4c_plan_knn: int32 index range exceeded"); return G4C_EUNSUPPORTED; }
    return knn_launch(*d, static_cast<cudaStream_t>(stream));
}

int g4c_debug_tc2(int32_t test, const float* A, const void* W_pack, float w_inv_scale, const float* P, float* D, int32_t flags, void* stream) {
    if (!A || !W_pack || !D || (test == 2 && !P)) { set_error("g4c_debug_tc2: NULL pointer"); return G4C_EINVAL; }
    return tc2_test_launch(test, A, W_pack, w_inv_scale, P, D, flags, static_cast<cudaStream_t>(stream));
}

int g4c_debug_profile(int32_t variant, uint64_t* out64) {
    if (!out64) { set_error("g4c_debug_profile: NULL pointer"); return G4C_EINVAL; }
    cudaDeviceSynchronize();
    return variant == G4C_EDGE_V3 ? edge_pair_profile(reinterpret_cast<unsigned long long*>(out64))
                                  : edge_v5_profile(reinterpret_cast<unsigned long long*>(out64));
}

int g4c_debug_tma(int32_t test, const float* src, int64_t rows, int32_t k, float* out, int32_t c0, int32_t j, int32_t n0, void* stream) {
    if ((test != 0 && test != 1 && test != 3) || !src || !out || rows < 1 || k < 1) { set_error("g4c_debug_tma: bad arguments"); return G4C_EINVAL; }
    return tma_test_launch(test, src, rows, k, out, c0, j, n0, static_cast<cudaStream_t>(stream));
}

int g4c_host_guillard(const int64_t* senders, int64_t n, int32_t k, uint8_t* coarse_mask) {
    if (!senders || !coarse_mask || n < 0 || k < 1) { set_error("g4c_host_guillard: bad arguments"); return G4C_EINVAL; }
    memset(coarse_mask, 1, (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (!coarse_mask[i]) continue;
        for (int m = 0; m < k; ++m) {
            const int64_t s = senders[i * k + m];
            if (s >= 0 && s < n) coarse_mask[s] = 0;
        }
    }
    return G4C_OK;
}

}  // extern "C"
