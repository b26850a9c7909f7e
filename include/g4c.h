/* g4c.h — C ABI of libg4c.so: the B200 (sm_100a) message-passing hot path of graphs4cfd.
 *
 * The reference (mario-linov/graphs4cfd) is pure Python and has no FFI of its own; its
 * boundary for this path is the Python class API of graphs4cfd/nn/blocks.py.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference
 * root).  The Python mirror of that class API lives in graphs4cfd_b200/blocks.py and calls
 * these functions through ctypes (see INTEGRATION.md for the binding a maintainer adds).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says host
 *  - feature matrices are row-major fp32 [rows, H]; index arrays are int32
 *  - the library never allocates, frees or retains device memory; nothing synchronises
 *  - every kernel is launched on the cudaStream_t passed as `void* stream` (CUDA-graph capturable)
 *  - return 0 on success, a G4C_E* code otherwise; g4c_last_error() gives the text (thread local)
 *  - weights: W_t[l] is the TRANSPOSE of torch's nn.Linear.weight, i.e. row-major [in_l, out_l],
 *    EXCEPT a final layer narrower than 16 outputs, which stays in torch layout [out, in]
 */
#ifndef G4C_H
#define G4C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G4C_VERSION 100

#if defined(__GNUC__)
#define G4C_API __attribute__((visibility("default")))
#else
#define G4C_API
#endif

enum { G4C_OK = 0, G4C_EINVAL = 1, G4C_EUNSUPPORTED = 2, G4C_ECUDA = 3 };
enum { G4C_ACT_NONE = 0, G4C_ACT_SELU = 1, G4C_ACT_TANH = 2 };
enum { G4C_AGGR_MEAN = 0, G4C_AGGR_SUM = 1 };
/* arithmetic of the dense layers */
enum { G4C_PREC_FP32 = 0,     /* fp32 FFMA on CUDA cores: exact-fp32 parity path            */
       G4C_PREC_FP16X3 = 1 }; /* tcgen05 tensor cores, operands split hi+lo fp16, 3 MMAs    */
/* kernel behind g4c_edge_aggr_fwd */
enum { G4C_EDGE_AUTO = 0,     /* v5 when the launch has a fixed in-degree and no permutations, v3 otherwise */
       G4C_EDGE_V3 = 1,       /* csrc/mp_edge_pair.cu: cp.async loaders, any topology        */
       G4C_EDGE_V5 = 2 };     /* csrc/mp_edge_v5.cu: TMA tiles, packed fp32 epilogues        */

#define G4C_MAX_LAYERS 3
#define G4C_MAX_SEGS 3

/* One reference `MLP` (graphs4cfd/nn/blocks.py:129-144): Linear+SELU chain, optional LayerNorm
 * (eps 1e-5, affine) after the last Linear.  Hidden widths must all equal `hidden`. */
typedef struct {
    int32_t n_layers;                 /* 2 or 3 (number of nn.Linear)                         */
    int32_t in_width;                 /* K of linear_1 = sum of the concatenated segments     */
    int32_t hidden;                   /* H: width of every layer but possibly the last        */
    int32_t out_width;                /* width of the last layer (== hidden, or < 16)         */
    const float* W_t[G4C_MAX_LAYERS]; /* see "weights" above                                  */
    const float* b[G4C_MAX_LAYERS];
    const float* ln_gamma;            /* NULL = no layer_norm                                 */
    const float* ln_beta;
} G4cMlp;

/* One input segment of a concatenation `torch.cat((seg0, seg1, ...), dim=-1)`. */
typedef struct {
    const float* ptr;                 /* [*, width] rows, row stride `stride` floats          */
    const int32_t* gather;            /* NULL: row r of the tile reads row r; else gather[r]  */
    int32_t width;
    int32_t stride;
    float scale;                      /* multiplies the segment (UpMP uses -1 on e_hl)        */
    int32_t _pad;
} G4cSeg;

/* out = act( MLP( cat(segs) ) ) over `rows` rows.
 * Replaces: MLP.forward (blocks.py:143) as used for encoders/decoders (nn/mus_gnn.py:317-318,369;
 * nn/remus_gnn.py:132-140,195), the MLP of DownMP (blocks.py:229), UpMP (blocks.py:285) and
 * UpEdgeMP (blocks.py:456).  Optional `residual` ([rows, out_width], stride res_stride) is added
 * after the MLP (nn/mus_gnn.py:373). */
typedef struct {
    int64_t rows;
    int32_t n_segs;
    int32_t act_out;
    G4cSeg seg[G4C_MAX_SEGS];
    G4cMlp mlp;
    float* out;                       /* [rows, out_width], row stride out_stride floats      */
    int32_t out_stride;
    int32_t res_stride;
    const float* residual;            /* NULL or [rows, >=out_width]                          */
} G4cRowMlpDesc;

/* Fused message-passing block:
 *   e' = edge_mlp( cat(e, S[src], T[tgt]) );  agg = aggr_{edges of tgt} e';  t' = node_mlp( cat(agg, T) )
 * Replaces: GNBlock.forward (blocks.py:175-186; S = T = v), EdgeMP.forward (blocks.py:322-333; rows are
 * angles, S = T = e), DownEdgeMP.forward (blocks.py:360-381; S = e1, T = e2, e' not returned).
 * Edges are consumed in AGGREGATION order: sorted by target; target n owns slots
 * [rowptr[n], rowptr[n+1]) or, with fixed_k > 0, [n*k, (n+1)*k).  `edge_perm`/`tgt_perm` map an
 * aggregation-order slot / target to its storage row (NULL = identity) so callers keep the
 * reference's own edge order at the API boundary.  act_* fold the model-level F.selu that
 * follows every block (nn/mus_gnn.py:321) into the store; agg always uses the un-activated e'. */
typedef struct {
    int32_t hidden;
    int32_t aggr;
    int32_t fixed_k;
    int32_t act_e_out;
    int32_t act_t_out;
    int32_t precision;
    int64_t n_targets;
    int64_t n_edges;
    const int32_t* rowptr;            /* [n_targets+1] when fixed_k == 0                      */
    const int32_t* src;               /* [n_edges] source row (in S) per aggregation slot     */
    const int32_t* edge_perm;         /* [n_edges] or NULL                                    */
    const int32_t* tgt_perm;          /* [n_targets] or NULL                                  */
    const float* e_in;                /* [n_edges, H]                                         */
    const float* src_feat;            /* S [*, H]                                             */
    const float* tgt_feat;            /* T [n_targets, H]                                     */
    float* e_out;                     /* [n_edges, H] or NULL (edge output discarded)         */
    float* t_out;                     /* [n_targets, H]                                       */
    G4cMlp edge_mlp;                  /* in_width = 3H                                        */
    G4cMlp node_mlp;                  /* in_width = 2H                                        */
} G4cMpDesc;

/* out[g] = act( reduce_{i in [ptr[g], ptr[g+1])} x[idx[i]] )   (idx NULL = identity)
 * Replaces: scatter(...,'mean')[mask] + activation of DownMP (blocks.py:231-233) and the dynamic half
 * of pool_edge/coalesce (blocks.py:64-67) once the pooled topology is cached at plan time. */
typedef struct {
    int64_t n_groups;
    int32_t width;
    int32_t aggr;
    int32_t act_out;
    int32_t _pad;
    const int32_t* ptr;
    const int32_t* idx;
    const float* x;
    float* out;
} G4cSegReduceDesc;

/* REMuS geometry helpers (all rows are fixed-k grouped by node):
 * project:  out[j, f] = V[col[j], 2f]*U[j,0] + V[col[j], 2f+1]*U[j,1]  for f < F, then the `extra`
 *           per-node scalars glob[col[j]], omega[col[j]] appended (nn/remus_gnn.py:124-130, blocks.py:453-454)
 * edge_to_node: V[n, 2f+c] = sum_m Uinv[n, c, m] * e[n*k+m, f]            (blocks.py:88-114)
 * interp:   y[i] = sum_m w[i*k+m]*x[x_idx[i*k+m]] / sum_m w[i*k+m], scattered to row y_row[i] of a
 *           zero-filled output (blocks.py:34-48, 443-451) */
typedef struct {
    int64_t n_edges;
    int32_t n_feat;                   /* F: V has 2F columns                                  */
    int32_t n_extra;                  /* 0..2 extra per-node scalars                          */
    const int32_t* col;               /* [n_edges] node row in V / extra                      */
    const float* V;                   /* [*, 2F]                                              */
    const float* U;                   /* [n_edges, 2]                                         */
    const float* extra[2];            /* [*, 1] each                                          */
    float* out;                       /* [n_edges, F + n_extra]                               */
} G4cProjectDesc;

typedef struct {
    int64_t n_nodes;
    int32_t k;
    int32_t n_feat;
    const float* Uinv;                /* [n_nodes, 2, k]                                      */
    const float* e;                   /* [n_nodes*k, F]                                       */
    float* V;                         /* [n_nodes, 2F], row stride out_stride                 */
    int32_t out_stride;
    int32_t res_stride;
    const float* residual;            /* optional [n_nodes, >=2F]: V = residual + ...         */
} G4cEdgeToNodeDesc;

typedef struct {
    int64_t n_out;                    /* interpolated rows                                    */
    int32_t k;
    int32_t width;
    const int32_t* x_idx;             /* [n_out*k]                                            */
    const float* w;                   /* [n_out*k]                                            */
    const int32_t* y_row;             /* [n_out] destination row or NULL (identity)           */
    const float* x;                   /* [*, width]                                           */
    float* y;                         /* [*, width]                                           */
} G4cInterpDesc;

/* Rollout state update (GNN.solve + shift_and_replace, nn/model.py:316-327):
 * outputs[:, t*nf:(t+1)*nf] = pred; field = cat(field[:, nf:], pred).  `field` is the first
 * field_width columns of the node-input matrix (row stride in_stride). */
typedef struct {
    int64_t n_nodes;
    int32_t nf;
    int32_t field_width;
    int32_t in_stride;
    int32_t out_stride;
    int32_t t;
    int32_t _pad;
    const float* pred;                /* [n_nodes, nf]                                        */
    float* node_in;
    float* outputs;                   /* [n_nodes, out_stride]                                */
} G4cStepUpdateDesc;

/* Halo exchange staging for the node-range partition (no reference counterpart: the reference
 * is single-device).  pack: buf[i] = x[idx[i]];  unpack: x[idx[i]] = buf[i]. */
typedef struct {
    int64_t n_rows;
    int32_t width;
    int32_t _pad;
    const int32_t* idx;
    const float* src;
    float* dst;
} G4cHaloDesc;

/* Edge half of the message-passing block on the tensor cores (hidden = 128, precision fp16x3), CTA-pair kernel
 * (csrc/mp_edge_pair.cu).  The first Linear of the edge MLP is split exactly,
 *     W1 cat(e, S[src], T[tgt]) + b1 = W1e e + P_r[src] + P_c[tgt],   P_r = S W1s^T,  P_c = T W1t^T + b1,
 * P_r / P_c being per-NODE products made by g4c_rowgemm_fwd.  Computes
 *     e' = LN(tail(selu(W1e e + P_r[src] + P_c[tgt])))        (blocks.py:181, 328, 376)
 *     agg[t] = mean|sum over the in-edges of t of e'           (blocks.py:183, 330, 378)
 * Topology fields have the meaning they have in G4cMpDesc.  Weights are "pair images" (ops.pack_weight_pair:
 * fp16 hi/lo split of s*W, 64 output rows per CTA, SWIZZLE_128B K-major); W_pair[0] holds the e-columns of
 * linear_1 with scale s; the kernel adds p_scale * (P_r[src] + P_c[tgt]) to s * W1e e, so either pass the plain
 * products with p_scale = s, or products already multiplied by s (an exact power of two) with p_scale = 1. */
typedef struct {
    int64_t n_targets, n_edges;
    int32_t fixed_k, n_layers, act_e_out, aggr;
    const int32_t* rowptr;
    const int32_t* src;
    const int32_t* edge_perm;
    const int32_t* tgt_perm;
    const float* e_in;                /* [n_edges, 128]                                        */
    const float* P_r;                 /* [*, 128] rows indexed by source id                    */
    const float* P_c;                 /* [*, 128] rows indexed by target storage row           */
    float* e_out;                     /* [n_edges, 128] or NULL                                */
    float* agg_out;                   /* [*, 128] rows indexed by target storage row           */
    const uint8_t* W[3];              /* pair images of linear_1[:, :128], linear_2, linear_3  */
    float inv_scale[3];               /* 1/s per layer                                         */
    float p_scale;                    /* s of layer 1                                          */
    const float* bias[3];             /* bias[0] unused (folded into P_c)                      */
    const float* gamma;               /* LayerNorm affine or NULL                              */
    const float* beta;
    int32_t variant;                  /* G4C_EDGE_AUTO (0) unless a test / benchmark pins a kernel */
    int32_t _pad;
} G4cEdgeDesc;

/* Row-tile MLP on the tensor cores (hidden = 128, precision fp16x3), CTA-pair kernel (csrc/mp_row_pair.cu):
 *     out = act( [LN]( MLP( cat(seg0[g0], seg1[g1], ...) ) ) [+ residual] )
 * Same role as g4c_rowmlp_fwd (MLP.forward, blocks.py:143, as used by encoders / decoders / DownMP / UpMP /
 * the node model of GNBlock, blocks.py:185, 229, 285) plus n_layers == 1 (a bare Linear: the per-node products
 * P_r, P_c of the split edge model).  Segments are 128 wide (two 64-wide K-blocks) or at most 16 wide (one
 * K-block, zero padded).  W[0] = pair images of linear_1 with its columns laid out K-block by K-block
 * ([128, 64*n_kblocks], narrow segments padded to 64 columns); W[l] = pair images of the later layers; when
 * out_width < 16 the last layer's rows are padded to 32 and stored as 16-row images (ops.pack_weight_pair with 32 rows).
 * At most 5 K-blocks in total (the resident weights plus the loaders' row rings must fit in 227 KiB of shared memory);
 * wide segments must be 16-byte aligned with a row stride that is a multiple of 4 floats. */
typedef struct {
    int64_t rows;
    int32_t n_segs;                   /* 1..3                                                  */
    int32_t n_layers;                 /* 1..3                                                  */
    int32_t act_out;
    int32_t out_width;                /* 128, or 1..15 (narrow last layer, no LayerNorm)       */
    int32_t out_stride, res_stride;
    G4cSeg seg[G4C_MAX_SEGS];
    const uint8_t* W[3];
    float inv_scale[3];
    int32_t _pad;
    const float* bias[3];
    const float* gamma;
    const float* beta;
    float* out;                       /* [rows, out_width], row stride out_stride floats       */
    const float* residual;            /* narrow output only: [rows, >= out_width] or NULL      */
    /* dual != 0: two bare Linears of the SAME 128-wide input in one pass (the per-node products P_r = S W1s^T and
     * P_c = S W1t^T + b1 of the split edge model when source and target features are the same tensor):
     * n_layers = 2, n_segs = 1 (width 128), out = x W[0]^T + bias[0], out2 = x W[1]^T + bias[1]; no LayerNorm,
     * no activation; the input is read once. */
    float* out2;                      /* [rows, 128], row stride out_stride floats             */
    int32_t dual, _pad2;
} G4cRowTcDesc;

/* Halo exchange as ONE kernel over NVLink peer memory (no reference counterpart; replaces pack kernel + NCCL all_to_all_single
 * + received-rows placement of the node-range partition).  Every rank owns a MAILBOX in symmetric memory (every rank maps every
 * other rank's): two halves of mail_stride floats.  One call = one exchange of the collective sequence:
 *   1. rows send_idx[seg_start[d] .. seg_start[d+1]) of `src` are stored straight into neighbour d's mailbox at dst[d] (peer
 *      pointer, already offset to where this rank's rows land in d's receive order) + the half of this exchange;
 *   2. the stores are fenced system-wide and the exchange number is published to every neighbour's flag slot (st.release.sys);
 *   3. the kernel waits until every neighbour's number has arrived in this rank's slots (ld.acquire.sys, local polling);
 *   4. the n_recv rows of this rank's mailbox half are copied to `ghost` (the ghost rows of the local feature array).
 * Halves alternate with the exchange number: a neighbour can run at most one exchange ahead (it needs this rank's flag to get
 * past step 3), so it never overwrites rows this rank has not copied out yet.  All ranks must issue the same sequence of calls;
 * neighbour sets must be symmetric and every neighbour is signalled even when no rows go to it.
 * `state` is rank-local device memory shared by ALL exchanges of one engine, zero at plan time: state[0] = exchanges completed,
 * state[1], state[2] = block counters.  A neighbour that does not answer within 60 s traps the kernel instead of hanging the
 * GPU.  CUDA-graph capturable.  The grid is at most one block per SM (all blocks wait in step 3, so they must be co-resident). */
#define G4C_MAX_PEERS 8
typedef struct {
    int64_t n_rows;                   /* rows sent to all neighbours together                 */
    int32_t width;                    /* floats per row, multiple of 4                        */
    int32_t n_peers;                  /* neighbours (sent to AND waited for), <= G4C_MAX_PEERS */
    const float* src;                 /* local feature array                                  */
    const int32_t* send_idx;          /* [n_rows] local rows, grouped by neighbour            */
    int32_t seg_start[G4C_MAX_PEERS + 1];
    int32_t _pad;
    float* dst[G4C_MAX_PEERS];        /* peer pointers into the neighbours' mailboxes, half 0 (NULL when nothing is sent to d) */
    uint64_t* peer_flag[G4C_MAX_PEERS];   /* this rank's slot in neighbour d's flag array     */
    const uint64_t* my_flag[G4C_MAX_PEERS];   /* neighbour d's slot in this rank's flag array */
    uint64_t* state;                  /* [3] rank-local                                       */
    int64_t mail_stride;              /* floats between the two halves of a mailbox (same on every rank) */
    const float* mail;                /* this rank's mailbox, half 0                          */
    float* ghost;                     /* [n_recv, width] where the received rows go           */
    int64_t n_recv;                   /* rows received from all neighbours together           */
} G4cHaloPutDesc;

/* Plan-time graph building on the device: exact 2-D k-nearest neighbours on a uniform cell grid.
 * Replaces: torch_cluster.knn_graph behind connect_knn (transforms/connect.py:58: k in-edges per node, neighbour -> centre,
 * ascending distance) and torch_cluster.knn behind get_knn_interpolate_weights (transforms/interpolate.py:125).
 * The caller sorts the data points by cell id (cell = floor((p - origin) / cell), id = cy * gx + cx) and passes the prefix
 * array; nbr[q, j] = index of the j-th nearest data point of query q (ascending distance, ties to the lower index; -1 when
 * fewer than k points exist).  exclude_self: query q IS data point q and must not be its own neighbour (knn_graph, loop=False). */
typedef struct {
    int64_t n_points, n_queries;
    int32_t k;                        /* 1..16                                                */
    int32_t exclude_self;
    const float* pos;                 /* [n_points, 2]                                        */
    const float* query;               /* [n_queries, 2]                                       */
    const int32_t* cell_start;        /* [gx*gy + 1]                                          */
    const int32_t* sorted_idx;        /* [n_points] point ids grouped by cell, ascending id   */
    float x0, y0, cell;               /* grid origin and cell size                            */
    int32_t gx, gy, _pad;
    int32_t* nbr;                     /* [n_queries, k] out                                   */
} G4cKnnDesc;

G4C_API int g4c_version(void);
G4C_API const char* g4c_last_error(void);

G4C_API int g4c_rowmlp_fwd(const G4cRowMlpDesc* d, void* stream);
G4C_API int g4c_mp_fwd(const G4cMpDesc* d, void* stream);
G4C_API int g4c_rowmlp_tc_fwd(const G4cRowTcDesc* d, void* stream);
G4C_API int g4c_edge_aggr_fwd(const G4cEdgeDesc* d, void* stream);
G4C_API int g4c_seg_reduce_fwd(const G4cSegReduceDesc* d, void* stream);
G4C_API int g4c_project_fwd(const G4cProjectDesc* d, void* stream);
G4C_API int g4c_edge_to_node_fwd(const G4cEdgeToNodeDesc* d, void* stream);
G4C_API int g4c_interp_fwd(const G4cInterpDesc* d, void* stream);
G4C_API int g4c_step_update(const G4cStepUpdateDesc* d, void* stream);
G4C_API int g4c_halo_pack(const G4cHaloDesc* d, void* stream);
G4C_API int g4c_halo_unpack(const G4cHaloDesc* d, void* stream);
G4C_API int g4c_halo_put(const G4cHaloPutDesc* d, void* stream);
G4C_API int g4c_plan_knn(const G4cKnnDesc* d, void* stream);

/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
G4C_API int64_t g4c_launch_count(void);
/* ... of which tensor-core (tcgen05) kernels: g4c_edge_aggr_fwd, g4c_rowmlp_tc_fwd */
G4C_API int64_t g4c_tc_launch_count(void);

/* self tests of the second-generation primitives (A operand in TMEM, tcgen05.cp, CTA pairs); see
 * graphs4cfd_b200/csrc/tc2_test.cu for the meaning of test / flags.  Used by tests/test_gpu_tc2.py. */
G4C_API int g4c_debug_tc2(int32_t test, const float* A, const void* W_pack, float w_inv_scale, const float* P, float* D,
                          int32_t flags, void* stream);

/* in-kernel phase profile of the edge kernel `variant` (G4C_EDGE_V3 / G4C_EDGE_V5; HOST pointer to 64 uint64; only in
 * builds with -DG4C_PROFILE) */
G4C_API int g4c_debug_profile(int32_t variant, uint64_t* out64);

/* one-warp hardware self tests of the bulk-tensor (TMA) copies csrc/mp_edge_v5.cu relies on (csrc/tma_test.cu):
 * 0 = 3-D tile load, 1 = 3-D tile store, 3 = tile load hanging over the end of the tensor.
 * src = [rows * k, 128] fp32; out = [32, 16] (test 1: [rows * k, 128]); tile origin (c0, j, n0). */
G4C_API int g4c_debug_tma(int32_t test, const float* src, int64_t rows, int32_t k, float* out, int32_t c0, int32_t j, int32_t n0,
                          void* stream);

/* host-side plan helper (HOST pointers): Guillard node-nested coarsening, the sequential sweep of
 * transforms/mugs.py:8-29.  senders = int64 [n, k]; coarse_mask = uint8 [n] (out). */
G4C_API int g4c_host_guillard(const int64_t* senders, int64_t n, int32_t k, uint8_t* coarse_mask);

#ifdef __cplusplus
}
#endif
#endif /* G4C_H */
