"""Tensor-level wrappers over the C ABI (include/g4c.h): PyTorch tensors in, PyTorch tensors out.
PyTorch only owns the memory and the stream; every computation below is a libg4c kernel."""
import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L


def _swizzled_images(M: torch.Tensor):
    """fp16 [R, K] (R % 8 == 0, K % 64 == 0) -> [K/64, R, 64] fp16, each a SWIZZLE_128B K-major image."""
    R, K = M.shape
    r = torch.arange(R, device=M.device)
    c = torch.arange(8, device=M.device)
    phys = c.unsqueeze(0) ^ (r.unsqueeze(1) & 7)
    blk = M.view(R, K // 64, 8, 8).permute(1, 0, 2, 3)
    out = torch.empty_like(blk)
    out[:, r.unsqueeze(1), phys, :] = blk
    return out.reshape(K // 64, R, 64)


# tcgen05.mma adds every product into its fp32 accumulator with round-toward-zero: each accumulating MMA pulls the running sum
# toward zero by a fraction of an ulp, a SYSTEMATIC relative shrink that grows with the number of MMAs of a GEMM.  Measured on
# B200 (tools/tc_bias.py, profiles/r2h_tc_bias.txt; the same for centred and offset inputs; torch's fp32 matmul: -5e-10):
#     3 MMAs (one K = 16 step of the 3-term split)  -0.9e-7      24 MMAs (K = 128)  -4.45e-7
#     48 MMAs (K = 256)                             -8.26e-7     72 MMAs (K = 384)  -1.21e-6      i.e.  6.2e-8 + 1.594e-8 n
# LayerNorm cancels a scale error, but encoders, decoders and hidden layers have none, and a rollout integrates the bias step
# after step (drift linear in the step count instead of a random walk: 6.1e-7 per step with the trained 3S-GNN).  The kernels
# multiply the accumulator by 1/s anyway, so the expected shrink is folded into that factor: exact on average, no instruction
# added (bias after compensation < 1.2e-7 even through three layers; 100-step rollout drift 9.0e-5 -> 3.6e-5 at a noise floor of
# 2.5e-5, profiles/r2h_rollout_drift.txt).
TC_RZ_SHRINK = (6.2e-8, 1.594e-8)          # relative shrink of an accumulator = a + b * (number of MMAs it received)


def rz_compensation(n_mma: int) -> float:
    """Factor that undoes the mean round-toward-zero shrink of an accumulator that received ``n_mma`` MMAs."""
    a, b = TC_RZ_SHRINK
    return 1.0 + (a + b * n_mma if n_mma > 0 else 0.0)


def weight_scale(W: torch.Tensor) -> float:
    """Power-of-two s keeping |s*W| below 1024 so the fp16 residual of s*W stays a normal number."""
    amax = float(W.abs().max())
    return 1.0 if amax == 0.0 else 2.0 ** min(14, math.floor(math.log2(1000.0 / amax)))


def pack_weight_pair(W: torch.Tensor, s: Optional[float] = None):
    """Operand images of W [N, K] (N = 128 or 32) for a CTA pair (cta_group::2): CTA r holds output rows
    [N/2*r, N/2*(r+1)).  Layout [r][kb][hi|lo][N/2 rows][64 fp16].  Returns (uint8 tensor, 1/s)."""
    N, K = W.shape
    assert N in (128, 32) and K % 64 == 0
    W = W.detach().float()
    s = weight_scale(W) if s is None else s
    Ws = W * s
    hi = Ws.half()
    lo = (Ws - hi.float()).half()
    parts = []
    for r in range(2):
        h = _swizzled_images(hi[N // 2 * r:N // 2 * (r + 1)].contiguous())      # [kb, N/2, 64]
        l = _swizzled_images(lo[N // 2 * r:N // 2 * (r + 1)].contiguous())
        parts.append(torch.stack([h, l], dim=1))                        # [kb, 2, 64, 64]
    pack = torch.stack(parts, dim=0).contiguous()                       # [2, kb, 2, 64, 64]
    return pack.view(torch.uint8).reshape(-1), 1.0 / s


class MlpPack:
    """Device-resident weights of one reference ``MLP`` (graphs4cfd/nn/blocks.py:129-144) in the
    layout the kernels read: every Linear transposed to [in, out] (a final layer narrower than 16
    stays [out, in]); optional LayerNorm affine."""

    def __init__(self, linears: Sequence[Tuple[torch.Tensor, torch.Tensor]], ln=None):
        assert 2 <= len(linears) <= L.MAX_LAYERS, "MLP must have 2 or 3 Linear layers"
        self.n_layers = len(linears)
        self.in_width = int(linears[0][0].shape[1])
        self.hidden = int(linears[0][0].shape[0])
        self.out_width = int(linears[-1][0].shape[0])
        for W, _ in linears[:-1]:
            if W.shape[0] != self.hidden:
                raise RuntimeError("graphs4cfd_b200: all hidden widths of an MLP must be equal")
        self._linears = [(W.detach().float(), b.detach().float()) for W, b in linears]
        self._ln_raw = None if ln is None else (ln[0].detach().float(), ln[1].detach().float())
        self._tc_row, self._tc_edge, self._tc_node_wide = {}, None, None
        self.W_t, self.b = [], []
        for i, (W, b) in enumerate(linears):
            W = W.detach().float()
            narrow_last = (i == self.n_layers - 1) and self.out_width != self.hidden
            self.W_t.append(W.contiguous().clone() if narrow_last else W.t().contiguous())
            self.b.append(b.detach().float().contiguous().clone())
        self.ln = None if ln is None else (ln[0].detach().float().contiguous().clone(),
                                           ln[1].detach().float().contiguous().clone())
        L.require_cuda_f32(*self.W_t, *self.b)
        self._struct = None

    # ---- tensor-core (fp16x3, CTA-pair kernels) operand images, built on first use
    def tc_row_ok(self, seg_widths) -> bool:
        return (self.hidden == 128 and sum(seg_widths) == self.in_width
                and all(w == 128 or 1 <= w <= 16 for w in seg_widths)
                and sum(2 if w == 128 else 1 for w in seg_widths) <= 5
                and (self.out_width == 128 or (self.out_width < 16 and self.ln is None and self.n_layers > 1)))

    def tc_row(self, seg_widths) -> "RowPairPack":
        key = tuple(seg_widths)
        if key not in self._tc_row:
            self._tc_row[key] = RowPairPack(self._linears, list(seg_widths), self._ln_raw)
        return self._tc_row[key]

    def tc_edge_ok(self, feat_width: int = 128) -> bool:
        """Edge model over cat(e [128], S[src] [feat_width], T[tgt] [feat_width]); feat_width 256 is the first block behind an
        up-sampling of the MuGS models (nn/mugs_gnn.py:34: mp121 takes cat(interpolated, skip))."""
        return (self.hidden == 128 and self.out_width == 128 and feat_width in (128, 256)
                and self.in_width == 128 + 2 * feat_width)

    def tc_node_ok(self, feat_width: int = 128) -> bool:
        """Node model over cat(aggregate [128], T [feat_width])."""
        if feat_width == 128:
            return self.tc_row_ok([128, 128])
        return (feat_width == 256 and self.hidden == 128 and self.out_width == 128 and self.in_width == 384
                and self.n_layers >= 2)

    def tc_node_wide(self):
        """Node model with 256-wide node features: the row kernel keeps linear_1 resident for at most two 128-wide segments,
        so linear_1 is applied in two launches that add up exactly like the single product would:
            Q   = T W1[:, 128:]^T + b1          (bare Linear over the two halves of T)
            out = MLP'(cat(agg, Q))   with linear_1' = [W1[:, :128] | I], bias 0, the later layers unchanged.
        Returns (RowPairPack of Q, RowPairPack of MLP')."""
        if self._tc_node_wide is None:
            W1, b1 = self._linears[0]
            W1 = W1.detach().float()
            eye = torch.eye(128, device=W1.device, dtype=torch.float32)
            first = (torch.cat([W1[:, :128], eye], dim=1).contiguous(), torch.zeros_like(b1.detach().float()))
            self._tc_node_wide = (RowPairPack([(W1[:, 128:].contiguous(), b1)], [128, 128]),
                                  RowPairPack([first] + list(self._linears[1:]), [128, 128], self._ln_raw))
        return self._tc_node_wide

    def tc_edge(self):
        """(EdgePairPack, projection of the source features, projection of the target features)."""
        if self._tc_edge is None:
            ep = EdgePairPack(self._linears, self._ln_raw)
            zero = torch.zeros_like(ep.b1)
            # P_r, P_c leave the row kernel already multiplied by the layer-1 scale s (a power of two: exact), so the
            # edge kernel's loaders only add them (G4cEdgeDesc.p_scale = 1)
            segs = [128] * (ep.feat_width // 128)
            self._tc_edge = (ep, RowPairPack([(ep.W1s, zero)], segs, out_scale=ep.p_scale),
                             RowPairPack([(ep.W1t, ep.b1)], segs, out_scale=ep.p_scale))
        return self._tc_edge

    @classmethod
    def from_module(cls, mlp_module):
        seq = mlp_module.MLP
        linears, i = [], 1
        while hasattr(seq, f"linear_{i}"):
            lin = getattr(seq, f"linear_{i}")
            linears.append((lin.weight, lin.bias))
            i += 1
        ln = (seq.layer_norm.weight, seq.layer_norm.bias) if hasattr(seq, "layer_norm") else None
        return cls(linears, ln)

    @classmethod
    def from_state(cls, params, prefix, device):
        linears, i = [], 1
        while f"{prefix}.MLP.linear_{i}.weight" in params:
            linears.append((params[f"{prefix}.MLP.linear_{i}.weight"].to(device),
                            params[f"{prefix}.MLP.linear_{i}.bias"].to(device)))
            i += 1
        g = params.get(f"{prefix}.MLP.layer_norm.weight")
        ln = None if g is None else (g.to(device), params[f"{prefix}.MLP.layer_norm.bias"].to(device))
        return cls(linears, ln)

    def struct(self) -> L.Mlp:
        if self._struct is None:
            m = L.Mlp()
            m.n_layers, m.in_width, m.hidden, m.out_width = self.n_layers, self.in_width, self.hidden, self.out_width
            for i in range(self.n_layers):
                m.W_t[i] = self.W_t[i].data_ptr()
                m.b[i] = self.b[i].data_ptr()
            if self.ln is not None:
                m.ln_gamma, m.ln_beta = self.ln[0].data_ptr(), self.ln[1].data_ptr()
            self._struct = m
        return self._struct


class EdgePairPack:
    """Tensor-core (CTA-pair) operand images of a reference edge MLP whose first Linear takes
    cat(e, S[src], T[tgt]) (blocks.py:181, 328, 376): linear_1 is split into its e / source / target column
    blocks sharing one power-of-two scale; the e block and the later layers are stored as pair images for
    g4c_edge_aggr_fwd, the source / target blocks as fp32 [128, 128] for the per-node products P_r, P_c."""

    def __init__(self, linears: Sequence[Tuple[torch.Tensor, torch.Tensor]], ln=None):
        assert 2 <= len(linears) <= 3
        W1, b1 = linears[0]
        assert W1.shape in ((128, 384), (128, 640)), "edge MLP of the tensor-core path: hidden 128, input cat(e, s, t)"
        self.feat_width = fw = (W1.shape[1] - 128) // 2
        W1 = W1.detach().float()
        s1 = weight_scale(W1)
        self.n_layers = len(linears)
        self.W_pair, self.inv_scale = [], []
        pk, inv = pack_weight_pair(W1[:, :128].contiguous(), s1)
        self.W_pair.append(pk)
        self.inv_scale.append(inv * rz_compensation(24))          # K = 128: 8 K-steps x 3 split products
        self.p_scale = s1
        self.W1s = W1[:, 128:128 + fw].contiguous()
        self.W1t = W1[:, 128 + fw:].contiguous()
        self.b1 = b1.detach().float().contiguous().clone()
        self.bias = [self.b1]
        for W, b in linears[1:]:
            assert W.shape == (128, 128)
            pk, inv = pack_weight_pair(W)
            self.W_pair.append(pk)
            self.inv_scale.append(inv * rz_compensation(24))
            self.bias.append(b.detach().float().contiguous().clone())
        self.ln = None if ln is None else (ln[0].detach().float().contiguous().clone(),
                                           ln[1].detach().float().contiguous().clone())


class RowPairPack:
    """Tensor-core (CTA-pair) operand images of a reference ``MLP`` (blocks.py:129-144) or of a bare Linear,
    for g4c_rowmlp_tc_fwd.  ``seg_widths`` are the widths of the concatenated input segments in order (each 128
    or <= 16): linear_1's columns are re-laid out K-block by K-block (narrow segments zero-padded to 64)."""

    def __init__(self, linears: Sequence[Tuple[torch.Tensor, torch.Tensor]], seg_widths: Sequence[int], ln=None,
                 out_scale: float = 1.0):
        assert 1 <= len(linears) <= 3
        assert out_scale == 1.0 or (len(linears) == 1 and ln is None), "out_scale: bare Linear only"
        self.out_scale = float(out_scale)
        dev = linears[0][0].device
        W1 = linears[0][0].detach().float()
        assert W1.shape[0] == 128 and W1.shape[1] == sum(seg_widths), "hidden width 128 and matching input width"
        cols, c0 = [], 0
        for w in seg_widths:
            assert w == 128 or 1 <= w <= 16, "segments must be 128 wide or at most 16"
            blk = W1[:, c0:c0 + w]
            if w != 128:
                blk = torch.cat([blk, torch.zeros(128, 64 - w, device=dev)], dim=1)
            cols.append(blk)
            c0 += w
        self.seg_widths = list(seg_widths)
        self.n_layers = len(linears)
        self.out_width = int(linears[-1][0].shape[0])
        self.W_pair, self.inv_scale, self.bias = [], [], []
        for i, (W, b) in enumerate(linears):
            W = torch.cat(cols, dim=1).contiguous() if i == 0 else W.detach().float()
            b = b.detach().float()
            if i == self.n_layers - 1 and self.out_width != 128:
                assert self.out_width < 16 and i > 0, "narrow output: at most 15 columns, not on a single Linear"
                W = torch.cat([W, torch.zeros(32 - self.out_width, W.shape[1], device=dev)], dim=0)
                b = torch.cat([b, torch.zeros(32 - self.out_width, device=dev)])
            else:
                assert W.shape[0] == 128
            pk, inv = pack_weight_pair(W.contiguous())
            self.W_pair.append(pk)
            # MMAs into this layer's accumulator: 3 split products per K = 16 step; a 128-wide segment has 8 steps, a narrow one 1
            n_mma = 3 * (sum(8 if w == 128 else 1 for w in seg_widths) if i == 0 else 8)
            self.inv_scale.append(inv * self.out_scale * rz_compensation(n_mma))           # out = out_scale * (x W^T + b)
            self.bias.append((b * self.out_scale).contiguous().clone())
        self.ln = None if ln is None else (ln[0].detach().float().contiguous().clone(),
                                           ln[1].detach().float().contiguous().clone())


def rowmlp_tc(pack: RowPairPack, segs, rows: Optional[int] = None, act=None, out=None, residual=None):
    """g4c_rowmlp_tc_fwd: out = act([LN](MLP(cat(segs)))) [+ residual]; segs = [(tensor, gather|None, scale)]."""
    d = L.RowTcDesc()
    tens = [s[0] for s in segs]
    assert [int(t.shape[1]) for t in tens] == pack.seg_widths, "segment widths do not match the packed linear_1"
    if rows is None:
        rows = int(segs[0][1].numel()) if segs[0][1] is not None else int(tens[0].shape[0])
    d.rows, d.n_segs, d.n_layers, d.act_out, d.out_width = rows, len(segs), pack.n_layers, L.ACTS[act], pack.out_width
    for i, (t, gather, scale) in enumerate(segs):
        if not (t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1):
            raise RuntimeError("rowmlp_tc: segments must be fp32 CUDA tensors with unit column stride")
        d.seg[i] = _seg_struct(t, gather, scale)
    for i in range(pack.n_layers):
        d.W[i] = pack.W_pair[i].data_ptr()
        d.inv_scale[i] = pack.inv_scale[i]
        d.bias[i] = pack.bias[i].data_ptr()
    if pack.ln is not None:
        d.gamma, d.beta = pack.ln[0].data_ptr(), pack.ln[1].data_ptr()
    if out is None:
        out = torch.empty(rows, pack.out_width, device=tens[0].device, dtype=torch.float32)
    d.out, d.out_stride = out.data_ptr(), int(out.stride(0))
    if residual is not None:
        d.residual, d.res_stride = residual.data_ptr(), int(residual.stride(0))
    L.launch("g4c_rowmlp_tc_fwd", d, out, residual, *tens, *[s[1] for s in segs])
    return out


def dual_linear_tc(pack_a: RowPairPack, pack_b: RowPairPack, x: torch.Tensor, out_a=None, out_b=None):
    """g4c_rowmlp_tc_fwd in dual mode: (x W_a^T + b_a, x W_b^T + b_b) with x [rows, 128] read once."""
    L.require_cuda_f32(x)
    assert pack_a.n_layers == 1 and pack_b.n_layers == 1 and pack_a.seg_widths == [128] and pack_b.seg_widths == [128]
    rows = int(x.shape[0])
    d = L.RowTcDesc()
    d.rows, d.n_segs, d.n_layers, d.act_out, d.out_width, d.dual = rows, 1, 2, L.ACTS[None], 128, 1
    d.seg[0] = _seg_struct(x, None, 1.0)
    for i, pk in enumerate((pack_a, pack_b)):
        d.W[i], d.inv_scale[i], d.bias[i] = pk.W_pair[0].data_ptr(), pk.inv_scale[0], pk.bias[0].data_ptr()
    if out_a is None:
        out_a = torch.empty(rows, 128, device=x.device, dtype=torch.float32)
    if out_b is None:
        out_b = torch.empty(rows, 128, device=x.device, dtype=torch.float32)
    assert out_a.stride(0) == out_b.stride(0)
    d.out, d.out2, d.out_stride = out_a.data_ptr(), out_b.data_ptr(), int(out_a.stride(0))
    L.launch("g4c_rowmlp_tc_fwd", d, x, out_a, out_b)
    return out_a, out_b


EDGE_VARIANTS = {"auto": L.EDGE_AUTO, "v3": L.EDGE_V3, "v5": L.EDGE_V5}
EDGE_VARIANT_DEFAULT = "auto"       # benchmarks may pin a kernel for every launch of the process (bench.py --edge-variant)


def edge_aggr(pack: EdgePairPack, topo: "MpTopo", e_in, P_r, P_c, aggr="mean", act_e=None, want_e=True,
              e_out=None, agg_out=None, p_prescaled=False, variant=None):
    """g4c_edge_aggr_fwd: returns (agg [n_targets,128], e_out|None).  ``p_prescaled``: P_r / P_c already carry the
    layer-1 scale (MlpPack.tc_edge's projections do); ``variant`` pins a kernel (tests, tools/bench_edge.py)."""
    L.require_cuda_f32(e_in, P_r, P_c)
    d = L.EdgeDesc()
    d.n_targets, d.n_edges, d.fixed_k, d.n_layers = topo.n_targets, topo.n_edges, topo.fixed_k, pack.n_layers
    d.act_e_out, d.aggr = L.ACTS[act_e], (L.AGGR_MEAN if aggr == "mean" else L.AGGR_SUM)
    d.rowptr = 0 if topo.rowptr is None else topo.rowptr.data_ptr()
    d.src = topo.src.data_ptr()
    d.edge_perm = 0 if topo.edge_perm is None else topo.edge_perm.data_ptr()
    d.tgt_perm = 0 if topo.tgt_perm is None else topo.tgt_perm.data_ptr()
    if want_e and e_out is None:
        e_out = torch.empty(topo.n_edges, 128, device=e_in.device, dtype=torch.float32)
    if agg_out is None:
        agg_out = torch.empty(P_c.shape[0], 128, device=e_in.device, dtype=torch.float32)
    d.e_in, d.P_r, d.P_c = e_in.data_ptr(), P_r.data_ptr(), P_c.data_ptr()
    d.e_out, d.agg_out = (e_out.data_ptr() if want_e else 0), agg_out.data_ptr()
    for i in range(pack.n_layers):
        d.W[i] = pack.W_pair[i].data_ptr()
        d.inv_scale[i] = pack.inv_scale[i]
        d.bias[i] = pack.bias[i].data_ptr()
    d.p_scale = 1.0 if p_prescaled else pack.p_scale
    d.variant = EDGE_VARIANTS[EDGE_VARIANT_DEFAULT if variant is None else variant]
    if pack.ln is not None:
        d.gamma, d.beta = pack.ln[0].data_ptr(), pack.ln[1].data_ptr()
    L.launch("g4c_edge_aggr_fwd", d, e_in, P_r, P_c, e_out, agg_out, topo.src, topo.rowptr, topo.edge_perm, topo.tgt_perm)
    return agg_out, (e_out if want_e else None)


FP16_SPLIT_MAX = 3.0e4      # |x| the fp16 (hi, lo) operand split accepts with a factor 2 of head-room (fp16 max 65504)


def check_fp16_range(t: torch.Tensor, what: str):
    """The tensor-core path splits activations into fp16 (hi, lo) without a per-tensor scale: LayerNorm / SELU / tanh
    outputs are O(1) by construction, but RAW inputs (fields, relative positions, encoder inputs) are whatever the caller
    passes.  |x| > 65504 would become inf (then NaN through the lo term).  Called once per plan / solve (it synchronises),
    never per step.  Values below 6e-5 keep an ABSOLUTE accuracy of 3e-8 (fp16 subnormal lo term), which is what a dot
    product with O(1) partners needs."""
    if t.numel() == 0:
        return
    amax = float(t.detach().abs().max())
    if not (amax <= FP16_SPLIT_MAX):          # also catches NaN
        raise RuntimeError(f"graphs4cfd_b200: {what} has |x| up to {amax:.3g}; the fp16x3 tensor-core path accepts raw inputs "
                           f"up to {FP16_SPLIT_MAX:.0e} (rescale the input, or use precision='fp32')")


class MpTopo:
    """Static topology of one message-passing level in AGGREGATION order (edges sorted by target).
    fixed_k > 0: target n owns slots [n*k, (n+1)*k) and the caller's order is already that order
    (kNN graphs, transforms/connect.py:58); otherwise CSR ``rowptr`` + ``edge_perm`` (slot -> caller row)."""

    def __init__(self, n_targets, n_edges, src, fixed_k=0, rowptr=None, edge_perm=None, tgt_perm=None):
        self.n_targets, self.n_edges = int(n_targets), int(n_edges)
        self.src, self.fixed_k, self.rowptr, self.edge_perm, self.tgt_perm = src, int(fixed_k), rowptr, edge_perm, tgt_perm

    @classmethod
    def from_edge_index(cls, edge_index: torch.Tensor, n_targets: int, device=None):
        device = edge_index.device if device is None else device
        row, col = edge_index[0], edge_index[1]
        E = int(row.numel())
        if E > 0 and n_targets > 0 and E % n_targets == 0:
            k = E // n_targets
            expect = torch.arange(n_targets, device=col.device).repeat_interleave(k)
            if torch.equal(col, expect):
                return cls(n_targets, E, row.to(device=device, dtype=torch.int32).contiguous(), fixed_k=k)
        perm = torch.sort(col, stable=True).indices
        counts = torch.bincount(col, minlength=n_targets)
        rowptr = torch.zeros(n_targets + 1, dtype=torch.int64, device=col.device)
        rowptr[1:] = counts.cumsum(0)
        return cls(n_targets, E, row[perm].to(device=device, dtype=torch.int32).contiguous(),
                   rowptr=rowptr.to(device=device, dtype=torch.int32).contiguous(),
                   edge_perm=perm.to(device=device, dtype=torch.int32).contiguous())


def _seg_struct(t: torch.Tensor, gather, scale) -> L.Seg:
    s = L.Seg()
    s.ptr, s.gather = t.data_ptr(), (0 if gather is None else gather.data_ptr())
    s.width, s.stride, s.scale = int(t.shape[1]), int(t.stride(0)), float(scale)
    return s


def rowmlp(pack: MlpPack, segs, rows: Optional[int] = None, act=None, out=None, residual=None, precision="auto"):
    """out = act(MLP(cat(segs, dim=-1)) [+ residual]); segs = [(tensor[*, w], gather_idx|None, scale)].
    precision "fp16x3" runs the tensor-core row kernel (hidden 128); "fp32" the CUDA-core one; "auto" picks the
    tensor-core kernel whenever it supports the shape."""
    if precision == "auto":
        widths = [int(s[0].shape[1]) for s in segs]
        ok = pack.tc_row_ok(widths) and not (residual is not None and pack.out_width == 128)
        precision = "fp16x3" if ok else "fp32"
    if precision == "fp16x3":
        widths = [int(s[0].shape[1]) for s in segs]
        if not pack.tc_row_ok(widths) or (residual is not None and pack.out_width == 128):
            raise RuntimeError(f"rowmlp: precision fp16x3 does not support hidden={pack.hidden}, segments={widths}, "
                               f"out_width={pack.out_width}; use precision='fp32'")
        return rowmlp_tc(pack.tc_row(widths), segs, rows=rows, act=act, out=out, residual=residual)
    if precision != "fp32":
        raise RuntimeError(f"rowmlp: unknown precision {precision!r}")
    d = L.RowMlpDesc()
    tens = [s[0] for s in segs]
    for t in tens:
        if not (t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1):
            raise RuntimeError("rowmlp: segments must be fp32 CUDA tensors with unit column stride")
    if rows is None:
        rows = int(segs[0][1].numel()) if segs[0][1] is not None else int(tens[0].shape[0])
    d.rows, d.n_segs, d.act_out = rows, len(segs), L.ACTS[act]
    for i, (t, gather, scale) in enumerate(segs):
        d.seg[i] = _seg_struct(t, gather, scale)
    d.mlp = pack.struct()
    if out is None:
        out = torch.empty(rows, pack.out_width, device=tens[0].device, dtype=torch.float32)
    d.out, d.out_stride = out.data_ptr(), int(out.stride(0))
    if residual is not None:
        d.residual, d.res_stride = residual.data_ptr(), int(residual.stride(0))
    L.launch("g4c_rowmlp_fwd", d, out, residual, *tens, *[s[1] for s in segs])
    return out


def mp(edge_pack: MlpPack, node_pack: MlpPack, topo: MpTopo, e_in, src_feat, tgt_feat, aggr="mean",
       act_e=None, act_t=None, want_e=True, precision="auto", e_out=None, t_out=None, ws=None):
    """Message-passing block.  precision "auto" (default): "fp16x3" when hidden = 128, else "fp32".  "fp32": one fused CUDA-core kernel (g4c_mp_fwd); "fp16x3": the tensor-core
    path = per-node products of the split first edge layer (g4c_rowmlp_tc_fwd x2), fused edge MLP + aggregation
    (g4c_edge_aggr_fwd), node model (g4c_rowmlp_tc_fwd).  ``ws`` = optional preallocated (P_r, P_c, agg).
    Returns (t_out, e_out|None)."""
    # node features: one matrix, or (MuGS, first block behind an up-sampling) a tuple of 128-wide matrices standing for their
    # concatenation cat(parts, dim=1) -- the kernels read the parts in place, nothing is concatenated
    def parts(x):
        if isinstance(x, (tuple, list)):
            L.require_cuda_f32(*x)
            return list(x)
        L.require_cuda_f32(x)
        return [x] if x.shape[1] <= 128 else [x[:, c:c + 128] for c in range(0, int(x.shape[1]), 128)]
    L.require_cuda_f32(e_in)
    same = src_feat is tgt_feat
    s_parts = parts(src_feat)
    t_parts = s_parts if same else parts(tgt_feat)
    H = edge_pack.hidden
    fw = sum(int(p.shape[1]) for p in t_parts)
    if sum(int(p.shape[1]) for p in s_parts) != fw:
        raise RuntimeError("mp: source and target features must have the same width")
    n_src, n_tgt = int(s_parts[0].shape[0]), int(t_parts[0].shape[0])
    tc_ok = edge_pack.tc_edge_ok(fw) and node_pack.tc_node_ok(fw) and all(int(p.shape[1]) == 128 for p in s_parts + t_parts)
    if precision == "auto":
        precision = "fp16x3" if tc_ok else "fp32"
    if precision == "fp16x3":
        if not tc_ok:
            raise RuntimeError(f"mp: precision fp16x3 needs hidden=128 and 128- or 256-wide node features (got hidden {H}, "
                               f"features {fw}); use precision='fp32'")
        ep, proj_s, proj_t = edge_pack.tc_edge()
        dev = e_in.device
        P_r, P_c, agg = ws if ws is not None else (None, None, None)
        segs = lambda ps: [(p, None, 1.0) for p in ps]
        if same and fw == 128:            # GNBlock / EdgeMP: one pass over the features makes both products
            if P_r is None:
                P_r = torch.empty(n_src, 128, device=dev, dtype=torch.float32)
            if P_c is None:
                P_c = torch.empty_like(P_r)
            dual_linear_tc(proj_s, proj_t, s_parts[0], out_a=P_r, out_b=P_c)
        else:                             # DownEdgeMP: sources and targets are different levels; MuGS: 256-wide features
            P_r = rowmlp_tc(proj_s, segs(s_parts), out=P_r)
            P_c = rowmlp_tc(proj_t, segs(t_parts), out=P_c)
        if agg is None:
            agg = torch.empty(n_tgt, 128, device=dev, dtype=torch.float32)
        agg, e_out = edge_aggr(ep, topo, e_in, P_r, P_c, aggr=aggr, act_e=act_e, want_e=want_e, e_out=e_out, agg_out=agg,
                                p_prescaled=True)
        if fw == 128:
            t_out = rowmlp_tc(node_pack.tc_row([128, 128]), [(agg, None, 1.0), (t_parts[0], None, 1.0)], act=act_t, out=t_out)
        else:
            # P_r is free again once the edge kernel has run: it holds Q, the target features' share of the node model's linear_1
            q_pack, wide_pack = node_pack.tc_node_wide()
            Q = rowmlp_tc(q_pack, segs(t_parts), out=P_r if n_src == n_tgt else None)
            t_out = rowmlp_tc(wide_pack, [(agg, None, 1.0), (Q, None, 1.0)], act=act_t, out=t_out)
        return t_out, e_out
    if isinstance(src_feat, (tuple, list)) or isinstance(tgt_feat, (tuple, list)):
        raise RuntimeError("mp: features given in parts run on the tensor-core path only")
    if fw != H:
        raise RuntimeError(f"mp: the CUDA-core block (precision fp32) takes node features as wide as hidden={H} (got {fw}); "
                           "256-wide features run on the tensor-core path only (hidden 128, precision 'auto' or 'fp16x3')")
    d = L.MpDesc()
    d.hidden, d.aggr, d.fixed_k = H, (L.AGGR_MEAN if aggr == "mean" else L.AGGR_SUM), topo.fixed_k
    d.act_e_out, d.act_t_out, d.precision = L.ACTS[act_e], L.ACTS[act_t], L.PRECISIONS[precision]
    d.n_targets, d.n_edges = topo.n_targets, topo.n_edges
    d.rowptr = 0 if topo.rowptr is None else topo.rowptr.data_ptr()
    d.src = topo.src.data_ptr()
    d.edge_perm = 0 if topo.edge_perm is None else topo.edge_perm.data_ptr()
    d.tgt_perm = 0 if topo.tgt_perm is None else topo.tgt_perm.data_ptr()
    if want_e and e_out is None:
        e_out = torch.empty(topo.n_edges, H, device=e_in.device, dtype=torch.float32)
    if t_out is None:
        t_out = torch.empty(tgt_feat.shape[0], H, device=e_in.device, dtype=torch.float32)
    d.e_in, d.src_feat, d.tgt_feat = e_in.data_ptr(), src_feat.data_ptr(), tgt_feat.data_ptr()
    d.e_out = e_out.data_ptr() if want_e else 0
    d.t_out = t_out.data_ptr()
    d.edge_mlp, d.node_mlp = edge_pack.struct(), node_pack.struct()
    L.launch("g4c_mp_fwd", d, e_in, src_feat, tgt_feat, e_out, t_out, topo.src, topo.rowptr, topo.edge_perm, topo.tgt_perm)
    return t_out, (e_out if want_e else None)


def seg_reduce(x, ptr, idx, n_groups, aggr="mean", act=None, out=None):
    L.require_cuda_f32(x)
    d = L.SegReduceDesc()
    d.n_groups, d.width = int(n_groups), int(x.shape[1])
    d.aggr, d.act_out = (L.AGGR_MEAN if aggr == "mean" else L.AGGR_SUM), L.ACTS[act]
    d.ptr, d.idx, d.x = ptr.data_ptr(), (0 if idx is None else idx.data_ptr()), x.data_ptr()
    if out is None:
        out = torch.empty(n_groups, x.shape[1], device=x.device, dtype=torch.float32)
    d.out = out.data_ptr()
    L.launch("g4c_seg_reduce_fwd", d, x, ptr, idx, out)
    return out


def project(V, col, U, extras=(), out=None):
    """(V[col].view(E,-1,2)*U.unsqueeze(1)).sum(-1) with per-node scalars appended."""
    L.require_cuda_f32(V, U, *extras)
    d = L.ProjectDesc()
    F = int(V.shape[1]) // 2
    d.n_edges, d.n_feat, d.n_extra = int(col.numel()), F, len(extras)
    d.col, d.V, d.U = col.data_ptr(), V.data_ptr(), U.data_ptr()
    for i, x in enumerate(extras):
        d.extra[i] = x.data_ptr()
    if out is None:
        out = torch.empty(col.numel(), F + len(extras), device=V.device, dtype=torch.float32)
    d.out = out.data_ptr()
    L.launch("g4c_project_fwd", d, V, col, U, out, *extras)
    return out


def edge_to_node(e, Uinv, out=None, residual=None):
    """edgeScalarToNodeVector with the precomputed pseudo-inverse: [N*k,F] -> [N,2F]."""
    L.require_cuda_f32(e, Uinv)
    d = L.EdgeToNodeDesc()
    n, _, k = Uinv.shape
    d.n_nodes, d.k, d.n_feat = int(n), int(k), int(e.shape[1])
    d.Uinv, d.e = Uinv.data_ptr(), e.data_ptr()
    if out is None:
        out = torch.empty(n, 2 * e.shape[1], device=e.device, dtype=torch.float32)
    d.V, d.out_stride = out.data_ptr(), int(out.stride(0))
    if residual is not None:
        d.residual, d.res_stride = residual.data_ptr(), int(residual.stride(0))
    L.launch("g4c_edge_to_node_fwd", d, e, Uinv, out, residual)
    return out


def interp(x, x_idx, w, k, n_out, y, y_row=None):
    L.require_cuda_f32(x, w, y)
    d = L.InterpDesc()
    d.n_out, d.k, d.width = int(n_out), int(k), int(x.shape[1])
    d.x_idx, d.w, d.x, d.y = x_idx.data_ptr(), w.data_ptr(), x.data_ptr(), y.data_ptr()
    d.y_row = 0 if y_row is None else y_row.data_ptr()
    L.launch("g4c_interp_fwd", d, x, x_idx, w, y, y_row)
    return y


def step_update(pred, node_in, field_width, outputs, t):
    d = L.StepUpdateDesc()
    d.n_nodes, d.nf, d.field_width = int(pred.shape[0]), int(pred.shape[1]), int(field_width)
    d.in_stride, d.out_stride, d.t = int(node_in.stride(0)), int(outputs.stride(0)), int(t)
    d.pred, d.node_in, d.outputs = pred.data_ptr(), node_in.data_ptr(), outputs.data_ptr()
    L.launch("g4c_step_update", d, pred, node_in, outputs)


def halo_pack(src, idx, dst):
    d = L.HaloDesc()
    d.n_rows, d.width = int(idx.numel()), int(src.shape[1])
    d.idx, d.src, d.dst = idx.data_ptr(), src.data_ptr(), dst.data_ptr()
    L.launch("g4c_halo_pack", d, src, idx, dst)
    return dst


def halo_unpack(buf, idx, dst):
    d = L.HaloDesc()
    d.n_rows, d.width = int(idx.numel()), int(dst.shape[1])
    d.idx, d.src, d.dst = idx.data_ptr(), buf.data_ptr(), dst.data_ptr()
    L.launch("g4c_halo_unpack", d, buf, idx, dst)
    return dst


def _pack_weight_single(W: torch.Tensor):
    """Single-CTA operand image of W [128, K] for the self tests 1 / 2 of csrc/tc2_test.cu: per 64-wide K-block the
    128-row hi image then the lo image (SWIZZLE_128B K-major).  Returns (uint8 tensor, 1/s)."""
    N, K = W.shape
    assert N == 128 and K % 64 == 0
    W = W.detach().float()
    s = weight_scale(W)
    Ws = W * s
    hi = Ws.half()
    lo = (Ws - hi.float()).half()
    pack = torch.stack([_swizzled_images(hi), _swizzled_images(lo)], dim=1).contiguous()        # [kb, hi|lo, 128, 64]
    return pack.view(torch.uint8).reshape(-1), 1.0 / s


def knn(pos: torch.Tensor, query: Optional[torch.Tensor], k: int, points_per_cell: float = 3.0) -> torch.Tensor:
    """g4c_plan_knn: the k nearest data points of every query, ascending distance, [n_queries, k] int64.  ``query=None``: the
    kNN graph of ``pos`` itself (a point is not its own neighbour).  Same result as the host k-d tree (mesh.knn_edges /
    knn_interp_weights; transforms/connect.py:58, transforms/interpolate.py:125).  Device tensors in, device tensor out."""
    L.require_cuda_f32(pos, query)
    self_graph = query is None
    q = pos if self_graph else query
    n, m = int(pos.shape[0]), int(q.shape[0])
    both = pos if self_graph else torch.cat([pos, q], dim=0)
    lo, hi = both.min(dim=0).values, both.max(dim=0).values
    ext = (hi - lo).clamp(min=1e-20)
    # the descriptor carries the cell size as fp32: bin the points with exactly that value
    cell = float(torch.sqrt(ext[0] * ext[1] * points_per_cell / max(n, 1)).clamp(min=1e-20).float())
    gx, gy = int(float(ext[0]) / cell) + 1, int(float(ext[1]) / cell) + 1
    x0, y0 = float(lo[0]), float(lo[1])
    # cell ids with the kernel's own arithmetic (double), points grouped by cell with ascending id inside a cell
    cx = ((pos[:, 0].double() - x0) / cell).floor().clamp_(0, gx - 1).long()
    cy = ((pos[:, 1].double() - y0) / cell).floor().clamp_(0, gy - 1).long()
    cid = cy * gx + cx
    order = torch.sort(cid, stable=True).indices
    start = torch.zeros(gx * gy + 1, dtype=torch.int64, device=pos.device)
    start[1:] = torch.bincount(cid, minlength=gx * gy).cumsum(0)
    d = L.KnnDesc()
    d.n_points, d.n_queries, d.k, d.exclude_self = n, m, int(k), int(self_graph)
    sorted_idx, cell_start = order.to(torch.int32).contiguous(), start.to(torch.int32).contiguous()
    nbr = torch.empty(m, k, dtype=torch.int32, device=pos.device)
    d.pos, d.query, d.cell_start, d.sorted_idx, d.nbr = pos.data_ptr(), q.data_ptr(), cell_start.data_ptr(), sorted_idx.data_ptr(), nbr.data_ptr()
    d.x0, d.y0, d.cell, d.gx, d.gy = x0, y0, cell, gx, gy
    L.launch("g4c_plan_knn", d, pos, q, nbr)
    return nbr.long()


def debug_tc2(test: int, A: torch.Tensor, W: torch.Tensor, P: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
    """Self tests of the TMEM-operand / tcgen05.cp / CTA-pair primitives (csrc/tc2_test.cu)."""
    L.require_cuda_f32(A, W, P)
    pk, inv = pack_weight_pair(W) if test == 3 else _pack_weight_single(W)
    D = torch.zeros(A.shape[0], 128, device=A.device, dtype=torch.float32)
    L.check(L.lib().g4c_debug_tc2(test, A.data_ptr(), pk.data_ptr(), inv, 0 if P is None else P.data_ptr(),
                                  D.data_ptr(), flags, L.stream_ptr()))
    return D
