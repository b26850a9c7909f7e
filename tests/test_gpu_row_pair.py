"""g4c_rowmlp_tc_fwd (CTA-pair tcgen05 row-tile MLP) against an fp64 restatement of MLP.forward
(graphs4cfd/nn/blocks.py:129-144) on concatenated, optionally gathered / scaled segments — the shapes the
hot path uses: node model (2 wide segments), DownMP (narrow + wide), UpMP (narrow, gathered wide, wide),
encoders (narrow only), decoder (narrow output + residual), and the bare Linear of the split edge model.
Tolerance 2e-5 rel-L2 (22-bit operands, ex2.approx SELU)."""
import pytest
import torch

from graphs4cfd_b200 import ops

gpu = pytest.mark.gpu
F = torch.nn.functional


def _lin(i, o, g, dev):
    return ((torch.rand(o, i, generator=g) * 2 - 1).div(i ** 0.5).to(dev), (torch.rand(o, generator=g) * 2 - 1).div(i ** 0.5).to(dev))


def _check(rows, seg_widths, widths, ln, act, gather_seg=None, scales=None, residual=False, seed=0):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(seed)
    dims = [sum(seg_widths)] + list(widths)
    lin = [_lin(dims[i], dims[i + 1], g, dev) for i in range(len(widths))]
    lnp = ((1 + 0.1 * torch.randn(128, generator=g)).to(dev), (0.1 * torch.randn(128, generator=g)).to(dev)) if ln else None
    scales = scales or [1.0] * len(seg_widths)
    segs, cat = [], []
    for i, w in enumerate(seg_widths):
        if gather_seg == i:
            src = torch.randn(rows // 3 + 1, w, generator=g).to(dev)
            idx = torch.randint(0, src.shape[0], (rows,), generator=g).to(dev).to(torch.int32)
            segs.append((src, idx, scales[i]))
            cat.append(src[idx.long()].double() * scales[i])
        else:
            t = torch.randn(rows, w, generator=g).to(dev)
            segs.append((t, None, scales[i]))
            cat.append(t.double() * scales[i])
    x = torch.cat(cat, dim=1)
    for i, (W, b) in enumerate(lin):
        x = x @ W.double().t() + b.double()
        if i < len(lin) - 1:
            x = F.selu(x)
    if ln:
        x = F.layer_norm(x, (128,), lnp[0].double(), lnp[1].double(), 1e-5)
    x = {"selu": F.selu, "tanh": torch.tanh, None: (lambda z: z)}[act](x)
    res = None
    if residual:
        res = torch.randn(rows, widths[-1] + 2, generator=g).to(dev)
        x = x + res[:, :widths[-1]].double()
    pack = ops.RowPairPack(lin, seg_widths, lnp)
    out = ops.rowmlp_tc(pack, segs, rows=rows, act=act, residual=res)
    torch.cuda.synchronize()
    err = float((out.double() - x).norm() / x.norm())
    assert err < 2e-5, f"rel-L2 {err:.3e}"


@gpu
@pytest.mark.parametrize("rows", [1, 255, 256, 1000, 70000])
def test_node_model(rows):
    _check(rows, [128, 128], [128, 128, 128], True, "selu")


@gpu
def test_node_model_two_layers():
    _check(1000, [128, 128], [128, 128], True, "selu")


@gpu
def test_two_layer_and_no_ln():
    _check(3001, [128], [128, 128], False, None)


@gpu
def test_bare_linear():
    _check(5000, [128], [128], False, None)


@gpu
def test_down_mlp():
    _check(4097, [2, 128], [128, 128, 128], True, None)


@gpu
def test_up_mlp_gathered_scaled():
    _check(9000, [2, 128, 128], [128, 128, 128], True, "tanh", gather_seg=1, scales=[-1.0, 1.0, 1.0])


@gpu
@pytest.mark.parametrize("kin", [2, 3, 5, 16])
def test_encoder_narrow_input(kin):
    _check(2500, [kin], [128, 128, 128], True, "selu")


@gpu
@pytest.mark.parametrize("nout", [1, 3])
def test_decoder_narrow_output_residual(nout):
    _check(7777, [128], [128, 128, nout], False, None, residual=True)


@gpu
@pytest.mark.parametrize("rows", [1, 300, 70001])
def test_dual_linear(rows):
    """Two bare Linears of the same input in one pass (P_r, P_c of the split edge model) against fp64."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(rows)
    la, lb = _lin(128, 128, g, dev), _lin(128, 128, g, dev)
    x = torch.randn(rows, 128, generator=g).to(dev)
    pa, pb = ops.RowPairPack([la], [128]), ops.RowPairPack([lb], [128])
    ya, yb = ops.dual_linear_tc(pa, pb, x)
    torch.cuda.synchronize()
    for y, (W, b) in ((ya, la), (yb, lb)):
        ref = x.double() @ W.double().t() + b.double()
        err = float((y.double() - ref).norm() / ref.norm())
        assert err < 2e-5, f"rel-L2 {err:.3e}"
    # and bit-identical to the single-output kernel
    assert torch.equal(ya, ops.rowmlp_tc(pa, [(x, None, 1.0)])) and torch.equal(yb, ops.rowmlp_tc(pb, [(x, None, 1.0)]))


@gpu
@pytest.mark.parametrize("magnitude", [2.0e4, 1.0, 1.0e-3, 1.0e-6])
def test_encoder_raw_input_range(magnitude):
    """The fp16 (hi, lo) operand split takes raw encoder inputs unscaled.  Up to the documented limit (3e4, ops.FP16_SPLIT_MAX)
    the product keeps its accuracy relative to the OUTPUT scale; tiny inputs (below 6e-5 the lo term is an fp16 subnormal)
    keep an absolute accuracy of ~3e-8 per operand, which a dot product against O(1) weights turns into < 1e-6."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(11)
    rows, kin = 3000, 5
    W, b = _lin(kin, 128, g, dev)
    x = ((torch.rand(rows, kin, generator=g) * 2 - 1) * magnitude).to(dev)          # |x| <= magnitude
    out = ops.rowmlp_tc(ops.RowPairPack([(W, b)], [kin]), [(x, None, 1.0)])
    ref = x.double() @ W.double().t() + b.double()
    err = float((out.double() - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= 2e-6 * scale + 1e-7, f"max abs error {err:.3e} at output scale {scale:.3e}"


@gpu
def test_raw_input_range_check():
    dev = torch.device("cuda")
    ops.check_fp16_range(torch.full((4, 4), 2.9e4, device=dev), "x")
    for bad in (torch.full((4, 4), 1.0e5, device=dev), torch.tensor([[float("nan")]], device=dev)):
        with pytest.raises(RuntimeError, match="fp16x3"):
            ops.check_fp16_range(bad, "x")


@gpu
@pytest.mark.parametrize("segs", [[128], [128, 128], [128, 128, 128], [2, 128]])
def test_round_toward_zero_shrink_is_compensated(segs):
    """tcgen05.mma accumulates with round-toward-zero: uncompensated, a K = 128 GEMM comes out 4.45e-7 too small (a pure scale
    error, measured in tools/tc_bias.py).  The packs fold the expected shrink into the accumulator scale: the remaining scale
    error <err, ref> / <ref, ref> must stay below 1.5e-7 (fp32 granularity of the factor: 6e-8) and far below the uncompensated
    value for the wide cases."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    rows, K = 20000, sum(segs)
    xs = [torch.randn(rows, w, generator=g).to(dev) for w in segs]
    W = (torch.randn(128, K, generator=g) / K ** 0.5).to(dev)
    b = (torch.randn(128, generator=g) * 0.1).to(dev)
    ref = torch.cat(xs, 1).double() @ W.double().t() + b.double()

    def bias(out):
        err = out.double() - ref
        return float((err * ref).sum() / (ref * ref).sum())

    comp = bias(ops.rowmlp_tc(ops.RowPairPack([(W, b)], segs), [(t, None, 1.0) for t in xs]))
    assert abs(comp) < 1.5e-7, comp
    saved = ops.TC_RZ_SHRINK
    try:
        ops.TC_RZ_SHRINK = (0.0, 0.0)
        raw = bias(ops.rowmlp_tc(ops.RowPairPack([(W, b)], segs), [(t, None, 1.0) for t in xs]))
    finally:
        ops.TC_RZ_SHRINK = saved
    assert raw < -3e-7 and abs(comp) < 0.35 * abs(raw), (raw, comp)
