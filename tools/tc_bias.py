"""Is the error of the fp16x3 tensor-core path random or systematic?  For a bare Linear, a 3-layer MLP without / with
LayerNorm and a whole message-passing block (trained weights), against fp64: rel-L2 error and the BIAS coefficient
b = <err, ref> / <ref, ref> (a pure scale error (1 + b) gives exactly b; unbiased rounding noise gives |b| << rel-L2),
plus the shrink statistic mean(|out| - |ref|) / mean|ref| (negative = results pulled toward zero, as round-toward-zero
accumulation would do)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from graphs4cfd_b200 import ops  # noqa: E402

if "--no-compensation" in sys.argv:
    ops.TC_RZ_SHRINK = (0.0, 0.0)
print(f"# TC_RZ_SHRINK = {ops.TC_RZ_SHRINK} (relative shrink of an accumulator = a + b * number of MMAs)")

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)


def stats(name, out, ref):
    out, ref = out.double(), ref.double()
    err = out - ref
    rel = float(err.norm() / ref.norm())
    b = float((err * ref).sum() / (ref * ref).sum())
    shrink = float((out.abs() - ref.abs()).mean() / ref.abs().mean())
    print(f"{name:58s} rel-L2 {rel:.2e}   bias coeff {b:+.2e}   shrink {shrink:+.2e}")


rows = 20000
for scale_in, label in ((1.0, "N(0,1) inputs"), (3.0, "3 N(0,1) + 2 (offset) inputs")):
    x = (torch.randn(rows, 128, generator=g) * scale_in + (2.0 if scale_in != 1.0 else 0.0)).to(dev)
    W = (torch.randn(128, 128, generator=g) * 0.09).to(dev)
    b = (torch.randn(128, generator=g) * 0.1).to(dev)
    out = ops.rowmlp_tc(ops.RowPairPack([(W, b)], [128]), [(x, None, 1.0)])
    stats(f"bare Linear K=128, {label}", out, x.double() @ W.double().t() + b.double())
    out32 = (x @ W.t() + b)
    stats(f"   (torch fp32 matmul of the same, for scale)", out32, x.double() @ W.double().t() + b.double())

lin = [((torch.randn(128, 128, generator=g) * 0.09).to(dev), (torch.randn(128, generator=g) * 0.1).to(dev)) for _ in range(3)]
x = torch.randn(rows, 128, generator=g).to(dev)
for ln in (None, (torch.ones(128, device=dev), torch.zeros(128, device=dev))):
    ref = x.double()
    for i, (W, b) in enumerate(lin):
        ref = ref @ W.double().t() + b.double()
        if i < 2:
            ref = F.selu(ref)
    if ln is not None:
        ref = F.layer_norm(ref, (128,), ln[0].double(), ln[1].double(), 1e-5)
    out = ops.rowmlp_tc(ops.RowPairPack(lin, [128], ln), [(x, None, 1.0)])
    stats(f"3-layer MLP, LayerNorm={'yes' if ln else 'no'}", out, ref)
    out32 = ops.rowmlp(ops.MlpPack(lin, ln), [(x, None, 1.0)], precision="fp32")
    stats(f"   (fp32 CUDA-core kernel of the same)", out32, ref)

# linearity in the number of MMAs: two and three 128-wide segments (K = 256, 384), a narrow segment (one K = 16 step)
for segs in ([128, 128], [128, 128, 128], [5], [2, 128]):
    K = sum(segs)
    xs = [torch.randn(rows, w, generator=g).to(dev) for w in segs]
    W = (torch.randn(128, K, generator=g) / K ** 0.5).to(dev)
    b = (torch.randn(128, generator=g) * 0.1).to(dev)
    out = ops.rowmlp_tc(ops.RowPairPack([(W, b)], segs), [(t, None, 1.0) for t in xs])
    stats(f"bare Linear, segments {segs}", out, torch.cat(xs, 1).double() @ W.double().t() + b.double())

# decoder-like: narrow output added to a residual
W3 = (torch.randn(3, 128, generator=g) * 0.09).to(dev)
b3 = (torch.randn(3, generator=g) * 0.1).to(dev)
lin_d = [lin[0], lin[1], (W3, b3)]
ref = x.double()
for i, (W, b) in enumerate(lin_d):
    ref = ref @ W.double().t() + b.double()
    if i < 2:
        ref = F.selu(ref)
out = ops.rowmlp_tc(ops.RowPairPack(lin_d, [128], None), [(x, None, 1.0)])
stats("decoder (128 -> 128 -> 128 -> 3), no residual", out, ref)
