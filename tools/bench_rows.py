"""Times the row-tile kernels of one level-1 MP (scratch): P_r / P_c products and the node model, 1M rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphs4cfd_b200 import ops

dev = torch.device("cuda"); torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
lin = lambda i, o: ((torch.rand(o, i, device=dev) * 2 - 1) / i ** 0.5, (torch.rand(o, device=dev) * 2 - 1) / i ** 0.5)
v, agg = torch.randn(n, 128, device=dev), torch.randn(n, 128, device=dev)
P = torch.empty(n, 128, device=dev); out = torch.empty(n, 128, device=dev)
proj = ops.RowPairPack([lin(128, 128)], [128])
node = ops.RowPairPack([lin(256, 128), lin(128, 128), lin(128, 128)], [128, 128], (torch.ones(128, device=dev), torch.zeros(128, device=dev)))

def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ms = t(lambda: ops.rowmlp_tc(proj, [(v, None, 1.0)], out=P))
print(f"bare Linear (P_r / P_c): {ms:.3f} ms, {2 * n * 512 / ms / 1e6:.0f} GB/s")
ms = t(lambda: ops.rowmlp_tc(node, [(agg, None, 1.0), (v, None, 1.0)], act="selu", out=out))
print(f"node model: {ms:.3f} ms, {3 * n * 512 / ms / 1e6:.0f} GB/s")
proj2 = ops.RowPairPack([lin(128, 128)], [128])
P2 = torch.empty(n, 128, device=dev)
ms = t(lambda: ops.dual_linear_tc(proj, proj2, v, out_a=P, out_b=P2))
print(f"dual Linear (P_r and P_c in one pass): {ms:.3f} ms, {3 * n * 512 / ms / 1e6:.0f} GB/s")
