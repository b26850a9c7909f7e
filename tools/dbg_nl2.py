"""Scratch: isolate the 2-layer irregular failure (edge kernel vs row kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_edge_pair as TE
import test_gpu_row_pair as TR

dev = torch.device("cuda")
def irregular(n_layers, aggr, seed=3, n=700, perm_edges=True, maxd=10):
    g = torch.Generator().manual_seed(seed)
    deg = torch.randint(0, maxd, (n,), generator=g)
    col = torch.arange(n).repeat_interleave(deg)
    if perm_edges:
        col = col[torch.randperm(col.numel(), generator=g)]
    col = col.to(dev)
    row = torch.randint(0, n, (col.numel(),), generator=g).to(dev)
    try:
        TE._run(n, row, col, n_layers, aggr, None)
        print(f"edge irregular nl={n_layers} {aggr} n={n} perm={perm_edges} maxd={maxd}: ok")
    except AssertionError as ex:
        print(f"edge irregular nl={n_layers} {aggr} n={n} perm={perm_edges} maxd={maxd}: FAIL {ex}")

for nl in (3, 2):
    for aggr in ("mean", "sum"):
        irregular(nl, aggr)
    irregular(nl, "mean", perm_edges=False)
    irregular(nl, "mean", n=100, maxd=3)
    irregular(nl, "mean", n=5000, maxd=8)
for args in [(1000, [128, 128], [128, 128], True, "selu"), (1000, [128, 128], [128, 128], True, None),
             (777, [128, 128], [128, 128], False, None), (777, [128, 128], [128, 128, 128], True, None),
             (777, [128], [128, 128], True, None)]:
    try:
        TR._check(*args)
        print("row", args, "ok")
    except AssertionError as ex:
        print("row", args, "FAIL", ex)
