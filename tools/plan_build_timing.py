"""Time of building the 1M-node benchmark meshes on the host (scipy k-d tree + torch CPU ops) and on the device
(g4c_plan_knn + torch device ops), and of the engine's plan on top of each.     python tools/plan_build_timing.py [--nodes N]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphs4cfd_b200 import Rollout, mesh as M  # noqa: E402
from graphs4cfd_b200.archs import init_params, mus_arch, remus_arch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=1_000_000)
a = ap.parse_args()
n = a.nodes


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


torch.zeros(1, device="cuda")
M.build_mus_mesh(2000, 6, M.auto_cells(2000, 3), device="cuda")            # warm-up (library load, kernels)
for name, build, arch in (("MuS-3", lambda dev: M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=0, device=dev), mus_arch(128, 3)),
                          ("REMuS", lambda dev: M.build_remus_mesh(n, 6, seed=0, device=dev), remus_arch(128))):
    gh, th = timed(lambda: build(None))
    gd, td = timed(lambda: build("cuda"))
    diff = {k: int((getattr(gd, k).cpu() != v).sum()) for k, v in gh.__dict__.items()
            if torch.is_tensor(v) and not v.is_floating_point() and getattr(gd, k).shape == v.shape and not torch.equal(getattr(gd, k).cpu(), v)}
    diff.update({k: "shape" for k, v in gh.__dict__.items() if torch.is_tensor(v) and getattr(gd, k).shape != v.shape})
    same = "all equal" if not diff else f"entries that differ: {diff}"
    if "edge_index" in diff:
        # equal-distance ties are the only legitimate difference: the device search sends them to the lower index
        bad = (gd.edge_index.cpu() != gh.edge_index).any(dim=0).nonzero().squeeze(1)
        pd_ = gh.pos.double()
        d_h = (pd_[gh.edge_index[0, bad]] - pd_[gh.edge_index[1, bad]]).pow(2).sum(1)
        d_d = (pd_[gd.edge_index.cpu()[0, bad]] - pd_[gh.edge_index[1, bad]]).pow(2).sum(1)
        same += f"; of the differing edges, {int((d_h == d_d).sum())} of {bad.numel()} have exactly equal distances (ties)"
    params = init_params(arch, seed=0)
    _, tp = timed(lambda: Rollout(params, gd, device="cuda", cuda_graph=False))
    print(f"{name}, {n} nodes: mesh on the host {th:.2f} s, on the device {td:.2f} s (integer layouts: {same}); engine plan on top {tp:.2f} s")
    del gh, gd
    torch.cuda.empty_cache()
