"""CPU emulation of the tensor-core arithmetic (fp16x3) over a long rollout, next to the fp32 noise floor.

Why: the north star asks for rollout outputs within 1e-4 rel-L2 of the reference's fp32 forward.  The rollout is a
chaotic-ish iterated map: even exact fp32 with the in-edges of every node summed in another order drifts away from the
reference.  This script measures, on the build container's CPU, with the oracle (oracle/restate.py):
  (a) fp32, in-edges re-ordered within each target   -> the noise floor of any correct fp32 implementation
  (b) every Linear computed as the kernels do:  x = hi + lo (fp16), s*W = hi + lo (fp16), y = (xh Wh + xl Wh + xh Wl) / s
      with fp32 accumulation      -> what the tcgen05 path adds
  (c) operands rounded to TF32 / BF16                 -> what a single-pass tensor-core path would add
against the unmodified oracle, step by step.  Weights: the reference's shipped 3S-GNN checkpoint when /root/reference
is present (trained weights amplify errors far more than a default init), else a seeded default init.

    python tools/precision_emulation.py [--nodes 6000] [--steps 30]
"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graphs4cfd_b200 import mesh as M                      # noqa: E402
from graphs4cfd_b200.archs import init_params, mus_arch    # noqa: E402
from oracle import restate as R                            # noqa: E402


def split16(x):
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi, lo


def linear_fp16x3(x, W, b):
    amax = float(W.abs().max())
    s = 1.0 if amax == 0.0 else 2.0 ** min(14, math.floor(math.log2(1000.0 / amax)))      # ops.weight_scale
    xh, xl = split16(x)
    wh, wl = split16(W * s)
    y = (xh @ wh.t() + xl @ wh.t() + xh @ wl.t()) / s
    return y + b


def round_mantissa(x, bits):
    """keep `bits` explicit mantissa bits of fp32 (round to nearest)"""
    i = x.contiguous().view(torch.int32)
    drop = 23 - bits
    i = (i + (1 << (drop - 1))) & ~((1 << drop) - 1)
    return i.view(torch.float32)


def make_linear(kind):
    if kind == "fp16x3":
        return linear_fp16x3
    bits = {"tf32": 10, "bf16": 7}[kind]
    return lambda x, W, b: round_mantissa(x, bits) @ round_mantissa(W, bits).t() + b


class patched_linear:
    def __init__(self, fn):
        self.fn = fn

    def __enter__(self):
        self.orig = F.linear
        F.linear = self.fn

    def __exit__(self, *exc):
        F.linear = self.orig


def permute_in_edges(g, seed):
    """the same graph with the in-edges of every target stored in another order (kNN layout: k consecutive per target)"""
    g2 = g.clone()
    E = g.edge_index.size(1)
    k = int((g.edge_index[1] == g.edge_index[1][0]).sum())
    gen = torch.Generator().manual_seed(seed)
    perm = (torch.arange(E).view(-1, k) .gather(1, torch.rand(E // k, k, generator=gen).argsort(dim=1))).reshape(-1)
    g2.edge_index = g.edge_index[:, perm]
    g2.edge_attr = g.edge_attr[perm]
    return g2


def rollout(params, g, steps):
    outs = []
    field0 = g.field
    with torch.no_grad():
        for _ in range(steps):
            pred = R.forward(params, g)
            outs.append(pred)
            g.field = torch.cat([g.field[:, pred.size(1):], pred], dim=1)
    g.field = field0
    return outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=6000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--model", default="mus", choices=["mus", "remus"])
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    weights = "seeded default init"
    have_ref = os.path.isdir("/root/reference/graphs4cfd")
    if have_ref:
        from oracle.pyg_stub import import_reference
        gfd = import_reference()
    if a.model == "remus":
        from graphs4cfd_b200.archs import remus_arch
        g = M.build_remus_mesh(a.nodes, 5, seed=0, points="uniform")       # k = 5: the shipped checkpoint's training setup
        params = init_params(remus_arch(128), seed=0)
        if have_ref:
            model = gfd.nn.NsRotEquiTreeScaleGNN(model="RE3S-GNN-NsEllipse-v1")
            params = {k: v.detach() for k, v in model.state_dict().items()}
            weights = "shipped RE3S-GNN-NsEllipse-v1 checkpoint"
        name = "3-scale REMuS-GNN"
    else:
        g = M.build_mus_mesh(a.nodes, 6, M.auto_cells(a.nodes, 3), seed=0)
        params = init_params(mus_arch(128, 3), seed=0)
        if have_ref:
            model = gfd.nn.NsThreeScaleGNN(model="3S-GNN-NsCircle-v1")
            params = {k: v.detach() for k, v in model.state_dict().items()}
            weights = "shipped 3S-GNN-NsCircle-v1 checkpoint"
        name = "3-scale MuS-GNN"
    print(f"{name}, hidden 128, {a.nodes}-node synthetic mesh, {weights}; rel-L2 of the prediction vs the fp32 oracle")
    ref = rollout(params, g.clone(), a.steps)
    runs = {}
    if a.model == "mus":       # (the REMuS angle layout is positional, transforms/remus.py:36-38: no free re-ordering)
        runs["fp32, in-edges re-ordered (noise floor)"] = rollout(params, permute_in_edges(g, 1), a.steps)
    for kind in ("fp16x3", "tf32", "bf16"):
        with patched_linear(make_linear(kind)):
            runs[f"{kind} operands, fp32 accumulate"] = rollout(params, g.clone(), a.steps)
    marks = [s for s in (1, 2, 5, 10, 20, 30, 50, 100) if s <= a.steps]
    print("| arithmetic | " + " | ".join(f"step {s}" for s in marks) + " |")
    print("|---|" + "---|" * len(marks))
    for name, outs in runs.items():
        cells = []
        for s in marks:
            cells.append(f"{float((outs[s - 1].double() - ref[s - 1].double()).norm() / ref[s - 1].double().norm()):.1e}")
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
