"""Per-step drift of long rollouts with the reference's trained 3S-GNN weights (staged by tools/stage_reference.py):
rel-L2 of the prediction at steps 1 / 10 / 30 / 50 / 100 against the truth of each size, for every arithmetic the engine
offers, next to the fp32 reorder-noise floor (the same fp32 arithmetic with every node's in-edges stored in reverse order).

    python tools/rollout_drift.py [--small 6000] [--large 200000] [--steps 100]
  small mesh: truth = the reference's own classes on the CPU;  large mesh: truth = the fp32 CUDA-core engine."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import graphs4cfd_b200 as g4  # noqa: E402
from graphs4cfd_b200 import mesh as M, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", type=int, default=6000)
ap.add_argument("--large", type=int, default=200_000)
ap.add_argument("--steps", type=int, default=100)
a = ap.parse_args()
CHECK = [t for t in (1, 2, 5, 10, 30, 50, 100, 200, 500) if t <= a.steps]

from oracle.pyg_stub import import_reference  # noqa: E402  (a measurement tool, like the tests: the oracle is the checker)
from conftest import shipped_model  # noqa: E402
from test_gpu_long_rollout import per_step_rel, reversed_in_edges  # noqa: E402

gfd = import_reference()
ref = shipped_model(gfd, "mus3")


def row(name, rel):
    print(f"  {name:58s} " + "  ".join(f"{rel[t - 1]:.2e}" for t in CHECK))


def engines(g, truth, floor):
    print("  " + " " * 58 + " " + "  ".join(f"step {t:<4d}" for t in CHECK))
    row("fp32 arithmetic, in-edges reversed (noise floor)", floor)
    row("fp32 CUDA-core kernels", per_step_rel(g4.Rollout(ref, g.clone(), precision="fp32", device="cuda").solve(a.steps), truth, 3))
    for variant, name in (("auto", "v5"), ("v3", "v3")):
        ops.EDGE_VARIANT_DEFAULT = variant
        row(f"fp16x3 tensor-core kernels, level-1 edge kernel {name}",
            per_step_rel(g4.Rollout(ref, g.clone(), device="cuda").solve(a.steps), truth, 3))
    ops.EDGE_VARIANT_DEFAULT = "auto"


with torch.no_grad():
    n = a.small
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=7)
    truth = ref.solve(g.clone(), a.steps)
    floor = per_step_rel(ref.solve(reversed_in_edges(g, 6), a.steps), truth, 3)
    print(f"# 3S-GNN-NsCircle-v1 weights, {n}-node mesh, truth = the reference's own classes on the CPU; rel-L2 of the prediction")
    engines(g, truth, floor)
    n = a.large
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=0)
    truth = g4.Rollout(ref, g.clone(), precision="fp32", device="cuda").solve(a.steps)
    floor = per_step_rel(g4.Rollout(ref, reversed_in_edges(g, 6), precision="fp32", device="cuda").solve(a.steps), truth, 3)
    print(f"# {n}-node mesh (configs[1]), truth = the fp32 CUDA-core engine")
    engines(g, truth, floor)
