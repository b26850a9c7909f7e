"""Per-operation device time of one single-GPU rollout step (eager, CUDA events between the plan's operations), the N = 1
companion of tools/partition_timeline.py.     python tools/step_timeline.py [--model mus|remus|mugs2|mugs4] [--nodes 1000000] [--reps 5]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="mus", choices=["mus", "remus", "mugs2", "mugs3", "mugs4"])
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--hidden", type=int, default=128)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()

from graphs4cfd_b200 import Rollout, ops  # noqa: E402
from graphs4cfd_b200 import mesh as M  # noqa: E402
from graphs4cfd_b200.archs import init_params, mugs_arch, mus_arch, remus_arch  # noqa: E402

dev = torch.device("cuda")
if a.model.startswith("mugs"):
    lv = int(a.model[4])
    g = M.build_mugs_mesh(a.nodes, 6, levels=lv, seed=0, edge_scale=(0.1, 0.25, 0.5, 1.0)[:lv], device=dev)
    params = init_params(mugs_arch(a.hidden, lv), seed=0)
elif a.model == "remus":
    g, params = M.build_remus_mesh(a.nodes, 6, seed=0), init_params(remus_arch(a.hidden), seed=0)
else:
    g, params = M.build_mus_mesh(a.nodes, 6, M.auto_cells(a.nodes, 3), seed=0), init_params(mus_arch(a.hidden, 3), seed=0)
eng = Rollout(params, g, device=dev, cuda_graph=False)


def label(op, s):
    if op == "mp":
        t = s["topo"]
        return f"mp targets={t.n_targets} edges={t.n_edges} {'fixed-k' if t.fixed_k else 'csr'} e_out={'yes' if s['e_out'] is not None else 'no'}"
    if op == "rowmlp":
        return f"rowmlp rows={s.get('rows') or s['segs'][0][0].shape[0]} segs={[int(x[0].shape[1]) for x in s['segs']]} out={s['pack'].out_width}"
    if op == "seg":
        return f"seg_reduce groups={s['n']} rows={int(s['idx'].numel())}"
    return s.get("label", op)


def run_one(op, s):
    saved = eng.steps
    eng.steps = [(op, s)]
    eng._run_step_eager()
    eng.steps = saved


for _ in range(2):
    eng._run_step_eager()
torch.cuda.synchronize()
n = len(eng.steps)
acc = [0.0] * n
for _ in range(a.reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i, (op, s) in enumerate(eng.steps):
        run_one(op, s)
        ev[i + 1].record()
    torch.cuda.synchronize()
    for i in range(n):
        acc[i] += ev[i].elapsed_time(ev[i + 1]) / a.reps
print(f"# single-GPU {a.model} step, {a.nodes} nodes, hidden {a.hidden}, eager, {a.reps} steps averaged; total {sum(acc):.3f} ms")
by = collections.defaultdict(float)
for (op, s), t in zip(eng.steps, acc):
    lab = label(op, s)
    by[lab] += t
    print(f"{t:8.3f} ms  {lab}")
print("# by operation shape:")
for lab, t in sorted(by.items(), key=lambda kv: -kv[1]):
    print(f"#   {t:8.3f} ms  {100 * t / sum(acc):5.1f} %  {lab}")
