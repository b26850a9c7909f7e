"""Scratch: per-launch-group durations of one MuS-3 rollout step at 1M nodes, in sequence (CUDA events, eager launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphs4cfd_b200 import Rollout, ops
from graphs4cfd_b200 import mesh as M
from graphs4cfd_b200.archs import init_params, mus_arch

n = 1_000_000
g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=0)
eng = Rollout(init_params(mus_arch(128, 3), seed=0), g, cuda_graph=False)
for _ in range(3):
    eng._run_step_eager()
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(eng.steps) + 1)]
tot = {}
for rep in range(5):
    evs[0].record()
    for i, (op, a) in enumerate(eng.steps):
        eng.steps, saved = [eng.steps[i]], eng.steps
        eng._run_step_eager()
        eng.steps = saved
        evs[i + 1].record()
    torch.cuda.synchronize()
    for i, (op, a) in enumerate(eng.steps):
        rows = a["topo"].n_targets if op == "mp" else (a["out"].shape[0] if "out" in a else 0)
        key = (i, op, rows)
        tot[key] = tot.get(key, 0.0) + evs[i].elapsed_time(evs[i + 1]) / 5
print(f"sum = {sum(tot.values()):.2f} ms")
for (i, op, rows), ms in tot.items():
    print(f"{i:3d} {op:7s} rows={rows:8d} {ms:7.3f} ms")
