"""MuGS-GNN at scale through the drop-in boundary: the reference's own NsTwoGuillardScaleGNN /
NsFourGuillardScaleGNN class with its shipped weights on a synthetic `--nodes`-node mesh (graphs4cfd_b200.mesh.build_mugs_mesh, the
layouts of the reference's transforms), driven by the reference's own GNN.solve on the GPU:
  reference   stock PyTorch eager (fp32, TF32 off) — the model untouched
  dropin      the same model after graphs4cfd_b200.accelerate(): tensor-core blocks (incl. the 256-wide ones), g4c_interp_fwd
  engine      graphs4cfd_b200.Rollout on the same model and mesh (rollout_mugs.py), one CUDA graph per step
Prints one JSON line with both rates, the rel-L2 between the two after `--steps` rollout steps and the libg4c launches per step.

    python tools/mugs_dropin_bench.py [--model mugs2|mugs4] [--nodes 1000000] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

MODELS = {"mugs2": ("NsTwoGuillardScaleGNN", "NsTwoGuillardScaleGNN.chk", "2GS-GNN-NsCircle-v1", 2),
          "mugs4": ("NsFourGuillardScaleGNN", "NsFourGuillardScaleGNN.chk", "4GS-GNN-NsCircle-v1", 4)}


def timed(model, g, steps):
    with torch.no_grad():
        model.solve(g.clone(), 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model.solve(g.clone(), steps)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mugs2", choices=list(MODELS))
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from oracle.pyg_stub import import_reference, staged_checkpoint      # the reference's model class IS the caller here
    gfd = import_reference()
    cls, chk, name, levels = MODELS[a.model]
    dev = torch.device("cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    t0 = time.perf_counter()
    g = M.build_mugs_mesh(a.nodes, 6, levels=levels, seed=0, edge_scale=(0.1, 0.25, 0.5, 1.0)[:levels], device=dev)
    torch.cuda.synchronize()
    t_mesh = time.perf_counter() - t0
    path = staged_checkpoint(chk)
    model = getattr(gfd.nn, cls)(checkpoint=path, device=dev) if path else getattr(gfd.nn, cls)(model=name, device=dev)
    out = {"model": name, "nodes": a.nodes, "edges": int(g.edge_index.shape[1]), "steps": a.steps, "mesh_build_s": round(t_mesh, 2),
           "level_nodes": [a.nodes] + [int(getattr(g, f"coarse_mask{l}").sum()) for l in range(2, levels + 1)]}
    try:
        dt_ref, want = timed(model, g, a.steps)
        out["reference_eager"] = {"steps_per_s": 1.0 / dt_ref, "ms_per_step": dt_ref * 1e3}
    except Exception as exc:                                            # e.g. out of memory
        want = None
        out["reference_eager"] = {"unavailable": f"{type(exc).__name__}: {str(exc)[:160]}"}
    torch.cuda.empty_cache()
    g4.accelerate(model)
    n0 = g4.ops.L.launch_count()
    dt, got = timed(model, g, a.steps)
    out["dropin"] = {"steps_per_s": 1.0 / dt, "ms_per_step": dt * 1e3,
                     "launches_per_step": (g4.ops.L.launch_count() - n0) / (a.steps + 1), "finite": bool(torch.isfinite(got).all())}
    if want is not None:
        out["rel_l2_vs_reference_eager"] = float((got.double() - want.double()).norm() / want.double().norm())
        out["speedup"] = dt_ref / dt
    # the plan-based engine on the same model and mesh (rollout_mugs.py), one CUDA graph per step, device-timed
    del got
    torch.cuda.empty_cache()
    eng = g4.Rollout(model, g, device=dev)
    res = eng.solve(a.steps)
    for _ in range(3):
        eng.step_only()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(4 * a.steps):
        eng.step_only()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / (4 * a.steps)
    out["engine"] = {"steps_per_s": 1e3 / ms, "ms_per_step": ms, "launches_per_step": eng.launches_per_step,
                     "buffer_gb": eng.buffer_bytes / 1e9}
    if want is not None:
        out["engine"]["rel_l2_vs_reference_eager"] = float((res.double() - want.double()).norm() / want.double().norm())
        out["engine"]["speedup"] = dt_ref * 1e3 / ms
    print(json.dumps(out))


if __name__ == "__main__":
    main()
