#!/usr/bin/env python
"""bench.py — rollout steps/s of the B200 message-passing hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm   (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's OWN classes on the host cores

One "step" = one rollout time step (`GNN.forward` + `shift_and_replace`, nn/model.py:316-320) of the
3-scale MuS-GNN on a synthetic mesh.  Workload at any N: the 1M-node / 6M-edge mesh, hidden=128, that
BASELINE.json's metric is quoted on (it fits one GPU), node-partitioned over N GPUs (strong scaling).
Weights: the reference's shipped 3S-GNN / RE3S-GNN checkpoints when their staged copies exist (baseline/_ref/weights,
tools/stage_reference.py) and the architecture is the shipped one (hidden 128, 3 scales), seeded default init otherwise.
Prints ONE JSON line (rank 0).

Reference arm: the UNMODIFIED reference classes (`NsThreeScaleGNN.solve`, imported under oracle/pyg_stub.py from
/root/reference or the staged copy baseline/_ref) with all host threads.  The 1M-node step takes ~25 s on 16 cores, so every
timed "step" is a bounded sample — one rollout step of the same model on a 50k-node mesh (1/20 of the workload's nodes) —
and `value` is the full-workload figure from a straight-line fit of seconds per step over {10k, 50k, 200k} nodes
(the reference's cost is linear in nodes and edges).  `ms_per_step` is what was really timed.  Its line also carries
`gpu_eager`: the same reference classes in stock PyTorch eager on this GPU at the full size (the honest GPU baseline).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="g4c", choices=["g4c", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--k", type=int, default=6)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--model", default="mus", choices=["mus", "remus"],
                    help="mus: the MuS-GNN workload of BASELINE.json's metric (default); remus: the 3-scale REMuS-GNN of "
                         "configs[2] / configs[4] (a measurement case, not the driver's bench line; N > 1 uses the edge-halo partition)")
    ap.add_argument("--edge-variant", default="auto", choices=["auto", "v3", "v5"],
                    help="kernel behind g4c_edge_aggr_fwd: auto (default: v5 on fixed in-degree launches, v3 otherwise) or pinned")
    ap.add_argument("--precision", default=os.environ.get("G4C_PRECISION", "auto"),
                    help="auto (fp16x3 tensor-core path when hidden=128, else fp32) | fp16x3 | fp32")
    ap.add_argument("--weights", default="auto", choices=["auto", "init"],
                    help="auto: the reference's shipped checkpoint when staged and the architecture matches; init: seeded default init")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-sample-nodes", type=int, default=50_000)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-parity", action="store_true", help="N > 1: skip the check against the single-GPU fp32 engine")
    ap.add_argument("--skip-gpu-eager", action="store_true")
    ap.add_argument("--halo", default="auto", choices=["auto", "nccl", "p2p"],
                    help="N > 1 (MuS): halo transport: NCCL all_to_all_single, or one kernel over NVLink peer memory (g4c_halo_put)")
    ap.add_argument("--overlap", action="store_true", help="N > 1 (MuS): halo exchange behind the first kernel of every block instead of in front of it")
    ap.add_argument("--rollout-check", type=int, default=0,
                    help="N > 1: additionally roll out this many steps and report rel-L2 against the single-GPU engine (configs[4])")
    return ap.parse_args()


def workload_name(a):
    if a.model == "remus":
        return (f"remus3-gnn rollout step, {a.nodes}-node/{a.nodes * a.k}-edge/{a.nodes * a.k * a.k}-angle synthetic kNN mesh, "
                f"hidden={a.hidden}")
    return f"mus{a.levels}-gnn rollout step, {a.nodes}-node/{a.nodes * a.k}-edge synthetic kNN mesh, hidden={a.hidden}"


def shipped_weights(a):
    """(state_dict, label) of the reference's trained checkpoint for this workload, or (None, ...): the staged weights-only
    copies (tools/stage_reference.py) hold hidden 128 models, 3 scales (3S-GNN-NsCircle-v1, RE3S-GNN-NsEllipse-v1)."""
    if a.weights != "auto" or a.hidden != 128 or (a.model == "mus" and a.levels != 3):
        return None, "seeded default init"
    name, label = (("NsRotEquiThreeScaleGNN.chk", "shipped RE3S-GNN-NsEllipse-v1") if a.model == "remus" else
                   ("NsThreeScaleGNN.chk", "shipped 3S-GNN-NsCircle-v1"))
    path = os.path.join(ROOT, "baseline", "_ref", "weights", name)
    if not os.path.exists(path):
        return None, "seeded default init (no staged checkpoint)"
    chk = torch.load(path, map_location="cpu", weights_only=False)
    return {k: v.float() for k, v in chk["weights"].items()}, label


def build_mesh(a, n):
    from graphs4cfd_b200 import mesh as M
    if a.model == "remus":
        return M.build_remus_mesh(n, a.k, seed=0)
    return M.build_mus_mesh(n, a.k, M.auto_cells(n, a.levels), seed=0)


def build_workload(a, n):
    """(mesh, parameters, weights label) of the benchmarked model on an n-node synthetic mesh."""
    from graphs4cfd_b200.archs import init_params, mus_arch, remus_arch
    params, label = shipped_weights(a)
    if params is None:
        params = init_params(remus_arch(a.hidden) if a.model == "remus" else mus_arch(a.hidden, a.levels), seed=0)
    return build_mesh(a, n), params, label


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, smax, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm / CPU baseline
def _model_name(a):
    return "3-scale REMuS-GNN" if a.model == "remus" else f"{a.levels}-scale MuS-GNN"


def reference_model(a, device):
    """The reference's own model class for this workload (nn/mus_gnn.py, nn/remus_gnn.py) built from its `arch` dict, with
    the shipped weights when staged.  None when no reference tree (and no staged copy) exists on this machine."""
    from oracle import pyg_stub                                   # the reference arm is the one place bench.py may use oracle/
    if pyg_stub.reference_root() is None:
        return None
    gfd = pyg_stub.import_reference()
    from graphs4cfd_b200.archs import init_params, mus_arch, remus_arch
    if a.model == "remus":
        cls, arch = gfd.nn.NsRotEquiTreeScaleGNN, remus_arch(a.hidden)
    else:
        cls = {1: gfd.nn.NsOneScaleGNN, 2: gfd.nn.NsTwoScaleGNN, 3: gfd.nn.NsThreeScaleGNN, 4: gfd.nn.NsFourScaleGNN}[a.levels]
        arch = mus_arch(a.hidden, a.levels)
    model = cls(arch=dict(arch), device=torch.device(device))
    params, _ = shipped_weights(a)
    model.load_state_dict(params if params is not None else init_params(arch, seed=0))
    return model.eval()


def _time_solver(solve, g, steps, warmup):
    """Seconds per rollout step of `solve(graph, n)` (GNN.solve: forward + shift_and_replace per step)."""
    with torch.no_grad():
        if warmup:
            solve(g.clone(), warmup)
        t0 = time.perf_counter()
        solve(g.clone(), steps)
        return (time.perf_counter() - t0) / steps


def cpu_steps_per_s(a, steps, warmup, sample_nodes, fit=False):
    """The reference's CPU path on a bounded sample of the workload (the same model on a `sample_nodes` mesh, every host
    thread), scaled to the full mesh.  kind "reference": the unmodified classes; "port": oracle/restate.py when no
    reference tree is on the machine.  fit: seconds per step additionally measured on a smaller and a larger mesh and fitted
    by a straight line in N; otherwise proportional scaling from the one sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    model = reference_model(a, "cpu")
    if model is not None:
        kind, solve = "reference", (lambda g, n: model.solve(g, n))
    else:
        from oracle import restate as R
        _, params, _ = build_workload(a, 64)
        kind, solve = "port", (lambda g, n: R.solve(params, g, n))
    n = min(sample_nodes, a.nodes)
    dt = _time_solver(solve, build_mesh(a, n), steps, warmup)
    points = {n: dt}
    full = dt * a.nodes / n
    how = f"steps/s scaled x{n / a.nodes:.4g} (proportional in nodes) to {a.nodes} nodes"
    if fit and n < a.nodes:
        lo, hi = max(n // 5, 1000), min(n * (4 if a.model == "mus" else 2), a.nodes)
        points[lo] = _time_solver(solve, build_mesh(a, lo), 3, 1)
        points[hi] = _time_solver(solve, build_mesh(a, hi), 2, 1)
        xs, ys = list(points.keys()), list(points.values())
        mx, my = sum(xs) / 3, sum(ys) / 3
        slope = sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / sum((x - mx) ** 2 for x in xs)
        full = my + slope * (a.nodes - mx)
        how = (f"seconds per step fitted by a straight line over {sorted(points)} nodes "
               f"({', '.join(f'{points[x]:.3f}' for x in sorted(points))} s) and evaluated at {a.nodes} nodes")
    cb = {"value": 1.0 / full, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": kind,
          "sample": f"{steps} steps of the same {_model_name(a)} (hidden={a.hidden}) on a {n}-node mesh, {dt:.3f} s/step measured; {how}",
          "seconds_per_step_by_nodes": {str(x): points[x] for x in sorted(points)}, "sample_fraction": n / a.nodes}
    return cb, dt


def gpu_eager_baseline(a, steps=3):
    """Stock PyTorch eager on this GPU running the reference's own classes at the full workload size (fp32, TF32 off: the
    reference's inference numerics) — the GPU baseline the fused kernels are up against."""
    if not torch.cuda.is_available():
        return None
    try:
        model = reference_model(a, "cuda")
        if model is None:
            return {"unavailable": "no reference tree on this machine"}
        g = build_mesh(a, a.nodes)
        torch.cuda.synchronize()
        with torch.no_grad():
            model.solve(g.clone(), 1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = model.solve(g.clone(), steps)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / steps
        return {"value": 1.0 / dt, "unit": "steps/s", "ms_per_step": dt * 1e3, "steps": steps, "nodes": a.nodes,
                "finite": bool(torch.isfinite(out).all()),
                "what": "the reference's own model class, stock PyTorch eager on this GPU, fp32 (TF32 off), H2D once"}
    except Exception as exc:                                        # e.g. out of memory at a size the reference cannot hold
        return {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}
    finally:
        torch.cuda.empty_cache()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt = cpu_steps_per_s(a, max(1, a.steps), max(0, a.warmup), a.cpu_sample_nodes, fit=a.steps >= 3)
    line = {"impl": "reference", "metric": "rollout_steps_per_s", "value": cb["value"], "unit": "steps/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": "host cores", "weights": shipped_weights(a)[1],
                       "step": f"each timed step is a bounded sample: one rollout step on a {min(a.cpu_sample_nodes, a.nodes)}-node mesh "
                               f"(ms_per_step); value = the fitted full-workload figure"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not a.skip_gpu_eager:
        line["gpu_eager"] = gpu_eager_baseline(a)
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_g4c(a):
    import torch.distributed as dist
    from graphs4cfd_b200 import Rollout, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries the ONE JSON line only: libraries write there too (NCCL prints its version banner on rank 0's
    # stdout), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ops.EDGE_VARIANT_DEFAULT = a.edge_variant
    g, params, weights_label = build_workload(a, a.nodes)
    if world > 1:
        # node-range partition; REMuS-GNN gets the edge-halo variant (graphs4cfd_b200/partition_remus.py)
        from graphs4cfd_b200.partition import partitioned_rollout
        eng = partitioned_rollout(params, g, rank=rank, world=world, precision=a.precision, device=dev, cuda_graph=not a.no_graph,
                                  overlap=a.overlap, halo=a.halo)
    else:
        eng = Rollout(params, g, precision=a.precision, device=dev, cuda_graph=not a.no_graph)
    N_local, nf, fw = eng.N, eng.nf, eng.field_width

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- N > 1: the partitioned result against the single-GPU fp32 engine (rank 0), before anything is timed
    parity = None
    if world > 1 and not a.skip_parity:
        parity = partition_parity(a, eng, g, params, rank, dev, dist)

    outputs = torch.empty(max(N_local, 1), nf * (a.steps + a.warmup), device=dev)

    def one_step(t):
        eng.step_only()
        ops.step_update(eng.pred, eng.node_in, fw, outputs, t)

    # ---- device-resident throughput
    for t in range(a.warmup):
        one_step(t)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ops.L.launch_count()
    barrier()
    ev0.record()
    for t in range(a.steps):
        one_step(a.warmup + t)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    launches = ops.L.launch_count() - launches0
    if not a.no_graph:
        launches = a.steps * eng.launches_per_step
    finite = bool(torch.isfinite(outputs[:N_local]).all())
    if world > 1:
        tms = torch.tensor([ms, 0.0 if finite else 1.0], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, finite = float(tms[0].item()), float(tms[1].item()) == 0.0
    ms_per_step = ms / a.steps

    # ---- end to end through the public API with HOST buffers: every step copies that step's input
    #      field host->device (pinned) and reads the prediction back device->host.
    field_host = torch.empty(max(N_local, 1), fw).pin_memory()
    field_host.copy_(eng.field0.cpu())
    pred_host = torch.empty(max(N_local, 1), nf).pin_memory()
    e2e_steps = max(3, a.steps // 2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(e2e_steps):
        eng.node_in[:, :fw].copy_(field_host, non_blocking=True)
        eng.step_only()
        pred_host.copy_(eng.pred, non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the caller consumes pred on the host
        field_host[:, fw - nf:] = pred_host               # host-side shift_and_replace (n_in = 1)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms.item())
    e2e_val = e2e_steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel: the level-1 fused MP launch (edge output kept), timed alone with
    #      CUDA events on the launching stream, inputs (3 GB of edge features) far larger than L2.
    roof = None
    if rank == 0:
        roof = roofline_mp(eng, a, dev, ms_per_step)

    if rank == 0:
        cb = None
        if not a.skip_cpu_baseline and world == 1:
            cb, _ = cpu_steps_per_s(a, 3, 1, a.cpu_sample_nodes)
        line = {"metric": "rollout_steps_per_s", "value": a.steps / (ms * 1e-3), "unit": "steps/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32" if eng.precision == "fp32" else eng.precision,
                "data": "synthetic",
                "config": {"workload": workload_name(a), "precision": eng.precision,
                           "parallelism": f"node-range partition x{world}" if world > 1 else "single GPU",
                           "cuda_graph": not a.no_graph, "weights": weights_label,
                           **({"edge_kernel": a.edge_variant} if a.edge_variant != "auto" else {}),
                           "l2": "inputs larger than L2 (level-1 %s features %.1f GB per buffer)" % (
                               ("angle", a.nodes * a.k * a.k * a.hidden * 4 / 1e9) if a.model == "remus"
                               else ("edge", a.nodes * a.k * a.hidden * 4 / 1e9))},
                "clocks": clocks, "finite": finite,
                "e2e": {"value": e2e_val, "unit": "steps/s", "h2d_bytes_per_step": N_local * fw * 4 * world,
                        "d2h_bytes_per_step": N_local * nf * 4 * world, "steps": e2e_steps},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb}
        if parity is not None:
            line["parity"] = parity
        if world > 1 and hasattr(eng, "exchanges_per_step"):
            line["config"]["halo_exchanges_per_step"] = int(eng.exchanges_per_step)
            line["config"]["halo_overlap"] = bool(getattr(eng, "overlap", False))
            line["config"]["halo_transport"] = getattr(eng, "halo", "nccl")
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # The step's CUDA graph captured NCCL kernels: drop the graph, drain the device and meet once more before the
        # communicator is torn down (destroying it while a captured graph is alive hung at N = 2 in round 1).
        eng.release_graph()
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        dist.destroy_process_group()


def partition_parity(a, eng, g, params, rank, dev, dist):
    """N > 1: rel-L2 of the gathered prediction after `warmup` rollout steps (and, with --rollout-check R, of an R-step rollout)
    against the single-GPU engine on rank 0: fp32 kernels for the warm-up check, the benchmark's precision for the long one."""
    from graphs4cfd_b200 import Rollout
    steps = max(1, a.warmup)
    out = {"steps": steps}
    mine = eng.gather(eng.solve(steps)[:, -eng.nf:].contiguous(), a.nodes)
    long_mine = eng.gather(eng.solve(a.rollout_check)[:, -eng.nf:].contiguous(), a.nodes) if a.rollout_check else None
    if rank == 0:
        try:
            single = Rollout(params, g, precision="fp32", device=dev, cuda_graph=False)
            want = single.solve(steps)[:, -eng.nf:]
            out["rel_l2_vs_single_gpu_fp32"] = float((mine - want).norm() / want.norm())
            out["max_abs"] = float((mine - want).abs().max())
            del single, want
            if long_mine is not None:
                single = Rollout(params, g, precision=eng.precision, device=dev)
                want = single.solve(a.rollout_check)[:, -eng.nf:]
                out["rollout_check"] = {"steps": a.rollout_check, "precision": eng.precision,
                                        "rel_l2_vs_single_gpu": float((long_mine - want).norm() / want.norm()),
                                        "finite": bool(torch.isfinite(long_mine).all())}
                del single, want
        except torch.OutOfMemoryError:
            out["unavailable"] = "the single-GPU engine of this workload does not fit one GPU"
        torch.cuda.empty_cache()
    dist.barrier()
    return out


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def _time_launch(fn, dev, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def step_algorithmic_bytes(eng, H):
    """Sum of the algorithmic bytes of the step's blocks (fp32 features, int32 indices, every tensor touched once;
    edge features that the model discards are not written).  Geometry helpers of the REMuS step are not counted (lower bound)."""
    total = 0
    for op, s in getattr(eng, "steps", []):
        if op == "mp":
            t = s["topo"]
            E, N = t.n_edges, t.n_targets
            total += 4 * H * ((2 if s["e_out"] is not None else 1) * E + 2 * N) + 4 * E + (0 if t.fixed_k else 4 * N)
        elif op == "rowmlp":
            rows = s.get("rows") or (int(s["segs"][0][1].numel()) if s["segs"][0][1] is not None else int(s["segs"][0][0].shape[0]))
            total += 4 * rows * (sum(int(x[0].shape[1]) for x in s["segs"]) + s["pack"].out_width)
            total += sum(4 * rows for x in s["segs"] if x[1] is not None)
        elif op == "seg":
            total += 4 * H * (int(s["idx"].numel()) + s["n"]) + 4 * int(s["idx"].numel()) + 4 * s["n"]
    return total


def recorded_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture of this
    workload (profiles/edge_kernel_traffic.json, written by tools/ncu_summary.py from the .ncu-rep), or None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "edge_kernel_traffic.json")))[kernel_key]
        return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"])
    except Exception:
        return None


def roofline_mp(eng, a, dev, ms_per_step):
    """Roofline of the dominant kernel, timed alone with CUDA events on the launching stream (torch's current
    stream, which libg4c launches on): the level-1 fused edge-MLP + aggregation launch that keeps e'.
    fp16x3: the kernel behind g4c_edge_aggr_fwd on the engine's own buffers; fp32: the fused `mp_kernel`.
    Inputs (3 GB of edge features per buffer at 1M nodes) are far larger than L2."""
    from graphs4cfd_b200 import ops
    mp_args = [s[1] for s in eng.steps if s[0] == "mp"] if hasattr(eng, "steps") else eng.mp_args   # single GPU / rank-local
    lvl1 = [m for m in mp_args if m["topo"].n_edges == mp_args[0]["topo"].n_edges]
    arg = next(m for m in lvl1 if m["e_out"] is not None)
    topo = arg["topo"]
    H = a.hidden
    E, N = topo.n_edges, topo.n_targets
    peaks, src = _peaks()
    n_with_e = sum(1 for m in lvl1 if m["e_out"] is not None)
    parts = {}
    if eng.precision == "fp16x3":
        ep, proj_s, proj_t = arg["ep"].tc_edge()
        s_in = arg.get("s_in", arg["v_in"])                 # REMuS partition: the sources are the edges incl. ghost edges
        n_rows = int(arg["v_in"].shape[0])                  # own + ghost rows on a partition
        P_r = torch.empty(int(s_in.shape[0]), 128, device=dev)
        P_c, agg = (torch.empty(n_rows, 128, device=dev) for _ in range(2))
        n_layers = ep.n_layers
        variant = a.edge_variant if a.edge_variant != "auto" else ("v5" if topo.fixed_k and topo.edge_perm is None and topo.tgt_perm is None else "v3")

        def launch():
            ops.edge_aggr(ep, topo, arg["e_in"], P_r, P_c, act_e="selu", want_e=True, e_out=arg["e_out"], agg_out=agg, p_prescaled=True)

        def launch_mp():
            ops.mp(arg["ep"], arg["np_"], topo, arg["e_in"], s_in, arg["v_in"], act_e="selu", act_t="selu",
                   want_e=True, precision=eng.precision, e_out=arg["e_out"], t_out=arg["v_out"], ws=(P_r, P_c, agg))

        dur_ms = _time_launch(launch, dev)
        n0 = ops.L.launch_count()
        launch_mp()
        mp_launches = ops.L.launch_count() - n0
        mp_ms = _time_launch(launch_mp, dev)
        # the two row-kernel launches of the block, each against its own algorithmic bytes
        dual_ms = _time_launch(lambda: ops.dual_linear_tc(proj_s, proj_t, arg["v_in"], out_a=P_c, out_b=agg), dev)
        node_ms = _time_launch(lambda: ops.rowmlp_tc(arg["np_"].tc_row([128, 128]), [(agg, None, 1.0), (arg["v_in"], None, 1.0)],
                                                     act="selu", out=arg["v_out"]), dev)
        for name, t_ms, nbytes in (("dual_linear (P_r, P_c)", dual_ms, 4 * H * 3 * n_rows), ("node_model", node_ms, 4 * H * 3 * N)):
            parts[name] = {"launch_ms": t_ms, "algorithmic_bytes": nbytes, "achieved": nbytes / t_ms / 1e6,
                           "frac": nbytes / t_ms / 1e6 / peaks["hbm_gbs"]}
        # each tensor touched once: read e, write e', read P_r, P_c, write agg (fp32 rows of H), read src ids
        alg_bytes = 4 * H * (2 * E + 3 * N) + 4 * E + (0 if topo.fixed_k else 4 * N)
        flops = 2 * E * n_layers * H * H              # K = H layers per edge (the gathered terms cost no MMA)
        kernel = (f"edge_{variant}_kernel (g4c_edge_aggr_fwd: level-1 fused edge MLP + LayerNorm + aggregation, e' kept)")
        extra = {"mp_block_ms": mp_ms, "mp_block_launches": mp_launches,
                 "mp_block_frac": (4 * H * (2 * E + 2 * N) + 4 * E) / mp_ms / 1e6 / peaks["hbm_gbs"],
                 "tensor": {"issued_fp16_tflops": 3 * flops / (dur_ms * 1e-3) / 1e12, "peak_bf16_tflops": peaks.get("bf16_tflops"),
                            "frac": 3 * flops / (dur_ms * 1e-3) / 1e12 / peaks.get("bf16_tflops", 1590.0),
                            "note": "fp16x3: every product is issued as 3 fp16 MMAs with fp32 accumulation"}}
        share = len(lvl1) * dur_ms / ms_per_step
        traffic = recorded_traffic(f"edge_{variant}:E={E}:N={N}:layers={n_layers}")
    else:
        def launch():
            ops.mp(arg["ep"], arg["np_"], topo, arg["e_in"], arg.get("s_in", arg["v_in"]), arg["v_in"], act_e="selu", act_t="selu",
                   want_e=True, precision=eng.precision, e_out=arg["e_out"], t_out=arg["v_out"])

        dur_ms = _time_launch(launch, dev, reps=5)
        alg_bytes = 4 * H * (2 * E + 2 * N) + 4 * E + (0 if topo.fixed_k else 4 * N)
        flops = 2 * E * (3 * H * H + 2 * H * H) + 2 * N * (2 * H * H + 2 * H * H)
        kernel = "mp_kernel (g4c_mp_fwd: level-1 fused edge MLP + aggregation + node MLP, fp32 FFMA)"
        extra = {}
        share = len(lvl1) * dur_ms / ms_per_step
        traffic = None
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9
    step_bytes = step_algorithmic_bytes(eng, H)
    step = None if not step_bytes else {"algorithmic_bytes": step_bytes, "achieved": step_bytes / ms_per_step / 1e6,
                                        "frac": step_bytes / ms_per_step / 1e6 / peaks["hbm_gbs"],
                                        "note": "sum of the blocks' algorithmic bytes / ms_per_step"}
    out = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
           "frac": achieved / peaks["hbm_gbs"], "peak_source": src, "traffic": traffic, "launch_ms": dur_ms,
           "algorithmic_bytes": alg_bytes, "algorithmic_tflop": flops / 1e12,
           "achieved_tflops": flops / (dur_ms * 1e-3) / 1e12, "level1_launches_per_step": len(lvl1),
           "share_of_step": share,
           "note": f"{n_with_e} of {len(lvl1)} level-1 launches write e'; share uses this launch's duration for all",
           "step": step, "parts": parts}
    out.update(extra)
    return out


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_g4c(args)
