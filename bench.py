#!/usr/bin/env python
"""bench.py — rollout steps/s of the B200 message-passing hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm   (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU path (oracle port)

One "step" = one rollout time step (`GNN.forward` + `shift_and_replace`, nn/model.py:316-320) of the
3-scale MuS-GNN on a synthetic mesh.  Workload at any N: the 1M-node / 6M-edge mesh, hidden=128, that
BASELINE.json's metric is quoted on (it fits one GPU), node-partitioned over N GPUs (strong scaling).
Prints ONE JSON line (rank 0).
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
                         "configs[2] (a measurement case, not the driver's bench line; N > 1 uses the edge-halo partition)")
    ap.add_argument("--edge-variant", default="auto", choices=["auto", "v3", "v5"],
                    help="kernel behind g4c_edge_aggr_fwd: auto (default: v5 on fixed in-degree launches, v3 otherwise) or pinned")
    ap.add_argument("--precision", default=os.environ.get("G4C_PRECISION", "auto"),
                    help="auto (fp16x3 tensor-core path when hidden=128, else fp32) | fp16x3 | fp32")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-sample-nodes", type=int, default=50_000)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    if a.model == "remus":
        return (f"remus3-gnn rollout step, {a.nodes}-node/{a.nodes * a.k}-edge/{a.nodes * a.k * a.k}-angle synthetic kNN mesh, "
                f"hidden={a.hidden}")
    return f"mus{a.levels}-gnn rollout step, {a.nodes}-node/{a.nodes * a.k}-edge synthetic kNN mesh, hidden={a.hidden}"


def build_workload(a, n):
    """(mesh, parameters) of the benchmarked model on an n-node synthetic mesh (seeded default-init weights)."""
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch, remus_arch
    if a.model == "remus":
        return M.build_remus_mesh(n, a.k, seed=0), init_params(remus_arch(a.hidden), seed=0)
    return M.build_mus_mesh(n, a.k, M.auto_cells(n, a.levels), seed=0), init_params(mus_arch(a.hidden, a.levels), seed=0)


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


# ----------------------------------------------------------------------------- CPU arm
def cpu_steps_per_s(a, steps, warmup, sample_nodes):
    """Oracle port (reference op sequence in plain torch, all host threads) on a bounded sample of the
    workload: the same model on a `sample_nodes` mesh; steps/s scaled linearly in N to the full mesh
    (the reference's cost is linear in nodes/edges: BASELINE.md §2)."""
    from oracle import restate as R
    torch.set_num_threads(os.cpu_count() or 1)
    n = min(sample_nodes, a.nodes)
    g, params = build_workload(a, n)
    with torch.no_grad():
        for _ in range(warmup):
            R.forward(params, g)
        t0 = time.perf_counter()
        for _ in range(steps):
            pred = R.forward(params, g)
            g.field = torch.cat([g.field[:, pred.size(1):], pred], dim=1)
        dt = (time.perf_counter() - t0) / steps
    scale = n / a.nodes
    return {"value": (1.0 / dt) * scale, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps of the same {'3-scale REMuS-GNN' if a.model == 'remus' else str(a.levels) + '-scale MuS-GNN'} (hidden={a.hidden}) on a {n}-node mesh, "
                      f"{dt:.3f} s/step measured; steps/s scaled x{scale:.4g} (linear in nodes) to {a.nodes} nodes"}, dt / scale


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
    cb, s_per_step = cpu_steps_per_s(a, steps, warmup, a.cpu_sample_nodes)
    line = {"impl": "reference", "metric": "rollout_steps_per_s", "value": cb["value"], "unit": "steps/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": "host cores", "timed_steps_on_sample": steps},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
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
    g, params = build_workload(a, a.nodes)
    if world > 1:
        # node-range partition; REMuS-GNN gets the edge-halo variant (graphs4cfd_b200/partition_remus.py)
        from graphs4cfd_b200.partition import partitioned_rollout
        eng = partitioned_rollout(params, g, rank=rank, world=world, precision=a.precision, device=dev, cuda_graph=not a.no_graph)
    else:
        eng = Rollout(params, g, precision=a.precision, device=dev, cuda_graph=not a.no_graph)
    N_local, nf, fw = eng.N, eng.nf, eng.field_width

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    outputs = torch.empty(N_local, nf * (a.steps + a.warmup), device=dev)

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
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / a.steps

    # ---- end to end through the public API with HOST buffers: every step copies that step's input
    #      field host->device (pinned) and reads the prediction back device->host.
    field_host = torch.empty(N_local, fw).pin_memory()
    field_host.copy_(eng.field0.cpu())
    pred_host = torch.empty(N_local, nf).pin_memory()
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
            cb, _ = cpu_steps_per_s(a, 2, 1, a.cpu_sample_nodes)
        line = {"metric": "rollout_steps_per_s", "value": a.steps / (ms * 1e-3), "unit": "steps/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32" if eng.precision == "fp32" else eng.precision,
                "data": "synthetic",
                "config": {"workload": workload_name(a), "precision": eng.precision,
                           "parallelism": f"node-range partition x{world}" if world > 1 else "single GPU",
                           "cuda_graph": not a.no_graph, "weights": "seeded default init",
                           **({"edge_kernel": a.edge_variant} if a.edge_variant != "auto" else {}),
                           "l2": "inputs larger than L2 (level-1 %s features %.1f GB per buffer)" % (
                               ("angle", a.nodes * a.k * a.k * a.hidden * 4 / 1e9) if a.model == "remus"
                               else ("edge", a.nodes * a.k * a.hidden * 4 / 1e9))},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "steps/s", "h2d_bytes_per_step": N_local * fw * 4 * world,
                        "d2h_bytes_per_step": N_local * nf * 4 * world, "steps": e2e_steps},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # Tearing the NCCL communicator down while CUDA graphs that captured its kernels are alive can hang
        # (seen at N=2: the line above printed, then destroy_process_group never returned).  Every rank has
        # finished its device work here: synchronise, meet once more, and leave without the teardown.
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


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


def roofline_mp(eng, a, dev, ms_per_step):
    """Roofline of the dominant kernel, timed alone with CUDA events on the launching stream (torch's current
    stream, which libg4c launches on): the level-1 fused edge-MLP + aggregation launch that keeps e'.
    fp16x3: `edge_pair_kernel` (g4c_edge_aggr_fwd) on the engine's own buffers; fp32: the fused `mp_kernel`.
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
    if eng.precision == "fp16x3":
        ep, _, _ = arg["ep"].tc_edge()
        n_rows = int(arg["v_in"].shape[0])                  # own + ghost rows on a partition
        P_r, P_c, agg = (torch.empty(n_rows, 128, device=dev) for _ in range(3))

        def launch():
            ops.edge_aggr(ep, topo, arg["e_in"], P_r, P_c, act_e="selu", want_e=True, e_out=arg["e_out"], agg_out=agg)

        def launch_mp():
            ops.mp(arg["ep"], arg["np_"], topo, arg["e_in"], arg["v_in"], arg["v_in"], act_e="selu", act_t="selu",
                   want_e=True, precision=eng.precision, e_out=arg["e_out"], t_out=arg["v_out"], ws=(P_r, P_c, agg))

        dur_ms = _time_launch(launch, dev)
        n0 = ops.L.launch_count()
        launch_mp()
        mp_launches = ops.L.launch_count() - n0
        mp_ms = _time_launch(launch_mp, dev)
        # each tensor touched once: read e, write e', read P_r, P_c, write agg (fp32 rows of H), read src ids
        alg_bytes = 4 * H * (2 * E + 3 * N) + 4 * E + (0 if topo.fixed_k else 4 * N)
        flops = 2 * E * 3 * H * H                     # three K = H layers per edge (the gathered terms cost no MMA)
        kernel = "edge_pair_kernel (g4c_edge_aggr_fwd: level-1 fused edge MLP + LayerNorm + aggregation, e' kept)"
        extra = {"mp_block_ms": mp_ms, "mp_block_launches": mp_launches,
                 "tensor": {"issued_fp16_tflops": 3 * flops / (dur_ms * 1e-3) / 1e12, "peak_bf16_tflops": peaks.get("bf16_tflops"),
                            "frac": 3 * flops / (dur_ms * 1e-3) / 1e12 / peaks.get("bf16_tflops", 1590.0),
                            "note": "fp16x3: every product is issued as 3 fp16 MMAs with fp32 accumulation"}}
        share = len(lvl1) * dur_ms / ms_per_step
        # ncu --set full capture of this launch (profiles/r1h_edge_pair_v3_ncu.txt): dram read + write bytes
        traffic = 8.055e9 if (E, N, H) == (6_000_000, 1_000_000, 128) else None
    else:
        def launch():
            ops.mp(arg["ep"], arg["np_"], topo, arg["e_in"], arg["v_in"], arg["v_in"], act_e="selu", act_t="selu",
                   want_e=True, precision=eng.precision, e_out=arg["e_out"], t_out=arg["v_out"])

        dur_ms = _time_launch(launch, dev, reps=5)
        alg_bytes = 4 * H * (2 * E + 2 * N) + 4 * E + (0 if topo.fixed_k else 4 * N)
        flops = 2 * E * (3 * H * H + 2 * H * H) + 2 * N * (2 * H * H + 2 * H * H)
        kernel = "mp_kernel (g4c_mp_fwd: level-1 fused edge MLP + aggregation + node MLP, fp32 FFMA)"
        extra = {}
        share = len(lvl1) * dur_ms / ms_per_step
        traffic = None
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9
    out = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
           "frac": achieved / peaks["hbm_gbs"], "peak_source": src, "traffic": traffic, "launch_ms": dur_ms,
           "algorithmic_bytes": alg_bytes, "algorithmic_tflop": flops / 1e12,
           "achieved_tflops": flops / (dur_ms * 1e-3) / 1e12, "level1_launches_per_step": len(lvl1),
           "share_of_step": share,
           "note": f"{n_with_e} of {len(lvl1)} level-1 launches write e'; share uses this launch's duration for all"}
    out.update(extra)
    return out


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_g4c(args)
