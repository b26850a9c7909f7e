"""N>1 path of the REMuS-GNN rollout on CPU: the plans and the step program of graphs4cfd_b200/partition_remus.py are
executed by gloo ranks with the oracle's block functions standing in for the kernels, edge / node-vector halos through
the same all_to_all_single the CUDA engine uses, and compared with the single-domain oracle (oracle/restate.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import ROOT, rel_l2


def _mesh_and_params(n=900, H=16, k=4, seed=5):
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, remus_arch
    return M.build_remus_mesh(n, k, seed=seed), init_params(remus_arch(H), seed=seed)


class CpuBackend:
    """torch/oracle implementation of the backend interface of partition_remus.run_step_program (tests only)."""

    def __init__(self, params, plan, field, glob, omega, use_dist=True):
        from oracle import restate as R
        from graphs4cfd_b200.partition_remus import SFX
        self.R, self.params, self.plan, self.use_dist = R, params, plan, use_dist
        self.H = params["edge_encoder.MLP.linear_1.weight"].shape[0]
        self.node_in, self.glob, self.omega = field, glob, omega
        self.field_width = field.shape[1]
        L, k = plan["levels"], plan["k"]
        self.k = k
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        self.a_static = {l: F.selu(R.mlp(params, "angle_encoder" + SFX[l], L[l]["angle_attr"])) for l in (1, 2, 3)}
        self.a_dn = {lo: F.selu(R.mlp(params, "angle_encoder" + name, L[lo + 1]["dn_attr"])) for lo, name in ((1, "12"), (2, "23"))}
        self.src = {l: t(L[l]["a_src"]) for l in (1, 2, 3)}
        self.src.update({("dn", lo): t(L[lo + 1]["dn_src"]) for lo in (1, 2)})
        self.col1 = {l: t(L[l]["col1"]) for l in (1, 2, 3)}
        self.vfull = torch.zeros(L[1]["n_own"], 2 * self.H)
        self.pred = torch.zeros(L[1]["n_own"], 2)

    def take(self, rows, width):
        return torch.full((int(rows), width), float("nan"))

    def give(self, t):
        pass

    def project(self, V, level, extras, out):
        col = self.col1[level]
        if col.numel() == 0:          # a rank that owns no node of this level
            return
        y = self.R.project_on_edges(V, col, self.plan["levels"][level]["U"])
        out[:] = torch.cat([y] + [x[col] for x in extras], dim=1)

    def rowmlp(self, prefix, segs, act, out, rows):
        x = torch.cat([scale * t[:rows] for t, _, scale in segs], dim=1)
        y = self.R.mlp(self.params, prefix, x)
        out[:rows] = F.selu(y) if act == "selu" else y

    def mp(self, name, key, a_in, s_in, t_in, a_out, t_out):
        src = self.src[key]
        n_t = src.numel() // self.k
        if n_t == 0:
            return
        idx = torch.stack([src, torch.arange(n_t).repeat_interleave(self.k)])
        if isinstance(key, tuple):
            e_new = self.R.down_edge_mp(self.params, name, torch.nan_to_num(s_in), torch.nan_to_num(t_in), a_in, idx)
        else:
            e_new, a_new = self.R.edge_mp(self.params, name, torch.nan_to_num(s_in), a_in, idx)
            if a_out is not None:
                a_out[:] = F.selu(a_new)
        t_out[:n_t] = F.selu(e_new[:n_t])

    def edge_to_node(self, e, level, out, residual):
        Uinv = self.plan["levels"][level]["Uinv"]
        n = Uinv.shape[0]
        if n == 0:
            return
        v = self.R.edge_scalar_to_node_vector(e[:n * self.k], Uinv)
        out[:n] = v if residual is None else residual + v

    def interp(self, v_lo, hi, vfull):
        P = self.plan["levels"][hi]
        n, ki = P["n_own"], P["it_k"]
        if n == 0:
            return
        y = self.R.knn_interpolate(v_lo, torch.arange(n).repeat_interleave(ki), torch.from_numpy(P["it_x"]), P["it_w"].unsqueeze(1))
        if hi == 1:
            vfull[:n] = y
        else:
            vfull[torch.from_numpy(P["row1"])] = y

    def xchg(self, buf, x):
        if not x.active or not self.use_dist:
            return
        send = buf[torch.from_numpy(x.send_idx)].contiguous()
        recv = torch.empty(x.n_recv, buf.shape[1])
        dist.all_to_all_single(recv, send, x.recv_splits, x.send_splits)
        buf[x.recv_off:x.recv_off + x.n_recv] = recv


def _golden_case():
    from conftest import load_golden, mesh_from
    d = load_golden("model_remus_h32")           # mesh built by the reference's own transforms, output of the unmodified reference
    return mesh_from(d["mesh"]), d["params"], d["out"], d["n_out"]


def _worker(rank, world, port, result_path, case):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphs4cfd_b200.partition_remus import build_remus_rank_plans, local_inputs, run_step_program
    if case == "golden":
        g, params, _, n_out = _golden_case()
    else:
        (g, params), n_out = _mesh_and_params(), 1
    plan = build_remus_rank_plans(g, world)[rank]
    be = CpuBackend(params, plan, *local_inputs(g, plan))
    outs = []
    with torch.no_grad():
        for t in range(n_out):                   # GNN.solve + shift_and_replace (nn/model.py:303-327) on the local rows
            pred = run_step_program(be, plan).clone()
            outs.append(pred)
            be.node_in = torch.cat([be.node_in[:, 2:], pred], dim=1)
    pred = torch.cat(outs, dim=1)
    full = torch.zeros(g.num_nodes, pred.shape[1])
    full[torch.from_numpy(plan["own1"])] = pred
    dist.all_reduce(full)
    if rank == 0:
        torch.save(full, result_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_remus_step_matches_single_domain_oracle(world, tmp_path):
    from oracle import restate as R
    port = 29900 + (os.getpid() % 400) + world
    path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(world, port, path, "synthetic"), nprocs=world, join=True)
    got = torch.load(path)
    g, params = _mesh_and_params()
    with torch.no_grad():
        want = R.forward(params, g)
    assert rel_l2(got, want) <= 1e-6, rel_l2(got, want)


@pytest.mark.parametrize("world", [2, 8])
def test_partitioned_remus_rollout_matches_reference_golden(world, tmp_path):
    """3-step rollout on the mesh the reference's transforms built, against the unmodified reference's output.
    At 8 ranks one rank owns no level-3 node at all (V_3 = 12 nodes): empty levels, empty exchanges."""
    port = 29900 + (os.getpid() % 400) + 7 + world
    path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(world, port, path, "golden"), nprocs=world, join=True)
    got = torch.load(path)
    _, _, want, _ = _golden_case()
    assert rel_l2(got, want) <= 1e-5, rel_l2(got, want)


def test_world1_program_matches_oracle():
    """world = 1: the same plan and program without exchanges (the configuration the CUDA engine is checked in on one GPU)."""
    from oracle import restate as R
    from graphs4cfd_b200.partition_remus import build_remus_rank_plans, local_inputs, run_step_program
    g, params = _mesh_and_params(n=500)
    plan = build_remus_rank_plans(g, 1)[0]
    be = CpuBackend(params, plan, *local_inputs(g, plan), use_dist=False)
    with torch.no_grad():
        pred = run_step_program(be, plan)
        want = R.forward(params, g)
    assert rel_l2(pred, want) <= 1e-6


@pytest.mark.parametrize("world", [2, 5, 8])
def test_remus_plan_invariants(world):
    from graphs4cfd_b200.partition_remus import build_remus_rank_plans
    g, _ = _mesh_and_params(n=2000)
    plans = build_remus_rank_plans(g, world)
    k = plans[0]["k"]
    for l in (1, 2, 3):
        owned = np.concatenate([plans[r]["levels"][l]["own"] for r in range(world)])
        assert np.array_equal(np.sort(owned), np.arange(owned.size))           # a partition of the level's nodes
        for key in ("mp_xchg", "down_xchg", "interp_xchg"):
            if key not in plans[0]["levels"][l]:
                continue
            for r in range(world):
                for q in range(world):
                    assert plans[r]["levels"][l][key].send_splits[q] == plans[q]["levels"][l][key].recv_splits[r]
        for r in range(world):
            P = plans[r]["levels"][l]
            assert (P["a_src"] >= 0).all() and (P["a_src"] < (P["n_own"] + P["n_ghost"]) * k).all()
            x = P["mp_xchg"]
            assert x.recv_off == P["n_own"] * k and x.n_recv == P["n_ghost"] * k
            if l > 1:
                Pl = plans[r]["levels"][l - 1]
                assert (P["dn_src"] < Pl["e_rows"]).all()
                assert Pl["down_xchg"].n_recv == Pl["n_dghost"] * k
                assert (Pl["it_x"] < P["n_own"] + P["n_ighost"]).all()       # level l-1 interpolates from level l


def test_partitioned_rollout_factory_dispatch(monkeypatch):
    """partition.partitioned_rollout picks the engine from the state-dict keys and forwards its arguments unchanged."""
    from graphs4cfd_b200 import partition, partition_remus
    calls = []
    monkeypatch.setattr(partition, "PartitionedRollout", lambda *a, **k: calls.append(("mus", a, k)) or "mus")
    monkeypatch.setattr(partition_remus, "PartitionedRemusRollout", lambda *a, **k: calls.append(("remus", a, k)) or "remus")
    kw = dict(precision="fp32", device="cuda:1", cuda_graph=True)
    assert partition.partitioned_rollout({"node_encoder.MLP.linear_1.weight": 0}, "g", rank=1, world=2, **kw) == "mus"
    assert partition.partitioned_rollout({"angle_encoder2.MLP.linear_1.weight": 0}, "g", rank=1, world=2, **kw) == "remus"
    assert [c[0] for c in calls] == ["mus", "remus"]
    for _, a, k in calls:
        assert a[1:] == ("g", 1, 2) and k == kw


def test_only_rank_plan_equals_full_plan():
    """only_rank skips the other ranks' heavy arrays but must give this rank exactly the plan of the full build."""
    from graphs4cfd_b200.partition_remus import build_remus_rank_plans
    from graphs4cfd_b200.partition import Xchg
    g, _ = _mesh_and_params(n=1500)
    world, r = 4, 2
    full = build_remus_rank_plans(g, world)[r]
    mine = build_remus_rank_plans(g, world, only_rank=r)[r]
    assert full["k"] == mine["k"] and np.array_equal(full["own1"], mine["own1"])
    for l in (1, 2, 3):
        A, B = full["levels"][l], mine["levels"][l]
        assert set(A) == set(B), set(A) ^ set(B)
        for key in A:
            a, b = A[key], B[key]
            if isinstance(a, Xchg):
                assert np.array_equal(a.send_idx, b.send_idx) and a.send_splits == b.send_splits
                assert a.recv_splits == b.recv_splits and a.recv_off == b.recv_off and a.active == b.active
            elif isinstance(a, torch.Tensor):
                assert torch.equal(a, b), key
            elif isinstance(a, np.ndarray):
                assert np.array_equal(a, b), key
            else:
                assert a == b, key
