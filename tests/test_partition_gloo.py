"""N>1 path on CPU: the node-partition plans and the step program of graphs4cfd_b200/partition.py are executed
by two gloo ranks with torch ops standing in for the kernels (the oracle's block functions), halo exchanges
through the same all_to_all_single the CUDA engine uses, and compared with the single-domain oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import ROOT, rel_l2


def _mesh_and_params(n=2400, H=16, levels=3, seed=3):
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, levels), seed=seed)
    return g, init_params(mus_arch(H, levels), seed=seed)


class CpuBackend:
    """torch/oracle implementation of the backend interface of partition.run_step_program (tests only)."""

    def __init__(self, params, plan, node_in, edge_attr, field_width, use_dist=True):
        from oracle import restate as R
        self.R, self.params, self.plan, self.use_dist = R, params, plan, use_dist
        L = plan["levels"]
        self.H = params["edge_encoder.MLP.linear_1.weight"].shape[0]
        self.node_in, self.field_width = node_in, field_width
        self.nf = params[[k for k in params if k.startswith("node_decoder") and k.endswith("weight")][-1]].shape[0]
        self.e0 = torch.zeros(L[0]["eglob"].size + L[0].get("n_edge_recv", 0), self.H)
        self.e0[:edge_attr.shape[0]] = F.selu(R.mlp(params, "edge_encoder", edge_attr))
        nl = len(L)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        self.e_hl = {l: L[l]["e_hl"] for l in range(nl - 1)}
        self.children = {l: (t(L[l]["children_ptr"]), t(L[l]["children_idx"])) for l in range(nl - 1)}
        self.pool = {l: (t(L[l]["pool_ptr"]), t(L[l]["pool_idx"])) for l in range(nl - 1)}
        self.parent_local = {l: t(L[l]["parent_local"]) for l in range(nl - 1)}
        self.pred = torch.zeros(L[0]["n_own"], self.nf)

    def alloc(self, level, kind):
        P = self.plan["levels"][level]
        rows = {"v": P["n_own"] + P["n_ghost"] + P["n_pghost"], "e": P["eglob"].size + P.get("n_edge_recv", 0),
                "x": P["n_own"] + P.get("n_child_recv", 0)}[kind]
        return torch.full((rows, self.H), float("nan"))

    def free(self, t):
        pass

    def rowmlp(self, prefix, segs, act, out, rows, residual=False):
        xs = [scale * (t[gather] if gather is not None else t[:rows])[:rows] for t, gather, scale in segs]
        y = self.R.mlp(self.params, prefix, torch.cat(xs, dim=1))
        if residual:
            y = y + self.node_in[:, self.field_width - self.nf:self.field_width]
        out[:rows] = {"selu": F.selu, "tanh": torch.tanh, None: (lambda z: z)}[act](y)

    def mp(self, name, level, e_in, v_in, e_out, v_out):
        P = self.plan["levels"][level]
        n_own, E = P["n_own"], P["eglob"].size
        counts = torch.from_numpy(np.diff(P["rowptr"]))
        ei = torch.stack([torch.from_numpy(P["src"]), torch.arange(n_own).repeat_interleave(counts)])
        v_new, e_new = self.R.gn_block(self.params, name, torch.nan_to_num(v_in), e_in[:E], ei)
        v_out[:n_own] = F.selu(v_new[:n_own])
        if e_out is not None:
            e_out[:E] = F.selu(e_new)

    def seg(self, x, csr, n, act, out):
        ptr, idx = csr
        gid = torch.arange(n).repeat_interleave(ptr[1:] - ptr[:-1])
        y = self.R.scatter_mean(x[idx], gid, n)
        out[:n] = torch.tanh(y) if act == "tanh" else y

    def xchg(self, buf, x):
        if not x.active:
            return
        send = buf[torch.from_numpy(x.send_idx)].contiguous()
        recv = torch.empty(x.n_recv, buf.shape[1])
        dist.all_to_all_single(recv, send, x.recv_splits, x.send_splits)
        buf[x.recv_off:x.recv_off + x.n_recv] = recv


def _worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphs4cfd_b200.partition import build_rank_plans, local_inputs, run_step_program
    g, params = _mesh_and_params()
    plans, prog = build_rank_plans(g, params, world)
    plan = plans[rank]
    node_in, edge_attr = local_inputs(g, plan)
    be = CpuBackend(params, plan, node_in, edge_attr, g.field.shape[1])
    with torch.no_grad():
        pred = run_step_program(be, plan, prog)
    full = torch.zeros(g.num_nodes, pred.shape[1])
    full[torch.from_numpy(plan["levels"][0]["own"])] = pred
    dist.all_reduce(full)
    if rank == 0:
        torch.save(full, result_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_partitioned_step_matches_single_domain_oracle(world, tmp_path):
    from oracle import restate as R
    port = 29500 + (os.getpid() % 400)
    path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(world, port, path), nprocs=world, join=True)
    got = torch.load(path)
    g, params = _mesh_and_params()
    with torch.no_grad():
        want = R.forward(params, g)
    assert rel_l2(got, want) <= 1e-6, rel_l2(got, want)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_plan_invariants(world):
    from graphs4cfd_b200.partition import build_rank_plans
    g, params = _mesh_and_params(n=3000)
    plans, _ = build_rank_plans(g, params, world)
    nl = len(plans[0]["levels"])
    for l in range(nl):
        owned = np.concatenate([plans[r]["levels"][l]["own"] for r in range(world)])
        assert np.array_equal(np.sort(owned), np.arange(owned.size))           # a partition of the level
        for key in ("mp_xchg", "up_xchg", "child_xchg", "edge_xchg"):
            if key not in plans[0]["levels"][l]:
                continue
            for r in range(world):
                for q in range(world):
                    assert plans[r]["levels"][l][key].send_splits[q] == plans[q]["levels"][l][key].recv_splits[r]
        for r in range(world):
            P = plans[r]["levels"][l]
            assert (P["src"] >= 0).all() and (P["src"] < P["n_own"] + P["n_ghost"]).all()
            assert P["rowptr"][-1] == P["eglob"].size
