"""Multi-GPU REMuS-GNN rollout: node partition of the static mesh with an EDGE halo (one process per GPU).

REMuS-GNN (nn/remus_gnn.py:119-199) passes messages from angles to edges, so the partition of SURVEY.md §8e
carries edge rows where the MuS-GNN partition (partition.py) carries node rows:

  * level-1 nodes are split into `world` equal strips along x (partition.strip_owners); the node sets of the
    coarser levels are subsets of the level-1 nodes (Guillard coarsening, transforms/remus.py:93-147), so a
    node has the same owner on every level.  A rank owns its nodes, their k in-edges on every level (edges are
    stored grouped by target, k per node: a contiguous block per node) and all angles into those edges.
  * EdgeMP on level l (blocks.py:322-333): angle (j, m) reads the m-th in-edge of the SOURCE node of edge j.
    Before every EdgeMP the k in-edges of every ghost node (a source owned elsewhere) are refreshed:
    pack kernel -> one all_to_all_single (NCCL, device buffers; or, with halo="p2p", the one-kernel exchange over NVLink peer
    memory of partition.PeerHalo) into the ghost tail of the edge array, which is
    laid out as [own nodes | ghost nodes | down ghosts] x k rows, so "local node index * k + m" addresses it.
  * DownEdgeMP lo -> lo+1 (blocks.py:360-381): the senders of coarse edge (j -> q) are the level-lo in-edges
    of j; for a level-(lo+1) ghost node j they are fetched into the "down ghost" region of the level-lo array.
  * UpEdgeMP (blocks.py:408-456): node vectors of the coarse level are computed for owned nodes, the
    interpolation sources owned elsewhere are fetched (rows of width 2H), the projection on the finer edges
    only reads owned nodes.

Plans are built from the full mesh on every rank with numpy (deterministic, no communication); the step
program is written against a small backend interface so that tests/test_partition_remus_gloo.py (world_size 2,
gloo, CPU) runs the same plan and program with torch ops as the CUDA engine does with libg4c kernels.
"""
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as LIB

from .partition import Xchg, _by_owner, _np, strip_owners
from .program import hidden_width

SFX = {1: "", 2: "2", 3: "3"}


# ------------------------------------------------------------------------------- global structure
def _level_structure(g, l):
    """(V_l node ids ascending [level-1 numbering], k, source of every edge as an index into V_l)."""
    ei = _np(getattr(g, "edge_index" + SFX[l])).astype(np.int64, copy=False)
    E = ei.shape[1]
    k = int((ei[1] == ei[1][0]).sum())
    if E % k or not (ei[1].reshape(-1, k) == ei[1].reshape(-1, k)[:, :1]).all():
        raise NotImplementedError("REMuS partition: edges must be stored as k in-edges per node, grouped by target")
    nodes = ei[1].reshape(-1, k)[:, 0].copy()
    if not (np.diff(nodes) > 0).all():
        raise NotImplementedError("REMuS partition: targets must be ascending")
    src = np.searchsorted(nodes, ei[0])
    assert (nodes[src] == ei[0]).all()
    return nodes, k, src


def _expand(idx: np.ndarray, k: int) -> np.ndarray:
    """rows of the k in-edges of every node of `idx` (node-major)."""
    return (idx[:, None] * k + np.arange(k, dtype=np.int64)[None, :]).reshape(-1)


def _halo(world, owner_of, need_lists, local_of_own, recv_off, unit=1):
    """Exchange objects for "rank r needs rows need_lists[r] (ids sorted by (owner, id)) from their owners".
    local_of_own[r][id] = local row of an owned id on rank r; `unit` rows are moved per id."""
    out = []
    for r in range(world):
        send_idx, send_splits, recv_splits = [], [], []
        for q in range(world):
            need = need_lists[q][owner_of[need_lists[q]] == r] if q != r else np.zeros(0, np.int64)
            loc = local_of_own[r][need]
            assert (loc >= 0).all()
            send_idx.append(_expand(loc, unit) if unit > 1 else loc)
            send_splits.append(need.size * unit)
            recv_splits.append(int((owner_of[need_lists[r]] == q).sum()) * unit if q != r else 0)
        out.append(Xchg(np.concatenate(send_idx), send_splits, recv_splits, recv_off[r]))
    active = any(x.n_send > 0 for x in out)
    for r, x in enumerate(out):
        x.active = active
        # rows of rank r land in rank q's receive order behind the rows of the ranks before r (peer-memory exchange)
        x.mail_base = [sum(out[q].recv_splits[:r]) for q in range(world)]
    return out


def build_remus_rank_plans(g, world: int, only_rank: Optional[int] = None):
    """Per-rank plans (numpy / CPU tensors); identical on every process.  The exchange lists need every rank's ghost sets,
    but the heavy per-rank arrays (angle topology and attributes, unit vectors, interpolation lists) are only built for
    `only_rank` when it is given (at 4M nodes the angle attributes of all ranks together are 2.3 GB per process)."""
    heavy = lambda r: only_rank is None or r == only_rank
    owner1 = strip_owners(g.pos, world)
    nodes, k_of, src_of, owner = {}, {}, {}, {}
    for l in (1, 2, 3):
        nodes[l], k_of[l], src_of[l] = _level_structure(g, l)
        owner[l] = owner1[nodes[l]]
    k = k_of[1]
    if k_of[2] != k or k_of[3] != k:
        raise NotImplementedError("REMuS partition: the same k on every level")
    if nodes[1].size != g.pos.shape[0]:
        raise NotImplementedError("REMuS partition: every level-1 node must have in-edges")
    plans = [dict(levels={}, k=k, world=world) for _ in range(world)]

    # ---- per level: own nodes, ghost sources, local numbering
    own, ghost, own_pos = {}, {}, {}
    for l in (1, 2, 3):
        own[l] = [np.nonzero(owner[l] == r)[0] for r in range(world)]
        ghost[l], own_pos[l] = [], []
        for r in range(world):
            s = src_of[l][_expand(own[l][r], k)]
            gh = np.unique(s[owner[l][s] != r])
            gh, _ = _by_owner(gh, owner[l])
            ghost[l].append(gh)
            pos = np.full(nodes[l].size, -1, dtype=np.int64)
            pos[own[l][r]] = np.arange(own[l][r].size)
            own_pos[l].append(pos)
    # down ghosts of level lo = the level-(lo+1) ghost nodes, as level-lo nodes (same (owner, id) order)
    dghost = {1: [], 2: [], 3: [np.zeros(0, np.int64) for _ in range(world)]}
    for lo in (1, 2):
        for r in range(world):
            ids = nodes[lo + 1][ghost[lo + 1][r]]
            d = np.searchsorted(nodes[lo], ids)
            assert (nodes[lo][d] == ids).all(), "node sets must be nested"
            dghost[lo].append(d)

    for l in (1, 2, 3):
        a_idx = _np(getattr(g, "angle_index" + SFX[l])).astype(np.int64, copy=False)
        E = nodes[l].size * k
        if a_idx.shape[1] != E * k or not (a_idx[1] == np.repeat(np.arange(E), k)).all():
            raise NotImplementedError("REMuS partition: angles must be stored as k per edge, grouped by edge")
        a_attr = getattr(g, "angle_attr" + SFX[l]).float()
        U = getattr(g, "edgeUnitVector" + SFX[l]).float()
        Uinv = getattr(g, "edgeUnitVectorInverse" + SFX[l]).float()
        for r in range(world):
            P = plans[r]["levels"][l] = {}
            o, gh, dg = own[l][r], ghost[l][r], dghost[l][r]
            P["own"], P["n_own"], P["n_ghost"], P["n_dghost"] = o, o.size, gh.size, dg.size
            P["e_rows"] = (o.size + gh.size + dg.size) * k
            if not heavy(r):
                continue
            g2l = own_pos[l][r].copy()
            g2l[gh] = o.size + np.arange(gh.size)
            own_e = _expand(o, k)                          # global ids of my edges, local order
            own_a = _expand(own_e, k)                      # ... of my angles
            srow = a_idx[0][own_a]                         # global edge row read by each angle
            loc = g2l[srow // k]
            assert (loc >= 0).all()
            P["a_src"] = loc * k + srow % k
            P["angle_attr"] = a_attr[torch.from_numpy(own_a)].contiguous()
            P["U"] = U[torch.from_numpy(own_e)].contiguous()
            P["Uinv"] = Uinv[torch.from_numpy(o)].contiguous()
            # target node of my edges as a LOCAL LEVEL-1 own row (projection / encoders read level-1 arrays)
            row1 = own_pos[1][r][nodes[l][o]]
            assert (row1 >= 0).all()
            P["row1"] = row1
            P["col1"] = np.repeat(row1, k)
        xs = _halo(world, owner[l], ghost[l], own_pos[l], [plans[r]["levels"][l]["n_own"] * k for r in range(world)], unit=k)
        for r in range(world):
            plans[r]["levels"][l]["mp_xchg"] = xs[r]

    # ---- DownEdgeMP lo -> lo+1
    for lo, name in ((1, "12"), (2, "23")):
        hi = lo + 1
        a_idx = _np(getattr(g, "angle_index" + name)).astype(np.int64, copy=False)
        a_attr = getattr(g, "angle_attr" + name).float()
        E_hi = nodes[hi].size * k
        order = np.argsort(a_idx[1], kind="stable")
        if a_idx.shape[1] != E_hi * k or not (a_idx[1][order] == np.repeat(np.arange(E_hi), k)).all():
            raise NotImplementedError("REMuS partition: k inter-level angles per coarse edge")
        by_edge = order.reshape(E_hi, k)                   # angle rows of every coarse edge, caller's order inside
        for r in range(world):
            if not heavy(r):
                continue
            P, Pl = plans[r]["levels"][hi], plans[r]["levels"][lo]
            d2l = own_pos[lo][r].copy()
            free = d2l[dghost[lo][r]] < 0                  # a down ghost that is also owned keeps its own row
            d2l[dghost[lo][r][free]] = Pl["n_own"] + Pl["n_ghost"] + np.nonzero(free)[0]
            arows = by_edge[_expand(own[hi][r], k)].reshape(-1)
            srow = a_idx[0][arows]
            loc = d2l[srow // k]
            assert (loc >= 0).all()
            P["dn_src"] = loc * k + srow % k
            P["dn_attr"] = a_attr[torch.from_numpy(arows)].contiguous()
        xs = _halo(world, owner[lo], dghost[lo], own_pos[lo],
                   [(plans[r]["levels"][lo]["n_own"] + plans[r]["levels"][lo]["n_ghost"]) * k for r in range(world)], unit=k)
        for r in range(world):
            plans[r]["levels"][lo]["down_xchg"] = xs[r]

    # ---- UpEdgeMP hi <- lo = hi+1: interpolation sources
    for hi, name in ((2, "32"), (1, "21")):
        lo = hi + 1
        y_idx = _np(getattr(g, "y_idx_" + name)).astype(np.int64, copy=False)
        x_idx = _np(getattr(g, "x_idx_" + name)).astype(np.int64, copy=False)
        w = getattr(g, "weights_" + name).float().reshape(-1)
        n_y = nodes[hi].size
        ki = y_idx.size // n_y
        if y_idx.size != n_y * ki or not (y_idx == np.repeat(np.arange(n_y), ki)).all():
            raise NotImplementedError("REMuS partition: interpolation lists must hold k entries per fine node, ascending")
        ighost = []
        for r in range(world):
            x = x_idx[_expand(own[hi][r], ki)]
            ig = np.unique(x[owner[lo][x] != r])
            ig, _ = _by_owner(ig, owner[lo])
            ighost.append(ig)
        xs = _halo(world, owner[lo], ighost, own_pos[lo], [plans[r]["levels"][lo]["n_own"] for r in range(world)])
        for r in range(world):
            P = plans[r]["levels"][hi]
            Plo = plans[r]["levels"][lo]
            Plo["n_ighost"] = ighost[r].size
            Plo["interp_xchg"] = xs[r]
            if not heavy(r):
                continue
            i2l = own_pos[lo][r].copy()
            i2l[ighost[r]] = Plo["n_own"] + np.arange(ighost[r].size)
            rows = _expand(own[hi][r], ki)
            P["it_x"] = i2l[x_idx[rows]]
            assert (P["it_x"] >= 0).all()
            P["it_w"] = w[torch.from_numpy(rows)].contiguous()
            P["it_k"] = ki
    # peer-memory halo (partition.PeerHalo): who talks to whom in ANY exchange, and the largest receive in units of H floats
    talk, mail_rows = [set() for _ in range(world)], 0
    for l in (1, 2, 3):
        for key, units in (("mp_xchg", 1), ("down_xchg", 1), ("interp_xchg", 2)):       # interp rows are node vectors [*, 2H]
            xs = [plans[r]["levels"][l].get(key) for r in range(world)]
            if xs[0] is None:
                continue
            for r, x in enumerate(xs):
                if x.active:
                    mail_rows = max(mail_rows, x.n_recv * units)
                for q in range(world):
                    if x.send_splits[q] or x.recv_splits[q]:
                        talk[r].add(q)
                        talk[q].add(r)
    for r in range(world):
        plans[r]["own1"] = own[1][r]
        plans[r]["neighbours"] = sorted(talk[r])
        plans[r]["mail_rows"] = int(mail_rows)
    return plans


def local_inputs(g, plan):
    """(field, glob, omega) rows of this rank's level-1 nodes, CPU tensors."""
    o = torch.from_numpy(plan["own1"])
    return g.field.float()[o].contiguous(), g.glob.float()[o].contiguous(), g.omega.float()[o].contiguous()


# ------------------------------------------------------------------------------- step program
RUNS = (("level", ["mp111", "mp112", "mp113", "mp114"], 1, False), ("down", "down_mp12", 1),
        ("level", ["mp211", "mp212"], 2, False), ("down", "down_mp23", 2),
        ("level", ["mp31", "mp32", "mp33", "mp34"], 3, True), ("up", "up_mp32", 2),
        ("level", ["mp221", "mp222"], 2, True), ("up", "up_mp21", 1),
        ("level", ["mp121", "mp122", "mp123", "mp124"], 1, True))


def run_step_program(be, plan):
    """Emit one time step of NsRotEquiTreeScaleGNN.forward (nn/remus_gnn.py:119-199) against backend `be`.
    Backend interface: take(rows, width) / give(t); project(V, level, extras, out); rowmlp(prefix, segs, act, out, rows);
    mp(name, key, a_in, s_in, t_in, a_out, t_out) with key = level or ("dn", lo); edge_to_node(e, level, out, residual);
    interp(v_lo, hi, vfull); xchg(buf, x); plus the static tensors a_static[l], a_dn[lo], vfull, node_in, glob, omega, pred."""
    L, k = plan["levels"], plan["k"]
    H = be.H
    F = be.field_width // 2
    e = {}
    for l in (1, 2, 3):
        n_e = L[l]["n_own"] * k
        proj = be.take(n_e, F + 2)
        be.project(be.node_in, l, (be.glob, be.omega), proj)
        e[l] = be.take(L[l]["e_rows"], H)
        be.rowmlp("edge_encoder" + SFX[l], [(proj, None, 1.0)], "selu", e[l], n_e)
        be.give(proj)
    a = {l: be.a_static[l] for l in (1, 2, 3)}
    for run in RUNS:
        if run[0] == "level":
            _, names, l, last_discards = run
            for i, name in enumerate(names):
                want_a = not (last_discards and i == len(names) - 1)
                be.xchg(e[l], L[l]["mp_xchg"])
                e_new = be.take(L[l]["e_rows"], H)
                a_new = be.take(L[l]["n_own"] * k * k, H) if want_a else None
                be.mp(name, l, a[l], e[l], e[l], a_new, e_new)
                be.give(e[l])
                if a[l] is not be.a_static[l]:
                    be.give(a[l])
                e[l], a[l] = e_new, a_new
        elif run[0] == "down":
            _, name, lo = run
            be.xchg(e[lo], L[lo]["down_xchg"])
            e_new = be.take(L[lo + 1]["e_rows"], H)
            be.mp(name, ("dn", lo), be.a_dn[lo], e[lo], e[lo + 1], None, e_new)
            be.give(e[lo + 1])
            e[lo + 1] = e_new
        else:
            _, name, hi = run
            lo = hi + 1
            v_lo = be.take(L[lo]["n_own"] + L[lo]["n_ighost"], 2 * H)
            be.edge_to_node(e[lo], lo, v_lo, None)
            be.xchg(v_lo, L[lo]["interp_xchg"])
            be.interp(v_lo, hi, be.vfull)
            be.give(v_lo)
            n_e = L[hi]["n_own"] * k
            proj = be.take(n_e, H)
            be.project(be.vfull, hi, (), proj)
            e_new = be.take(L[hi]["e_rows"], H)
            be.rowmlp(name + ".up_mlp", [(proj, None, 1.0), (e[hi], None, 1.0)], "selu", e_new, n_e)
            be.give(proj)
            be.give(e[hi])
            be.give(e[lo])
            e[hi] = e_new
    n_e = L[1]["n_own"] * k
    dec = be.take(n_e, 1)
    be.rowmlp("edge_decoder", [(e[1], None, 1.0)], None, dec, n_e)
    be.edge_to_node(dec, 1, be.pred, be.node_in[:, be.field_width - 2:be.field_width])
    return be.pred


# ------------------------------------------------------------------------------- CUDA engine
class _CudaBackend:
    def __init__(self, eng):
        self.eng = eng
        self.H, self.field_width = eng.H, eng.field_width
        self.pool: Dict = {}
        self.ws: Dict = {}
        self.steps = []

    def take(self, rows, width):
        rows = max(int(rows), 1)
        lst = self.pool.setdefault((rows, width), [])
        if lst:
            return lst.pop()
        t = torch.zeros(rows, width, device=self.eng.device, dtype=torch.float32)
        self.eng.buffer_bytes += t.numel() * 4
        return t

    def give(self, t):
        if t is not None:
            self.pool.setdefault((int(t.shape[0]), int(t.shape[1])), []).append(t)

    def project(self, V, level, extras, out):
        from . import ops
        col, U = self.eng.col1[level], self.eng.U[level]
        if col.numel() == 0:              # this rank owns no node of the level: nothing to launch
            return
        self.steps.append(lambda: ops.project(V, col, U, extras, out=out))

    def rowmlp(self, prefix, segs, act, out, rows):
        from . import ops
        if rows == 0:
            return
        pack = self.eng.pack(prefix)
        precision = "auto" if self.eng.precision == "fp16x3" else "fp32"      # "auto": tensor-core kernel where it supports the shape
        self.steps.append(lambda: ops.rowmlp(pack, segs, rows=rows, act=act, out=out, precision=precision))

    def _workspace(self, n_src, n_tgt):
        """(P_r, P_c, agg) of the tensor-core path, shared by every block of the same size."""
        if self.eng.precision != "fp16x3":
            return None
        ws = self.ws.get((n_src, n_tgt))
        if ws is None:
            mk = lambda n: torch.empty(max(n, 1), 128, device=self.eng.device, dtype=torch.float32)
            ws = self.ws[(n_src, n_tgt)] = (mk(n_src), mk(n_tgt), mk(n_tgt))
            self.eng.buffer_bytes += (n_src + 2 * n_tgt) * 512
        return ws

    def mp(self, name, key, a_in, s_in, t_in, a_out, t_out):
        from . import ops
        eng = self.eng
        ep, npk, topo = eng.pack(name + ".angle_mlp"), eng.pack(name + ".edge_mlp"), eng.topos[key]
        if topo.n_targets == 0:
            return
        eng.mp_args.append(dict(ep=ep, np_=npk, topo=topo, e_in=a_in, s_in=s_in, v_in=t_in, e_out=a_out, v_out=t_out))
        ws = self._workspace(int(s_in.shape[0]), int(t_in.shape[0]))
        self.steps.append(lambda: ops.mp(ep, npk, topo, a_in, s_in, t_in, act_e="selu", act_t="selu",
                                         want_e=a_out is not None, precision=eng.precision, e_out=a_out, t_out=t_out, ws=ws))

    def edge_to_node(self, e, level, out, residual):
        from . import ops
        Uinv = self.eng.Uinv[level]
        if Uinv.shape[0] == 0:
            return
        self.steps.append(lambda: ops.edge_to_node(e, Uinv, out=out, residual=residual))

    def interp(self, v_lo, hi, vfull):
        from . import ops
        it = self.eng.interp[hi]
        if it["n_y"] == 0:
            return
        self.steps.append(lambda: ops.interp(v_lo, it["x_idx"], it["w"], it["k"], it["n_y"], vfull, it["y_row"]))

    def xchg(self, buf, x: Xchg):
        if not x.active:
            return
        from . import ops
        import torch.distributed as dist
        eng = self.eng
        send_idx = torch.from_numpy(x.send_idx).to(eng.device, torch.int32)
        if eng.p2p is not None:          # one kernel over NVLink peer memory (partition.PeerHalo, g4c_halo_put)
            self.steps.append(eng.p2p.exchange(buf, x, send_idx))
            eng.exchanges_per_step += 1
            return
        stage = torch.empty(max(x.n_send, 1), buf.shape[1], device=eng.device, dtype=torch.float32)
        eng.buffer_bytes += stage.numel() * 4

        def run():
            if x.n_send:
                ops.halo_pack(buf, send_idx, stage)
            dist.all_to_all_single(buf[x.recv_off:x.recv_off + x.n_recv], stage[:x.n_send], x.recv_splits, x.send_splits)

        self.steps.append(run)
        eng.exchanges_per_step += 1


class PartitionedRemusRollout:
    """Rank-local slice of a REMuS-GNN rollout.  API mirrors partition.PartitionedRollout (solve / step_only /
    gather / pred / node_in).  world = 1 runs the same plan and program without any exchange."""

    def __init__(self, params, graph, rank: int, world: int, precision="auto", device="cuda", cuda_graph=False, halo="auto"):
        """halo: "nccl" = pack kernel + all_to_all_single; "p2p" = every exchange is one kernel over NVLink peer memory
        (partition.PeerHalo, g4c_halo_put); "auto" = nccl: the 20 exchanges are 0.6 % of a REMuS step and the two transports
        measured the same on 2 B200 (10.0 vs 10.2 steps/s, profiles/r2w_*), unlike the MuS partition where p2p wins at 8 GPUs."""
        from . import ops
        self.device = dev = LIB.cuda_device(device)
        self.rank, self.world, self.precision = rank, world, precision
        self.params = {k: v.to(dev) for k, v in params.items()}
        self.H = hidden_width(self.params)
        if self.precision == "auto":
            self.precision = "fp16x3" if self.H == 128 else "fp32"
        self.packs = {}
        self.plan = plan = build_remus_rank_plans(graph, world, only_rank=rank)[rank]
        L, k = plan["levels"], plan["k"]
        field, glob, omega = local_inputs(graph, plan)
        self.node_in = field.to(dev)
        self.field_width = int(field.shape[1])
        self.field0 = self.node_in.clone()
        self.N = int(L[1]["n_own"])
        self.nf = 2
        self.own = torch.from_numpy(plan["own1"])
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, torch.int32)
        f32 = lambda t: t.to(dev, torch.float32).contiguous()
        self.topos, self.col1, self.U, self.Uinv, self.interp = {}, {}, {}, {}, {}
        self.buffer_bytes = 0
        self.exchanges_per_step = 0
        self.mp_args = []
        from .partition import PeerHalo
        if halo == "auto":
            halo = "nccl"
        if halo not in ("p2p", "nccl"):
            raise ValueError(f"halo={halo!r} (auto, nccl, p2p)")
        self.halo = halo if world > 1 else "nccl"
        self.p2p = None
        if self.halo == "p2p":
            import torch.distributed as dist
            self.p2p = PeerHalo(self, dist.group.WORLD, plan["mail_rows"] * self.H, plan["neighbours"])
        be = _CudaBackend(self)
        be.a_static, be.a_dn = {}, {}
        row_prec = "auto" if self.precision == "fp16x3" else "fp32"
        for l in (1, 2, 3):
            P = L[l]
            n_e = P["n_own"] * k
            self.topos[l] = ops.MpTopo(n_e, n_e * k, i32(P["a_src"]), fixed_k=k)
            self.col1[l], self.U[l], self.Uinv[l] = i32(P["col1"]), f32(P["U"]), f32(P["Uinv"])
            be.a_static[l] = torch.zeros(max(n_e * k, 1), self.H, device=dev)
            if n_e:
                ops.rowmlp(self.pack("angle_encoder" + SFX[l]), [(f32(P["angle_attr"]), None, 1.0)], act="selu", out=be.a_static[l],
                           precision=row_prec)
            self.buffer_bytes += be.a_static[l].numel() * 4
        for lo, name in ((1, "12"), (2, "23")):
            P = L[lo + 1]
            n_e = P["n_own"] * k
            self.topos[("dn", lo)] = ops.MpTopo(n_e, n_e * k, i32(P["dn_src"]), fixed_k=k)
            be.a_dn[lo] = torch.zeros(max(n_e * k, 1), self.H, device=dev)
            if n_e:
                ops.rowmlp(self.pack("angle_encoder" + name), [(f32(P["dn_attr"]), None, 1.0)], act="selu", out=be.a_dn[lo],
                           precision=row_prec)
            self.buffer_bytes += be.a_dn[lo].numel() * 4
        for hi in (2, 1):
            P = L[hi]
            self.interp[hi] = dict(x_idx=i32(P["it_x"]), w=f32(P["it_w"]), k=int(P["it_k"]), n_y=int(P["n_own"]),
                                   y_row=None if hi == 1 else i32(P["row1"]))
        be.vfull = torch.zeros(max(self.N, 1), 2 * self.H, device=dev)      # UpEdgeMP scratch (blocks.py:443)
        be.node_in, be.glob, be.omega = self.node_in, glob.to(dev), omega.to(dev)
        self.pred = be.pred = torch.empty(max(self.N, 1), 2, device=dev, dtype=torch.float32)
        run_step_program(be, plan)
        self._steps, self._be = be.steps, be
        self.launches_per_step = len(be.steps) + 1
        self.use_graph, self._graph = cuda_graph, None

    def pack(self, prefix):
        from . import ops
        p = self.packs.get(prefix)
        if p is None:
            p = self.packs[prefix] = ops.MlpPack.from_state(self.params, prefix, self.device)
        return p

    def step_only(self):
        if not self.use_graph:
            for fn in self._steps:
                fn()
            return
        if self._graph is None:
            from . import ops
            n0 = ops.L.launch_count()
            for fn in self._steps:
                fn()
            self.launches_per_step = ops.L.launch_count() - n0 + 1          # libg4c kernels per step (+ step_update)
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            with LIB.graph_capture(self._graph, self.device):
                for fn in self._steps:
                    fn()
        self._graph.replay()

    def release_graph(self):
        """Drop the captured step graph (it holds NCCL kernels; do this before the process group is destroyed)."""
        self._graph = None

    def solve(self, n_out: int) -> torch.Tensor:
        """Local rows of the rollout output [n_own, 2*n_out]; `gather` assembles the global tensor."""
        from . import ops
        with torch.no_grad(), torch.cuda.device(self.device):
            self.node_in.copy_(self.field0)
            out = torch.empty(max(self.N, 1), self.nf * n_out, device=self.device, dtype=torch.float32)
            for t in range(n_out):
                self.step_only()
                ops.step_update(self.pred, self.node_in, self.field_width, out, t)
            self.node_in.copy_(self.field0)
        return out[:self.N]

    def gather(self, local_out: torch.Tensor, n_total: int) -> Optional[torch.Tensor]:
        """All ranks call; every rank gets the global [n_total, width] tensor in original node order."""
        full = torch.zeros(n_total, local_out.shape[1], device=self.device, dtype=torch.float32)
        full[self.own.to(self.device)] = local_out
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(full)
        return full
