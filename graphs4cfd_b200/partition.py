"""Multi-GPU rollout: node partition of the static mesh + halo exchange over NVLink (one process per GPU).

The reference is single-device (no collective anywhere), so this layer has no reference
counterpart; it must only reproduce the single-device result.  The per-step message pass has a 1-hop
dependency on a static graph, so the mesh shards naturally:

  * level-1 nodes are split into `world` equal strips along x (contiguous ranges after sorting by x);
    a coarse node is owned by the owner of its first child.  A rank owns its nodes, ALL in-edges of its
    nodes (edge features never travel for message passing) and computes every block for them.
  * before every MP the ghost source rows (1-hop neighbours owned elsewhere) are refreshed, either by ONE kernel that packs the
    rows, stores them into the neighbours' mailboxes over NVLink peer memory, signals, waits and unpacks (PeerHalo /
    g4c_halo_put, the default on one node), or by a pack kernel -> one `all_to_all_single` (NCCL, device buffers) straight
    into the ghost tail of the feature array.
  * DownMP: children / fine edges whose parent (coarse edge) is owned elsewhere are shipped to that owner
    ("reverse halo") and appended after the local rows, so the segmented means run locally in the
    reference's summation order.  UpMP: coarse rows of remote parents are fetched the same way.

Plans are built from the full mesh on every rank with numpy (deterministic, no communication); the step
program is written once against a small backend interface so that the CPU test (tests/test_partition_gloo.py,
world_size 2, gloo) executes the very same plan and program with torch ops as the CUDA engine does with
libg4c kernels.
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as LIB

from .program import block_program, hidden_width


# ------------------------------------------------------------------------------- global structure
def _np(t):
    return t.detach().cpu().numpy()


def _ranges(start: np.ndarray, counts: np.ndarray):
    """(ptr, flat) with flat = concatenation of arange(start[i], start[i]+counts[i])."""
    ptr = np.zeros(counts.size + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    flat = np.repeat(start - ptr[:-1], counts) + np.arange(int(ptr[-1]), dtype=np.int64)
    return ptr, flat


def build_global_levels(g, n_down: int):
    """Global (whole-mesh) multi-level topology in aggregation order, mirroring Rollout._plan_mus."""
    levels = []
    row, col = _np(g.edge_index[0]).astype(np.int64), _np(g.edge_index[1]).astype(np.int64)
    n = int(g.pos.shape[0])
    order = np.argsort(col, kind="stable")
    first = dict(n=n, row=row[order], col=col[order], eperm=order)
    levels.append(first)
    for l in range(1, n_down + 1):
        cur = levels[-1]
        parent = _np(getattr(g, f"idx{l}_to_idx{l + 1}")).astype(np.int64)
        cur["parent"] = parent
        cur["e_hl"] = getattr(g, f"e_{l}{l + 1}").float()
        n_l = int(parent.max()) + 1
        # children CSR (ascending fine id inside a group)
        corder = np.argsort(parent, kind="stable")
        cptr = np.zeros(n_l + 1, dtype=np.int64)
        np.cumsum(np.bincount(parent, minlength=n_l), out=cptr[1:])
        cur["child_ptr"], cur["child_idx"] = cptr, corder
        # pooled edges: remap, drop self loops, coalesce order (row-major), then sort by target
        pr, pc = parent[cur["row"]], parent[cur["col"]]
        keep = np.nonzero(pr != pc)[0]
        key = pr[keep] * n_l + pc[keep]
        korder = np.argsort(key, kind="stable")
        uniq, start, counts = np.unique(key[korder], return_index=True, return_counts=True)
        crow, ccol = uniq // n_l, uniq % n_l
        torder = np.argsort(ccol, kind="stable")
        members = keep[korder]                       # fine edge ids grouped by coarse edge (coalesce order)
        pptr, flat = _ranges(start[torder], counts[torder])
        cur["pool_ptr"], cur["pool_idx"] = pptr, members[flat]
        levels.append(dict(n=n_l, row=crow[torder], col=ccol[torder]))
    return levels


def strip_owners(pos: torch.Tensor, world: int) -> np.ndarray:
    """Equal-size strips along x (ties broken by node id)."""
    n = pos.shape[0]
    order = np.argsort(_np(pos[:, 0]), kind="stable")
    owner = np.empty(n, dtype=np.int64)
    bounds = [(n * r) // world for r in range(world + 1)]
    for r in range(world):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


def level_owners(levels, owner1: np.ndarray):
    owners = [owner1]
    for l in range(len(levels) - 1):
        lv = levels[l]
        first_child = lv["child_idx"][lv["child_ptr"][:-1]]
        owners.append(owners[-1][first_child])
    return owners


class Xchg:
    """One halo exchange: rows `send_idx` of the buffer (grouped by destination rank) are packed and sent,
    received rows land contiguously at row `recv_off` (grouped by source rank)."""

    def __init__(self, send_idx, send_splits, recv_splits, recv_off):
        self.send_idx = np.asarray(send_idx, dtype=np.int64)
        self.send_splits = [int(x) for x in send_splits]
        self.recv_splits = [int(x) for x in recv_splits]
        self.recv_off = int(recv_off)
        self.n_send, self.n_recv = int(sum(self.send_splits)), int(sum(self.recv_splits))
        self.active = True      # set False when no rank sends anything (then every rank skips the collective)
        self.mail_base = None   # per destination rank: the row of ITS receive order where this rank's rows land (peer-memory exchange)


def _by_owner(ids: np.ndarray, owner: np.ndarray):
    """ids sorted by (owner, id) and the per-owner counts."""
    o = owner[ids]
    order = np.lexsort((ids, o))
    return ids[order], o[order]


def build_rank_plans(g, params, world: int):
    """Per-rank plans (numpy) for every rank; identical on every process."""
    prog = block_program(params)
    n_down = sum(1 for _, k in prog if k == "down")
    levels = build_global_levels(g, n_down)
    owners = level_owners(levels, strip_owners(g.pos, world))
    nl = len(levels)
    plans = [dict(levels=[dict() for _ in range(nl)]) for _ in range(world)]

    for l, lv in enumerate(levels):
        own_lists = [np.nonzero(owners[l] == r)[0] for r in range(world)]
        e_owner = owners[l][lv["col"]]
        # ---- per-rank local sets
        ghosts, eglobs = [], []
        for r in range(world):
            eglob = np.nonzero(e_owner == r)[0]
            src = lv["row"][eglob]
            gh = np.unique(src[owners[l][src] != r])
            gh, _ = _by_owner(gh, owners[l])
            ghosts.append(gh)
            eglobs.append(eglob)
        # parent ghosts (rows of THIS level fetched for the UpMP of level l-1)
        pghosts = []
        for r in range(world):
            if l == 0:
                pghosts.append(np.zeros(0, np.int64))
                continue
            fine_own = np.nonzero(owners[l - 1] == r)[0]
            par = np.unique(levels[l - 1]["parent"][fine_own])
            pg, _ = _by_owner(par[owners[l][par] != r], owners[l])
            pghosts.append(pg)
        for r in range(world):
            P = plans[r]["levels"][l]
            own = own_lists[r]
            P["own"], P["ghost"], P["pghost"], P["eglob"] = own, ghosts[r], pghosts[r], eglobs[r]
            P["n_own"], P["n_ghost"], P["n_pghost"] = own.size, ghosts[r].size, pghosts[r].size
            g2l = np.full(lv["n"], -1, dtype=np.int64)
            g2l[own] = np.arange(own.size)
            g2l[ghosts[r]] = own.size + np.arange(ghosts[r].size)
            P["g2l"] = g2l
            eg = eglobs[r]
            P["src"] = g2l[lv["row"][eg]]
            counts = np.bincount(g2l[lv["col"][eg]], minlength=own.size)[:own.size]
            rowptr = np.zeros(own.size + 1, dtype=np.int64)
            np.cumsum(counts, out=rowptr[1:])
            P["rowptr"] = rowptr
            k = int(counts[0]) if own.size else 0
            P["fixed_k"] = k if own.size and k > 0 and bool((counts == k).all()) else 0
        # ---- MP halo: ghost rows <- owners
        for r in range(world):
            P = plans[r]["levels"][l]
            send_idx, send_splits, recv_splits = [], [], []
            for q in range(world):
                need = ghosts[q][owners[l][ghosts[q]] == r] if q != r else np.zeros(0, np.int64)
                send_idx.append(P["g2l"][need])
                send_splits.append(need.size)
                recv_splits.append(int((owners[l][ghosts[r]] == q).sum()) if q != r else 0)
            P["mp_xchg"] = Xchg(np.concatenate(send_idx), send_splits, recv_splits, P["n_own"])
        # ---- UpMP halo at this (coarse) level: parent rows <- owners, stored after the MP ghosts
        if l > 0:
            for r in range(world):
                P = plans[r]["levels"][l]
                send_idx, send_splits, recv_splits = [], [], []
                for q in range(world):
                    need = pghosts[q][owners[l][pghosts[q]] == r] if q != r else np.zeros(0, np.int64)
                    send_idx.append(P["g2l"][need])
                    send_splits.append(need.size)
                    recv_splits.append(int((owners[l][pghosts[r]] == q).sum()) if q != r else 0)
                P["up_xchg"] = Xchg(np.concatenate(send_idx), send_splits, recv_splits, P["n_own"] + P["n_ghost"])
                # local index of every own fine node's parent inside this level's local array
                Pf = plans[r]["levels"][l - 1]
                par = levels[l - 1]["parent"][Pf["own"]]
                loc = P["g2l"][par].copy()                      # own parents (ghost ids are not valid here)
                remote = owners[l][par] != r
                pg_index = {int(gid): i for i, gid in enumerate(pghosts[r])}
                if remote.any():
                    loc[remote] = P["n_own"] + P["n_ghost"] + np.array([pg_index[int(x)] for x in par[remote]], dtype=np.int64)
                Pf["parent_local"] = loc

    # ---- DownMP reverse halos (children rows and fine-edge rows shipped to the owner of the parent)
    for l in range(nl - 1):
        lv = levels[l]
        parent = lv["parent"]
        own_f, own_c = owners[l], owners[l + 1]
        # children
        child_owner_of_parent = own_c[parent]                    # per fine node: who aggregates it
        edge_owner = own_f[lv["col"]]
        # per fine edge: the rank that pools it = owner of the target of its coarse edge (-1: dropped self loop)
        edge_consumer = np.full(lv["row"].size, -1, dtype=np.int64)
        pcnt = lv["pool_ptr"][1:] - lv["pool_ptr"][:-1]
        edge_consumer[lv["pool_idx"]] = np.repeat(own_c[levels[l + 1]["col"]], pcnt)
        for r in range(world):
            P, Pc = plans[r]["levels"][l], plans[r]["levels"][l + 1]
            own = P["own"]
            pos_in_own = np.full(lv["n"], -1, dtype=np.int64)
            pos_in_own[own] = np.arange(own.size)
            # rows I receive: fine nodes owned elsewhere whose parent I own, ordered by (owner, id)
            recv_nodes = np.nonzero((child_owner_of_parent == r) & (own_f != r))[0]
            recv_nodes, recv_o = _by_owner(recv_nodes, own_f)
            send_idx, send_splits, recv_splits = [], [], []
            for q in range(world):
                mine_for_q = np.nonzero((child_owner_of_parent == q) & (own_f == r))[0] if q != r else np.zeros(0, np.int64)
                send_idx.append(pos_in_own[mine_for_q])
                send_splits.append(mine_for_q.size)
                recv_splits.append(int((recv_o == q).sum()) if q != r else 0)
            P["child_xchg"] = Xchg(np.concatenate(send_idx), send_splits, recv_splits, own.size)
            ext = pos_in_own.copy()
            ext[recv_nodes] = own.size + np.arange(recv_nodes.size)
            own_c_ids = Pc["own"]
            cptr, cidx = lv["child_ptr"], lv["child_idx"]
            lptr, flat = _ranges(cptr[own_c_ids], cptr[own_c_ids + 1] - cptr[own_c_ids])
            P["children_ptr"], P["children_idx"] = lptr, ext[cidx[flat]]
            assert (P["children_idx"] >= 0).all()
            P["n_child_recv"] = recv_nodes.size
            # fine edges
            eg = P["eglob"]
            pos_in_eg = np.full(lv["row"].size, -1, dtype=np.int64)
            pos_in_eg[eg] = np.arange(eg.size)
            pool_ptr, pool_idx = lv["pool_ptr"], lv["pool_idx"]
            ce_mine = Pc["eglob"]                                  # coarse edges I own (target-sorted global ids)
            lptr, flat = _ranges(pool_ptr[ce_mine], pool_ptr[ce_mine + 1] - pool_ptr[ce_mine])
            needed = pool_idx[flat]
            recv_edges = np.nonzero((edge_consumer == r) & (edge_owner != r))[0]
            recv_edges, recv_eo = _by_owner(recv_edges, edge_owner)
            send_idx, send_splits, recv_splits = [], [], []
            for q in range(world):
                need_q = np.nonzero((edge_consumer == q) & (edge_owner == r))[0] if q != r else np.zeros(0, np.int64)
                send_idx.append(pos_in_eg[need_q])
                send_splits.append(need_q.size)
                recv_splits.append(int((recv_eo == q).sum()) if q != r else 0)
            P["edge_xchg"] = Xchg(np.concatenate(send_idx), send_splits, recv_splits, eg.size)
            eext = pos_in_eg.copy()
            eext[recv_edges] = eg.size + np.arange(recv_edges.size)
            P["pool_ptr"], P["pool_idx"] = lptr, eext[needed]
            assert (P["pool_idx"] >= 0).all()
            P["n_edge_recv"] = recv_edges.size
            P["e_hl"] = lv["e_hl"][torch.from_numpy(own)]
    # exchanges in which nobody sends anything are skipped by every rank
    talk = [set() for _ in range(world)]      # ranks that exchange rows with rank r in ANY exchange (symmetric by construction)
    mail_rows = 0                             # most rows any rank receives in one exchange (size of a mailbox half, PeerHalo)
    for l in range(nl):
        for key in ("mp_xchg", "up_xchg", "child_xchg", "edge_xchg"):
            xs = [plans[r]["levels"][l].get(key) for r in range(world)]
            if xs[0] is None:
                continue
            active = any(x.n_send > 0 for x in xs)
            for r, x in enumerate(xs):
                x.active = active
                # rows of rank r land in rank q's receive order behind the rows of the ranks before r (grouped by source)
                x.mail_base = [sum(xs[q].recv_splits[:r]) for q in range(world)]
                mail_rows = max(mail_rows, x.n_recv if active else 0)
                for q in range(world):
                    if x.send_splits[q] or x.recv_splits[q]:
                        talk[r].add(q)
                        talk[q].add(r)
    for r in range(world):
        plans[r]["eperm0"] = levels[0]["eperm"]
        plans[r]["neighbours"] = sorted(talk[r])
        plans[r]["mail_rows"] = int(mail_rows)
    return plans, prog


# ------------------------------------------------------------------------------- step program
def run_step_program(be, plan, prog, params_have_loc=None):
    """Emit one time step against backend `be` (see CudaBackend / the CPU backend of the gloo test).
    Returns the prediction buffer [n_own, nf]."""
    L = plan["levels"]
    body = [(n, k) for n, k in prog if k != "mlp"]
    v = be.alloc(0, "v")
    be.rowmlp("node_encoder", [(be.node_in, None, 1.0)], "selu", v, L[0]["n_own"])
    e = be.e0
    level = 0
    saved = {}
    for i, (name, kind) in enumerate(body):
        nxt = body[i + 1][1] if i + 1 < len(body) else "decoder"
        P = L[level]
        if kind == "mp":
            want_e = nxt not in ("up", "decoder")
            be.xchg(v, P["mp_xchg"])
            v_new = be.alloc(level, "v")
            e_new = be.alloc(level, "e") if want_e else None
            be.mp(name, level, e, v, e_new, v_new)
            be.free(v)
            if e is not be.e0 and not any(e is s[1] for s in saved.values()):
                be.free(e)
            v, e = v_new, e_new
        elif kind == "down":
            saved[level] = (v, e)
            x = be.alloc(level, "x")
            be.rowmlp(name + ".down_mlp", [(be.e_hl[level], None, 1.0), (v, None, 1.0)], None, x, P["n_own"])
            be.xchg(x, P["child_xchg"])
            v_l = be.alloc(level + 1, "v")
            be.seg(x, be.children[level], L[level + 1]["n_own"], "tanh", v_l)
            be.free(x)
            be.xchg(e, P["edge_xchg"])
            e_l = be.alloc(level + 1, "e")
            be.seg(e, be.pool[level], L[level + 1]["eglob"].size, None, e_l)
            v, e = v_l, e_l
            level += 1
        elif kind == "up":
            v_old, e_old = saved.pop(level - 1)
            Ph = L[level - 1]
            be.xchg(v, L[level]["up_xchg"])
            v_new = be.alloc(level - 1, "v")
            be.rowmlp(name + ".up_mlp", [(be.e_hl[level - 1], None, -1.0), (v, be.parent_local[level - 1], 1.0), (v_old, None, 1.0)],
                      "tanh", v_new, Ph["n_own"])
            be.free(v)
            be.free(v_old)
            if e is not None:
                be.free(e)
            v, e = v_new, e_old
            level -= 1
        else:
            raise ValueError(f"unexpected block kind {kind} in a MuS-GNN")
    be.rowmlp("node_decoder", [(v, None, 1.0)], None, be.pred, L[0]["n_own"], residual=True)
    be.free(v)
    return be.pred


def local_inputs(g, plan):
    """(node_in [n_own, W], edge_attr rows of the local level-1 edges) of this rank, CPU tensors."""
    own = torch.from_numpy(plan["levels"][0]["own"])
    parts = [getattr(g, a) for a in ("field", "loc", "glob", "omega") if hasattr(g, a)]
    node_in = torch.cat([p.float() for p in parts], dim=1)[own].contiguous()
    eperm = torch.from_numpy(plan["eperm0"])
    eg = torch.from_numpy(plan["levels"][0]["eglob"])
    edge_attr = g.edge_attr.float()[eperm][eg].contiguous()
    return node_in, edge_attr


# ------------------------------------------------------------------------------- CUDA engine
class _CudaBackend:
    def __init__(self, eng):
        self.eng = eng
        self.pool: Dict = {}
        self.steps = []
        self.labels = []          # one label per step function (tools/partition_timeline.py)
        self._pending = None      # a halo exchange not emitted yet: the message-passing block that follows may overlap it
        self._ws: Dict = {}       # per level: (P_r, P_c, agg) of the tensor-core block

    def _flush(self):
        if self._pending is not None:
            _, _, start, label = self._pending
            self._pending = None
            self.steps.append(lambda: start(False))
            self.labels.append(label)

    def alloc(self, level, kind):
        P = self.eng.plan["levels"][level]
        rows = {"v": P["n_own"] + P["n_ghost"] + P["n_pghost"],
                "e": P["eglob"].size + P.get("n_edge_recv", 0),
                "x": P["n_own"] + P.get("n_child_recv", 0)}[kind]
        lst = self.pool.setdefault((rows, kind, level), [])
        if lst:
            return lst.pop()
        t = torch.zeros(max(rows, 1), self.eng.H, device=self.eng.device, dtype=torch.float32)
        t._g4c_key = (rows, kind, level)
        self.eng.buffer_bytes += t.numel() * 4
        return t

    def free(self, t):
        self.pool.setdefault(t._g4c_key, []).append(t)

    def rowmlp(self, prefix, segs, act, out, rows, residual=False):
        from . import ops
        eng = self.eng
        res = eng.node_in[:, eng.field_width - eng.nf:eng.field_width] if residual else None
        pack = eng.pack(prefix)
        precision = "auto" if eng.precision == "fp16x3" else "fp32"      # "auto": tensor-core kernel where it supports the shape
        self._flush()
        self.steps.append(lambda: ops.rowmlp(pack, segs, rows=rows, act=act, out=out, residual=res, precision=precision))
        self.labels.append(f"rowmlp {prefix} rows={rows}")

    def mp(self, name, level, e_in, v_in, e_out, v_out):
        from . import ops
        eng = self.eng
        ep, npk, topo = eng.pack(name + ".edge_mlp"), eng.pack(name + ".node_mlp"), eng.topos[level]
        eng.mp_args.append(dict(ep=ep, np_=npk, topo=topo, e_in=e_in, v_in=v_in, e_out=e_out, v_out=v_out))    # for bench.py's roofline
        pend = self._pending
        if (pend is not None and pend[0] is v_in and eng.overlap and eng.precision == "fp16x3" and ep.tc_edge_ok()
                and npk.tc_row_ok([128, 128]) and pend[1].recv_off == topo.n_targets):
            # Tensor-core block with its halo exchange in flight behind the first kernel: the ghost rows are only SOURCES, and a
            # source enters the edge model through its product P_r = S W1s^T (mp_edge_pair.cu), so
            #   pack + all_to_all (NCCL stream)   ||   P_r, P_c of the OWN rows (one pass over v, compute stream)
            #   then P_r of the ghost rows (a few hundred rows), the fused edge kernel, the node model of the own rows.
            self._pending = None
            _, x, start, _ = pend
            n_own, g0, g1 = topo.n_targets, pend[1].recv_off, pend[1].recv_off + pend[1].n_recv
            epk, proj_s, proj_t = ep.tc_edge()
            node_pk = npk.tc_row([128, 128])
            ws = self._ws.get(level)
            if ws is None:
                mk = lambda n: torch.empty(max(n, 1), 128, device=eng.device, dtype=torch.float32)
                ws = self._ws[level] = (mk(v_in.shape[0]), mk(n_own), mk(n_own))
                eng.buffer_bytes += sum(t.numel() * 4 for t in ws)
            P_r, P_c, agg = ws

            def run():
                work = start(True)
                if n_own:
                    ops.dual_linear_tc(proj_s, proj_t, v_in[:n_own], out_a=P_r[:n_own], out_b=P_c[:n_own])
                if work is not None:
                    work.wait()
                if g1 > g0:
                    ops.rowmlp_tc(proj_s, [(v_in[g0:g1], None, 1.0)], out=P_r[g0:g1])
                if n_own:
                    ops.edge_aggr(epk, topo, e_in, P_r, P_c, aggr="mean", act_e="selu", want_e=e_out is not None,
                                  e_out=e_out, agg_out=agg, p_prescaled=True)
                    ops.rowmlp_tc(node_pk, [(agg[:n_own], None, 1.0), (v_in[:n_own], None, 1.0)], act="selu", out=v_out[:n_own])

            self.steps.append(run)
            self.labels.append(f"mp+halo {name} level={level} targets={topo.n_targets} edges={topo.n_edges} ghosts={g1 - g0}")
            return
        self._flush()
        self.steps.append(lambda: ops.mp(ep, npk, topo, e_in, v_in, v_in, act_e="selu", act_t="selu",
                                         want_e=e_out is not None, precision=eng.precision, e_out=e_out, t_out=v_out))
        self.labels.append(f"mp {name} level={level} targets={topo.n_targets} edges={topo.n_edges}")

    def seg(self, x, csr, n, act, out):
        from . import ops
        ptr, idx = csr
        self._flush()
        self.steps.append(lambda: ops.seg_reduce(x, ptr, idx, n, "mean", act, out=out))
        self.labels.append(f"seg_reduce groups={n}")

    def xchg(self, buf, x: Xchg):
        if not x.active:
            return
        from . import ops
        import torch.distributed as dist
        eng = self.eng
        send_idx = torch.from_numpy(x.send_idx).to(eng.device, torch.int32)
        if eng.p2p is not None:
            put = eng.p2p.exchange(buf, x, send_idx)
            self._flush()
            self.steps.append(put)
            self.labels.append(f"halo put (peer memory) send={x.n_send} recv={x.n_recv} rows")
            eng.exchanges_per_step += 1
            return
        stage = torch.empty(max(x.n_send, 1), eng.H, device=eng.device, dtype=torch.float32)
        eng.buffer_bytes += stage.numel() * 4

        def start(async_op):
            if x.n_send:
                ops.halo_pack(buf, send_idx, stage)
            return dist.all_to_all_single(buf[x.recv_off:x.recv_off + x.n_recv], stage[:x.n_send], x.recv_splits, x.send_splits,
                                          async_op=async_op)

        self._flush()
        self._pending = (buf, x, start, f"halo exchange send={x.n_send} recv={x.n_recv} rows")
        eng.exchanges_per_step += 1


def peer_memory_available(world: int) -> bool:
    """True when the process group is NCCL over GPUs of ONE node (every rank can map every other rank's memory) and torch
    has the symmetric-memory allocator.  Decided from facts every rank sees alike, so all ranks take the same transport."""
    import os
    import torch.distributed as dist
    if world < 2 or not dist.is_available() or not dist.is_initialized() or dist.get_backend() != "nccl":
        return False
    if int(os.environ.get("LOCAL_WORLD_SIZE", world)) != world or world > torch.cuda.device_count():
        return False
    try:
        import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
    except ImportError:
        return False
    return True


class PeerHalo:
    """Halo exchange over NVLink peer memory.  Every rank owns a mailbox (two halves of ``mail_floats`` floats: the largest receive
    of any rank in any exchange) and a flag
    array in symmetric memory (torch.distributed._symmetric_memory: every rank maps every other rank's allocation); one
    kernel per exchange (g4c_halo_put, include/g4c.h) packs the rows, stores them straight into the neighbours' mailboxes,
    publishes the exchange number to their flags, waits for theirs and copies the received rows behind the own rows of the
    feature array.  No NCCL call and no separate pack or unpack launch; the feature arrays themselves stay ordinary memory."""

    def __init__(self, eng, group, mail_floats, neighbours):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.eng, self.rank, self.world = eng, eng.rank, eng.world
        self.peers = list(neighbours)
        if len(self.peers) > _lib.MAX_PEERS:
            raise RuntimeError(f"peer-memory halo: {len(self.peers)} neighbours (max {_lib.MAX_PEERS})")
        self.mail_stride = (max(int(mail_floats), 4) + 3) // 4 * 4            # floats per half, 16-byte multiple
        self.mail = symm_mem.empty(2 * self.mail_stride, dtype=torch.float32, device=eng.device)
        self.flags = symm_mem.empty(self.world, dtype=torch.int64, device=eng.device)
        self.mail.zero_()
        self.flags.zero_()
        torch.cuda.synchronize(eng.device)
        self._handles = (symm_mem.rendezvous(self.mail, group), symm_mem.rendezvous(self.flags, group))
        self.mail_ptrs = list(self._handles[0].buffer_ptrs)
        self.flag_ptrs = list(self._handles[1].buffer_ptrs)
        self.state = torch.zeros(4, dtype=torch.int64, device=eng.device)
        eng.buffer_bytes += self.mail.numel() * 4
        dist.barrier(group)          # nobody stores into a mailbox or a flag that is not zeroed and mapped yet

    def exchange(self, buf, x, send_idx):
        """The g4c_halo_put launch of exchange ``x`` on the feature array ``buf`` (returns the callable)."""
        from . import _lib
        width = int(buf.shape[1])
        if x.n_recv * width > self.mail_stride:
            raise RuntimeError("peer-memory halo: the exchange does not fit the mailbox")
        d = _lib.HaloPutDesc()
        d.n_rows, d.width, d.n_peers = int(x.n_send), width, len(self.peers)
        d.src, d.send_idx = buf.data_ptr(), send_idx.data_ptr()
        d.state = self.state.data_ptr()
        d.mail_stride, d.mail = self.mail_stride, self.mail.data_ptr()
        d.ghost, d.n_recv = buf.data_ptr() + x.recv_off * width * 4, int(x.n_recv)
        off = 0
        for i, q in enumerate(self.peers):
            d.seg_start[i] = off
            n = x.send_splits[q]
            d.dst[i] = (self.mail_ptrs[q] + x.mail_base[q] * width * 4) if n else None
            d.peer_flag[i] = self.flag_ptrs[q] + self.rank * 8
            d.my_flag[i] = self.flag_ptrs[self.rank] + q * 8
            off += n
        d.seg_start[len(self.peers)] = off
        if off != x.n_send or any(x.recv_splits[q] for q in range(self.world) if q not in self.peers):
            raise RuntimeError("peer-memory halo: rows travel between ranks outside the neighbour set")
        keep = (d, buf, send_idx)

        def run():
            _lib.launch("g4c_halo_put", keep[0], buf, send_idx, self.state, self.mail)

        return run


class PartitionedRollout:
    """Rank-local slice of a MuS-GNN rollout.  API mirrors Rollout (solve / step_only / pred / node_in)."""

    def __init__(self, params, graph, rank: int, world: int, precision="auto", device="cuda", cuda_graph=False, renumber=True,
                 overlap=False, halo="auto"):
        """overlap: the halo exchange of a tensor-core message-passing block runs on NCCL's stream behind the block's first
        kernel (the per-node products of the own rows) instead of in front of the block.  Off by default: measured on B200
        (profiles/r2i_*, r2j_*) the extra launch for the ghost rows' products costs more than the hidden latency saves."""
        from . import ops
        self.overlap = overlap
        self.p2p = None
        self.device = LIB.cuda_device(device)
        self.rank, self.world, self.precision = rank, world, precision
        node_perm = None
        if renumber and graph.pos.shape[0] > 1:
            # same plan-time Morton renumbering as Rollout (mesh.morton_order): rows owned by a rank are then stored along the
            # curve inside its strip; `own` (used by gather) maps back to the caller's node ids
            from .mesh import morton_order, permute_mus_nodes
            node_perm = morton_order(graph.pos)
            graph = permute_mus_nodes(graph, node_perm)
        self.params = {k: v.to(self.device) for k, v in params.items()}
        self.H = hidden_width(self.params)
        if self.precision == "auto":
            self.precision = "fp16x3" if self.H == 128 else "fp32"
        self.packs = {}
        plans, self.prog = build_rank_plans(graph, params, world)
        self.plan = plan = plans[rank]
        L = plan["levels"]
        dev = self.device
        node_in, edge_attr = local_inputs(graph, plan)
        self.node_in = node_in.to(dev)
        self.field_width = int(graph.field.shape[1])
        self.field0 = self.node_in[:, :self.field_width].clone()
        self.N = int(L[0]["n_own"])
        self.nf = int(self.pack("node_decoder").out_width)
        self.own = torch.from_numpy(L[0]["own"] if node_perm is None else node_perm[L[0]["own"]])
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, torch.int32)
        self.topos = []
        for P in L:
            self.topos.append(ops.MpTopo(P["n_own"], P["eglob"].size, i32(P["src"]), fixed_k=P["fixed_k"],
                                         rowptr=None if P["fixed_k"] else i32(P["rowptr"])))
        self.e_hl = {l: L[l]["e_hl"].to(dev).contiguous() for l in range(len(L) - 1)}
        self.children = {l: (i32(L[l]["children_ptr"]), i32(L[l]["children_idx"])) for l in range(len(L) - 1)}
        self.pool = {l: (i32(L[l]["pool_ptr"]), i32(L[l]["pool_idx"])) for l in range(len(L) - 1)}
        self.parent_local = {l: i32(L[l]["parent_local"]) for l in range(len(L) - 1)}
        self.buffer_bytes = 0
        self.exchanges_per_step = 0
        self.mp_args = []
        e0_rows = L[0]["eglob"].size + L[0].get("n_edge_recv", 0)
        self.e0 = torch.zeros(max(e0_rows, 1), self.H, device=dev, dtype=torch.float32)
        row_prec = "auto" if self.precision == "fp16x3" else "fp32"
        if self.precision == "fp16x3":
            ops.check_fp16_range(edge_attr, "edge_attr")
            ops.check_fp16_range(self.node_in, "the node inputs (field, loc, glob, omega)")
        ops.rowmlp(self.pack("edge_encoder"), [(edge_attr.to(dev), None, 1.0)], act="selu", out=self.e0, precision=row_prec)
        self.e0._g4c_key = ("e0",)
        self.pred = torch.empty(max(self.N, 1), self.nf, device=dev, dtype=torch.float32)
        # halo transport: "nccl" = pack kernel + all_to_all_single; "p2p" = one kernel over NVLink peer memory (PeerHalo);
        # "auto" = p2p when every rank of the job is a GPU of this node (measured on 8 B200: 181.3 against 176.3 steps/s,
        # profiles/r2r_*), else nccl
        if halo == "auto":
            halo = "p2p" if peer_memory_available(world) and not overlap else "nccl"
        self.halo = halo if world > 1 else "nccl"
        if self.halo == "p2p":
            if overlap:
                raise ValueError("overlap=True moves the exchange to NCCL's stream: it needs halo='nccl'")
            import torch.distributed as dist
            self.p2p = PeerHalo(self, dist.group.WORLD, plan["mail_rows"] * self.H, plan["neighbours"])
        elif self.halo != "nccl":
            raise ValueError(f"halo={halo!r} (nccl, p2p)")
        be = _CudaBackend(self)
        be.node_in, be.e0, be.e_hl, be.children, be.pool, be.parent_local, be.pred = \
            self.node_in, self.e0, self.e_hl, self.children, self.pool, self.parent_local, self.pred
        run_step_program(be, plan, self.prog)
        be._flush()
        self._steps, self.step_labels = be.steps, be.labels
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
        """Local rows of the rollout output [n_own, nf*n_out]; `gather` assembles the global tensor."""
        from . import ops
        with torch.no_grad(), torch.cuda.device(self.device):
            self.node_in[:, :self.field_width].copy_(self.field0)
            out = torch.empty(max(self.N, 1), self.nf * n_out, device=self.device, dtype=torch.float32)
            for t in range(n_out):
                self.step_only()
                ops.step_update(self.pred, self.node_in, self.field_width, out, t)
            self.node_in[:, :self.field_width].copy_(self.field0)
        return out[:self.N]

    def gather(self, local_out: torch.Tensor, n_total: int) -> Optional[torch.Tensor]:
        """All ranks call; every rank gets the global [n_total, width] tensor in original node order."""
        import torch.distributed as dist
        width = local_out.shape[1]
        full = torch.zeros(n_total, width, device=self.device, dtype=torch.float32)
        full[self.own.to(self.device)] = local_out
        dist.all_reduce(full)
        return full


def partitioned_rollout(params, graph, rank: int, world: int, **kw):
    """Rank-local rollout engine for either model family: REMuS-GNN state dicts (they hold ``angle_encoder*`` keys,
    nn/remus_gnn.py:19-57) get the edge-halo partition of partition_remus.py, everything else the node-halo one."""
    if any(k.startswith("angle_encoder") for k in params):
        from .partition_remus import PartitionedRemusRollout
        kw = {k: v for k, v in kw.items() if k not in ("overlap", "renumber")}      # MuS-engine options
        return PartitionedRemusRollout(params, graph, rank, world, **kw)
    return PartitionedRollout(params, graph, rank, world, **kw)
