"""Rollout engine: ``Rollout(model_or_state_dict, graph).solve(n_out)`` == ``GNN.solve`` (nn/model.py:303-321).

At construction the engine turns (state_dict, static mesh attributes) into a plan:
  * the block program (program.py) with the model-level ``F.selu``/``tanh`` folded into kernel epilogues
    (nn/mus_gnn.py:317-367) and discarded edge outputs never written (nn/mus_gnn.py:346,354,366);
  * every level's topology in aggregation order (fixed-k level 1; CSR coarse levels stored sorted by
    target, so no permutation is needed inside the rollout), pooled-edge and children CSRs
    (the per-step ``coalesce`` sort and ``.item()`` syncs of blocks.py:45,63,109 disappear);
  * the static encoders evaluated once (edge encoder of MuS, nn/mus_gnn.py:317; angle encoders of REMuS,
    nn/remus_gnn.py:136-140);
  * pre-allocated, liveness-shared activation buffers, and one CUDA graph of a whole time step.
``solve`` then replays the graph n_out times; nothing returns to the host inside the loop.
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as LIB

from . import ops
from .blocks import children_csr, pooled_edges
from .program import block_program, hidden_width, is_remus


def _state_of(model_or_params) -> Dict[str, torch.Tensor]:
    if isinstance(model_or_params, dict):
        return model_or_params
    return {k: v.detach() for k, v in model_or_params.state_dict().items()}


class _Pool:
    """Shape-keyed free list: buffers are handed out at plan time, released at last use."""

    def __init__(self, device):
        self.device, self.free, self.bytes = device, {}, 0

    def take(self, rows, width):
        lst = self.free.setdefault((rows, width), [])
        if lst:
            return lst.pop()
        self.bytes += rows * width * 4
        return torch.empty(rows, width, device=self.device, dtype=torch.float32)

    def give(self, t):
        self.free.setdefault((t.shape[0], t.shape[1]), []).append(t)


class _Level:
    """Static data of one MuS level."""

    def __init__(self):
        self.n = 0
        self.topo: Optional[ops.MpTopo] = None
        self.e_hl = None        # relative position to the parent cell [n, 2]  (e_{l,l+1})
        self.parent = None      # int32 [n] parent id at level l+1
        self.children = None    # (ptr, idx) CSR of this level's nodes per parent
        self.pool = None        # (ptr, idx, n_coarse_edges) CSR of this level's edges per coarse edge


class Rollout:
    def __init__(self, model_or_params, graph, precision: str = "auto", device="cuda", cuda_graph: bool = True,
                 renumber: bool = True):
        """renumber: MuS-GNN plans renumber the level-1 nodes along a Morton curve of ``graph.pos`` (mesh.morton_order) so
        that gathered source rows are near in memory whatever order the mesh came in; inputs (``set_field``) and outputs
        (``solve``) stay in the caller's node order, ``node_in`` / ``pred`` / ``field0`` are in ENGINE order
        (``node_perm[i]`` = caller's index of engine row i)."""
        self.device = LIB.cuda_device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Rollout needs a CUDA device; graphs4cfd_b200 has no CPU path")
        self.params = {k: v.to(self.device) for k, v in _state_of(model_or_params).items()}
        self.H = hidden_width(self.params)
        if precision == "auto":           # tensor cores (fp16x3) for the shipped hidden width, CUDA cores otherwise
            precision = "fp16x3" if self.H == 128 else "fp32"
        self.precision = precision
        self.prog = block_program(self.params)
        self.packs = {}
        self._ws = {}
        self.use_graph = cuda_graph
        self._graph = None
        self.launches_per_step = 0
        self.node_perm = None
        from .rollout_mugs import is_mugs, plan_mugs
        if is_remus(self.params):
            from .rollout_remus import plan_remus
            plan_remus(self, graph)
        elif is_mugs(self.params):
            plan_mugs(self, graph)           # nodes keep the caller's order (no plan-time renumbering for MuGS meshes)
        else:
            if renumber and hasattr(graph, "pos") and graph.pos is not None and graph.pos.shape[0] > 1:
                from .mesh import morton_order, permute_mus_nodes
                perm = morton_order(graph.pos)
                if not np.array_equal(perm, np.arange(perm.size)):
                    graph = permute_mus_nodes(graph, perm)
                    self.node_perm = torch.from_numpy(perm).to(self.device)
            self._plan_mus(graph)

    # ------------------------------------------------------------------ helpers
    def pack(self, prefix) -> ops.MlpPack:
        p = self.packs.get(prefix)
        if p is None:
            p = self.packs[prefix] = ops.MlpPack.from_state(self.params, prefix, self.device)
        return p

    def _dev(self, t, dtype=None):
        t = t.to(self.device)
        if dtype is not None:
            t = t.to(dtype)
        return t.contiguous()

    def check_raw(self, t, what):
        """Plan-time range check of a RAW input of the fp16x3 path (ops.check_fp16_range; synchronises, once per plan)."""
        if self.precision == "fp16x3":
            ops.check_fp16_range(t, what)

    def static_mlp(self, prefix, x, act="selu", out=None):
        """Encoder of a static input, evaluated once per plan with the ENGINE's precision (nn/mus_gnn.py:317,
        nn/remus_gnn.py:136-140 recompute it every step)."""
        pack, segs = self.pack(prefix), [(x, None, 1.0)]
        return ops.rowmlp(pack, segs, act=act, out=out, precision=self._row_precision(pack, segs, None))

    # ------------------------------------------------------------------ MuS plan
    def _plan_mus(self, g):
        dev = self.device
        node_parts = [getattr(g, a) for a in ("field", "loc", "glob", "omega") if hasattr(g, a)]
        self.field_width = int(g.field.shape[1])
        self.node_in = self._dev(torch.cat([p.float() for p in node_parts], dim=1))
        self.field0 = self.node_in[:, :self.field_width].clone()
        self.N = int(self.node_in.shape[0])
        self.nf = int(self.pack("node_decoder").out_width)

        # ---- levels
        n_down = sum(1 for _, k in self.prog if k == "down")
        levels: List[_Level] = []
        ei = g.edge_index.to(dev)
        lv = _Level()
        lv.n = self.N
        lv.topo = ops.MpTopo.from_edge_index(ei, self.N)
        levels.append(lv)
        # static edge encoder output in level-1 aggregation order
        ea = self._dev(g.edge_attr.float())
        if lv.topo.edge_perm is not None:
            ea = ea[lv.topo.edge_perm.long()].contiguous()
            ei = ei[:, lv.topo.edge_perm.long()]
            lv.topo.edge_perm = None
        self.check_raw(ea, "edge_attr")
        self.check_raw(self.node_in, "the node inputs (field, loc, glob, omega)")
        self.e0 = self.static_mlp("edge_encoder", ea)
        for l in range(1, n_down + 1):
            idx = getattr(g, f"idx{l}_to_idx{l + 1}").to(dev)
            cur = levels[-1]
            cur.e_hl = self._dev(getattr(g, f"e_{l}{l + 1}").float())
            self.check_raw(cur.e_hl, f"e_{l}{l + 1}")
            cur.parent = idx.to(torch.int32).contiguous()
            n_l, cptr, cidx = children_csr(idx)
            cur.children = (cptr, cidx)
            ei_l, eptr, eidx = pooled_edges(idx, ei)
            # Store the coarse edges in aggregation order (storage order == slot order).  Coarse in-degrees vary (2..7), and a
            # unit of 128 targets costs its LARGEST in-degree in slots, so the units are formed from the targets sorted by
            # in-degree (uniform units, ~30 % fewer slots): unit row n is node node_order[n] (MpTopo.tgt_perm); node features
            # keep their own storage order.
            deg_l = torch.bincount(ei_l[1], minlength=n_l)
            node_order = torch.sort(deg_l, stable=True).indices
            rank = torch.empty_like(node_order)
            rank[node_order] = torch.arange(n_l, device=dev)
            order = torch.sort(rank[ei_l[1]], stable=True).indices
            counts = (eptr[1:] - eptr[:-1]).long()[order]
            new_ptr = torch.zeros(order.numel() + 1, dtype=torch.int64, device=dev)
            new_ptr[1:] = counts.cumsum(0)
            starts = eptr[:-1].long()[order]
            gather = torch.repeat_interleave(starts - new_ptr[:-1], counts) + torch.arange(int(new_ptr[-1]), device=dev)
            cur.pool = (new_ptr.to(torch.int32), eidx[gather].contiguous(), int(order.numel()))
            ei = ei_l[:, order]
            nxt = _Level()
            nxt.n = n_l
            rowptr = torch.zeros(n_l + 1, dtype=torch.int64, device=dev)
            rowptr[1:] = deg_l[node_order].cumsum(0)
            nxt.topo = ops.MpTopo(n_l, int(ei.size(1)), ei[0].to(torch.int32).contiguous(),
                                  rowptr=rowptr.to(torch.int32).contiguous(), tgt_perm=node_order.to(torch.int32).contiguous())
            levels.append(nxt)
        self.levels = levels

        # ---- step program with static buffer assignment (liveness-shared)
        body = [(n, k) for n, k in self.prog if k != "mlp"]
        pool = _Pool(dev)
        H = self.H
        steps = []
        v = pool.take(self.N, H)
        steps.append(("rowmlp", dict(pack=self.pack("node_encoder"), segs=[(self.node_in, None, 1.0)], act="selu", out=v)))
        e = self.e0                      # static: never released, never written
        level = 0
        saved = {}
        for i, (name, kind) in enumerate(body):
            nxt = body[i + 1][1] if i + 1 < len(body) else "decoder"
            L = levels[level]
            if kind == "mp":
                want_e = nxt not in ("up", "decoder")
                v_new = pool.take(L.n, H)
                e_new = pool.take(L.topo.n_edges, H) if want_e else None
                steps.append(("mp", dict(ep=self.pack(name + ".edge_mlp"), np_=self.pack(name + ".node_mlp"), topo=L.topo,
                                         e_in=e, v_in=v, e_out=e_new, v_out=v_new)))
                pool.give(v)
                if e is not self.e0 and not any(e is s[1] for s in saved.values()):
                    pool.give(e)
                v, e = v_new, e_new
            elif kind == "down":
                saved[level] = (v, e)
                x = pool.take(L.n, H)
                steps.append(("rowmlp", dict(pack=self.pack(name + ".down_mlp"), segs=[(L.e_hl, None, 1.0), (v, None, 1.0)],
                                             act=None, out=x)))
                nl = levels[level + 1]
                v_l = pool.take(nl.n, H)
                steps.append(("seg", dict(x=x, ptr=L.children[0], idx=L.children[1], n=nl.n, act="tanh", out=v_l)))
                pool.give(x)
                e_l = pool.take(L.pool[2], H)
                steps.append(("seg", dict(x=e, ptr=L.pool[0], idx=L.pool[1], n=L.pool[2], act=None, out=e_l)))
                v, e = v_l, e_l
                level += 1
            elif kind == "up":
                v_old, e_old = saved.pop(level - 1)
                Lh = levels[level - 1]
                v_new = pool.take(Lh.n, H)
                steps.append(("rowmlp", dict(pack=self.pack(name + ".up_mlp"),
                                             segs=[(Lh.e_hl, None, -1.0), (v, Lh.parent, 1.0), (v_old, None, 1.0)],
                                             act="tanh", out=v_new, rows=Lh.n)))
                pool.give(v)
                pool.give(v_old)
                if e is not None:
                    pool.give(e)
                v, e = v_new, e_old
                level -= 1
            else:
                raise ValueError(f"unexpected block kind {kind} in a MuS-GNN")
        self.pred = torch.empty(self.N, self.nf, device=dev, dtype=torch.float32)
        resid = self.node_in[:, self.field_width - self.nf:self.field_width]
        steps.append(("rowmlp", dict(pack=self.pack("node_decoder"), segs=[(v, None, 1.0)], act=None, out=self.pred,
                                     residual=resid)))
        self.steps = steps
        self.buffer_bytes = pool.bytes
        self.launches_per_step = len(steps) + 1      # + step_update

    # ------------------------------------------------------------------ execution
    def _row_precision(self, pack, segs, residual):
        """fp16x3 runs every row MLP the tensor-core kernel supports on it; the rest stays on the fp32 kernel."""
        if self.precision != "fp16x3":
            return "fp32"
        widths = [int(t.shape[1]) for t, _, _ in segs]
        ok = pack.tc_row_ok(widths) and not (residual is not None and pack.out_width == 128)
        return "fp16x3" if ok else "fp32"

    def _mp_workspace(self, n_src, n_tgt):
        """(P_r, P_c, agg) of the tensor-core message-passing path, shared by every block of the same size."""
        if self.precision != "fp16x3":
            return None
        ws = self._ws.get((n_src, n_tgt))
        if ws is None:
            mk = lambda n: torch.empty(n, 128, device=self.device, dtype=torch.float32)
            ws = self._ws[(n_src, n_tgt)] = (mk(n_src), mk(n_tgt), mk(n_tgt))
        return ws

    def _run_step_eager(self):
        for op, a in self.steps:
            if op == "rowmlp":
                ops.rowmlp(a["pack"], a["segs"], rows=a.get("rows"), act=a["act"], out=a["out"], residual=a.get("residual"),
                           precision=self._row_precision(a["pack"], a["segs"], a.get("residual")))
            elif op == "mp":
                s_in = a.get("s_in", a["v_in"])
                ops.mp(a["ep"], a["np_"], a["topo"], a["e_in"], s_in, a["v_in"],
                       aggr=a.get("aggr", "mean"), act_e=a.get("act_e", "selu"), act_t=a.get("act_t", "selu"),
                       want_e=a["e_out"] is not None, precision=self.precision, e_out=a["e_out"], t_out=a["v_out"],
                       ws=self._mp_workspace(int(s_in.shape[0]), int(a["v_in"].shape[0])))
            elif op == "seg":
                ops.seg_reduce(a["x"], a["ptr"], a["idx"], a["n"], "mean", a["act"], out=a["out"])
            elif op == "call":
                a["fn"]()
            else:
                raise ValueError(op)

    def _step(self):
        if not self.use_graph:
            self._run_step_eager()
            return
        if self._graph is None:
            # warm up on a side stream (lazy CUDA init, cudaFuncSetAttribute) before capture
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                n0 = ops.L.launch_count()
                self._run_step_eager()
                self.launches_per_step = ops.L.launch_count() - n0 + 1      # + step_update
            torch.cuda.current_stream(self.device).wait_stream(s)
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            with LIB.graph_capture(self._graph, self.device):
                self._run_step_eager()
        self._graph.replay()

    def set_field(self, field: torch.Tensor):
        """Load a new initial field [N, nf*n_in] (host or device) given in the CALLER's node order."""
        field = field.to(self.device, torch.float32)
        self.node_in[:, :self.field_width].copy_(field if self.node_perm is None else field[self.node_perm])

    def _set_field_engine_order(self, field: torch.Tensor):
        self.node_in[:, :self.field_width].copy_(field)

    def solve(self, n_out: int, field: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Roll the model out for n_out steps; returns [N, nf*n_out] on the device (original node order).
        The engine's own input state is restored afterwards, like GNN.solve restores graph.field."""
        assert n_out > 0, "n_out must be greater than 0."
        with torch.no_grad(), torch.cuda.device(self.device):
            if field is None:
                self._set_field_engine_order(self.field0)
            else:
                self.set_field(field)
            outputs = torch.empty(self.N, self.nf * n_out, device=self.device, dtype=torch.float32)
            if self.precision == "fp16x3":
                ops.check_fp16_range(self.node_in[:, :self.field_width], "the initial field")
            for t in range(n_out):
                self._step()
                ops.step_update(self.pred, self.node_in, self.field_width, outputs, t)
            self._set_field_engine_order(self.field0)
            if self.precision == "fp16x3":
                # every prediction was the next step's raw input: one deferred check instead of a sync per step
                ops.check_fp16_range(outputs, "a predicted field of this rollout")
            if self.node_perm is not None:            # back to the caller's node order
                ordered = torch.empty_like(outputs)
                ordered[self.node_perm] = outputs
                outputs = ordered
        return outputs

    def step_only(self):
        """One time step without the output bookkeeping (used by the benchmark's timed loop)."""
        self._step()

    def release_graph(self):
        self._graph = None
