"""MuGS-GNN plan for the rollout engine: one time step of ``NsTwoGuillardScaleGNN`` / ``NsThreeGuillardScaleGNN`` /
``NsFourGuillardScaleGNN.forward`` (nn/mugs_gnn.py:81-133, 219-295, 394-489) as a static list of libg4c launches.

The block sequence is read off the state_dict (program.py): ``mp<level>...`` names carry their level in the first digit
(``mp111`` level 1, ``mp21`` / ``mp211`` level 2, ...).  Between consecutive blocks
  * level + 1: ``restriction`` (blocks.py:9-32) — the node rows of the coarse mask are gathered (g4c_halo_pack, a row gather),
    the level's statically encoded edges take over; node and edge features of the level that is left are kept (skip);
  * level - 1: ``knn_interpolate`` (blocks.py:34-48, g4c_interp_fwd) to the finer level; the block that follows takes
    cat(interpolated, skip) (256 wide) — handed to ops.mp as two 128-wide parts, nothing is concatenated.
Static at plan time: every level's edge encoder (their inputs never change during a rollout, nn/mugs_gnn.py:87-88 recompute them
each step), topologies in aggregation order with level-local node ids, int32 index copies, the interpolation lists.
Model-level ``F.selu`` are folded into the kernels' epilogues; edge outputs the model discards (``_``) are never written.
"""
import re

import torch

from . import ops


def is_mugs(params) -> bool:
    return "edge_encoder2.MLP.linear_1.weight" in params and not any(k.startswith(("down_mp", "angle_encoder")) for k in params)


def _level_of(name: str) -> int:
    m = re.fullmatch(r"mp(\d)\d+", name)
    if m is None:
        raise ValueError(f"unexpected block {name!r} in a MuGS-GNN")
    return int(m.group(1))


def plan_mugs(eng, g):
    from .blocks import interp_layout
    from .rollout import _Pool
    dev, H = eng.device, eng.H
    f32 = lambda t: t.to(dev, torch.float32).contiguous()
    i32 = lambda t: t.to(dev).to(torch.int32).contiguous()

    node_parts = [getattr(g, a) for a in ("field", "loc", "glob", "omega") if hasattr(g, a)]
    eng.field_width = int(g.field.shape[1])
    eng.node_in = f32(torch.cat([p.float() for p in node_parts], dim=1))
    eng.field0 = eng.node_in[:, :eng.field_width].clone()
    eng.N = int(eng.node_in.shape[0])
    eng.nf = int(eng.pack("node_decoder").out_width)
    eng.check_raw(eng.node_in, "the node inputs (field, loc, glob, omega)")

    body = [name for name, kind in eng.prog if kind == "mp"]
    n_levels = max(_level_of(n) for n in body)
    sfx = lambda l: "" if l == 1 else str(l)
    # ---- levels: node sets (level-1 ids), level-local topologies, statically encoded edges, restriction / interpolation lists
    ids = {1: torch.arange(eng.N, device=dev)}
    n, topo, e_static, restrict, interp = {1: eng.N}, {}, {}, {}, {}
    for l in range(1, n_levels + 1):
        if l > 1:
            mask = getattr(g, f"coarse_mask{l}").to(dev)
            ids[l] = mask.nonzero().squeeze(1)
            n[l] = int(ids[l].numel())
            restrict[l] = i32(mask[ids[l - 1]].nonzero().squeeze(1))          # rows of level l-1 kept at level l
            name = f"{l}{l - 1}"
            y_idx = getattr(g, "y_idx_" + name).to(dev)
            n_y, k_it = interp_layout(y_idx)                                   # refuses lists that are not uniform-k and sorted
            if n_y != n[l - 1]:
                raise RuntimeError(f"MuGS plan: y_idx_{name} covers {n_y} nodes, level {l - 1} has {n[l - 1]}")
            interp[l - 1] = dict(x_idx=i32(getattr(g, "x_idx_" + name)), w=f32(getattr(g, "weights_" + name)).reshape(-1),
                                 k=k_it, n_y=n_y)
        ei = getattr(g, "edge_index" + sfx(l)).to(dev)
        if l > 1:                                                              # level-1 ids -> level-l ids (restriction)
            local = torch.full((eng.N,), -1, dtype=torch.long, device=dev)
            local[ids[l]] = torch.arange(n[l], device=dev)
            ei = local[ei]
            if bool((ei < 0).any()):
                raise RuntimeError(f"MuGS plan: edge_index{l} names nodes outside coarse_mask{l}")
        topo[l] = ops.MpTopo.from_edge_index(ei, n[l])
        ea = f32(getattr(g, "edge_attr" + sfx(l)))
        if topo[l].edge_perm is not None:                                      # store the edges in aggregation order
            ea = ea[topo[l].edge_perm.long()].contiguous()
            topo[l].edge_perm = None
        eng.check_raw(ea, "edge_attr" + sfx(l))
        e_static[l] = eng.static_mlp("edge_encoder" + sfx(l), ea)

    # ---- step program with static buffer assignment (liveness-shared)
    pool = _Pool(dev)
    steps = []
    v = pool.take(eng.N, H)
    steps.append(("rowmlp", dict(pack=eng.pack("node_encoder"), segs=[(eng.node_in, None, 1.0)], act="selu", out=v)))
    level, e = 1, e_static[1]
    skip = {}                       # level -> (node features, edge features) kept while the coarser levels run
    wide = None                     # the interpolated half of the next block's 256-wide input
    statics = {id(t) for t in e_static.values()}
    for i, name in enumerate(body):
        l = _level_of(name)
        if l == level + 1:                                                     # restriction
            skip[level] = (v, e)
            v_c = pool.take(n[l], H)
            steps.append(("call", dict(fn=(lambda src=v, idx=restrict[l], dst=v_c: ops.halo_pack(src, idx, dst)),
                                       label=f"restriction (row gather) level {level} -> {l} rows={n[l]}")))
            v, e, level = v_c, e_static[l], l
        elif l == level - 1:                                                   # interpolation to the finer level
            up = pool.take(n[l], H)
            it = interp[l]
            steps.append(("call", dict(fn=(lambda it=it, src=v, dst=up: ops.interp(src, it["x_idx"], it["w"], it["k"], it["n_y"], dst)),
                                       label=f"knn_interpolate level {level} -> {l} rows={n[l]}")))
            pool.give(v)
            if e is not None and id(e) not in statics:
                pool.give(e)
            v, e = skip.pop(l)
            wide, level = up, l
        elif l != level:
            raise ValueError(f"MuGS plan: block {name} jumps from level {level} to {l}")
        nxt = _level_of(body[i + 1]) if i + 1 < len(body) else 0
        want_e = nxt >= level                                                  # the model drops it in front of an up-sampling / the decoder
        v_new = pool.take(n[level], H)
        e_new = pool.take(topo[level].n_edges, H) if want_e else None
        ep, npk = eng.pack(name + ".edge_mlp"), eng.pack(name + ".node_mlp")
        if wide is not None:
            feats = (wide, v)
            steps.append(("call", dict(fn=(lambda ep=ep, npk=npk, tp=topo[level], e_in=e, feats=feats, e_out=e_new, v_out=v_new, nl=n[level]:
                                           ops.mp(ep, npk, tp, e_in, feats, feats, act_e="selu", act_t="selu", want_e=e_out is not None,
                                                  precision=eng.precision, e_out=e_out, t_out=v_out,
                                                  ws=eng._mp_workspace(nl, nl))),
                                       label=f"mp (256-wide features) targets={n[level]} edges={topo[level].n_edges} "
                                             f"e_out={'yes' if e_new is not None else 'no'}")))
            pool.give(wide)
            wide = None
        else:
            steps.append(("mp", dict(ep=ep, np_=npk, topo=topo[level], e_in=e, v_in=v, e_out=e_new, v_out=v_new)))
        kept = any(v is s[0] for s in skip.values())
        if not kept:
            pool.give(v)
        if id(e) not in statics and not any(e is s[1] for s in skip.values()):
            pool.give(e)
        v, e = v_new, e_new
    if level != 1:
        raise ValueError("MuGS plan: the block sequence does not return to level 1")
    eng.pred = torch.empty(eng.N, eng.nf, device=dev, dtype=torch.float32)
    resid = eng.node_in[:, eng.field_width - eng.nf:eng.field_width]
    steps.append(("rowmlp", dict(pack=eng.pack("node_decoder"), segs=[(v, None, 1.0)], act=None, out=eng.pred, residual=resid)))
    eng.steps = steps
    eng.buffer_bytes = pool.bytes
    eng.launches_per_step = len(steps) + 1
    eng.level_nodes = [n[l] for l in range(1, n_levels + 1)]
    eng._keep = (topo, e_static, restrict, interp, ids)
