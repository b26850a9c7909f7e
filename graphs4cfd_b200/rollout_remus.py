"""REMuS-GNN plan for the rollout engine: one time step of ``NsRotEquiTreeScaleGNN.forward``
(nn/remus_gnn.py:119-199) as a static list of libg4c launches.

Static at plan time: the five angle encoders (nn/remus_gnn.py:136-140 — their inputs never change during a
rollout), every angle topology in aggregation order (level-l angles are fixed-k groups by construction,
transforms/remus.py:36-38; the inter-level angles of ``angleIndexDownMP`` are re-ordered once so that they are
fixed-k groups in coarse-edge order), int32 index copies, the zero-filled node-vector scratch of UpEdgeMP.
"""
import torch

from . import ops


def _fixed_k_topo(angle_index, n_targets, device):
    """angle topology + permutation of the angle rows into aggregation order (None if already in order)."""
    topo = ops.MpTopo.from_edge_index(angle_index.to(device), n_targets)
    perm = None
    if topo.edge_perm is not None:
        perm = topo.edge_perm.long()
        deg = (topo.rowptr[1:] - topo.rowptr[:-1])
        k = int(deg[0]) if deg.numel() else 0
        if deg.numel() and bool((deg == k).all()):
            topo = ops.MpTopo(n_targets, topo.n_edges, topo.src, fixed_k=k)
        else:
            topo.edge_perm = None          # rows will be stored permuted, CSR stays
    return topo, perm


def plan_remus(eng, g):
    from .rollout import _Pool
    dev, H = eng.device, eng.H
    f32 = lambda t: t.to(dev, torch.float32).contiguous()
    i32 = lambda t: t.to(dev).to(torch.int32).contiguous()

    eng.field_width = int(g.field.shape[1])
    eng.node_in = f32(g.field).clone()
    eng.field0 = eng.node_in.clone()
    eng.N = int(eng.node_in.shape[0])
    eng.nf = 2
    glob, omega = f32(g.glob), f32(g.omega)
    sfx = {1: "", 2: "2", 3: "3"}
    col, U, Uinv, E, topo, a_static = {}, {}, {}, {}, {}, {}
    for l in (1, 2, 3):
        ei = getattr(g, "edge_index" + sfx[l])
        col[l] = i32(ei[1])
        U[l] = f32(getattr(g, "edgeUnitVector" + sfx[l]))
        Uinv[l] = f32(getattr(g, "edgeUnitVectorInverse" + sfx[l]))
        E[l] = int(ei.size(1))
        topo[l], perm = _fixed_k_topo(getattr(g, "angle_index" + sfx[l]), E[l], dev)
        attr = f32(getattr(g, "angle_attr" + sfx[l]))
        if perm is not None:
            attr = attr[perm].contiguous()
        eng.check_raw(attr, "angle_attr" + sfx[l])
        a_static[l] = eng.static_mlp("angle_encoder" + sfx[l], attr)
    topo_dn, a_dn = {}, {}
    for lo, name in ((1, "12"), (2, "23")):
        topo_dn[lo], perm = _fixed_k_topo(getattr(g, "angle_index" + name), E[lo + 1], dev)
        attr = f32(getattr(g, "angle_attr" + name))
        if perm is not None:
            attr = attr[perm].contiguous()
        eng.check_raw(attr, "angle_attr" + name)
        a_dn[lo] = eng.static_mlp("angle_encoder" + name, attr)
    interp = {}
    for hi, name, mask in ((2, "32", g.coarse_mask2), (1, "21", None)):
        from .blocks import interp_layout
        n_y, k_it = interp_layout(getattr(g, "y_idx_" + name))          # refuses lists that are not uniform-k and sorted
        interp[hi] = dict(x_idx=i32(getattr(g, "x_idx_" + name)), w=f32(getattr(g, "weights_" + name)).reshape(-1),
                          k=k_it, n_y=n_y,
                          y_row=None if mask is None else i32(mask.nonzero().squeeze(1)))
    vfull = torch.zeros(eng.N, 2 * H, device=dev, dtype=torch.float32)   # UpEdgeMP scratch (blocks.py:443)

    pool = _Pool(dev)
    steps = []
    F = eng.field_width // 2
    e = {}
    for l in (1, 2, 3):
        proj = pool.take(E[l], F + 2)
        steps.append(("call", dict(fn=(lambda l=l, proj=proj: ops.project(eng.node_in, col[l], U[l], (glob, omega), out=proj)))))
        e[l] = pool.take(E[l], H)
        steps.append(("rowmlp", dict(pack=eng.pack("edge_encoder" + sfx[l]), segs=[(proj, None, 1.0)], act="selu", out=e[l])))
        pool.give(proj)
    a = dict(a_static)

    def run_level(names, l, last_discards):
        for i, name in enumerate(names):
            want_a = not (last_discards and i == len(names) - 1)
            e_new = pool.take(E[l], H)
            a_new = pool.take(topo[l].n_edges, H) if want_a else None
            steps.append(("mp", dict(ep=eng.pack(name + ".angle_mlp"), np_=eng.pack(name + ".edge_mlp"), topo=topo[l],
                                     e_in=a[l], v_in=e[l], e_out=a_new, v_out=e_new)))
            pool.give(e[l])
            if a[l] is not a_static[l]:
                pool.give(a[l])
            e[l] = e_new
            if want_a:
                a[l] = a_new
            else:
                a[l] = None

    def down(name, lo):
        e_new = pool.take(E[lo + 1], H)
        steps.append(("mp", dict(ep=eng.pack(name + ".angle_mlp"), np_=eng.pack(name + ".edge_mlp"), topo=topo_dn[lo],
                                 e_in=a_dn[lo], s_in=e[lo], v_in=e[lo + 1], e_out=None, v_out=e_new)))
        pool.give(e[lo + 1])
        e[lo + 1] = e_new

    def up(name, hi):
        lo = hi + 1
        v_lo = pool.take(Uinv[lo].shape[0], 2 * H)
        steps.append(("call", dict(fn=(lambda lo=lo, v_lo=v_lo, src=e[lo]: ops.edge_to_node(src, Uinv[lo], out=v_lo)))))
        it = interp[hi]
        steps.append(("call", dict(fn=(lambda it=it, v_lo=v_lo: ops.interp(v_lo, it["x_idx"], it["w"], it["k"], it["n_y"], vfull, it["y_row"])))))
        pool.give(v_lo)
        proj = pool.take(E[hi], H)
        steps.append(("call", dict(fn=(lambda hi=hi, proj=proj: ops.project(vfull, col[hi], U[hi], (), out=proj)))))
        e_new = pool.take(E[hi], H)
        steps.append(("rowmlp", dict(pack=eng.pack(name + ".up_mlp"), segs=[(proj, None, 1.0), (e[hi], None, 1.0)],
                                     act="selu", out=e_new)))
        pool.give(proj)
        pool.give(e[hi])
        pool.give(e[lo])
        e[hi] = e_new

    run_level(["mp111", "mp112", "mp113", "mp114"], 1, False)
    down("down_mp12", 1)
    run_level(["mp211", "mp212"], 2, False)
    down("down_mp23", 2)
    run_level(["mp31", "mp32", "mp33", "mp34"], 3, True)
    up("up_mp32", 2)
    run_level(["mp221", "mp222"], 2, True)
    up("up_mp21", 1)
    run_level(["mp121", "mp122", "mp123", "mp124"], 1, True)
    dec = pool.take(E[1], 1)
    steps.append(("rowmlp", dict(pack=eng.pack("edge_decoder"), segs=[(e[1], None, 1.0)], act=None, out=dec)))
    eng.pred = torch.empty(eng.N, 2, device=dev, dtype=torch.float32)
    resid = eng.node_in[:, eng.field_width - 2:eng.field_width]
    steps.append(("call", dict(fn=(lambda: ops.edge_to_node(dec, Uinv[1], out=eng.pred, residual=resid)))))
    eng.steps = steps
    eng.buffer_bytes = pool.bytes
    eng.launches_per_step = len(steps) + 1
    eng._keep = (col, U, Uinv, topo, topo_dn, a_static, a_dn, interp, vfull, glob, omega)
