"""TEST INFRASTRUCTURE — CPU restatement of the graphs4cfd message-passing hot path.

This is the ORACLE: a plain-torch (CPU, fp32, dense ATen ops) restatement of the
reference algorithm, written functionally over a flat ``state_dict``.  It keeps the
reference's op sequence (concat -> Linear/SELU chain -> LayerNorm -> scatter) so that it
is also a fair stand-in for the reference's CPU cost (``bench.py`` ``cpu_baseline`` /
``--impl reference``, kind "port").  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it; nothing under ``graphs4cfd_b200/`` does.

Parity pin: the reference has NO tests or golden vectors of its own, so
the pin is the reference itself: ``oracle/make_golden.py`` imports the UNMODIFIED
``/root/reference/graphs4cfd`` (under ``oracle/pyg_stub.py``) in the build container and
writes ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` holds this file to those
vectors (bit-for-bit where the op order is identical, 1e-6 otherwise) and
``tests/test_oracle_vs_reference.py`` re-runs the live comparison when the reference tree
is present.

Every function cites the reference lines it follows (paths relative to /root/reference/).
"""
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ------------------------------------------------------------- PyG utility semantics
def scatter_sum(src, index, dim_size):
    """torch_geometric.utils.scatter(reduce='sum') on dim 0 (call sites blocks.py:46-47)."""
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def scatter_mean(src, index, dim_size):
    """torch_geometric.utils.scatter(reduce='mean'): sum / clamp(count, 1)
    (call sites blocks.py:183,231,330,378)."""
    total = scatter_sum(src, index, dim_size)
    count = torch.zeros(dim_size, dtype=src.dtype).index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
    return total / count.clamp(min=1).view((-1,) + (1,) * (src.dim() - 1))


def pooled_edge_topology(idx_hr_to_lr, edge_index):
    """Static half of pool_edge (blocks.py:51-68): remap endpoints, drop self loops,
    sort remaining (row, col) pairs row-major and merge duplicates.
    Returns (edge_index_lr[2,E_l], keep_mask[E_h], group_id[kept]) so that the dynamic
    half is a mean of the kept fine-edge features over ``group_id``."""
    num_nodes = int(idx_hr_to_lr.max()) + 1
    ei = idx_hr_to_lr[edge_index.reshape(-1)].view(2, -1)
    keep = ei[0] != ei[1]
    ei = ei[:, keep]
    key = ei[0] * num_nodes + ei[1]
    uniq, group = torch.unique(key, sorted=True, return_inverse=True)
    ei_lr = torch.stack([uniq // num_nodes, uniq % num_nodes])
    return ei_lr, keep, group


def pool_edge(idx_hr_to_lr, edge_index, edge_attr):
    """blocks.py:51-68 with aggr='mean'."""
    ei_lr, keep, group = pooled_edge_topology(idx_hr_to_lr, edge_index)
    if ei_lr.size(1) == 0:
        return ei_lr, edge_attr[keep]
    return ei_lr, scatter_mean(edge_attr[keep], group, ei_lr.size(1))


# --------------------------------------------------------------------------- blocks
def mlp(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """MLP.forward (blocks.py:129-144): linear_1, selu, ..., linear_L [, layer_norm]."""
    n_lin = 0
    while f"{prefix}.MLP.linear_{n_lin + 1}.weight" in p:
        n_lin += 1
    assert n_lin >= 2, f"no MLP under {prefix}"
    for i in range(1, n_lin + 1):
        x = F.linear(x, p[f"{prefix}.MLP.linear_{i}.weight"], p[f"{prefix}.MLP.linear_{i}.bias"])
        if i < n_lin:
            x = F.selu(x)
    g = p.get(f"{prefix}.MLP.layer_norm.weight")
    if g is not None:
        x = F.layer_norm(x, (x.size(-1),), g, p[f"{prefix}.MLP.layer_norm.bias"], 1e-5)
    return x


def gn_block(p: Params, name: str, v, e, edge_index, aggr: str = "mean"):
    """GNBlock.forward (blocks.py:175-186)."""
    row, col = edge_index[0], edge_index[1]
    e = mlp(p, f"{name}.edge_mlp", torch.cat((e, v[row], v[col]), dim=-1))
    red = scatter_mean if aggr == "mean" else scatter_sum
    agg = red(e, col, v.size(0))
    v = mlp(p, f"{name}.node_mlp", torch.cat((agg, v), dim=-1))
    return v, e


def down_mp(p: Params, name: str, field_h, e_hl, idx_h_to_l, edge_index_h, edge_attr_h,
            activation: Optional[Callable] = torch.tanh):
    """DownMP.forward (blocks.py:219-237).  ``scatter(e, cluster)[mask]`` equals a segmented
    mean over ``idx_h_to_l`` because mask lists the non-empty clusters in ascending order
    (transforms/mus.py:27-31)."""
    x = mlp(p, f"{name}.down_mlp", torch.cat((e_hl, field_h), dim=-1))
    n_l = int(idx_h_to_l.max()) + 1
    field_l = scatter_mean(x, idx_h_to_l, n_l)
    if activation is not None:
        field_l = activation(field_l)
    ei_l, ea_l = pool_edge(idx_h_to_l, edge_index_h, edge_attr_h)
    return field_l, ei_l, ea_l


def up_mp(p: Params, name: str, field_l, field_h_old, e_hl, idx_h_to_l,
          activation: Optional[Callable] = torch.tanh):
    """UpMP.forward (blocks.py:265-290)."""
    x = mlp(p, f"{name}.up_mlp", torch.cat((-e_hl, field_l[idx_h_to_l], field_h_old), dim=-1))
    return activation(x) if activation is not None else x


def edge_mp(p: Params, name: str, e, a, angle_index, aggr: str = "mean"):
    """EdgeMP.forward (blocks.py:322-333)."""
    row, col = angle_index[0], angle_index[1]
    a = mlp(p, f"{name}.angle_mlp", torch.cat((a, e[row], e[col]), dim=1))
    red = scatter_mean if aggr == "mean" else scatter_sum
    agg = red(a, col, e.size(0))
    e = mlp(p, f"{name}.edge_mlp", torch.cat((agg, e), dim=1))
    return e, a


def down_edge_mp(p: Params, name: str, e1, e2, a12, angle_index12):
    """DownEdgeMP.forward (blocks.py:360-381)."""
    row, col = angle_index12[0], angle_index12[1]
    a12 = mlp(p, f"{name}.angle_mlp", torch.cat((a12, e1[row], e2[col]), dim=1))
    agg = scatter_mean(a12, col, e2.size(0))
    return mlp(p, f"{name}.edge_mlp", torch.cat((agg, e2), dim=1))


def edge_scalar_to_node_vector(edge_attr, unit_inverse):
    """edgeScalarToNodeVector (blocks.py:88-114) with the precomputed pseudo-inverse:
    per node  pinv(U)[2,k] @ e[k,F]  -> [2,F] -> interleaved (f0x,f0y,f1x,...)."""
    n, _, k = unit_inverse.shape
    v = unit_inverse @ edge_attr.view(n, k, edge_attr.size(1))
    return v.transpose(1, 2).flatten(1, 2)


def knn_interpolate(x, y_idx, x_idx, weights):
    """knn_interpolate (blocks.py:34-48)."""
    ny = int(y_idx.max()) + 1
    num = scatter_sum(x[x_idx] * weights, y_idx, ny)
    den = scatter_sum(weights, y_idx, ny)
    return num / den


def project_on_edges(node_vec, col, unit):
    """(v[col].view(E,-1,2) * U.unsqueeze(1)).sum(-1)  (remus_gnn.py:124-126, blocks.py:453-454)."""
    return (node_vec[col].reshape(col.size(0), -1, 2) * unit.unsqueeze(1)).sum(dim=-1)


def up_edge_mp(p: Params, name: str, total_num_nodes, y_idx, x_idx, weights, e2, unit_inverse2,
               e1, col1, unit1, coarse_mask1=None):
    """UpEdgeMP.forward (blocks.py:408-456)."""
    v2 = edge_scalar_to_node_vector(e2, unit_inverse2)
    v1 = torch.zeros(total_num_nodes, 2 * e2.size(1))
    interp = knn_interpolate(v2, y_idx, x_idx, weights)
    if coarse_mask1 is None:
        v1[:] = interp
    else:
        v1[coarse_mask1] = interp
    proj = project_on_edges(v1, col1, unit1)
    return mlp(p, f"{name}.up_mlp", torch.cat([proj, e1], dim=1))


# --------------------------------------------------------------------------- programs
def block_program(p: Params) -> List[Tuple[str, str]]:
    """Ordered (name, kind) list of the model's blocks.  state_dict order is module
    registration order, which in every reference ``load_arch`` (e.g. nn/mus_gnn.py:274-310,
    nn/remus_gnn.py:75-117) is also the execution order."""
    names: List[str] = []
    for key in p:
        top = key.split(".")[0]
        if top not in names:
            names.append(top)
    remus = any(n.startswith("angle_encoder") for n in names)
    prog = []
    for n in names:
        sub = {key.split(".")[1] for key in p if key.startswith(n + ".")}
        if "MLP" in sub:
            kind = "mlp"
        elif {"edge_mlp", "node_mlp"} <= sub:
            kind = "mp"
        elif "down_mlp" in sub:
            kind = "down"
        elif "up_mlp" in sub:
            kind = "up_edge" if remus else "up"
        elif {"angle_mlp", "edge_mlp"} <= sub:
            kind = "down_edge" if n.startswith("down") else "edge_mp"
        else:
            raise ValueError(f"unrecognised block {n}: {sorted(sub)}")
        prog.append((n, kind))
    return prog


def mus_forward(p: Params, g) -> torch.Tensor:
    """One time step of any MuS-GNN (nn/mus_gnn.py:68-97,173-218,312-373,485-562,615-636,
    704-741,827-880,984-1053): the fixed block sequence with F.selu after each encoder/MP,
    tanh on pool/unpool, and the residual update of the newest time slice."""
    prog = block_program(p)
    node_in = torch.cat([getattr(g, a) for a in ("field", "loc", "glob", "omega") if hasattr(g, a)], dim=1)
    e = F.selu(mlp(p, "edge_encoder", g.edge_attr))
    v = F.selu(mlp(p, "node_encoder", node_in))
    edge_index = g.edge_index
    saved = {}
    level = 1
    body = [(n, k) for n, k in prog if k != "mlp"]
    for i, (name, kind) in enumerate(body):
        nxt = body[i + 1][1] if i + 1 < len(body) else "decoder"
        if kind == "mp":
            v, e_new = gn_block(p, name, v, e, edge_index)
            v = F.selu(v)
            if nxt in ("up", "decoder"):
                e = None          # edge output discarded (e.g. nn/mus_gnn.py:346,354,366)
            else:
                e = F.selu(e_new)
        elif kind == "down":
            saved[level] = (v, edge_index, e)
            idx = getattr(g, f"idx{level}_to_idx{level + 1}")
            v, edge_index, e = down_mp(p, name, v, getattr(g, f"e_{level}{level + 1}"), idx, edge_index, e)
            level += 1
        elif kind == "up":
            v_old, edge_index_h, e_h = saved[level - 1]
            idx = getattr(g, f"idx{level - 1}_to_idx{level}")
            v = up_mp(p, name, v, v_old, getattr(g, f"e_{level - 1}{level}"), idx)
            edge_index, e = edge_index_h, e_h
            level -= 1
        else:
            raise ValueError(kind)
    out = mlp(p, "node_decoder", v)
    nf = out.size(1)
    return g.field[:, -nf:] + out


def remus_forward(p: Params, g) -> torch.Tensor:
    """One time step of NsRotEquiTreeScaleGNN (nn/remus_gnn.py:119-199)."""
    sfx = {1: "", 2: "2", 3: "3"}
    col = {l: getattr(g, "edge_index" + sfx[l])[1] for l in (1, 2, 3)}
    e = {}
    for l in (1, 2, 3):
        proj = project_on_edges(g.field, col[l], getattr(g, "edgeUnitVector" + sfx[l]))
        x = torch.cat([proj, g.glob[col[l]], g.omega[col[l]]], dim=1)
        e[l] = F.selu(mlp(p, "edge_encoder" + sfx[l], x))
    a = {l: F.selu(mlp(p, "angle_encoder" + sfx[l], getattr(g, "angle_attr" + sfx[l]))) for l in (1, 2, 3)}
    a12 = F.selu(mlp(p, "angle_encoder12", g.angle_attr12))
    a23 = F.selu(mlp(p, "angle_encoder23", g.angle_attr23))
    aidx = {l: getattr(g, "angle_index" + sfx[l]) for l in (1, 2, 3)}

    def run(names, l, last_discards):
        for i, n in enumerate(names):
            e[l], a_new = edge_mp(p, n, e[l], a[l], aidx[l])
            e[l] = F.selu(e[l])
            if not (last_discards and i == len(names) - 1):
                a[l] = F.selu(a_new)

    run(["mp111", "mp112", "mp113", "mp114"], 1, False)
    e[2] = F.selu(down_edge_mp(p, "down_mp12", e[1], e[2], a12, g.angle_index12))
    run(["mp211", "mp212"], 2, False)
    e[3] = F.selu(down_edge_mp(p, "down_mp23", e[2], e[3], a23, g.angle_index23))
    run(["mp31", "mp32", "mp33", "mp34"], 3, True)
    n_total = g.pos.size(0)
    e[2] = F.selu(up_edge_mp(p, "up_mp32", n_total, g.y_idx_32, g.x_idx_32, g.weights_32, e[3],
                             g.edgeUnitVectorInverse3, e[2], col[2], g.edgeUnitVector2, g.coarse_mask2))
    run(["mp221", "mp222"], 2, True)
    e[1] = F.selu(up_edge_mp(p, "up_mp21", n_total, g.y_idx_21, g.x_idx_21, g.weights_21, e[2],
                             g.edgeUnitVectorInverse2, e[1], col[1], g.edgeUnitVector))
    run(["mp121", "mp122", "mp123", "mp124"], 1, True)
    dec = mlp(p, "edge_decoder", e[1])
    out = edge_scalar_to_node_vector(dec, g.edgeUnitVectorInverse)
    return g.field[:, -2:] + out


def mugs_forward(p: Params, g) -> torch.Tensor:
    """One time step of NsTwoGuillardScaleGNN / NsThreeGuillardScaleGNN / NsFourGuillardScaleGNN
    (nn/mugs_gnn.py:81-133, 219-295, 394-489): blocks named mp<level>..., ``restriction`` (blocks.py:9-32) when the level
    grows, ``knn_interpolate`` (blocks.py:34-48) + cat with the skipped features when it shrinks."""
    body = [n for n, k in block_program(p) if k == "mp"]
    level_of = lambda n: int(n[2])
    n_levels = max(level_of(n) for n in body)
    sfx = lambda l: "" if l == 1 else str(l)
    num_nodes = g.pos.size(0)
    node_in = torch.cat([getattr(g, a) for a in ("field", "loc", "glob", "omega") if hasattr(g, a)], dim=1)
    e_enc = {l: F.selu(mlp(p, "edge_encoder" + sfx(l), getattr(g, "edge_attr" + sfx(l)))) for l in range(1, n_levels + 1)}
    v = F.selu(mlp(p, "node_encoder", node_in))
    masks = {1: torch.ones(num_nodes, dtype=torch.bool)}
    for l in range(2, n_levels + 1):
        masks[l] = getattr(g, f"coarse_mask{l}")
    level, e, edge_index = 1, e_enc[1], g.edge_index
    saved = {}
    for i, name in enumerate(body):
        l = level_of(name)
        if l == level + 1:
            saved[level] = (v, edge_index, e)
            v = v[masks[l][masks[level]]]                       # nn/mugs_gnn.py:104, 117-118
            mask2idx = -torch.ones(num_nodes, dtype=torch.long)
            mask2idx[masks[l]] = torch.arange(v.size(0))
            edge_index, e, level = mask2idx[getattr(g, f"edge_index{l}")], e_enc[l], l
        elif l == level - 1:
            tag = f"{level}{l}"
            up = knn_interpolate(v, getattr(g, "y_idx_" + tag), getattr(g, "x_idx_" + tag), getattr(g, "weights_" + tag))
            v_old, edge_index, e = saved.pop(l)
            v, level = torch.cat([up, v_old], dim=1), l
        v, e_new = gn_block(p, name, v, e, edge_index)
        v = F.selu(v)
        nxt = level_of(body[i + 1]) if i + 1 < len(body) else 0
        e = F.selu(e_new) if nxt >= level else None             # discarded edge output (e.g. nn/mugs_gnn.py:116, 127)
    out = mlp(p, "node_decoder", v)
    nf = out.size(1)
    return g.field[:, -nf:] + out


def forward(p: Params, g) -> torch.Tensor:
    if any(k.startswith("angle_encoder") for k in p):
        return remus_forward(p, g)
    if "edge_encoder2.MLP.linear_1.weight" in p and not any(k.startswith("down_mp") for k in p):
        return mugs_forward(p, g)
    return mus_forward(p, g)


def solve(p: Params, g, n_out: int) -> torch.Tensor:
    """GNN.solve + shift_and_replace (nn/model.py:303-327): rollout, output [N, nf*n_out]."""
    assert n_out > 0
    field0 = g.field
    outs = []
    with torch.no_grad():
        for t in range(n_out):
            pred = forward(p, g)
            outs.append(pred)
            if t + 1 < n_out:
                nf = pred.size(1)
                g.field = torch.cat([g.field[:, nf:], pred], dim=1)
    g.field = field0
    return torch.cat(outs, dim=1)
