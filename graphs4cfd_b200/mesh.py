"""Synthetic CFD meshes and the static per-mesh attributes the hot path reads.

The reference builds these attributes with its transforms (datasets/transforms stay
in the reference; they are OUT of the replaced path).  The benchmark and the tests
still need 10k ... 4M-node inputs in exactly the reference's layouts, and the
reference versions are O(E^2) Python loops (transforms/remus.py:36,159-161) or need
torch_cluster, so the layouts are rebuilt here in vectorised numpy:

* ``knn_edges``          = transforms/connect.py:9-71  (non-periodic branch): edges
                           (neighbour -> centre), grouped by centre, k per centre
* ``grid_clustering``    = transforms/mus.py:9-37
* ``guillard_coarsening``= transforms/mugs.py:8-29
* ``extend_graph``       = transforms/remus.py:9-44    (closed form of angle_index)
* ``angle_index_down``   = transforms/remus.py:151-176
* ``knn_interp_weights`` = transforms/interpolate.py:110-129

``tests/test_oracle_vs_reference.py`` (``test_mus_layouts_match_reference_transforms``,
``test_remus_layouts_match_reference_transforms``) checks them against the reference's own
transforms on the same points wherever the reference tree or its staged copy exists.
"""
import math
import os
from typing import Sequence

import numpy as np
import torch


class Mesh:
    """Attribute bag with the slice of ``torch_geometric.data.Data`` that
    ``GNN.solve`` (nn/model.py:303-321) touches: ``num_nodes``, ``to`` and plain attributes."""

    def __init__(self, **kwargs):
        for key, val in kwargs.items():
            setattr(self, key, val)

    @property
    def num_nodes(self):
        return self.pos.size(0)

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def to(self, device, *args, **kwargs):
        for key, val in list(self.__dict__.items()):
            if torch.is_tensor(val):
                setattr(self, key, val.to(device))
        return self

    def clone(self):
        out = Mesh()
        for key, val in self.__dict__.items():
            out.__dict__[key] = val.clone() if torch.is_tensor(val) else val
        return out

    def keys(self):
        return list(self.__dict__)


# --------------------------------------------------------------------------- plan-time spatial renumbering
def morton_order(pos: torch.Tensor, bits: int = 24) -> np.ndarray:
    """Permutation that sorts 2-D points along the Z-order (Morton) curve: perm[i] = index of the point that comes i-th.
    The hot path gathers P_r[src] rows for every edge; kNN sources are near in SPACE, so a space-filling order makes them
    near in MEMORY whatever order the caller's mesh came in (hard part 4).  Keys are 2 x `bits` bits of the
    positions quantised on the bounding box; ties (coincident points) keep their given order."""
    p = pos[:, :2].detach().double().cpu().numpy()
    lo = p.min(axis=0)
    span = max(float((p.max(axis=0) - lo).max()), 1e-30)
    q = np.minimum((p - lo) / span * float(2 ** bits - 1), float(2 ** bits - 1)).astype(np.uint64)

    def spread(x):                       # abcd -> 0a0b0c0d
        x = (x | (x << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        x = (x | (x << np.uint64(2))) & np.uint64(0x3333333333333333)
        x = (x | (x << np.uint64(1))) & np.uint64(0x5555555555555555)
        return x

    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    return np.argsort(key, kind="stable")


def permute_mus_nodes(g: "Mesh", perm) -> "Mesh":
    """The same MuS mesh with its level-1 nodes renumbered: new node i = old node perm[i].  Every per-node attribute follows
    its node, edges are regrouped by their new target keeping each target's in-edges in their given relative order (so a
    node's aggregation order, hence its arithmetic, is unchanged), coarser levels keep their numbering."""
    perm = torch.as_tensor(np.asarray(perm), dtype=torch.long)
    n = int(g.pos.shape[0])
    inv = torch.empty(n, dtype=torch.long)
    inv[perm] = torch.arange(n)
    out = Mesh()
    ei = g.edge_index
    order = torch.sort(inv[ei[1].cpu()], stable=True).indices
    # attributes indexed by level-1 node (transforms/mus.py:9-37 stores the level-1 -> level-2 maps per FINE node)
    level1 = {"pos", "field", "glob", "omega", "loc", "target", "bound", "batch", "cluster_2", "idx1_to_idx2", "e_12"}
    for key, val in g.__dict__.items():
        if not torch.is_tensor(val):
            out.__dict__[key] = val
        elif key == "edge_index":
            out.edge_index = inv[ei.cpu()[:, order]].to(ei.device)
        elif key == "edge_attr":
            out.edge_attr = val[order.to(val.device)]
        elif key in level1:
            assert val.shape[0] == n, key
            out.__dict__[key] = val[perm.to(val.device)]
        else:
            out.__dict__[key] = val
    return out


# --------------------------------------------------------------------------- points
def jittered_points(n: int, seed: int = 0, box=(4.0, 1.0), jitter: float = 0.35) -> torch.Tensor:
    """n points on a jittered lattice in [0,box_x]x[0,box_y], numbered column by column
    (x-major strips), so a contiguous node range is a vertical strip of the domain."""
    rng = np.random.default_rng(seed)
    ny = max(1, int(round(math.sqrt(n * box[1] / box[0]))))
    nx = (n + ny - 1) // ny
    h = box[1] / ny
    ix, iy = np.divmod(np.arange(n, dtype=np.int64), ny)
    pos = np.stack([(ix + 0.5) * (box[0] / nx), (iy + 0.5) * h], axis=1)
    pos += rng.uniform(-jitter, jitter, size=pos.shape) * np.array([box[0] / nx, h])
    return torch.from_numpy(pos.astype(np.float32))


def uniform_points(n: int, seed: int = 0, box=(1.0, 1.0)) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    return torch.from_numpy((rng.uniform(0, 1, size=(n, 2)) * np.array(box)).astype(np.float32))


# ----------------------------------------------------------------------------- kNN
def knn_edges(pos: torch.Tensor, k: int):
    """(edge_index int64[2, N*k], edge_attr fp32[N*k, 2]); edge j*k+m points from the
    m-th nearest neighbour of node j to node j (all in-edges of a node are contiguous).
    CUDA positions: the search runs on the device (g4c_plan_knn); host positions: scipy's k-d tree."""
    if pos.is_cuda:
        from . import ops
        n = pos.shape[0]
        nbr = ops.knn(pos.contiguous(), None, k)
        centre = torch.arange(n, device=pos.device).repeat_interleave(k)
        edge_index = torch.stack([nbr.reshape(-1), centre])
        return edge_index, pos[edge_index[1]] - pos[edge_index[0]]
    from scipy.spatial import cKDTree
    pts = pos.double().numpy()
    n = pts.shape[0]
    _, nbr = cKDTree(pts).query(pts, k=k + 1, workers=-1)
    nbr = nbr.astype(np.int64)
    self_col = nbr == np.arange(n)[:, None]
    # drop the node itself (normally column 0; with duplicated points fall back to the last hit)
    none = ~self_col.any(axis=1)
    self_col[none, k] = True
    nbr = nbr[~self_col].reshape(n, k)
    centre = np.repeat(np.arange(n, dtype=np.int64), k)
    edge_index = torch.from_numpy(np.stack([nbr.reshape(-1), centre]))
    edge_attr = pos[edge_index[1]] - pos[edge_index[0]]
    return edge_index, edge_attr


def knn_interp_weights(pos_x: torch.Tensor, pos_y: torch.Tensor, k: int):
    """for every y its k nearest x: (y_idx, x_idx, 1/max(d^2,1e-16))."""
    if pos_x.is_cuda:
        from . import ops
        x_idx = ops.knn(pos_x.contiguous(), pos_y.contiguous(), k).reshape(-1)
        y_idx = torch.arange(pos_y.size(0), device=pos_y.device).repeat_interleave(k)
        diff = pos_x[x_idx] - pos_y[y_idx]
        return y_idx, x_idx, 1.0 / torch.clamp((diff * diff).sum(dim=-1, keepdim=True), min=1e-16)
    from scipy.spatial import cKDTree
    _, nbr = cKDTree(pos_x.double().numpy()).query(pos_y.double().numpy(), k=k, workers=-1)
    nbr = np.asarray(nbr, dtype=np.int64).reshape(pos_y.size(0), k)
    y_idx = torch.from_numpy(np.repeat(np.arange(pos_y.size(0), dtype=np.int64), k))
    x_idx = torch.from_numpy(nbr.reshape(-1))
    diff = pos_x[x_idx] - pos_y[y_idx]
    weights = 1.0 / torch.clamp((diff * diff).sum(dim=-1, keepdim=True), min=1e-16)
    return y_idx, x_idx, weights


# ------------------------------------------------------------------ MuS (grid) levels
def grid_clustering(pos_1: torch.Tensor, cell_size: float):
    """(pos_2, cluster_2, mask_2, idx1_to_idx2, e_12) with the reference's numbering:
    coarse nodes are the non-empty cells in ascending cell id (x fastest)."""
    p = pos_1
    start = p.min(dim=0).values
    end = p.max(dim=0).values
    size = torch.tensor([cell_size, cell_size], dtype=p.dtype, device=p.device)
    num_voxels = ((end - start) / size).to(torch.long) + 1
    coord = ((p - start) / size).to(torch.long)
    cluster = coord[:, 0] + coord[:, 1] * num_voxels[0]
    mask, inverse = torch.unique(cluster, sorted=True, return_inverse=True)
    n2 = mask.numel()
    summed = torch.zeros(n2, 2, dtype=p.dtype, device=p.device).index_add_(0, inverse, p)
    count = torch.zeros(n2, dtype=p.dtype, device=p.device).index_add_(0, inverse, torch.ones(p.size(0), dtype=p.dtype, device=p.device))
    pos_2 = summed / count.clamp(min=1).unsqueeze(1)
    e_12 = (pos_2[inverse] - p) / cell_size
    return pos_2, cluster, mask, inverse, e_12


# -------------------------------------------------------------- REMuS (Guillard) levels
def guillard_coarsening(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """Sequential node-nested coarsening: visiting nodes in order, a node still marked
    coarse removes its k senders from the coarse set."""
    k = int((edge_index[1] == 0).sum())
    dev = edge_index.device                    # the sweep is sequential: it runs on the host whatever the device of the mesh
    senders = edge_index[0].view(-1, k).cpu().numpy()
    # Host-side work either way (mesh synthesis for tests and the benchmark, not the product path): the C helper of libg4c
    # when the library has been built, the same sweep in Python when it has not (e.g. oracle-only test runs).
    from . import _lib
    if os.path.exists(_lib.LIB_PATH):
        return torch.from_numpy(_lib.host_guillard(senders, int(num_nodes))).to(dev)
    coarse = np.ones(int(num_nodes), dtype=bool)
    for i in range(senders.shape[0]):
        if coarse[i]:
            coarse[senders[i]] = False
    return torch.from_numpy(coarse).to(dev)


def _local_index(edge_index: torch.Tensor):
    """Level-l edge_index is stored in LEVEL-1 node numbering (transforms/remus.py:120-122);
    map it back to 0..V_l-1 through the sorted target ids (targets are grouped ascending)."""
    k = int((edge_index[1] == edge_index[1][0]).sum())
    owners = edge_index[1].view(-1, k)[:, 0].contiguous()
    local = torch.searchsorted(owners, edge_index.reshape(-1)).view(2, -1)
    return local, owners, k


def extend_graph(edge_index: torch.Tensor, edge_attr: torch.Tensor, k: int):
    """(edgeUnitVector[E,2], angle_index[2,kE], angle_attr[kE,4]).  Angle j*k+m goes from
    the m-th in-edge of the SOURCE node of edge j into edge j: row = src_local(j)*k+m, col = j."""
    num_edges = edge_index.size(1)
    local, _, _ = _local_index(edge_index)
    size = edge_attr.norm(2, dim=1, keepdim=True)
    unit = edge_attr / size
    dev = edge_index.device
    row = (local[0].unsqueeze(1) * k + torch.arange(k, device=dev)).reshape(-1)
    col = torch.arange(num_edges, device=dev).repeat_interleave(k)
    cos = (unit[row] * unit[col]).sum(dim=1)
    sin = unit[row, 0] * unit[col, 1] - unit[row, 1] * unit[col, 0]
    angle_attr = torch.cat([size[row], size[col], cos.unsqueeze(1), sin.unsqueeze(1)], dim=1)
    return unit, torch.stack([row, col]), angle_attr


def angle_index_down(edge_index1, edge_attr1, edge_index2, edge_attr2, coarse_index2, k):
    """Inter-level angles: for every coarse node j (ascending) and every level-2 edge (j->q)
    leaving it (ascending edge id), the k level-1 in-edges of j are the senders."""
    local1, owners1, _ = _local_index(edge_index1)
    # position of each coarse node among the level-1 targets -> its k in-edges
    pos_in_l1 = torch.searchsorted(owners1, coarse_index2)
    in_edges = pos_in_l1.unsqueeze(1) * k + torch.arange(k, device=edge_index1.device)                      # [V2, k]
    # level-2 edges grouped by source node, ascending node then ascending edge id
    src2 = torch.searchsorted(coarse_index2, edge_index2[0])
    order = torch.sort(src2, stable=True).indices                                # out_edges_index2
    row = in_edges[src2[order]].reshape(-1)
    col = order.repeat_interleave(k)
    size1 = edge_attr1.norm(2, dim=1, keepdim=True)
    size2 = edge_attr2.norm(2, dim=1, keepdim=True)
    u1, u2 = edge_attr1 / size1, edge_attr2 / size2
    cos = (u1[row] * u2[col]).sum(dim=1)
    sin = u1[row, 0] * u2[col, 1] - u1[row, 1] * u2[col, 0]
    attr = torch.cat([size1[row], size2[col], cos.unsqueeze(1), sin.unsqueeze(1)], dim=1)
    return torch.stack([row, col]), attr


# ------------------------------------------------------------------- synthetic fields
def _fields(pos: torch.Tensor, num_fields: int, n_in: int = 1):
    x, y = pos[:, 0], pos[:, 1]
    base = [0.5 + 0.3 * torch.sin(3 * x) * torch.cos(5 * y),
            0.2 * torch.cos(2 * x) * torch.sin(4 * y),
            -0.1 + 0.4 * torch.sin(x + 2 * y)]
    cols = []
    for t in range(n_in):
        for f in range(num_fields):
            cols.append(base[f % 3] * (1.0 - 0.02 * t))
    return torch.stack(cols, dim=1).contiguous()


def _omega(pos: torch.Tensor):
    x, y = pos[:, 0], pos[:, 1]
    box_x = float(x.max())
    om = ((x < 0.01 * box_x) | (((x - 0.25 * box_x) ** 2 + (y - 0.5) ** 2) < 0.01)).float()
    return om.unsqueeze(1)


def build_mus_mesh(n: int, k: int = 6, cells: Sequence[float] = (), seed: int = 0,
                   points: str = "jittered", num_fields: int = 3, edge_scale=None,
                   glob: float = 0.3, device=None) -> Mesh:
    """MuS-GNN input: level-1 kNN graph + ``len(cells)`` grid-clustered levels.
    ``cells`` in units of the mean edge length when given as ("auto", ratio...) is not
    supported; pass absolute sizes (see ``auto_cells``).
    ``device``: build on that CUDA device (kNN by g4c_plan_knn, everything else torch device ops); the same points give
    the same index arrays as the host build (tests/test_gpu_plan.py)."""
    pos = jittered_points(n, seed) if points == "jittered" else uniform_points(n, seed)
    if device is not None:
        pos = pos.to(device)
    edge_index, edge_attr = knn_edges(pos, k)
    r = float(edge_attr.norm(dim=1).mean()) if edge_scale is None else edge_scale
    edge_attr = edge_attr / (2 * r)
    m = Mesh(pos=pos, edge_index=edge_index, edge_attr=edge_attr,
             field=_fields(pos, num_fields), glob=torch.full((n, 1), glob, device=pos.device), omega=_omega(pos))
    p = pos
    for lvl, cell in enumerate(cells, start=2):
        pos_l, cluster, mask, idx, e = grid_clustering(p, cell)
        setattr(m, f'pos_{lvl}', pos_l)
        setattr(m, f'cluster_{lvl}', cluster)
        setattr(m, f'mask_{lvl}', mask)
        setattr(m, f'idx{lvl - 1}_to_idx{lvl}', idx)
        setattr(m, f'e_{lvl - 1}{lvl}', e)
        p = pos_l
    return m


def build_mugs_mesh(n: int, k: int = 6, levels: int = 2, seed: int = 0, points: str = "jittered", interp_k: int = None,
                    edge_scale=None, device=None) -> Mesh:
    """MuGS-GNN input with ``levels`` (2..4) node-nested levels in the layouts of GuillardCoarseningAndConnectKNN
    (transforms/mugs.py:58-88: level-l edges in LEVEL-1 node ids, coarse masks over the level-1 nodes) and BuildKnnInterpWeights
    (transforms/interpolate.py:147-155).  ``edge_scale``: per-level characteristic edge length (default: the level's mean)."""
    assert 2 <= levels <= 4
    interp_k = k if interp_k is None else interp_k
    edge_scale = (None,) * levels if edge_scale is None else edge_scale
    pos = jittered_points(n, seed) if points == "jittered" else uniform_points(n, seed)
    if device is not None:
        pos = pos.to(device)
    m = Mesh(pos=pos, field=_fields(pos, 3), glob=torch.full((n, 1), 0.3, device=pos.device), omega=_omega(pos))

    def scaled(ei, ea, s):
        r = float(ea.norm(dim=1).mean()) if s is None else s
        return ei, ea / (2 * r)

    m.edge_index, m.edge_attr = scaled(*knn_edges(pos, k), edge_scale[0])
    ei_l, mask_prev, idx_prev = m.edge_index, None, torch.arange(n, device=pos.device)
    for l in range(2, levels + 1):
        keep = guillard_coarsening(ei_l, idx_prev.numel())              # over the nodes of level l-1
        mask = torch.zeros(n, dtype=torch.bool, device=pos.device)
        mask[idx_prev[keep]] = True
        idx = mask.nonzero().squeeze(1)                                 # level-1 ids of the level-l nodes
        ei_l, ea = scaled(*knn_edges(pos[idx], k), edge_scale[l - 1])
        setattr(m, f"coarse_mask{l}", mask)
        setattr(m, f"edge_index{l}", idx[ei_l])
        setattr(m, f"edge_attr{l}", ea)
        y, x, w = knn_interp_weights(pos[idx], pos[idx_prev], interp_k)
        setattr(m, f"y_idx_{l}{l - 1}", y)
        setattr(m, f"x_idx_{l}{l - 1}", x)
        setattr(m, f"weights_{l}{l - 1}", w)
        idx_prev = idx
    return m


def collate(meshes: Sequence[Mesh], cells: Sequence[float] = (), interp_k: int = None) -> Mesh:
    """Several meshes as ONE input, the way the reference's loader batches graphs (loader.py:14-56): every tensor is concatenated
    along its first dimension, except attributes whose name contains 'index', which are concatenated along the last one and
    shifted — node lists (edge_index*) by the node count of the graphs before, the REMuS angle lists by the EDGE count of their
    level (angle_index<l>: level l; angle_index<l><l+1>: row 0 level l, row 1 level l+1; the loader's correction,
    loader.py:18-51).  ``batch`` [N] is the graph id of every node.
    Layouts that are built on the whole batch in the reference (batch-level transforms, loader.py:57-60) are rebuilt here too:
    ``cells`` -> the MuS grid clustering of the concatenated points (transforms/mus.py:9-37 ignores graph ids: nodes of different
    graphs in the same cell share a parent, exactly as in the reference); ``interp_k`` -> the interpolation lists between the
    Guillard levels, neighbours searched inside each graph (transforms/interpolate.py:147-155)."""
    first = meshes[0]
    keys = [k for k, v in first.__dict__.items() if torch.is_tensor(v)]
    n_off = [0]
    for m in meshes:
        n_off.append(n_off[-1] + m.pos.size(0))

    def edge_off(level):                     # edges of `level` in the graphs before each one
        name = "edge_index" + ("" if level == 1 else str(level))
        off = [0]
        for m in meshes:
            off.append(off[-1] + getattr(m, name).size(1))
        return off

    out = Mesh()
    for key in keys:
        vals = [getattr(m, key) for m in meshes]
        if key.startswith("angle_index"):
            lv = key[len("angle_index"):] or "1"
            rows = (int(lv[0]), int(lv[1])) if len(lv) == 2 else (int(lv), int(lv))
            o0, o1 = edge_off(rows[0]), edge_off(rows[1])
            vals = [torch.stack([v[0] + o0[i], v[1] + o1[i]]) for i, v in enumerate(vals)]
            setattr(out, key, torch.cat(vals, dim=1))
        elif "index" in key:
            setattr(out, key, torch.cat([v + n_off[i] for i, v in enumerate(vals)], dim=-1))
        else:
            setattr(out, key, torch.cat(vals, dim=0))
    out.batch = torch.cat([torch.full((m.pos.size(0),), i, dtype=torch.long, device=m.pos.device) for i, m in enumerate(meshes)])
    p = out.pos
    for lvl, cell in enumerate(cells, start=2):
        pos_l, cluster, mask, idx, e = grid_clustering(p, cell)
        for name, val in ((f"pos_{lvl}", pos_l), (f"cluster_{lvl}", cluster), (f"mask_{lvl}", mask),
                          (f"idx{lvl - 1}_to_idx{lvl}", idx), (f"e_{lvl - 1}{lvl}", e)):
            setattr(out, name, val)
        p = pos_l
    if interp_k is not None:
        lo_mask, l = torch.ones_like(out.batch, dtype=torch.bool), 2
        while hasattr(out, f"coarse_mask{l}"):
            hi_mask = getattr(out, f"coarse_mask{l}")
            ys, xs, ws = [], [], []
            for b in range(len(meshes)):     # neighbours inside each graph; indices are positions within the level's node list
                in_b = out.batch == b
                x_ids = (hi_mask & in_b)[hi_mask].nonzero().squeeze(1)
                y_ids = (lo_mask & in_b)[lo_mask].nonzero().squeeze(1)
                y, x, w = knn_interp_weights(out.pos[hi_mask & in_b], out.pos[lo_mask & in_b], interp_k)
                ys.append(y_ids[y]); xs.append(x_ids[x]); ws.append(w)
            setattr(out, f"y_idx_{l}{l - 1}", torch.cat(ys))
            setattr(out, f"x_idx_{l}{l - 1}", torch.cat(xs))
            setattr(out, f"weights_{l}{l - 1}", torch.cat(ws))
            lo_mask, l = hi_mask, l + 1
    return out


def auto_cells(n: int, levels: int, box=(4.0, 1.0), ratios=(5.0, 20.0, 80.0)):
    """cell sizes giving N_2 ~ N/5, N_3 ~ N/20, N_4 ~ N/80 (the reference example's ratios)."""
    area = box[0] * box[1]
    return [math.sqrt(area * ratios[i] / n) for i in range(levels - 1)]


def build_remus_mesh(n: int, k: int = 6, seed: int = 0, points: str = "jittered",
                     interp_k: int = None, edge_scale=(None, None, None), device=None) -> Mesh:
    """REMuS-GNN 3-level input in the layouts of transforms/remus.py:93-147 +
    transforms/interpolate.py:147-155.  ``device``: build on that CUDA device (kNN searches by g4c_plan_knn, the closed-form
    angle lists as torch device ops; only the sequential Guillard sweep visits the host)."""
    interp_k = k if interp_k is None else interp_k
    pos = jittered_points(n, seed) if points == "jittered" else uniform_points(n, seed)
    if device is not None:
        pos = pos.to(device)
    m = Mesh(pos=pos, field=_fields(pos, 2), glob=torch.full((n, 1), 0.3, device=pos.device), omega=_omega(pos))

    def scaled(ei, ea, s):
        r = float(ea.norm(dim=1).mean()) if s is None else s
        return ei, ea / (2 * r)

    m.edge_index, m.edge_attr = scaled(*knn_edges(pos, k), edge_scale[0])
    m.coarse_mask2 = guillard_coarsening(m.edge_index, n)
    ci2 = m.coarse_mask2.nonzero().squeeze(1)
    ei2, m.edge_attr2 = scaled(*knn_edges(pos[ci2], k), edge_scale[1])
    m.coarse_mask3 = torch.zeros_like(m.coarse_mask2)
    m.coarse_mask3[m.coarse_mask2] = guillard_coarsening(ei2, ci2.numel())
    ci3 = m.coarse_mask3.nonzero().squeeze(1)
    ei3, m.edge_attr3 = scaled(*knn_edges(pos[ci3], k), edge_scale[2])
    m.edge_index2, m.edge_index3 = ci2[ei2], ci3[ei3]
    for sfx, ei, ea, nl in (("", m.edge_index, m.edge_attr, n), ("2", m.edge_index2, m.edge_attr2, ci2.numel()),
                            ("3", m.edge_index3, m.edge_attr3, ci3.numel())):
        unit, ai, aa = extend_graph(ei, ea, k)
        setattr(m, "edgeUnitVector" + sfx, unit)
        setattr(m, "angle_index" + sfx, ai)
        setattr(m, "angle_attr" + sfx, aa)
        setattr(m, "edgeUnitVectorInverse" + sfx, torch.linalg.pinv(unit.view(nl, k, 2)))
    m.angle_index12, m.angle_attr12 = angle_index_down(m.edge_index, m.edge_attr, m.edge_index2, m.edge_attr2, ci2, k)
    m.angle_index23, m.angle_attr23 = angle_index_down(m.edge_index2, m.edge_attr2, m.edge_index3, m.edge_attr3, ci3, k)
    m.y_idx_21, m.x_idx_21, m.weights_21 = knn_interp_weights(pos[m.coarse_mask2], pos, interp_k)
    m.y_idx_32, m.x_idx_32, m.weights_32 = knn_interp_weights(pos[m.coarse_mask3], pos[m.coarse_mask2], interp_k)
    return m
