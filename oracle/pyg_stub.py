"""TEST INFRASTRUCTURE — not part of the product path.

Minimal stand-ins for the reference's un-vendored third-party imports so that
``/root/reference/graphs4cfd`` can be imported *verbatim* in this container
(used only by ``oracle/make_golden.py`` and the ``reference``-marked tests to
pin ``oracle/restate.py``).  Nothing under ``graphs4cfd_b200/`` imports this.

What is restated (dependency: ``torch_geometric`` — unpinned in the reference's
pyproject.toml:35, any release with ``torch_geometric.utils.scatter``, i.e.
>= 2.3; ``torch_cluster`` — unpinned git master, pyproject.toml:37):

* ``torch_geometric.utils.scatter``  (call sites: graphs4cfd/nn/blocks.py:46,47,183,231,330,378)
    sum  : ``new_zeros(size).scatter_add_(dim, broadcast(index), src)``
    mean : the sum divided by ``scatter_add_`` of ones clamped to min 1
    ``dim_size`` defaults to ``index.max()+1``
* ``torch_geometric.utils.remove_self_loops`` (blocks.py:65): mask ``row != col``
* ``torch_geometric.utils.coalesce`` (blocks.py:67): stable sort of
  ``row*num_nodes+col`` (output sorted by source), duplicates reduced with
  ``scatter(reduce)``
* ``torch_geometric.nn.knn_graph / knn / voxel_grid`` (pre-processing only:
  transforms/connect.py:58, interpolate.py:125, mus.py:25) through scipy cKDTree
  and integer arithmetic
* ``torch_geometric.data.Data / Dataset``: attribute bags; ``Batch.from_data_list`` with Data's default collation rules
  (loader.py:53: 'index' attributes are concatenated along the last dimension and offset by the node counts)
* ``matplotlib``, ``h5py``, ``torch_geometric.loader``: empty placeholders (never called
  on the hot path)
"""
import sys
import types

import numpy as np
import torch


# --------------------------------------------------------------------- utils
def _broadcast(index, src, dim):
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter(src, index, dim=0, dim_size=None, reduce='sum'):
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim] = dim_size
    if reduce in ('sum', 'add'):
        return src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
    if reduce == 'mean':
        count = src.new_zeros(dim_size)
        count.scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
        shape = [1] * out.dim()
        shape[dim] = -1
        return out / count.view(shape)
    raise ValueError(f"stub scatter: reduce={reduce!r} is not used by graphs4cfd")


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[mask]


def coalesce(edge_index, edge_attr='???', num_nodes=None, reduce='sum',
             is_sorted=False, sort_by_row=True):
    nnz = edge_index.size(1)
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if nnz > 0 else 0
    idx = edge_index.new_empty(nnz + 1)
    idx[0] = -1
    idx[1:] = edge_index[1 - int(sort_by_row)]
    idx[1:].mul_(num_nodes).add_(edge_index[int(sort_by_row)])
    has_attr = not isinstance(edge_attr, str)
    if not is_sorted:
        idx[1:], perm = torch.sort(idx[1:], stable=True)
        edge_index = edge_index[:, perm]
        if has_attr and edge_attr is not None:
            edge_attr = edge_attr[perm]
    mask = idx[1:] > idx[:-1]
    if bool(mask.all()):
        return (edge_index, edge_attr) if has_attr else edge_index
    edge_index = edge_index[:, mask]
    if not has_attr:
        return edge_index
    if edge_attr is None:
        return edge_index, None
    dim_size = edge_index.size(1)
    gid = torch.arange(0, nnz, device=edge_index.device)
    gid.sub_(mask.logical_not().cumsum(dim=0))
    return edge_index, scatter(edge_attr, gid, 0, dim_size, reduce)


# ---------------------------------------------------------------- cluster ops
def knn_graph(x, k, batch=None, loop=False, flow='source_to_target', **_):
    """edges (neighbour -> centre), grouped by centre ascending, k per centre."""
    from scipy.spatial import cKDTree
    pts = x.detach().cpu().double().numpy()
    tree = cKDTree(pts)
    _, nbr = tree.query(pts, k=k + (0 if loop else 1))
    n = pts.shape[0]
    if not loop:
        # drop self (first hit unless exact duplicates reorder it)
        out = np.empty((n, k), dtype=np.int64)
        for i in range(n):
            row = nbr[i]
            keep = row[row != i]
            out[i] = keep[:k]
        nbr = out
    centre = np.repeat(np.arange(n, dtype=np.int64), k)
    src = nbr.reshape(-1)
    ei = np.stack([src, centre]) if flow == 'source_to_target' else np.stack([centre, src])
    return torch.from_numpy(ei).to(x.device)


def knn(x, y, k, batch_x=None, batch_y=None, **_):
    """for each y the k nearest x (of the same graph when batch vectors are given); returns [y_index; x_index], y ascending."""
    from scipy.spatial import cKDTree
    if batch_x is not None and batch_y is not None and int(batch_x.max()) > 0:
        parts = []
        for b in range(int(batch_x.max()) + 1):
            ix, iy = (batch_x == b).nonzero().squeeze(1), (batch_y == b).nonzero().squeeze(1)
            sub = knn(x[ix], y[iy], k)
            parts.append(torch.stack([iy[sub[0]], ix[sub[1]]]))
        return torch.cat(parts, dim=1)
    tree = cKDTree(x.detach().cpu().double().numpy())
    _, nbr = tree.query(y.detach().cpu().double().numpy(), k=k)
    nbr = np.asarray(nbr).reshape(y.size(0), k)
    yi = np.repeat(np.arange(y.size(0), dtype=np.int64), k)
    return torch.from_numpy(np.stack([yi, nbr.reshape(-1).astype(np.int64)])).to(x.device)


def voxel_grid(pos, size, batch=None, start=None, end=None):
    pos = pos.detach()
    if not torch.is_tensor(size):
        size = torch.tensor([float(size)] * pos.size(1), dtype=pos.dtype)
    size = size.to(pos.dtype).expand(pos.size(1)) if size.dim() == 0 else size.to(pos.dtype)
    if batch is not None:
        pos = torch.cat([pos, batch.to(pos.dtype).view(-1, 1)], dim=1)
        size = torch.cat([size, size.new_ones(1)])
    start = pos.min(dim=0).values if start is None else start
    end = pos.max(dim=0).values if end is None else end
    num_voxels = ((end - start) / size).to(torch.long) + 1
    stride = torch.cat([num_voxels.new_ones(1), num_voxels.cumprod(0)[:-1]])
    coord = ((pos - start) / size).to(torch.long)
    return (coord * stride).sum(dim=1)


# ------------------------------------------------------------------ containers
class Data:
    def __init__(self, **kwargs):
        for key, val in kwargs.items():
            setattr(self, key, val)

    @property
    def num_nodes(self):
        for key in ('pos', 'field', 'x'):
            t = self.__dict__.get(key)
            if torch.is_tensor(t):
                return t.size(0)
        return None

    @property
    def num_edges(self):
        ei = self.__dict__.get('edge_index')
        return ei.size(1) if torch.is_tensor(ei) else 0

    def keys(self):
        return [k for k in self.__dict__]

    def to(self, device, *args, **kwargs):
        for key, val in list(self.__dict__.items()):
            if torch.is_tensor(val):
                setattr(self, key, val.to(device))
        return self

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for key, val in self.__dict__.items():
            out.__dict__[key] = val.clone() if torch.is_tensor(val) else val
        return out

    def __contains__(self, key):
        return key in self.__dict__


class Batch(Data):
    """``torch_geometric.data.Batch.from_data_list`` with ``Data``'s DEFAULT collation rules (the reference's Graph class does
    not override them, graph.py:6-10): a tensor attribute whose name contains 'index' (or is 'face') is concatenated along its
    LAST dimension with every graph's values incremented by the number of nodes of the graphs before it
    (``Data.__cat_dim__`` / ``Data.__inc__``); every other tensor is concatenated along dimension 0 unchanged; ``batch`` [N] holds
    the graph id of every node and ``ptr`` the node offsets."""

    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        keys = [k for k in data_list[0].__dict__ if torch.is_tensor(data_list[0].__dict__[k])]
        offsets = [0]
        for d in data_list:
            offsets.append(offsets[-1] + d.num_nodes)
        for key in keys:
            vals = [d.__dict__[key] for d in data_list]
            if 'index' in key or key == 'face':
                setattr(out, key, torch.cat([v + off for v, off in zip(vals, offsets)], dim=-1))
            elif vals[0].dim() == 0:
                setattr(out, key, torch.stack(vals))
            else:
                setattr(out, key, torch.cat(vals, dim=0))
        out.batch = torch.cat([torch.full((d.num_nodes,), i, dtype=torch.long) for i, d in enumerate(data_list)])
        out.ptr = torch.tensor(offsets, dtype=torch.long)
        return out


class Dataset:
    def __init__(self, *a, **k):
        pass


def install():
    """Register the stand-ins in sys.modules (idempotent)."""
    if 'torch_geometric' in sys.modules and getattr(sys.modules['torch_geometric'], '_g4c_stub', False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    tg = mod('torch_geometric', _g4c_stub=True)
    tg.utils = mod('torch_geometric.utils', scatter=scatter, remove_self_loops=remove_self_loops,
                   coalesce=coalesce)
    tg.data = mod('torch_geometric.data', Data=Data, Batch=Batch, Dataset=Dataset)
    tg.nn = mod('torch_geometric.nn', knn_graph=knn_graph, knn=knn, voxel_grid=voxel_grid)
    tg.loader = mod('torch_geometric.loader', DataLoader=object)
    tg.transforms = mod('torch_geometric.transforms', Compose=object)
    mod('torch_cluster', knn_graph=knn_graph, knn=knn)
    for name in ('h5py',):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name, File=object)
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
        import matplotlib.tri  # noqa: F401
    except Exception:
        mpl = mod('matplotlib')
        mpl.pyplot = mod('matplotlib.pyplot')
        mpl.tri = mod('matplotlib.tri', Triangulation=object)
        mpl.collections = mod('matplotlib.collections', LineCollection=object)
        mpl.colors = mod('matplotlib.colors')
        mpl.cm = mod('matplotlib.cm')


REFERENCE_ROOT = '/root/reference'
# copy staged by tools/stage_reference.py (git-ignored; travels to the GPU box with the repo snapshot)
STAGED_ROOT = __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))),
                                         'baseline', '_ref')


def reference_root():
    """Directory holding the unmodified reference package: the read-only tree in the build container, else the staged copy."""
    import os
    for root in (REFERENCE_ROOT, STAGED_ROOT):
        if os.path.isdir(os.path.join(root, 'graphs4cfd')):
            return root
    return None


def staged_checkpoint(name):
    """Path of a weights-only copy of a shipped checkpoint (tools/stage_reference.py), or None."""
    import os
    p = os.path.join(STAGED_ROOT, 'weights', name)
    return p if os.path.exists(p) else None


def import_reference():
    """Import the UNMODIFIED reference package under the stub: from /root/reference in the build container, from the
    staged byte-for-byte copy (baseline/_ref, see tools/stage_reference.py) elsewhere."""
    root = reference_root()
    if root is None:
        raise ImportError("no reference tree: neither /root/reference nor baseline/_ref (python tools/stage_reference.py) is present")
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    import graphs4cfd  # noqa: E402
    return graphs4cfd
