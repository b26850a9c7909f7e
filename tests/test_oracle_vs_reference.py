"""Live pin: restatement and mesh-layout builders against the reference's own classes / transforms imported verbatim
(from /root/reference in the build container, from the staged byte-for-byte copy baseline/_ref elsewhere)."""
import pytest
import torch

from conftest import rel_l2, shipped_model

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def gfd():
    from oracle.pyg_stub import import_reference
    return import_reference()


def test_shipped_3s_checkpoint_one_step(gfd):
    from graphs4cfd_b200 import mesh as M
    from oracle import restate as R
    model = shipped_model(gfd, "mus3")
    n = 2000
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=3)
    ref = model.solve(g.clone(), 2)
    out = R.solve({k: v.detach() for k, v in model.state_dict().items()}, g.clone(), 2)
    assert rel_l2(out, ref) <= 1e-6


def test_shipped_remus_checkpoint_one_step(gfd):
    from graphs4cfd_b200 import mesh as M
    from oracle import restate as R
    model = shipped_model(gfd, "remus")
    g = M.build_remus_mesh(400, 5, seed=5, points="uniform")
    ref = model.solve(g.clone(), 2)
    out = R.solve({k: v.detach() for k, v in model.state_dict().items()}, g.clone(), 2)
    assert rel_l2(out, ref) <= 1e-6


def test_mus_layouts_match_reference_transforms(gfd):
    from graphs4cfd_b200 import mesh as M
    n = 1500
    cells = M.auto_cells(n, 4)
    g = M.build_mus_mesh(n, 6, cells, seed=2)
    ei, ea = gfd.transforms.connect_knn(g.pos, 6)
    assert torch.equal(ei, g.edge_index)
    ref = gfd.transforms.GridClustering(cells)(M.Mesh(pos=g.pos.clone()))
    for lvl in (2, 3, 4):
        for name in (f"pos_{lvl}", f"cluster_{lvl}", f"mask_{lvl}", f"idx{lvl-1}_to_idx{lvl}", f"e_{lvl-1}{lvl}"):
            a, b = getattr(g, name), getattr(ref, name)
            assert torch.equal(a, b) if not a.is_floating_point() else torch.allclose(a, b, atol=1e-6), name


def test_remus_layouts_match_reference_transforms(gfd):
    from graphs4cfd_b200 import mesh as M
    k = 5
    g = M.build_remus_mesh(300, k, seed=9, points="uniform", edge_scale=(0.1, 0.2, 0.4))
    ref = M.Mesh(pos=g.pos.clone(), field=g.field.clone())
    ref = gfd.transforms.BuildRemusGraph(num_levels=3, k=k, scale_edge_length=(0.1, 0.2, 0.4))(ref)
    ref = gfd.transforms.BuildKnnInterpWeights(k)(ref)
    for name in ("edge_index", "edge_index2", "edge_index3", "coarse_mask2", "coarse_mask3", "angle_index",
                 "angle_index2", "angle_index3", "angle_index12", "angle_index23", "y_idx_21", "x_idx_21",
                 "y_idx_32", "x_idx_32"):
        assert torch.equal(getattr(g, name), getattr(ref, name)), name
    for name in ("edge_attr", "edge_attr2", "edge_attr3", "edgeUnitVector", "edgeUnitVector2", "edgeUnitVector3",
                 "edgeUnitVectorInverse", "edgeUnitVectorInverse2", "edgeUnitVectorInverse3", "angle_attr",
                 "angle_attr2", "angle_attr3", "angle_attr12", "angle_attr23", "weights_21", "weights_32"):
        assert rel_l2(getattr(g, name), getattr(ref, name)) <= 1e-5, name


@pytest.mark.parametrize("levels", [2, 4])
def test_mugs_layouts_match_reference_transforms(gfd, levels):
    """build_mugs_mesh against GuillardCoarseningAndConnectKNN (transforms/mugs.py:58-88) + BuildKnnInterpWeights
    (transforms/interpolate.py:147-155) on the same points."""
    from graphs4cfd_b200 import mesh as M
    k, scale = 6, (0.1, 0.25, 0.5, 1.0)[:levels]
    g = M.build_mugs_mesh(900 if levels == 2 else 6000, k, levels=levels, seed=12, points="uniform", edge_scale=scale)
    ref = M.Mesh(pos=g.pos.clone(), field=g.field.clone())
    ref = gfd.transforms.GuillardCoarseningAndConnectKNN(k=(k,) * levels, scale_edge_attr=scale)(ref)
    ref = gfd.transforms.BuildKnnInterpWeights(k)(ref)
    assert torch.equal(g.edge_index, ref.edge_index) and rel_l2(g.edge_attr, ref.edge_attr) <= 1e-6
    for l in range(2, levels + 1):
        for name in (f"coarse_mask{l}", f"edge_index{l}", f"y_idx_{l}{l-1}", f"x_idx_{l}{l-1}"):
            assert torch.equal(getattr(g, name), getattr(ref, name)), name
        for name in (f"edge_attr{l}", f"weights_{l}{l-1}"):
            assert rel_l2(getattr(g, name), getattr(ref, name)) <= 1e-5, name


@pytest.mark.parametrize("levels,n", [(2, 1500), (3, 4000), (4, 9000)])
def test_mugs_restatement_matches_reference_models(gfd, levels, n):
    """oracle.restate.mugs_forward against the reference's own MuGS classes (nn/mugs_gnn.py), seeded default init, 2 steps."""
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import mugs_arch
    from oracle import restate as R
    cls = {2: "NsTwoGuillardScaleGNN", 3: "NsThreeGuillardScaleGNN", 4: "NsFourGuillardScaleGNN"}[levels]
    torch.manual_seed(50 + levels)
    model = getattr(gfd.nn, cls)(arch=mugs_arch(32, levels), device=torch.device("cpu"))
    g = M.build_mugs_mesh(n, 6, levels=levels, seed=13, edge_scale=(0.1, 0.25, 0.5, 1.0)[:levels])
    ref = model.solve(g.clone(), 2)
    out = R.solve({k: v.detach() for k, v in model.state_dict().items()}, g.clone(), 2)
    assert rel_l2(out, ref) <= 1e-6


def _assert_same_mesh(a, b, float_tol=1e-6):
    keys = [k for k, v in b.__dict__.items() if torch.is_tensor(v) and k != "ptr"]
    assert set(keys) <= set(a.__dict__), set(keys) - set(a.__dict__)
    for k in keys:
        x, y = getattr(a, k), getattr(b, k)
        assert x.shape == y.shape, k
        assert (rel_l2(x, y) <= float_tol) if x.is_floating_point() else torch.equal(x, y), k


def _graphs(gfd, meshes):
    """The reference's own Graph objects (its Collater only takes torch_geometric Data instances, loader.py:16)."""
    return [gfd.Graph(**{k: v.clone() for k, v in m.__dict__.items() if torch.is_tensor(v)}) for m in meshes]


def test_collate_matches_reference_loader_remus(gfd):
    """mesh.collate against the reference's Collater (loader.py:14-56, its angle-index correction) + the batch-level
    BuildKnnInterpWeights transform; PyG's Batch.from_data_list is restated in oracle/pyg_stub.py (default Data rules)."""
    from graphs4cfd_b200 import mesh as M
    k = 5
    gs = [M.build_remus_mesh(n, k, seed=s, points="uniform", edge_scale=(0.1, 0.2, 0.4)) for n, s in ((260, 1), (300, 2), (240, 3))]
    ours = M.collate([g.clone() for g in gs], interp_k=k)
    ref = gfd.loader.Collater(gfd.transforms.BuildKnnInterpWeights(k))(_graphs(gfd, gs))
    _assert_same_mesh(ours, ref, 1e-5)


def test_collate_matches_reference_loader_mus_and_mugs(gfd):
    from graphs4cfd_b200 import mesh as M
    gs = [M.build_mus_mesh(n, 6, (), seed=s, edge_scale=0.1) for n, s in ((700, 4), (900, 5))]
    cells = (0.12, 0.3)
    ours = M.collate([g.clone() for g in gs], cells=cells)
    ref = gfd.loader.Collater(gfd.transforms.GridClustering(cells))(_graphs(gfd, gs))
    _assert_same_mesh(ours, ref)
    gs = [M.build_mugs_mesh(n, 6, levels=3, seed=s, edge_scale=(0.1, 0.25, 0.5)) for n, s in ((2500, 6), (3000, 7))]
    ours = M.collate([g.clone() for g in gs], interp_k=6)
    ref = gfd.loader.Collater(gfd.transforms.BuildKnnInterpWeights(6))(_graphs(gfd, gs))
    _assert_same_mesh(ours, ref, 1e-5)
