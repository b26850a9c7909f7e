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
