"""Plan-time graph building on the device.  `g4c_plan_knn` (exact 2-D kNN on a cell grid) against the host
k-d tree the reference's connectivity comes from (torch_cluster.knn_graph / knn behind transforms/connect.py:58 and
transforms/interpolate.py:125; restated with scipy's cKDTree in mesh.knn_edges): index arrays must be EQUAL.  Then the whole
MuS / REMuS mesh built on the GPU against the host build (itself held to the reference's transforms in
tests/test_oracle_vs_reference.py): integer layouts equal, float attributes to 1e-6."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k,points", [(5000, 6, "jittered"), (3000, 5, "uniform"), (40, 6, "uniform"), (100000, 6, "jittered")])
def test_knn_graph_equals_host_kd_tree(n, k, points):
    from graphs4cfd_b200 import mesh as M
    pos = M.jittered_points(n, 3) if points == "jittered" else M.uniform_points(n, 3)
    ei_h, ea_h = M.knn_edges(pos, k)
    ei_d, ea_d = M.knn_edges(pos.cuda(), k)
    assert torch.equal(ei_d.cpu(), ei_h)
    assert torch.equal(ea_d.cpu(), ea_h)


def test_knn_interpolation_lists_equal_host():
    from graphs4cfd_b200 import mesh as M
    pos_y = M.uniform_points(4000, 1)
    pos_x = pos_y[torch.randperm(4000, generator=torch.Generator().manual_seed(0))[:700]]     # a coarse subset, queries outside its hull
    y_h, x_h, w_h = M.knn_interp_weights(pos_x, pos_y, 5)
    y_d, x_d, w_d = M.knn_interp_weights(pos_x.cuda(), pos_y.cuda(), 5)
    assert torch.equal(y_d.cpu(), y_h) and torch.equal(x_d.cpu(), x_h)
    assert rel_l2(w_d.cpu(), w_h) <= 1e-6


def _same(a, b, name):
    if a.is_floating_point():
        assert rel_l2(a.cpu(), b) <= 1e-5, name
    else:
        assert torch.equal(a.cpu(), b), name


def test_mus_mesh_built_on_device_equals_host_build():
    from graphs4cfd_b200 import mesh as M
    n = 20000
    host = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=2)
    dev = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=2, device="cuda")
    assert dev.edge_index.is_cuda
    for key, val in host.__dict__.items():
        _same(getattr(dev, key), val, key)


def test_remus_mesh_built_on_device_equals_host_build_and_runs():
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, remus_arch
    host = M.build_remus_mesh(3000, 6, seed=4)
    dev = M.build_remus_mesh(3000, 6, seed=4, device="cuda")
    for key, val in host.__dict__.items():
        _same(getattr(dev, key), val, key)
    params = init_params(remus_arch(32), seed=0)
    a = g4.Rollout(params, host, device="cuda").solve(2)
    b = g4.Rollout(params, dev, device="cuda").solve(2)
    assert rel_l2(b.cpu(), a.cpu()) <= 1e-5
