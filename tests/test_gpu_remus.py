"""GPU parity tests of the REMuS-GNN path (EdgeMP / DownEdgeMP / UpEdgeMP / edgeScalarToNodeVector and the
rollout engine) against golden vectors from the unmodified reference and the oracle."""
import pytest
import torch

from conftest import load_golden, mesh_from, rel_l2

pytestmark = pytest.mark.gpu
TOL_BLOCK = 2e-6


def dev(t):
    return t.cuda().contiguous()


def load_into(module, params, prefix):
    module.load_state_dict({k[len(prefix) + 1:]: v for k, v in params.items() if k.startswith(prefix + ".")})
    return module.cuda().eval()


def test_remus_blocks_golden():
    import graphs4cfd_b200 as g4
    d = load_golden("remus_blocks_h32")
    H = 32
    g = mesh_from(d["mesh"]).to("cuda")
    p = d["params"]
    emp = load_into(g4.EdgeMP((3 * H, (H, H), True), (2 * H, (H, H), True)), p, "emp")
    dmp = load_into(g4.DownEdgeMP((3 * H, (H, H), True), (2 * H, (H, H), True)), p, "dmp")
    ump = load_into(g4.UpEdgeMP((2 * H, (H, H, H), True)), p, "ump")
    e1, a1, e2, a12, e3 = (dev(d[k]) for k in ("e1", "a1", "e2", "a12", "e3"))
    with torch.no_grad():
        e1o, a1o = emp(e1, a1, g.angle_index)
        assert rel_l2(e1o.cpu(), d["e1_out"]) <= TOL_BLOCK and rel_l2(a1o.cpu(), d["a1_out"]) <= TOL_BLOCK
        e2o = dmp(e1, e2, a12, g.angle_index12)
        assert rel_l2(e2o.cpu(), d["e2_down"]) <= TOL_BLOCK
        e1u = ump(g.pos, g.y_idx_21, g.x_idx_21, g.weights_21, e2, g.edge_index2, g.edgeUnitVectorInverse2,
                  g.coarse_mask2, e1, g.edge_index, g.edgeUnitVector)
        assert rel_l2(e1u.cpu(), d["e1_up"]) <= TOL_BLOCK
        e2u = ump(g.pos, g.y_idx_32, g.x_idx_32, g.weights_32, e3, g.edge_index3, g.edgeUnitVectorInverse3,
                  g.coarse_mask3, e2, g.edge_index2, g.edgeUnitVector2, g.coarse_mask2)
        assert rel_l2(e2u.cpu(), d["e2_up"]) <= TOL_BLOCK
        nv = g4.edgeScalarToNodeVector(e1, g.edge_index, edgeUnitVectorInverse=g.edgeUnitVectorInverse)
        assert rel_l2(nv.cpu(), d["node_vec"]) <= TOL_BLOCK


def test_remus_edgemp_sum_golden():
    """EdgeMP(aggr='sum') (blocks.py:307-333): fixture written by the unmodified reference."""
    import graphs4cfd_b200 as g4
    d = load_golden("remus_edgemp_sum_h32")
    H = 32
    g = mesh_from(d["mesh"]).to("cuda")
    emp = load_into(g4.EdgeMP((3 * H, (H, H), True), (2 * H, (H, H), True), aggr="sum"), d["params"], "emp")
    with torch.no_grad():
        e1o, a1o = emp(dev(d["e1"]), dev(d["a1"]), g.angle_index)
    assert rel_l2(e1o.cpu(), d["e1_out"]) <= TOL_BLOCK and rel_l2(a1o.cpu(), d["a1_out"]) <= TOL_BLOCK


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_remus_rollout_golden(cuda_graph):
    import graphs4cfd_b200 as g4
    d = load_golden("model_remus_h32")
    eng = g4.Rollout(d["params"], mesh_from(d["mesh"]), cuda_graph=cuda_graph)
    out = eng.solve(d["n_out"])
    assert out.shape == d["out"].shape
    assert rel_l2(out.cpu(), d["out"]) <= 5e-5, rel_l2(out.cpu(), d["out"])


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-5), ("fp16x3", 1e-4)])
def test_remus_rollout_h128_vs_oracle(precision, tol):
    """REMuS-GNN at the benchmark's width (hidden=128, k=6) on a small mesh, both arithmetic paths."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, remus_arch
    from oracle import restate as R
    g = M.build_remus_mesh(1500, 6, seed=2)
    params = init_params(remus_arch(128), seed=3)
    want = R.solve(params, g.clone(), 2)
    got = g4.Rollout(params, g, precision=precision).solve(2).cpu()
    assert rel_l2(got, want) <= tol, rel_l2(got, want)


@pytest.mark.parametrize("hidden", [64, 256])
def test_remus_rollout_other_widths_vs_oracle(hidden):
    """configs[4] names hidden = 256: widths other than 128 run on the exact-fp32 CUDA-core kernels (precision 'auto')."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, remus_arch
    from oracle import restate as R
    g = M.build_remus_mesh(900, 6, seed=7)
    params = init_params(remus_arch(hidden), seed=4)
    want = R.solve(params, g.clone(), 2)
    eng = g4.Rollout(params, g)
    assert eng.precision == "fp32"
    got = eng.solve(2).cpu()
    assert rel_l2(got, want) <= 5e-5, rel_l2(got, want)
