"""MuGS-GNN through the drop-in boundary: the reference's OWN NsTwoGuillardScaleGNN / NsFourGuillardScaleGNN
classes (nn/mugs_gnn.py, imported verbatim under oracle/pyg_stub.py) with the shipped trained weights, ``accelerate()``d, driven
by the reference's unmodified ``GNN.solve`` on the GPU, against the same classes untouched on the CPU.

What the models need beyond the MuS blocks:
  * ``MP`` with 256-wide node features (the first block behind every up-sampling, nn/mugs_gnn.py:34, 121-123) — tensor-core
    path only: ops.mp splits linear_1 of both MLPs (MlpPack.tc_edge / tc_node_wide);
  * the module-level helpers ``knn_interpolate`` (blocks.py:34-48 -> g4c_interp_fwd) and ``restriction`` (blocks.py:9-32, cached).
Tolerance: fp16x3 tensor-core path, 2e-5 per step budget -> 1e-4 rel-L2 over a 3-step rollout (the bound of test_gpu_dropin.py)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

MODELS = {"mugs2": ("NsTwoGuillardScaleGNN", "NsTwoGuillardScaleGNN.chk", "2GS-GNN-NsCircle-v1", 2, 2500),
          "mugs4": ("NsFourGuillardScaleGNN", "NsFourGuillardScaleGNN.chk", "4GS-GNN-NsCircle-v1", 4, 9000)}


def _shipped(gfd, kind, device):
    from oracle.pyg_stub import staged_checkpoint
    cls, chk, name, _, _ = MODELS[kind]
    path = staged_checkpoint(chk)
    dev = torch.device(device)
    return getattr(gfd.nn, cls)(checkpoint=path, device=dev) if path else getattr(gfd.nn, cls)(model=name, device=dev)


@pytest.mark.reference
@pytest.mark.parametrize("kind", ["mugs2", "mugs4"])
def test_mugs_trained_checkpoint_dropin(kind):
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from oracle.pyg_stub import import_reference
    gfd = import_reference()
    levels, n = MODELS[kind][3], MODELS[kind][4]
    g = M.build_mugs_mesh(n, 6, levels=levels, seed=31, edge_scale=(0.1, 0.25, 0.5, 1.0)[:levels])
    ref = _shipped(gfd, kind, "cpu")
    with torch.no_grad():
        want = ref.solve(g.clone(), 3)
    fast = _shipped(gfd, kind, "cuda")
    keys = list(fast.state_dict().keys())
    g4.accelerate(fast)
    assert list(fast.state_dict().keys()) == keys == list(ref.state_dict().keys())
    assert all(type(m).__module__.startswith("graphs4cfd_b200") for m in fast.children())
    n0, tc0 = g4.ops.L.launch_count(), g4.ops.L.tc_launch_count()
    with torch.no_grad():
        got = fast.solve(g.clone(), 3).cpu()
    per_step = (g4.ops.L.launch_count() - n0) / 3
    assert g4.ops.L.tc_launch_count() > tc0
    assert torch.isfinite(got).all()
    err = rel_l2(got, want)
    print(f"{kind}: {n} nodes, 3 steps, rel-L2 vs the reference on the CPU {err:.3e}, {per_step:.0f} libg4c launches per step")
    assert err <= 1e-4, err
    # the CPU model of the same module still runs the reference's own helpers (dispatch wrappers, not replacements)
    with torch.no_grad():
        again = ref.solve(g.clone(), 1)
    assert torch.equal(again, want[:, :again.shape[1]])


@pytest.mark.reference
@pytest.mark.parametrize("levels,n", [(2, 2500), (3, 6000), (4, 9000)])
def test_mugs_rollout_engine_matches_reference(levels, n):
    """The plan-based engine (rollout_mugs.py: static encoders, folded activations, no concatenation, one CUDA graph per step)
    against the reference's own class on the CPU; seeded default-init weights (the 3-scale model ships no checkpoint), 3 steps,
    eager and captured."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import mugs_arch
    from oracle.pyg_stub import import_reference
    gfd = import_reference()
    cls = {2: "NsTwoGuillardScaleGNN", 3: "NsThreeGuillardScaleGNN", 4: "NsFourGuillardScaleGNN"}[levels]
    torch.manual_seed(40 + levels)
    ref = getattr(gfd.nn, cls)(arch=mugs_arch(128, levels), device=torch.device("cpu"))
    g = M.build_mugs_mesh(n, 6, levels=levels, seed=33, edge_scale=(0.1, 0.25, 0.5, 1.0)[:levels])
    with torch.no_grad():
        want = ref.solve(g.clone(), 3)
    for graph in (False, True):
        eng = g4.Rollout(ref, g.clone(), device=torch.device("cuda"), cuda_graph=graph)
        assert eng.level_nodes[0] == n and len(eng.level_nodes) == levels
        got = eng.solve(3).cpu()
        err = rel_l2(got, want)
        print(f"mugs{levels} engine (cuda_graph={graph}): rel-L2 vs the reference on the CPU {err:.3e}, {eng.launches_per_step} launches per step")
        assert err <= 1e-4, err
        again = eng.solve(2).cpu()                   # the engine restores its input state (GNN.solve does, nn/model.py:320)
        assert torch.equal(again, got[:, :again.shape[1]])


@pytest.mark.parametrize("name", ["model_mugs2_h128", "model_mugs3_h128"])
def test_mugs_rollout_engine_matches_golden(name):
    """Engine against the rollouts the unmodified reference wrote on the CPU (tests/golden, oracle/make_golden.py --mugs); needs no
    reference tree.  Parameters are regenerated from the stored seed."""
    import graphs4cfd_b200 as g4
    from conftest import load_golden, mesh_from
    from graphs4cfd_b200.archs import init_params, mugs_arch
    d = load_golden(name)
    params = init_params(mugs_arch(d["hidden"], d["levels"]), seed=d["param_seed"])
    eng = g4.Rollout(params, mesh_from(d["mesh"]), device=torch.device("cuda"))
    err = rel_l2(eng.solve(d["n_out"]).cpu(), d["out"])
    assert err <= 1e-4, err
    with pytest.raises(RuntimeError, match="256-wide|tensor-core"):
        g4.Rollout(params, mesh_from(d["mesh"]), device=torch.device("cuda"), precision="fp32").solve(1)


@pytest.mark.parametrize("aggr", ["mean", "sum"])
def test_mp_block_with_256_wide_node_features(aggr):
    """ops.mp with 256-wide node features against an fp64 restatement of GNBlock.forward (blocks.py:176-186)."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M, ops
    dev = torch.device("cuda")
    torch.manual_seed(5)
    n, k, H, F = 700, 6, 128, 256
    ei, _ = M.knn_edges(M.uniform_points(n, 3), k)
    edge = g4.MLP(H + 2 * F, (H, H, H), True).to(dev)
    node = g4.MLP(H + F, (H, H, H), True).to(dev)
    v, e = torch.randn(n, F, device=dev), torch.randn(n * k, H, device=dev)
    blk = g4.MP((H + 2 * F, (H, H, H), True), (H + F, (H, H, H), True), aggr=aggr).to(dev)
    blk.edge_mlp, blk.node_mlp = edge, node
    with torch.no_grad():
        v_new, e_new = blk(v, e, ei.to(dev))

    def mlp64(m, x):
        seq = m.MLP
        x = torch.nn.functional.selu(x @ seq.linear_1.weight.double().T + seq.linear_1.bias.double())
        x = torch.nn.functional.selu(x @ seq.linear_2.weight.double().T + seq.linear_2.bias.double())
        x = x @ seq.linear_3.weight.double().T + seq.linear_3.bias.double()
        return torch.nn.functional.layer_norm(x, (H,), seq.layer_norm.weight.double(), seq.layer_norm.bias.double())

    row, col = ei.to(dev)
    v64, e64 = v.double(), e.double()
    with torch.no_grad():
        e_ref = mlp64(edge, torch.cat([e64, v64[row], v64[col]], dim=1))
        agg = torch.zeros(n, H, device=dev, dtype=torch.float64).index_add_(0, col, e_ref)
        if aggr == "mean":
            agg /= k
        v_ref = mlp64(node, torch.cat([agg, v64], dim=1))
    assert rel_l2(e_new.double(), e_ref) <= 2e-6, rel_l2(e_new.double(), e_ref)
    assert rel_l2(v_new.double(), v_ref) <= 2e-6, rel_l2(v_new.double(), v_ref)
    with pytest.raises(RuntimeError, match="256-wide"):
        blk.precision = "fp32"
        blk(v, e, ei.to(dev))
