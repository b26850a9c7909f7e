"""Drop-in proof on hardware (north_star: "blocks drop into the existing graphs4cfd.nn model classes unchanged").

The reference's OWN model classes (imported verbatim under oracle/pyg_stub.py from /root/reference or from the staged
copy baseline/_ref, tools/stage_reference.py) run with our blocks swapped in, two ways:
  * ``accelerate(model)``        on an already-built model: module swap in place, Parameters shared;
  * ``patch_reference(gfd)``     before construction: ``load_arch`` (nn/mus_gnn.py:274-310, nn/remus_gnn.py:75-117) builds
                                 the model from our classes.
Then the reference's unmodified ``GNN.solve`` / ``forward`` (nn/model.py:303-321, nn/mus_gnn.py:312-373,
nn/remus_gnn.py:119-199) drive them.  Checked: rollout output against the goldens the unmodified reference wrote on the
CPU, against the plan-based ``Rollout`` engine, against the reference model itself with its trained weights (hidden 128,
tensor-core path); ``state_dict()`` keys AND their order; a ``.chk``-style ``load_state_dict`` round trip; that
``precision="fp32"`` launches no tensor-core kernel.
Tolerances: fp32 kernels 1e-5 rel-L2 on 2-3 step rollouts; fp16x3 tensor-core path 2e-5 per step budget -> 1e-4."""
import pytest
import torch

from conftest import load_golden, mesh_from, rel_l2, shipped_model

pytestmark = [pytest.mark.gpu, pytest.mark.reference]


@pytest.fixture(scope="module")
def gfd():
    from oracle.pyg_stub import import_reference
    return import_reference()


def _hidden(params):
    return next(v for k, v in params.items() if k.endswith("linear_1.weight")).shape[0]


def _build(gfd, d, dev):
    from graphs4cfd_b200 import archs
    H = _hidden(d["params"])
    if d["cls"] == "NsRotEquiTreeScaleGNN":
        model = gfd.nn.NsRotEquiTreeScaleGNN(arch=archs.remus_arch(H), device=dev)
    else:
        model = getattr(gfd.nn, d["cls"])(arch=archs.mus_arch(H, 3), device=dev)
    model.load_state_dict(d["params"])
    return model


@pytest.mark.parametrize("name", ["model_ns3_h32", "model_remus_h32"])
def test_accelerate_then_reference_solve_matches_golden(gfd, name):
    import graphs4cfd_b200 as g4
    dev = torch.device("cuda")
    d = load_golden(name)
    model = _build(gfd, d, dev)
    keys_before = list(model.state_dict().keys())
    ptrs_before = {k: v.data_ptr() for k, v in model.state_dict().items()}
    g4.accelerate(model)
    assert list(model.state_dict().keys()) == keys_before, "state_dict key ORDER must survive the module swap"
    assert {k: v.data_ptr() for k, v in model.state_dict().items()} == ptrs_before, "Parameters must be shared, not copied"
    blocks = [m for m in model.children()]
    assert blocks and all(type(m).__module__.startswith("graphs4cfd_b200") for m in blocks)
    n0 = g4.ops.L.launch_count()
    out = model.solve(mesh_from(d["mesh"]), d["n_out"]).cpu()          # the reference's own GNN.solve
    assert g4.ops.L.launch_count() > n0, "no libg4c kernel ran"
    assert rel_l2(out, d["out"]) <= 1e-5
    eng = g4.Rollout(model, mesh_from(d["mesh"]), device=dev)          # plan-based engine from the same (accelerated) model
    assert rel_l2(eng.solve(d["n_out"]).cpu(), out) <= 1e-5


@pytest.mark.parametrize("name", ["model_ns3_h32", "model_remus_h32"])
def test_patch_reference_builds_models_from_our_blocks(gfd, name):
    import graphs4cfd_b200 as g4
    dev = torch.device("cuda")
    d = load_golden(name)
    mods = [gfd.nn.mus_gnn, gfd.nn.remus_gnn, gfd.nn.mugs_gnn]
    names = ("MLP", "MP", "DownMP", "UpMP", "EdgeMP", "DownEdgeMP", "UpEdgeMP", "edgeScalarToNodeVector", "knn_interpolate",
             "restriction")
    saved = [{n: getattr(m, n) for n in names if hasattr(m, n)} for m in mods]
    try:
        g4.patch_reference(gfd)
        model = _build(gfd, d, dev)
        assert all(type(m).__module__.startswith("graphs4cfd_b200") for m in model.children())
        ref_keys = list(d["params"].keys())
        assert list(model.state_dict().keys()) == ref_keys, "same keys, same order as the reference's state_dict"
        out = model.solve(mesh_from(d["mesh"]), d["n_out"]).cpu()
        assert rel_l2(out, d["out"]) <= 1e-5
    finally:
        for m, sv in zip(mods, saved):
            for n, v in sv.items():
                setattr(m, n, v)


@pytest.mark.parametrize("kind,n_nodes,k", [("mus3", 3000, 6), ("remus", 600, 5)])
def test_trained_checkpoint_dropin_tensor_core_path(gfd, kind, n_nodes, k):
    """Hidden 128, the reference's trained weights: the accelerated model on the GPU (fp16x3 tensor-core kernels) against
    the SAME classes unaccelerated on the CPU, 3 rollout steps; plus the .chk-style load_state_dict round trip."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    dev = torch.device("cuda")
    ref = shipped_model(gfd, kind)
    g = M.build_mus_mesh(n_nodes, k, M.auto_cells(n_nodes, 3), seed=3) if kind == "mus3" else \
        M.build_remus_mesh(n_nodes, k, seed=5, points="uniform")
    with torch.no_grad():
        want = ref.solve(g.clone(), 3)
    fast = g4.accelerate(shipped_model(gfd, kind, device="cuda"))
    sd = {k_: v.clone() for k_, v in ref.state_dict().items()}
    fast.load_state_dict(sd)                                            # the keys a shipped .chk holds
    back = fast.state_dict()
    assert list(back.keys()) == list(sd.keys()) and all(torch.equal(back[k_].cpu(), sd[k_]) for k_ in sd)
    tc0 = g4.ops.L.tc_launch_count()
    with torch.no_grad():
        got = fast.solve(g.clone(), 3).cpu()
    assert g4.ops.L.tc_launch_count() > tc0, "hidden 128 under precision='auto' must take the tensor-core kernels"
    assert rel_l2(got, want) <= 1e-4, rel_l2(got, want)
    # the plan-based engine on the same model and mesh
    eng = g4.Rollout(fast, g.clone(), device=dev)
    assert rel_l2(eng.solve(3).cpu(), want) <= 1e-4


def test_precision_fp32_launches_no_tensor_core_kernel(gfd):
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    ref = shipped_model(gfd, "mus3")
    g = M.build_mus_mesh(1500, 6, M.auto_cells(1500, 3), seed=4)
    with torch.no_grad():
        want = ref.solve(g.clone(), 2)
    fast = g4.accelerate(shipped_model(gfd, "mus3", device="cuda"), precision="fp32")
    tc0 = g4.ops.L.tc_launch_count()
    with torch.no_grad():
        got = fast.solve(g.clone(), 2).cpu()
        eng_out = g4.Rollout(fast, g.clone(), precision="fp32", device=torch.device("cuda")).solve(2).cpu()
    assert g4.ops.L.tc_launch_count() == tc0, "precision='fp32' must stay on the CUDA-core kernels at hidden 128 too"
    assert rel_l2(got, want) <= 1e-5 and rel_l2(eng_out, want) <= 1e-5


def test_up_edge_mp_refuses_irregular_interpolation_lists():
    from graphs4cfd_b200.blocks import interp_layout
    dev = torch.device("cuda")
    ok = torch.arange(5, device=dev).repeat_interleave(3)
    assert interp_layout(ok) == (5, 3)
    with pytest.raises(RuntimeError, match="y_idx"):
        interp_layout(ok.flip(0))
    with pytest.raises(RuntimeError, match="y_idx"):
        interp_layout(torch.tensor([0, 0, 1, 1, 1, 2], device=dev))
