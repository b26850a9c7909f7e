"""Plan-time spatial renumbering (hard part 4): the rollout engine renumbers the level-1 nodes along a Morton
curve of their positions, so a mesh whose nodes arrive in ANY order is executed in the same memory order; inputs and outputs
stay in the caller's order.  Property: a randomly shuffled copy of a mesh gives, after un-shuffling, BIT-identical outputs
(every node keeps the relative order of its in-edges, so every row sees the same arithmetic)."""
import pytest
import torch

from conftest import load_golden, mesh_from, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,hidden", [("fp32", 32), ("fp16x3", 128)])
def test_shuffled_mesh_is_bit_identical_after_unshuffle(precision, hidden):
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    n = 5000
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=3)
    params = init_params(mus_arch(hidden, 3), seed=1)
    shuffle = torch.randperm(n, generator=torch.Generator().manual_seed(0))
    gs = M.permute_mus_nodes(g, shuffle.numpy())                      # node i of gs = node shuffle[i] of g
    a = g4.Rollout(params, g.clone(), precision=precision, device="cuda").solve(3)
    eng = g4.Rollout(params, gs, precision=precision, device="cuda")
    assert eng.node_perm is not None
    b = eng.solve(3)
    assert torch.equal(b, a[shuffle.cuda()])
    # and renumbering is a pure re-layout: same result as the engine that keeps the given order (up to the summation order
    # of the pooled coarse edges, whose fine members are visited in node order)
    c = g4.Rollout(params, g.clone(), precision=precision, device="cuda", renumber=False).solve(3)
    assert rel_l2(a, c) <= (2e-6 if precision == "fp32" else 1e-5)


def test_renumbered_engine_matches_reference_golden():
    import graphs4cfd_b200 as g4
    d = load_golden("model_ns3_h32")
    eng = g4.Rollout(d["params"], mesh_from(d["mesh"]), device="cuda")
    assert rel_l2(eng.solve(d["n_out"]).cpu(), d["out"]) <= 1e-5


def test_set_field_takes_the_callers_order():
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    n = 2000
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=4)
    params = init_params(mus_arch(32, 3), seed=2)
    eng = g4.Rollout(params, g.clone(), device="cuda")
    base = eng.solve(2)
    field2 = g.field * 0.5 + 0.1
    g2 = g.clone()
    g2.field = field2
    want = g4.Rollout(params, g2, device="cuda").solve(2)
    assert torch.equal(eng.solve(2, field=field2), want)
    assert torch.equal(eng.solve(2), base), "solve() restores the engine's own initial field"
