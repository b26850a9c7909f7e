"""Batched graphs: several meshes collated the way the reference's loader does (loader.py:14-56; restated by
graphs4cfd_b200.mesh.collate and pinned to the reference's own Collater in tests/test_oracle_vs_reference.py) run as ONE input
through the rollout engines and the drop-in blocks.
Checked against the oracle on the same collated input and — where graphs cannot interact (REMuS, MuGS, one-scale MuS: every
index list stays inside its graph) — against the graphs solved one by one: a batch is a block-diagonal system, so each graph's
rows must come out as they do alone.  Tolerances: fp32 kernels 1e-5, fp16x3 1e-4 (rel-L2, 2-3 steps)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


def test_remus_batch_matches_oracle_and_single_graphs():
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, remus_arch
    from oracle import restate as R
    k = 5
    gs = [M.build_remus_mesh(n, k, seed=s, points="uniform") for n, s in ((300, 1), (260, 2), (340, 3))]
    batch = M.collate([g.clone() for g in gs], interp_k=k)
    for H, tol in ((32, 1e-5), (128, 1e-4)):
        params = init_params(remus_arch(H), seed=7)
        want = R.solve(params, batch.clone(), 2)
        got = g4.Rollout(params, batch.clone(), device=DEV).solve(2).cpu()
        assert rel_l2(got, want) <= tol, (H, rel_l2(got, want))
        singles = torch.cat([g4.Rollout(params, g.clone(), device=DEV).solve(2).cpu() for g in gs])
        assert rel_l2(got, singles) <= tol / 10, (H, rel_l2(got, singles))


def test_mus_batch_with_batch_level_clustering_matches_oracle():
    """MuS levels are clustered on the whole batch (transforms/mus.py:25 ignores the graph ids, as the reference does), so the
    graphs DO meet on the coarse levels: only the oracle on the same input is the yardstick."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    from oracle import restate as R
    gs = [M.build_mus_mesh(n, 6, (), seed=s, edge_scale=0.05) for n, s in ((1500, 4), (1100, 5))]
    batch = M.collate([g.clone() for g in gs], cells=M.auto_cells(1300, 3))
    for H, tol in ((32, 1e-5), (128, 1e-4)):
        params = init_params(mus_arch(H, 3), seed=8)
        want = R.solve(params, batch.clone(), 3)
        got = g4.Rollout(params, batch.clone(), device=DEV).solve(3).cpu()
        assert rel_l2(got, want) <= tol, (H, rel_l2(got, want))
    # one-scale model: no coarse level, the graphs never meet
    params = init_params(mus_arch(128, 1), seed=9)
    flat = M.collate([g.clone() for g in gs])
    got = g4.Rollout(params, flat.clone(), device=DEV).solve(2).cpu()
    singles = torch.cat([g4.Rollout(params, g.clone(), device=DEV).solve(2).cpu() for g in gs])
    assert rel_l2(got, singles) <= 1e-5


def test_mugs_batch_matches_oracle_and_single_graphs():
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mugs_arch
    from oracle import restate as R
    gs = [M.build_mugs_mesh(n, 6, levels=3, seed=s, edge_scale=(0.1, 0.25, 0.5)) for n, s in ((3000, 6), (2600, 7))]
    batch = M.collate([g.clone() for g in gs], interp_k=6)
    params = init_params(mugs_arch(128, 3), seed=10)
    want = R.solve(params, batch.clone(), 2)
    got = g4.Rollout(params, batch.clone(), device=DEV).solve(2).cpu()
    assert rel_l2(got, want) <= 1e-4, rel_l2(got, want)
    singles = torch.cat([g4.Rollout(params, g.clone(), device=DEV).solve(2).cpu() for g in gs])
    assert rel_l2(got, singles) <= 1e-5, rel_l2(got, singles)


@pytest.mark.reference
def test_batch_through_the_reference_model_with_our_blocks():
    """The reference's own NsThreeScaleGNN (shipped weights), accelerate()d, solving a collated batch on the GPU, against the same
    class on the CPU."""
    import graphs4cfd_b200 as g4
    from conftest import shipped_model
    from graphs4cfd_b200 import mesh as M
    from oracle.pyg_stub import import_reference
    gfd = import_reference()
    gs = [M.build_mus_mesh(n, 6, (), seed=s, edge_scale=0.05) for n, s in ((1500, 4), (1100, 5))]
    batch = M.collate([g.clone() for g in gs], cells=M.auto_cells(1300, 3))
    with torch.no_grad():
        want = shipped_model(gfd, "mus3").solve(batch.clone(), 2)
        got = g4.accelerate(shipped_model(gfd, "mus3", device="cuda")).solve(batch.clone(), 2).cpu()
    assert rel_l2(got, want) <= 1e-4, rel_l2(got, want)
