"""Host-side mesh logic that needs neither the GPU nor the reference tree: collation of several meshes into one input
(mesh.collate, the layout of loader.py:14-56) checked through the oracle — a batch is a block-diagonal system, so every graph's
rows must come out as they do alone — plus the MuGS mesh builder's invariants and the transport choice of the partition."""
import torch

from conftest import rel_l2


def test_collate_is_block_diagonal_for_the_oracle():
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mugs_arch, remus_arch
    from oracle import restate as R
    k = 5
    gs = [M.build_remus_mesh(n, k, seed=s, points="uniform") for n, s in ((220, 1), (180, 2), (260, 3))]
    batch = M.collate([g.clone() for g in gs], interp_k=k)
    assert batch.batch.tolist() == [0] * 220 + [1] * 180 + [2] * 260
    e1 = [g.edge_index.size(1) for g in gs]
    assert int(batch.angle_index[:, e1[0] * k:].min()) >= e1[0], "angle lists are offset by EDGE counts (loader.py:18-27)"
    params = init_params(remus_arch(16), seed=7)
    whole = R.solve(params, batch.clone(), 2)
    parts = torch.cat([R.solve(params, g.clone(), 2) for g in gs])
    assert rel_l2(whole, parts) <= 1e-6

    gs = [M.build_mugs_mesh(n, 6, levels=3, seed=s, edge_scale=(0.1, 0.25, 0.5)) for n, s in ((2200, 6), (1800, 7))]
    batch = M.collate([g.clone() for g in gs], interp_k=6)
    params = init_params(mugs_arch(16, 3), seed=10)
    whole = R.solve(params, batch.clone(), 2)
    parts = torch.cat([R.solve(params, g.clone(), 2) for g in gs])
    assert rel_l2(whole, parts) <= 1e-6


def test_mugs_mesh_invariants():
    """Layouts the MuGS plan relies on (rollout_mugs.py): nested masks, level-l edges in level-1 ids inside the mask with k
    in-edges per node grouped by target, interpolation lists uniform and sorted."""
    from graphs4cfd_b200 import mesh as M
    n, k = 3000, 6
    g = M.build_mugs_mesh(n, k, levels=3, seed=4)
    m2, m3 = g.coarse_mask2, g.coarse_mask3
    assert m2.dtype == torch.bool and bool((m3 <= m2).all()) and 0 < int(m3.sum()) < int(m2.sum()) < n
    for l, mask in ((2, m2), (3, m3)):
        ei = getattr(g, f"edge_index{l}")
        ids = mask.nonzero().squeeze(1)
        assert bool(mask[ei].all()) and ei.size(1) == ids.numel() * k
        assert torch.equal(ei[1], ids.repeat_interleave(k))
        lo = n if l == 2 else int(m2.sum())
        assert torch.equal(getattr(g, f"y_idx_{l}{l - 1}"), torch.arange(lo).repeat_interleave(k))
        assert int(getattr(g, f"x_idx_{l}{l - 1}").max()) < ids.numel()


def test_peer_memory_halo_is_not_chosen_without_an_nccl_job():
    from graphs4cfd_b200.partition import peer_memory_available
    assert peer_memory_available(1) is False
    assert peer_memory_available(2) is False          # no process group in this process
