"""GPU parity tests (run on the B200 box): CUDA path through the C ABI vs golden vectors written by
the unmodified reference, and vs the oracle restatement on seeded inputs."""
import pytest
import torch

from conftest import load_golden, mesh_from, rel_l2

pytestmark = pytest.mark.gpu

TOL_BLOCK = 2e-6     # fp32 FFMA path: same arithmetic, different summation order
TOL_ROLLOUT = 5e-5   # a few chained steps of ~30 blocks


def dev(t):
    return t.cuda().contiguous()


def load_into(module, params, prefix):
    sd = {k[len(prefix) + 1:]: v for k, v in params.items() if k.startswith(prefix + ".")}
    module.load_state_dict(sd)
    return module.cuda().eval()


@pytest.mark.parametrize("name,args", [("mlp_enc", (5, (32, 32, 32), False)), ("mlp_ln2", (4, (32, 32), True)),
                                       ("mlp_dec", (32, (32, 32, 3), False)), ("mlp_dec1", (32, (32, 1), False))])
def test_mlp(name, args):
    import graphs4cfd_b200 as g4
    d = load_golden(name)
    m = load_into(g4.MLP(*args), d["params"], "m")
    with torch.no_grad():
        y = m(dev(d["x"]))
    assert rel_l2(y.cpu(), d["y"]) <= TOL_BLOCK


@pytest.mark.parametrize("name,H", [("mp_trained_h128", 128), ("mp_irregular_mean_h32", 32),
                                    ("mp_irregular_sum_h32", 32), ("mp_h64", 64)])
def test_gn_block(name, H):
    import graphs4cfd_b200 as g4
    d = load_golden(name)
    blk = load_into(g4.MP((3 * H, (H, H, H), True), (2 * H, (H, H, H), True), aggr=d["aggr"]), d["params"], "mp")
    with torch.no_grad():
        v, e = blk(dev(d["v"]), dev(d["e"]), d["edge_index"].cuda())
    assert rel_l2(v.cpu(), d["v_out"]) <= TOL_BLOCK
    assert rel_l2(e.cpu(), d["e_out"]) <= TOL_BLOCK


def test_state_dict_keys_match_reference():
    import graphs4cfd_b200 as g4
    d = load_golden("mp_trained_h128")
    blk = g4.MP((384, (128, 128, 128), True), (256, (128, 128, 128), True))
    assert set("mp." + k for k in blk.state_dict()) == set(d["params"])


def test_down_up():
    import graphs4cfd_b200 as g4
    d = load_golden("down_up_h32")
    H = 32
    g = mesh_from(d["mesh"]).to("cuda")
    hr_field, hr_pos = g.field, g.pos
    dn = load_into(g4.DownMP((2 + H, (H, H, H), True), 1), d["params"], "down")
    up = load_into(g4.UpMP((2 + 2 * H, (H, H, H), True), 2), d["params"], "up")
    with torch.no_grad():
        g = dn(g, activation=torch.tanh)
        assert torch.equal(g.edge_index.cpu(), d["edge_index_l"])
        assert rel_l2(g.field.cpu(), d["field_l"]) <= TOL_BLOCK
        assert rel_l2(g.edge_attr.cpu(), d["edge_attr_l"]) <= TOL_BLOCK
        g = up(g, hr_field, hr_pos, activation=torch.tanh)
    assert rel_l2(g.field.cpu(), d["field_h_up"]) <= TOL_BLOCK


@pytest.mark.parametrize("name", ["model_ns1_h16", "model_ns2_h16", "model_ns3_h32", "model_ns4_h16", "model_adv1_h16", "model_adv2_h16",
                                  "model_adv3_h16", "model_adv4_h16"])
@pytest.mark.parametrize("cuda_graph", [False, True])
def test_mus_rollout_golden(name, cuda_graph):
    import graphs4cfd_b200 as g4
    d = load_golden(name)
    eng = g4.Rollout(d["params"], mesh_from(d["mesh"]), cuda_graph=cuda_graph)
    out = eng.solve(d["n_out"])
    assert out.shape == d["out"].shape
    assert rel_l2(out.cpu(), d["out"]) <= TOL_ROLLOUT
    # engine state is restored: a second solve gives the same answer
    assert torch.equal(eng.solve(d["n_out"]), out)


def test_config1_block_vs_oracle():
    """BASELINE config 1: single MP block forward, 10k-node / 60k-edge random mesh, hidden=64."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from oracle import restate as R
    torch.manual_seed(0)
    H, n, k = 64, 10000, 6
    blk = g4.MP((3 * H, (H, H, H), True), (2 * H, (H, H, H), True))
    ei, _ = M.knn_edges(M.uniform_points(n, 0), k)
    v, e = torch.randn(n, H), torch.randn(n * k, H)
    params = {"mp." + a: b.detach() for a, b in blk.state_dict().items()}
    with torch.no_grad():
        v_ref, e_ref = R.gn_block(params, "mp", v, e, ei)
        blk = blk.cuda()
        v_out, e_out = blk(dev(v), dev(e), ei.cuda())
    assert rel_l2(v_out.cpu(), v_ref) <= TOL_BLOCK and rel_l2(e_out.cpu(), e_ref) <= TOL_BLOCK


def test_empty_and_ragged_inputs():
    import graphs4cfd_b200 as g4
    from oracle import restate as R
    torch.manual_seed(1)
    H = 16
    blk = g4.MP((3 * H, (H, H, H), True), (2 * H, (H, H, H), True))
    params = {"mp." + a: b.detach() for a, b in blk.state_dict().items()}
    blk = blk.cuda()
    # (a) no edges at all: every node aggregates zeros (count clamped to 1)
    v = torch.randn(37, H)
    ei = torch.zeros(2, 0, dtype=torch.long)
    e = torch.zeros(0, H)
    with torch.no_grad():
        v_ref, _ = R.gn_block(params, "mp", v, e, ei)
        v_out, e_out = blk(dev(v), dev(e), ei.cuda())
    assert e_out.shape == (0, H) and rel_l2(v_out.cpu(), v_ref) <= TOL_BLOCK
    # (b) one hub node receiving 300 edges, the rest none (tile with a very ragged degree profile)
    n = 700
    v = torch.randn(n, H)
    ei = torch.stack([torch.randint(0, n, (300,)), torch.full((300,), 5)])
    e = torch.randn(300, H)
    with torch.no_grad():
        v_ref, e_ref = R.gn_block(params, "mp", v, e, ei)
        v_out, e_out = blk(dev(v), dev(e), ei.cuda())
    assert rel_l2(v_out.cpu(), v_ref) <= 1e-5 and rel_l2(e_out.cpu(), e_ref) <= TOL_BLOCK


def test_rejects_cpu_tensors_loudly():
    import graphs4cfd_b200 as g4
    blk = g4.MLP(4, (16, 16), False)
    with pytest.raises(RuntimeError):
        blk(torch.randn(3, 4))


@pytest.mark.parametrize("hidden", [64, 256])
def test_mus_rollout_other_widths_vs_oracle(hidden):
    """BASELINE.json configs[0] (hidden 64) and configs[4] (hidden 256): the exact-fp32 CUDA-core path."""
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    from oracle import restate as R
    n = 2500
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=11)
    params = init_params(mus_arch(hidden, 3), seed=5)
    want = R.solve(params, g.clone(), 2)
    got = g4.Rollout(params, g).solve(2).cpu()
    assert rel_l2(got, want) <= 5e-5, rel_l2(got, want)
