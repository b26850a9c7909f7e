"""TEST INFRASTRUCTURE — writes tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
Every fixture stores the inputs, the reference parameters (state_dict) and the outputs the
reference's own classes produced on CPU/fp32, so that the oracle restatement and the CUDA
path can be held to them on machines where the reference tree does not exist.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyg_stub import import_reference  # noqa: E402
from graphs4cfd_b200 import mesh as M  # noqa: E402

gfd = import_reference()
B = gfd.nn.blocks
OUT = os.path.join(ROOT, "tests", "golden")


def mesh_dict(g):
    return {k: v for k, v in g.__dict__.items() if torch.is_tensor(v)}


def sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def mus_arch(H, levels, adv=False, nf=3, node_in=5):
    mp = lambda: ((3 * H, (H, H, H), True), (2 * H, (H, H, H), True))
    a = {"edge_encoder": (2, (H, H, H), False), "node_encoder": (node_in, (H, H, H), False)}
    n1 = 2 if adv else 4
    if levels == 1:
        names = [f"mp1{i}" for i in range(1, 9)] if not adv else ["mp111", "mp112", "mp121", "mp122"]
        for n in names:
            a[n] = mp()
    else:
        for i in range(1, n1 + 1):
            a[f"mp11{i}"] = mp()
        a["down_mp12"] = (2 + H, (H, H, H), True)
        if levels == 2:
            for i in range(1, 5):
                a[f"mp2{i}"] = mp()
        else:
            a["mp211"], a["mp212"] = mp(), mp()
            a["down_mp23"] = (2 + H, (H, H, H), True)
            if levels == 3:
                for i in range(1, 5):
                    a[f"mp3{i}"] = mp()
            else:
                a["mp311"], a["mp312"] = mp(), mp()
                a["down_mp34"] = (2 + H, (H, H, H), True)
                for i in range(1, 5):
                    a[f"mp4{i}"] = mp()
                a["up_mp43"] = (2 + 2 * H, (H, H, H), True)
                a["mp321"], a["mp322"] = mp(), mp()
            a["up_mp32"] = (2 + 2 * H, (H, H, H), True)
            a["mp221"], a["mp222"] = mp(), mp()
        a["up_mp21"] = (2 + 2 * H, (H, H, H), True)
        for i in range(1, n1 + 1):
            a[f"mp12{i}"] = mp()
    a["decoder"] = (H, (H, H, nf), False)
    return a


def remus_arch(H):
    mp = lambda: ((3 * H, (H, H), True), (2 * H, (H, H), True))
    a = {}
    for n in ("angle_encoder", "angle_encoder12", "angle_encoder2", "angle_encoder23", "angle_encoder3"):
        a[n] = (4, (H, H), True)
    for n in ("edge_encoder", "edge_encoder2", "edge_encoder3"):
        a[n] = (3, (H, H), True)
    for n in ("mp111", "mp112", "mp113", "mp114", "down_mp12", "mp211", "mp212", "down_mp23",
              "mp31", "mp32", "mp33", "mp34"):
        a[n] = mp()
    a["up_mp32"] = (2 * H, (H, H, H), True)
    a["mp221"], a["mp222"] = mp(), mp()
    a["up_mp21"] = (2 * H, (H, H, H), True)
    for n in ("mp121", "mp122", "mp123", "mp124"):
        a[n] = mp()
    a["decoder"] = (H, (H, 1), False)
    return a


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(1234)
    fx = {}

    # ---- 1. GNBlock with TRAINED weights (mp111 of the shipped 3S-GNN), fixed-k graph
    chk = torch.load(os.path.join(gfd.nn.__path__[0], "weights/NsMuSGNN/NsThreeScaleGNN.chk"),
                     map_location="cpu", weights_only=False)
    H = 128
    blk = B.MP(*chk["arch"]["mp111"])
    blk.load_state_dict({k[len("mp111."):]: v for k, v in chk["weights"].items() if k.startswith("mp111.")})
    blk.eval()
    n, k = 160, 6
    ei, _ = M.knn_edges(M.uniform_points(n, 3), k)
    v, e = torch.randn(n, H), torch.randn(n * k, H)
    with torch.no_grad():
        vo, eo = blk(v, e, ei)
    fx["mp_trained_h128"] = dict(params={"mp." + a: b for a, b in sd(blk).items()}, v=v, e=e, edge_index=ei,
                                 v_out=vo, e_out=eo, aggr="mean")

    # ---- 2. GNBlock on an irregular graph (random degrees incl. isolated targets), mean and sum
    for aggr in ("mean", "sum"):
        Hs = 32
        blk = B.MP((3 * Hs, (Hs, Hs, Hs), True), (2 * Hs, (Hs, Hs, Hs), True), aggr=aggr).eval()
        n, E = 211, 903
        ei = torch.stack([torch.randint(0, n, (E,)), torch.randint(0, n - 7, (E,))])   # last 7 nodes isolated
        v, e = torch.randn(n, Hs), torch.randn(E, Hs)
        with torch.no_grad():
            vo, eo = blk(v, e, ei)
        fx[f"mp_irregular_{aggr}_h32"] = dict(params={"mp." + a: b for a, b in sd(blk).items()}, v=v, e=e,
                                              edge_index=ei, v_out=vo, e_out=eo, aggr=aggr)

    # ---- 3. config-1 shaped block (H=64, 3-layer) at reduced size, default init
    Hs = 64
    blk = B.MP((3 * Hs, (Hs, Hs, Hs), True), (2 * Hs, (Hs, Hs, Hs), True)).eval()
    n, k = 300, 6
    ei, _ = M.knn_edges(M.uniform_points(n, 5), k)
    v, e = torch.randn(n, Hs), torch.randn(n * k, Hs)
    with torch.no_grad():
        vo, eo = blk(v, e, ei)
    fx["mp_h64"] = dict(params={"mp." + a: b for a, b in sd(blk).items()}, v=v, e=e, edge_index=ei,
                        v_out=vo, e_out=eo, aggr="mean")

    # ---- 4. MLP variants (2-layer / 3-layer / LN / tiny in & out widths)
    for tag, args in (("enc", (5, (32, 32, 32), False)), ("ln2", (4, (32, 32), True)), ("dec", (32, (32, 32, 3), False)),
                      ("dec1", (32, (32, 1), False))):
        m = B.MLP(*args).eval()
        x = torch.randn(333, args[0])
        with torch.no_grad():
            y = m(x)
        fx[f"mlp_{tag}"] = dict(params={"m." + a: b for a, b in sd(m).items()}, x=x, y=y)

    # ---- 5. DownMP / UpMP / pool_edge on a 2-level grid-clustered mesh
    Hs = 32
    g = M.build_mus_mesh(500, 6, M.auto_cells(500, 2), seed=2)
    g.field = torch.randn(500, Hs)
    g.edge_attr = torch.randn(g.edge_index.size(1), Hs)
    dn = B.DownMP((2 + Hs, (Hs, Hs, Hs), True), 1).eval()
    up = B.UpMP((2 + 2 * Hs, (Hs, Hs, Hs), True), 2).eval()
    gg = g.clone()
    with torch.no_grad():
        gg = dn(gg, activation=torch.tanh)
        f2, ei2, ea2 = gg.field.clone(), gg.edge_index.clone(), gg.edge_attr.clone()
        gg = up(gg, g.field, g.pos, activation=torch.tanh)
    fx["down_up_h32"] = dict(mesh=mesh_dict(g), params={**{"down." + a: b for a, b in sd(dn).items()},
                                                        **{"up." + a: b for a, b in sd(up).items()}},
                             field_l=f2, edge_index_l=ei2, edge_attr_l=ea2, field_h_up=gg.field.clone())

    # ---- 6. REMuS blocks
    k = 5
    g = M.build_remus_mesh(130, k, seed=4, points="uniform")
    E1, E2 = g.edge_index.size(1), g.edge_index2.size(1)
    emp = B.EdgeMP((3 * Hs, (Hs, Hs), True), (2 * Hs, (Hs, Hs), True)).eval()
    dmp = B.DownEdgeMP((3 * Hs, (Hs, Hs), True), (2 * Hs, (Hs, Hs), True)).eval()
    ump = B.UpEdgeMP((2 * Hs, (Hs, Hs, Hs), True)).eval()
    e1, a1 = torch.randn(E1, Hs), torch.randn(E1 * k, Hs)
    e2, a12 = torch.randn(E2, Hs), torch.randn(E2 * k, Hs)
    with torch.no_grad():
        e1o, a1o = emp(e1, a1, g.angle_index)
        e2o = dmp(e1, e2, a12, g.angle_index12)
        e1u = ump(g.pos, g.y_idx_21, g.x_idx_21, g.weights_21, e2, g.edge_index2, g.edgeUnitVectorInverse2,
                  g.coarse_mask2, e1, g.edge_index, g.edgeUnitVector)
        e3 = torch.randn(g.edge_index3.size(1), Hs)
        e2u = ump(g.pos, g.y_idx_32, g.x_idx_32, g.weights_32, e3, g.edge_index3, g.edgeUnitVectorInverse3,
                  g.coarse_mask3, e2, g.edge_index2, g.edgeUnitVector2, g.coarse_mask2)
        nv = B.edgeScalarToNodeVector(e1, g.edge_index, edgeUnitVectorInverse=g.edgeUnitVectorInverse)
    fx["remus_blocks_h32"] = dict(mesh=mesh_dict(g), k=k,
                                  params={**{"emp." + a: b for a, b in sd(emp).items()},
                                          **{"dmp." + a: b for a, b in sd(dmp).items()},
                                          **{"ump." + a: b for a, b in sd(ump).items()}},
                                  e1=e1, a1=a1, e2=e2, a12=a12, e3=e3, e1_out=e1o, a1_out=a1o, e2_down=e2o,
                                  e1_up=e1u, e2_up=e2u, node_vec=nv)

    # ---- 7. whole models, seeded default init, short rollouts
    def model_fixture(cls, arch, g, n_out):
        model = cls(arch=arch)
        out = model.solve(g.clone(), n_out)
        return dict(mesh=mesh_dict(g), params=sd(model), out=out, n_out=n_out, cls=cls.__name__)

    g3 = M.build_mus_mesh(1200, 6, M.auto_cells(1200, 3), seed=1)
    fx["model_ns3_h32"] = model_fixture(gfd.nn.NsThreeScaleGNN, mus_arch(32, 3), g3, 3)
    g1 = M.build_mus_mesh(600, 6, (), seed=6)
    fx["model_ns1_h16"] = model_fixture(gfd.nn.NsOneScaleGNN, mus_arch(16, 1), g1, 2)
    g2 = M.build_mus_mesh(800, 6, M.auto_cells(800, 2), seed=7)
    fx["model_ns2_h16"] = model_fixture(gfd.nn.NsTwoScaleGNN, mus_arch(16, 2), g2, 2)
    g4 = M.build_mus_mesh(1600, 6, M.auto_cells(1600, 4), seed=8)
    fx["model_ns4_h16"] = model_fixture(gfd.nn.NsFourScaleGNN, mus_arch(16, 4), g4, 2)
    ga = M.build_mus_mesh(900, 6, M.auto_cells(900, 3), seed=9, num_fields=1)
    del ga.glob
    ga.loc = torch.randn(900, 2) * 0.3
    fx["model_adv3_h16"] = model_fixture(gfd.nn.AdvThreeScaleGNN, mus_arch(16, 3, adv=True, nf=1, node_in=4), ga, 2)
    gr = M.build_remus_mesh(300, 5, seed=11, points="uniform")
    fx["model_remus_h32"] = model_fixture(gfd.nn.NsRotEquiTreeScaleGNN, remus_arch(32), gr, 3)

    total = 0
    for name, d in fx.items():
        path = os.path.join(OUT, name + ".pt")
        torch.save(d, path)
        total += os.path.getsize(path)
        print(f"{name:28s} {os.path.getsize(path) / 1e6:7.2f} MB")
    print(f"total {total / 1e6:.2f} MB")


def extra():
    """Round-2 additions, written WITHOUT touching the fixtures above (own seeds): the remaining advection models
    (nn/mus_gnn.py:68-97, 173-218, 984-1053) and the REMuS angle -> edge blocks with aggr='sum' (blocks.py:307-333)."""
    torch.manual_seed(1234)
    fx = {}

    def model_fixture(cls, arch, g, n_out):
        model = cls(arch=arch)
        out = model.solve(g.clone(), n_out)
        return dict(mesh=mesh_dict(g), params=sd(model), out=out, n_out=n_out, cls=cls.__name__)

    for levels, cls, n, seed in ((1, gfd.nn.AdvOneScaleGNN, 500, 21), (2, gfd.nn.AdvTwoScaleGNN, 700, 22),
                                 (4, gfd.nn.AdvFourScaleGNN, 1600, 24)):
        ga = M.build_mus_mesh(n, 6, M.auto_cells(n, levels) if levels > 1 else (), seed=seed, num_fields=1)
        del ga.glob
        ga.loc = torch.randn(n, 2) * 0.3
        fx[f"model_adv{levels}_h16"] = model_fixture(cls, mus_arch(16, levels, adv=True, nf=1, node_in=4), ga, 2)

    Hs, k = 32, 5
    g = M.build_remus_mesh(130, k, seed=4, points="uniform")
    E1 = g.edge_index.size(1)
    emp = B.EdgeMP((3 * Hs, (Hs, Hs), True), (2 * Hs, (Hs, Hs), True), aggr="sum").eval()
    e1, a1 = torch.randn(E1, Hs), torch.randn(E1 * k, Hs)
    with torch.no_grad():
        e1o, a1o = emp(e1, a1, g.angle_index)
    fx["remus_edgemp_sum_h32"] = dict(mesh=mesh_dict(g), k=k, params={"emp." + a: b for a, b in sd(emp).items()},
                                      e1=e1, a1=a1, e1_out=e1o, a1_out=a1o)
    for name, d in fx.items():
        path = os.path.join(OUT, name + ".pt")
        torch.save(d, path)
        print(f"{name:28s} {os.path.getsize(path) / 1e6:7.2f} MB")


def mugs():
    """MuGS-GNN rollouts (nn/mugs_gnn.py).  Hidden 128 (the 256-wide blocks exist on the tensor-core path only), so the
    parameters are NOT stored (7-13 MB each): they are graphs4cfd_b200.archs.init_params(mugs_arch(128, levels), seed), loaded
    into the reference's class here and regenerated by the tests from the stored seed."""
    from graphs4cfd_b200.archs import init_params, mugs_arch
    for levels, cls, n, seed in ((2, gfd.nn.NsTwoGuillardScaleGNN, 1200, 61), (3, gfd.nn.NsThreeGuillardScaleGNN, 3500, 62)):
        arch = mugs_arch(128, levels)
        params = init_params(arch, seed=seed)
        model = cls(arch=arch)
        model.load_state_dict(params)
        g = M.build_mugs_mesh(n, 6, levels=levels, seed=seed, edge_scale=(0.1, 0.25, 0.5)[:levels])
        out = model.solve(g.clone(), 3)
        name = f"model_mugs{levels}_h128"
        path = os.path.join(OUT, name + ".pt")
        torch.save(dict(mesh=mesh_dict(g), param_seed=seed, levels=levels, hidden=128, out=out, n_out=3, cls=cls.__name__), path)
        print(f"{name:28s} {os.path.getsize(path) / 1e6:7.2f} MB")


if __name__ == "__main__":
    if "--mugs" in sys.argv:
        mugs()
    elif "--extra" in sys.argv:
        extra()
    else:
        main()
