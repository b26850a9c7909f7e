"""Architecture dictionaries of the reference models (same keys and tuple forms as the ``arch`` dicts in
nn/mus_gnn.py:226-258, nn/remus_gnn.py:16-58 and nn/mugs_gnn.py:16-41) and seeded default-init parameters for them.
Used by the benchmark and the tests (no checkpoints can be downloaded; the shipped ones stay in the reference)."""
from collections import OrderedDict

import torch
from torch import nn

from .blocks import MLP


def mus_arch(H: int = 128, levels: int = 3, node_in: int = 5, nf: int = 3, adv: bool = False):
    mp = lambda: ((3 * H, (H, H, H), True), (2 * H, (H, H, H), True))
    down = lambda: (2 + H, (H, H, H), True)
    up = lambda: (2 + 2 * H, (H, H, H), True)
    a = OrderedDict()
    a["edge_encoder"] = (2, (H, H, H), False)
    a["node_encoder"] = (node_in, (H, H, H), False)
    n1 = 2 if adv else 4
    if levels == 1:
        for n in (["mp111", "mp112", "mp121", "mp122"] if adv else [f"mp1{i}" for i in range(1, 9)]):
            a[n] = mp()
    else:
        for i in range(1, n1 + 1):
            a[f"mp11{i}"] = mp()
        a["down_mp12"] = down()
        if levels == 2:
            for i in range(1, 5):
                a[f"mp2{i}"] = mp()
        else:
            a["mp211"], a["mp212"] = mp(), mp()
            a["down_mp23"] = down()
            if levels == 3:
                for i in range(1, 5):
                    a[f"mp3{i}"] = mp()
            else:
                a["mp311"], a["mp312"] = mp(), mp()
                a["down_mp34"] = down()
                for i in range(1, 5):
                    a[f"mp4{i}"] = mp()
                a["up_mp43"] = up()
                a["mp321"], a["mp322"] = mp(), mp()
            a["up_mp32"] = up()
            a["mp221"], a["mp222"] = mp(), mp()
        a["up_mp21"] = up()
        for i in range(1, n1 + 1):
            a[f"mp12{i}"] = mp()
    a["decoder"] = (H, (H, H, nf), False)
    return a


def mugs_arch(H: int = 128, levels: int = 2, node_in: int = 5, nf: int = 3):
    """MuGS-GNN (nn/mugs_gnn.py:16-41, 140-171, 302-340): the first block behind every up-sampling takes 2H-wide node features."""
    mp = lambda: ((3 * H, (H, H, H), True), (2 * H, (H, H, H), True))
    wide = lambda: ((5 * H, (H, H, H), True), (3 * H, (H, H, H), True))
    a = OrderedDict()
    for l in range(1, levels + 1):
        a["edge_encoder" + ("" if l == 1 else str(l))] = (2, (H, H, H), False)
    a["node_encoder"] = (node_in, (H, H, H), False)
    for i in range(1, 5):
        a[f"mp11{i}"] = mp()
    for l in range(2, levels):
        a[f"mp{l}11"], a[f"mp{l}12"] = mp(), mp()
    for i in range(1, 5):
        a[f"mp{levels}{i}"] = mp()
    for l in range(levels - 1, 1, -1):
        a[f"mp{l}21"], a[f"mp{l}22"] = wide(), mp()
    a["mp121"] = wide()
    for i in range(2, 5):
        a[f"mp12{i}"] = mp()
    a["decoder"] = (H, (H, H, nf), False)
    return a


def remus_arch(H: int = 128):
    mp = lambda: ((3 * H, (H, H), True), (2 * H, (H, H), True))
    a = OrderedDict()
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


def init_params(arch, seed: int = 0):
    """state_dict (CPU fp32) with torch's default nn.Linear / nn.LayerNorm initialisation, keys named as
    the reference model classes name them."""
    remus = any(k.startswith("angle_encoder") for k in arch)
    gen = torch.random.fork_rng()
    with gen:
        torch.manual_seed(seed)
        out = OrderedDict()

        def add(prefix, args):
            m = MLP(*args)
            for k, v in m.state_dict().items():
                out[f"{prefix}.{k}"] = v.detach().clone()

        for name, spec in arch.items():
            if name == "decoder":
                add("edge_decoder" if remus else "node_decoder", spec)
            elif isinstance(spec[0], tuple):
                first, second = ("angle_mlp", "edge_mlp") if remus else ("edge_mlp", "node_mlp")
                add(f"{name}.{first}", spec[0])
                add(f"{name}.{second}", spec[1])
            elif name.startswith("down_mp"):
                add(f"{name}.down_mlp", spec)
            elif name.startswith("up_mp"):
                add(f"{name}.up_mlp", spec)
            else:
                add(name, spec)
    return out
