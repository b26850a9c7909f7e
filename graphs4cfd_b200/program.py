"""Block program of a graphs4cfd model, derived from its state_dict.

Every reference ``load_arch`` registers its blocks in execution order (nn/mus_gnn.py:274-310,
nn/remus_gnn.py:75-117), so the ordered top-level prefixes of the state_dict ARE the per-step
sequence of SURVEY.md Appendix A; the sub-module names give the block kind."""
from typing import Dict, List, Tuple

import torch


def top_level_names(params: Dict[str, torch.Tensor]) -> List[str]:
    names: List[str] = []
    for key in params:
        top = key.split(".", 1)[0]
        if top not in names:
            names.append(top)
    return names


def is_remus(params) -> bool:
    return any(k.startswith("angle_encoder") for k in params)


def block_program(params: Dict[str, torch.Tensor]) -> List[Tuple[str, str]]:
    """[(block name, kind)], kind in {mlp, mp, down, up, edge_mp, down_edge, up_edge}."""
    remus = is_remus(params)
    prog = []
    for name in top_level_names(params):
        subs = {k.split(".")[1] for k in params if k.startswith(name + ".")}
        if "MLP" in subs:
            kind = "mlp"
        elif {"edge_mlp", "node_mlp"} <= subs:
            kind = "mp"
        elif "down_mlp" in subs:
            kind = "down"
        elif "up_mlp" in subs:
            kind = "up_edge" if remus else "up"
        elif {"angle_mlp", "edge_mlp"} <= subs:
            kind = "down_edge" if name.startswith("down") else "edge_mp"
        else:
            raise ValueError(f"unrecognised block {name!r} with sub-modules {sorted(subs)}")
        prog.append((name, kind))
    return prog


def hidden_width(params) -> int:
    for k, v in params.items():
        if k.endswith("MLP.linear_1.weight"):
            return int(v.shape[0])
    raise ValueError("no MLP in state_dict")
