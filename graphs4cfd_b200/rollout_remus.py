"""REMuS-GNN plan for the rollout engine (nn/remus_gnn.py:119-199) — filled in by plan_remus()."""


def plan_remus(engine, graph):
    raise NotImplementedError("REMuS rollout plan not built yet")
