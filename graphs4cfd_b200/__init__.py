"""graphs4cfd_b200 — B200-native (sm_100a) message-passing hot path for graphs4cfd models.

    from graphs4cfd_b200 import accelerate, Rollout
    accelerate(model)                      # swap the reference blocks for fused CUDA blocks, in place
    out = Rollout(model, graph).solve(100) # whole rollout on the device, CUDA-graphed time step

See DESIGN.md for the kernels and INTEGRATION.md for the binding into the reference package."""
from . import mesh  # noqa: F401
from .blocks import (MLP, MP, DownEdgeMP, DownMP, EdgeMP, GNBlock, UpEdgeMP, UpMP, accelerate,  # noqa: F401
                     edgeScalarToNodeVector, patch_reference)
from .rollout import Rollout  # noqa: F401

__all__ = ["MLP", "GNBlock", "MP", "DownMP", "UpMP", "EdgeMP", "DownEdgeMP", "UpEdgeMP", "edgeScalarToNodeVector",
           "accelerate", "patch_reference", "Rollout", "mesh"]
