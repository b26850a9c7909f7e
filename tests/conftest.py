import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the read-only tree of the build container, or the byte-for-byte copy staged by tools/stage_reference.py (travels to the GPU box)
HAVE_REFERENCE = os.path.isdir("/root/reference/graphs4cfd") or os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "graphs4cfd"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference tree (/root/reference or its staged copy baseline/_ref)")


def pytest_collection_modifyitems(config, items):
    skip_ref = pytest.mark.skip(reason="no reference tree on this machine (python tools/stage_reference.py)")
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "reference" in item.keywords and not HAVE_REFERENCE:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords and not torch.cuda.is_available():
            item.add_marker(skip_gpu)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def mesh_from(d):
    from graphs4cfd_b200.mesh import Mesh
    return Mesh(**{k: v.clone() for k, v in d.items()})


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))


@pytest.fixture
def golden():
    return load_golden


def shipped_model(gfd, kind, device="cpu"):
    """The reference's own model class with the reference's trained weights: kind "mus3" = 3S-GNN-NsCircle-v1
    (nn/mus_gnn.py:267), "remus" = RE3S-GNN-NsEllipse-v1 (nn/remus_gnn.py:66).  Loaded from the weights-only copies staged
    by tools/stage_reference.py when they exist (the GPU box), else from the reference tree itself."""
    from oracle.pyg_stub import staged_checkpoint
    cls, chk, name = {"mus3": (gfd.nn.NsThreeScaleGNN, "NsThreeScaleGNN.chk", "3S-GNN-NsCircle-v1"),
                      "remus": (gfd.nn.NsRotEquiTreeScaleGNN, "NsRotEquiThreeScaleGNN.chk", "RE3S-GNN-NsEllipse-v1")}[kind]
    path = staged_checkpoint(chk)
    dev = torch.device(device)
    return cls(checkpoint=path, device=dev) if path else cls(model=name, device=dev)
