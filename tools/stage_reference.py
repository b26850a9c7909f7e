#!/usr/bin/env python
"""Stage the UNMODIFIED reference Python package (no datasets, no optimiser state) under the git-ignored
``baseline/_ref/`` so that it travels to the GPU box with the repo snapshot (``/root/reference`` does not exist there).

    python tools/stage_reference.py            # run in the build container; __graft_entry__.build() calls it too

What is staged (about 0.3 MB of sources + 45 MB of weights):
  baseline/_ref/graphs4cfd/**.py     byte-for-byte copies of /root/reference/graphs4cfd/**.py
  baseline/_ref/weights/*.chk        {'arch', 'weights'} of the shipped 3S-GNN, RE3S-GNN, 2GS-GNN and 4GS-GNN checkpoints (what
                                     GNN(checkpoint=...) reads, nn/model.py:122-129; optimiser / scheduler state dropped)
  baseline/_ref/STAGED.json          sha256 of every staged source file next to the original's (the proof of "unmodified")

The reference cannot be pip-installed here: its build backend (flit_core, pyproject.toml:1-3) and its torch_geometric /
torch_cluster dependencies are absent from the image and there is no network; it is imported under oracle/pyg_stub.py.
Users: tests/test_gpu_dropin.py (drop-in proof on hardware), tests/test_gpu_long_rollout.py (trained-weight rollouts),
bench.py --impl reference (the reference's own classes on the host cores and, as `gpu_eager`, on the GPU).
Nothing under graphs4cfd_b200/ reads it.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
CHECKPOINTS = {"NsThreeScaleGNN.chk": "graphs4cfd/nn/weights/NsMuSGNN/NsThreeScaleGNN.chk",
               "NsRotEquiThreeScaleGNN.chk": "graphs4cfd/nn/weights/NsREMuSGNN/NsRotEquiThreeScaleGNN.chk",
               "NsTwoGuillardScaleGNN.chk": "graphs4cfd/nn/weights/NsMuGSGNN/NsTwoGuillardScaleGNN.chk",
               "NsFourGuillardScaleGNN.chk": "graphs4cfd/nn/weights/NsMuGSGNN/NsFourGuillardScaleGNN.chk"}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "graphs4cfd")):
        return False
    manifest = {"source": SRC, "files": {}, "checkpoints": {}}
    for base, _, files in os.walk(os.path.join(SRC, "graphs4cfd")):
        for f in files:
            if not f.endswith(".py"):
                continue
            src = os.path.join(base, f)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest["files"][rel] = {"sha256": sha(dst), "sha256_original": sha(src)}
    import torch
    os.makedirs(os.path.join(DST, "weights"), exist_ok=True)
    for name, rel in CHECKPOINTS.items():
        dst = os.path.join(DST, "weights", name)
        if not os.path.exists(dst):
            chk = torch.load(os.path.join(SRC, rel), map_location="cpu", weights_only=False)
            torch.save({"arch": chk["arch"], "weights": {k: v.clone() for k, v in chk["weights"].items()}}, dst)
        manifest["checkpoints"][name] = {"from": rel, "bytes": os.path.getsize(dst)}
    json.dump(manifest, open(os.path.join(DST, "STAGED.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest['files'])} source files and {len(manifest['checkpoints'])} checkpoints under {DST}")
    return True


if __name__ == "__main__":
    if not stage():
        print(f"{SRC} is not present on this machine: nothing staged", file=sys.stderr)
        sys.exit(1)
