"""CUDA engine of the partitioned REMuS-GNN rollout (graphs4cfd_b200/partition_remus.py).  world = 1 runs the rank plan
and the step program through the libg4c kernels on one GPU (no exchange); the NCCL case needs >= 2 visible GPUs.
(The exchange program itself is covered on CPU by tests/test_partition_remus_gloo.py.)"""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden, mesh_from, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_world1_rollout_golden(cuda_graph):
    from graphs4cfd_b200.partition_remus import PartitionedRemusRollout
    d = load_golden("model_remus_h32")
    g = mesh_from(d["mesh"])
    eng = PartitionedRemusRollout(d["params"], g, rank=0, world=1, cuda_graph=cuda_graph)
    out = eng.gather(eng.solve(d["n_out"]), g.num_nodes)
    assert out.shape == d["out"].shape
    assert rel_l2(out.cpu(), d["out"]) <= 5e-5, rel_l2(out.cpu(), d["out"])


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-5), ("fp16x3", 1e-4)])
def test_world1_rollout_h128_vs_oracle(precision, tol):
    import graphs4cfd_b200 as g4
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, remus_arch
    from graphs4cfd_b200.partition_remus import PartitionedRemusRollout
    from oracle import restate as R
    g = M.build_remus_mesh(1500, 6, seed=2)
    params = init_params(remus_arch(128), seed=3)
    want = R.solve(params, g.clone(), 2)
    eng = PartitionedRemusRollout(params, g, rank=0, world=1, precision=precision)
    got = eng.gather(eng.solve(2), g.num_nodes).cpu()
    assert rel_l2(got, want) <= tol, rel_l2(got, want)
    single = g4.Rollout(params, g, precision=precision).solve(2).cpu()
    assert rel_l2(got, single) <= tol, rel_l2(got, single)


@pytest.mark.parametrize("halo,graph", [("nccl", "0"), ("p2p", "1")])
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_partitioned_remus_rollout_nccl(precision, halo, graph):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, G4C_PRECISION=precision, G4C_MODEL="remus", G4C_HALO=halo, G4C_GRAPH=graph)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0
