"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): node-partitioned rollout over NCCL."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("halo,graph", [("nccl", "0"), ("p2p", "0"), ("p2p", "1")])
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_partitioned_rollout_matches_single_gpu(precision, halo, graph):
    """halo = nccl: pack kernel + all_to_all_single; p2p: one kernel over NVLink peer memory (g4c_halo_put), eager and
    captured in the step's CUDA graph."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, G4C_PRECISION=precision, G4C_HALO=halo, G4C_GRAPH=graph)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0


def test_blocks_follow_their_tensors_device():
    """A model on cuda:1 while cuda:0 is PyTorch's current device: kernels, their attributes and tensor maps must go to the
    tensors' device and stream (the launch helper makes that device current for the call)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import graphs4cfd_b200 as g4
    from conftest import load_golden, mesh_from, rel_l2
    torch.cuda.set_device(0)
    d = load_golden("model_ns3_h32")
    out = g4.Rollout(d["params"], mesh_from(d["mesh"]), device="cuda:1").solve(d["n_out"])
    assert out.device.index == 1 and rel_l2(out.cpu(), d["out"]) <= 1e-5
    d = load_golden("mp_trained_h128")
    blk = g4.MP((384, (128, 128, 128), True), (256, (128, 128, 128), True))
    blk.load_state_dict({k[3:]: v for k, v in d["params"].items()})
    blk = blk.to("cuda:1")
    with torch.no_grad():
        v, e = blk(d["v"].to("cuda:1"), d["e"].to("cuda:1"), d["edge_index"].to("cuda:1"))
    assert torch.cuda.current_device() == 0 and v.device.index == 1 and torch.isfinite(v).all() and torch.isfinite(e).all()
    with pytest.raises(RuntimeError, match="different devices"):
        blk(d["v"].to("cuda:0"), d["e"].to("cuda:1"), d["edge_index"].to("cuda:1"))
