"""torchrun entry (one rank per GPU): PartitionedRollout (G4C_MODEL=remus: PartitionedRemusRollout) vs the single-domain oracle and vs single-GPU Rollout.
Launched by tests/test_gpu_multi.py:  python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from graphs4cfd_b200 import Rollout
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    from graphs4cfd_b200.partition import PartitionedRollout
    from oracle import restate as R
    precision = os.environ.get("G4C_PRECISION", "fp32")
    H = 128 if precision != "fp32" else 32
    steps = 3
    if os.environ.get("G4C_MODEL", "mus") == "remus":        # edge-halo partition of the REMuS-GNN (partition_remus.py)
        from graphs4cfd_b200.archs import remus_arch
        from graphs4cfd_b200.partition_remus import PartitionedRemusRollout
        n = 4000
        g = M.build_remus_mesh(n, 6, seed=5)
        params = init_params(remus_arch(H), seed=2)
        eng = PartitionedRemusRollout(params, g, rank, world, precision=precision, device=dev,
                                      cuda_graph=os.environ.get("G4C_GRAPH", "0") == "1", halo=os.environ.get("G4C_HALO", "auto"))
    else:
        n = 6000
        g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=5)
        params = init_params(mus_arch(H, 3), seed=2)
        eng = PartitionedRollout(params, g, rank, world, precision=precision, device=dev,
                                 cuda_graph=os.environ.get("G4C_GRAPH", "0") == "1", halo=os.environ.get("G4C_HALO", "auto"))
    out = eng.gather(eng.solve(steps), n).cpu()
    ok = True
    if rank == 0:
        want = R.solve(params, g.clone(), steps)
        rel = float((out - want).norm() / want.norm())
        single = Rollout(params, g, precision=precision, device=dev).solve(steps).cpu()
        rel1 = float((out - single).norm() / single.norm())
        print(f"world={world} precision={precision} halo={getattr(eng, 'halo', 'nccl')} graph={os.environ.get('G4C_GRAPH', '0')}: "
              f"rel-L2 vs oracle {rel:.3e}, vs single-GPU engine {rel1:.3e}, exchanges/step={eng.exchanges_per_step}")
        ok = rel <= (5e-5 if precision == "fp32" else 2e-4) and rel1 <= (5e-6 if precision == "fp32" else 2e-4)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
