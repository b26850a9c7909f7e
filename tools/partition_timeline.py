"""Per-operation device time of one partitioned MuS-GNN rollout step (eager, no CUDA graph), rank 0's view: every step function
of PartitionedRollout bracketed by CUDA events, averaged over `--reps` steps after warm-up.  Shows how much of the step is halo
exchange (pack kernel + NCCL all_to_all_single) and how much is kernels, per level.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 \
        tools/partition_timeline.py [--nodes 1000000] [--reps 5] [--halo nccl|p2p]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--halo", default="nccl", choices=["nccl", "p2p"])
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    from graphs4cfd_b200.partition import PartitionedRollout
    g = M.build_mus_mesh(a.nodes, 6, M.auto_cells(a.nodes, 3), seed=0)
    eng = PartitionedRollout(init_params(mus_arch(a.hidden, 3), seed=0), g, rank, world, device=dev, cuda_graph=False, halo=a.halo)
    for _ in range(3):
        eng.step_only()
    torch.cuda.synchronize(dev)
    n = len(eng._steps)
    acc = [0.0] * n
    total = 0.0
    for _ in range(a.reps):
        dist.barrier()
        torch.cuda.synchronize(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for i, fn in enumerate(eng._steps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize(dev)
        for i in range(n):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
        total += ev[0].elapsed_time(ev[n])
    if rank == 0:
        by_kind = collections.defaultdict(float)
        print(f"# partitioned MuS-3 step, {a.nodes} nodes, hidden {a.hidden}, world {world}, halo {a.halo}, rank 0, eager (events between step functions), "
              f"{a.reps} steps averaged; step total {total / a.reps:.3f} ms")
        for lab, t in zip(eng.step_labels, acc):
            t /= a.reps
            by_kind[lab.split()[0] + (" " + lab.split()[2] if lab.startswith("mp") else "")] += t
            print(f"{t:8.3f} ms  {lab}")
        print("# by kind:")
        for kind, t in sorted(by_kind.items(), key=lambda kv: -kv[1]):
            print(f"#   {t:8.3f} ms  {100 * t * a.reps / total:5.1f} %  {kind}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
