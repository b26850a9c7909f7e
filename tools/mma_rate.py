"""Scratch: tcgen05.mma issue rate of the CTA-pair M=256 N=128 K=16 fp16 MMA with the A operand in TMEM (csrc/tc2_test.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphs4cfd_b200 import ops
dev = torch.device("cuda")
A = torch.randn(256, 128, device=dev) * 0.01
W = torch.randn(128, 128, device=dev) * 0.01
for name, fl in (("one accumulator (chained)", 4), ("two accumulators (alternating)", 4 | 8),
                 ("chained + 16 warps tcgen05.ld", 4 | 16), ("chained + 16 warps tcgen05.st", 4 | 32),
                 ("chained + 16 warps ld.shared.v4", 4 | 64), ("chained + ld + st + lds", 4 | 16 | 32 | 64)):
    for reps in (400,):
        D = ops.debug_tc2(3, A, W, flags=fl | (reps << 8))
        torch.cuda.synchronize()
        print(f"{name}, {24 * reps} MMAs: {float(D[0, 0]):.1f} cycles per MMA (64 at the dense fp16 peak)")
