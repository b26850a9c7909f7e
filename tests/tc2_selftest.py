"""Hardware self tests of csrc/tc2_core.cuh, one per process (a failing tcgen05 kernel traps and poisons
the CUDA context, so tests/test_gpu_tc2.py runs each case in its own interpreter).

    python tests/tc2_selftest.py <case>      cases: ts_st, cp_roundtrip, cp_gemm, pair_st, pair_cp
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphs4cfd_b200 import ops  # noqa: E402


def main(case):
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    rows = 256 if case.startswith("pair") else 128
    A = torch.randn(rows, 128, device=dev)
    W = torch.randn(128, 128, device=dev) * 0.1
    P = torch.randn(rows, 128, device=dev)
    ref = (A.double() @ W.double().t())
    if case == "ts_st":
        D = ops.debug_tc2(1, A, W)
    elif case == "cp_roundtrip":
        D = ops.debug_tc2(2, A, W, P, flags=1)
        ref = P.double()
    elif case == "cp_gemm":
        D = ops.debug_tc2(2, A, W, P)
        ref = ref + P.double()
    elif case == "pair_st":
        D = ops.debug_tc2(3, A, W)
    elif case == "pair_cp":
        D = ops.debug_tc2(3, A, W, flags=1)
    else:
        raise SystemExit(f"unknown case {case}")
    torch.cuda.synchronize()
    err = float((D.double() - ref).norm() / ref.norm())
    print(f"tc2_selftest {case}: rel-L2 = {err:.3e}")
    if not err < 2e-6:
        bad = (D.double() - ref).abs() > 1e-3
        print("  mismatching rows:", bad.any(1).nonzero().flatten()[:16].tolist(), " cols:", bad.any(0).nonzero().flatten()[:16].tolist())
        raise SystemExit(1)


if __name__ == "__main__":
    main(sys.argv[1])
