"""The arithmetic of the tensor-core path (fp16 (hi, lo) split of both operands, three products, fp32
accumulation), emulated on the CPU through the oracle: a short rollout stays at the fp32 noise floor, where single-pass
TF32 operands do not.  (tools/precision_emulation.py runs the long version with the reference's trained checkpoint.)"""
import os
import sys

import torch

from conftest import ROOT, rel_l2

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_fp16x3_emulation_tracks_fp32_over_a_short_rollout():
    import precision_emulation as PE
    from graphs4cfd_b200 import mesh as M
    from graphs4cfd_b200.archs import init_params, mus_arch
    n, steps = 1200, 3
    g = M.build_mus_mesh(n, 6, M.auto_cells(n, 3), seed=1)
    params = init_params(mus_arch(128, 3), seed=1)
    ref = PE.rollout(params, g.clone(), steps)
    with PE.patched_linear(PE.make_linear("fp16x3")):
        x3 = PE.rollout(params, g.clone(), steps)
    with PE.patched_linear(PE.make_linear("tf32")):
        tf = PE.rollout(params, g.clone(), steps)
    e3, et = rel_l2(x3[-1], ref[-1]), rel_l2(tf[-1], ref[-1])
    assert e3 <= 5e-6, e3
    assert e3 <= 0.05 * et, (e3, et)


def test_split_is_exact_to_22_bits():
    import precision_emulation as PE
    x = torch.randn(4096) * 3
    hi, lo = PE.split16(x)
    assert float(((hi + lo) - x).abs().max() / x.abs().max()) <= 2.0 ** -21
