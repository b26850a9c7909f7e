"""TMEM-operand MMA, tcgen05.cp and CTA-pair (cta_group::2) primitives against an fp64 matmul.
Each case runs in its own process (see tests/tc2_selftest.py).  Tolerance: the 3-term fp16 split keeps
22 significant bits per operand -> rel-L2 < 2e-6 on a K = 128 product."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ts_st", "cp_roundtrip", "cp_gemm", "pair_st", "pair_cp"])
def test_tc2_primitive(case):
    r = subprocess.run([sys.executable, os.path.join(HERE, "tc2_selftest.py"), case], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
