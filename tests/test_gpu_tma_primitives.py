"""Hardware self tests of the bulk-tensor (TMA) copies behind the fixed-in-degree edge kernel (csrc/tma_test.cu,
csrc/mp_edge_v5.cu): tile load / tile store with the swizzles and read-back formulas the kernel uses."""
import pytest
import torch

from graphs4cfd_b200 import _lib as L

pytestmark = [pytest.mark.gpu]


def _call(test, src, rows, k, out, c0, j, n0):
    L.check(L.lib().g4c_debug_tma(test, src.data_ptr(), rows, k, out.data_ptr(), c0, j, n0, L.stream_ptr()))
    torch.cuda.synchronize()


def test_tile_load_swizzle64():
    dev = torch.device("cuda")
    rows, k = 100, 6
    src = torch.arange(rows * k * 128, device=dev, dtype=torch.float32).view(rows * k, 128)
    out = torch.full((32, 16), -1.0, device=dev)
    _call(0, src, rows, k, out, 32, 2, 40)
    want = src.view(rows, k, 128)[40:72, 2, 32:48]
    assert torch.equal(out, want)


def test_tile_load_past_the_end_is_zero_filled():
    dev = torch.device("cuda")
    rows, k = 50, 5
    src = torch.randn(rows * k, 128, device=dev)
    out = torch.full((32, 16), -1.0, device=dev)
    _call(3, src, rows, k, out, 112, 4, 40)              # rows 40..71 of 50
    want = torch.zeros(32, 16, device=dev)
    want[:10] = src.view(rows, k, 128)[40:50, 4, 112:128]
    assert torch.equal(out, want)


def test_tile_store_swizzle32():
    dev = torch.device("cuda")
    rows, k = 60, 6
    vals = torch.randn(32, 8, device=dev)
    out = torch.full((rows * k, 128), -7.0, device=dev)
    _call(1, vals, rows, k, out, 24, 3, 40)              # rows 40..71 of 60: the last 12 are clipped
    want = torch.full((rows, k, 128), -7.0, device=dev)
    want[40:60, 3, 24:32] = vals[:20]
    assert torch.equal(out.view(rows, k, 128), want)
