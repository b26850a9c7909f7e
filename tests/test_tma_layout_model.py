"""Shared-memory layout arithmetic of csrc/mp_edge_pair_tma.cu against the canonical definition of the TMA swizzles.

The hardware's SWIZZLE_{32,64,128}B modes are Swizzle<B,4,3> on the shared-memory byte address (B = 1, 2, 3): the 16-byte
chunk index (address bits 4..4+B-1) is XORed with address bits 7..7+B-1.  The kernel writes / reads the same tiles with
hand-written address formulas (transcribed below from the .cu); this test checks, on the CPU, that every formula agrees
with that definition, i.e. that a tile written by one side is read correctly by the other.  (The hardware behaviour itself is
checked by tests/test_gpu_tma_primitives.py.)"""
import numpy as np


def swizzle(addr, bits):
    """Swizzle<bits,4,3>: physical byte address of logical byte address `addr` (tile base aligned to 128 << bits... at least 1 KiB)."""
    return addr ^ (((addr >> 7) & ((1 << bits) - 1)) << 4)


def tma_tile_image(tile, row_bytes, bits):
    """bytes of a [rows, row_bytes] tile as the TMA engine lays it out in shared memory with the given swizzle"""
    rows = tile.shape[0]
    img = np.full(rows * row_bytes, -1, dtype=np.int64)
    for r in range(rows):
        for b in range(row_bytes):
            img[swizzle(r * row_bytes + b, bits)] = tile[r, b]
    assert (img >= 0).all()
    return img


def test_loader_readback_matches_swizzle64_tile():
    # a 32-row x 64-byte tile (16 fp32 columns) delivered by TMA with SWIZZLE_64B; loader lane = row reads chunk c at
    #   st + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)                        (mp_edge_pair_tma.cu, read-back loop)
    tile = np.arange(32 * 64).reshape(32, 64)
    img = tma_tile_image(tile, 64, 2)
    for lane in range(32):
        swz = (lane >> 1) & 3
        for c in range(4):
            off = lane * 64 + ((c ^ swz) << 4)
            assert (img[off:off + 16] == tile[lane, 16 * c:16 * c + 16]).all()


def test_cp_async_destination_matches_the_same_layout():
    # cp.async lanes: sub = lane >> 2 (row within a group of 8), piece = lane & 3 (16-byte chunk); rows 8 i + sub, i = 0..3:
    #   dst = stage + sub * 64 + ((piece ^ ((sub >> 1) & 3)) << 4) + 512 i    (dst_off and the "+ 512 i" offsets of the asm block)
    tile = np.arange(32 * 64).reshape(32, 64)
    img = tma_tile_image(tile, 64, 2)
    mine = np.full(32 * 64, -1, dtype=np.int64)
    for lane in range(32):
        sub, piece = lane >> 2, lane & 3
        for i in range(4):
            dst = sub * 64 + ((piece ^ ((sub >> 1) & 3)) << 4) + 512 * i
            mine[dst:dst + 16] = tile[8 * i + sub, 16 * piece:16 * piece + 16]
    assert (mine == img).all()


def test_gather4_destination_matches_the_same_layout():
    # lanes 0-7 issue one gather4 each: four 64-byte row pieces to dst = array + 256 * lane, swizzled by ADDRESS like any tile
    tile = np.arange(32 * 64).reshape(32, 64)
    img = tma_tile_image(tile, 64, 2)
    mine = np.full(32 * 64, -1, dtype=np.int64)
    for lane in range(8):
        for i in range(4):
            for b in range(64):
                mine[swizzle(256 * lane + 64 * i + b, 2)] = tile[4 * lane + i, b]
    assert (mine == img).all()


def test_epilogue_staging_matches_swizzle32_tile():
    # e' staging: 32 rows x 32 bytes (8 fp32 columns), SWIZZLE_32B; thread lane = row writes columns 0-3 at
    #   ost0 = base + lane * 32 + (((lane >> 2) & 1) << 4)   and columns 4-7 at ost0 ^ 16
    tile = np.arange(32 * 32).reshape(32, 32)
    img = tma_tile_image(tile, 32, 1)
    mine = np.full(32 * 32, -1, dtype=np.int64)
    for lane in range(32):
        ost0 = lane * 32 + (((lane >> 2) & 1) << 4)
        mine[ost0:ost0 + 16] = tile[lane, 0:16]
        mine[(ost0 ^ 16):(ost0 ^ 16) + 16] = tile[lane, 16:32]
    assert (mine == img).all()


def test_stage_and_staging_offsets_keep_the_swizzle_phase():
    # the swizzles act on absolute shared-memory addresses: every tile base must be a multiple of the pattern's period
    # (512 B for SWIZZLE_64B, 256 B for SWIZZLE_32B).  Offsets of struct Smem in mp_edge_pair_tma.cu:
    ring, arr, stg, n_warps, n_stg = 96 * 1024, 2048, 6144, 8, 2
    for w in range(n_warps):
        for s in range(n_stg):
            for a in range(3):
                assert (ring + (w * n_stg + s) * stg + a * arr) % 512 == 0
    ostage = ring + n_warps * n_stg * stg
    assert ostage == 192 * 1024
    for w in range(16):
        assert (ostage + w * 1024) % 256 == 0 and (ostage + w * 1024) % 1024 == 0      # "& ~1023" recovers the tile base
    # two-layer MLPs double-buffer the staging in the unused third-layer weight space: w[2] = 2 * 4 * HIMG bytes into the struct
    w2 = 2 * 4 * 64 * 128
    for w in range(16):
        for buf in range(2):
            assert (w2 + w * 2048 + buf * 1024) % 1024 == 0
    assert w2 + 16 * 2048 == 3 * 4 * 64 * 128        # exactly fills w[2]
