"""Host side of the sparse heatmap transfer (okp_host_pack_tiles_f32, csrc/okp_host_pack.cpp; no GPU needed):
the packed tiles are what a NumPy restatement marks, and -- the claim the transfer rests on -- a map in which
everything outside the marked tiles is replaced by zero decodes (C oracle) to exactly the same tables."""
import ctypes

import numpy as np
import pytest

from object_keypoints_b200 import _lib

TH, TW = 4, 16


def pack(heat, threshold=0.5, capacity=None, threads=2):
    L = _lib.lib()
    heat = np.ascontiguousarray(heat, np.float32)
    maps, H, W = heat.shape
    tiles = -(-H // TH) * -(-W // TW)
    capacity = maps * tiles if capacity is None else capacity
    scratch = np.zeros(L.okp_host_pack_scratch_bytes(maps, H, W), np.uint8)
    offsets = np.zeros(maps + 1, np.int64)
    ids = np.full(max(capacity, 1), -1, np.int32)
    packed = np.full((max(capacity, 1), TH * TW), np.float32(7.0))
    count = ctypes.c_longlong()
    rc = L.okp_host_pack_tiles_f32(heat.ctypes.data, maps, H, W, ctypes.c_float(threshold), scratch.ctypes.data,
                                   offsets.ctypes.data, ids.ctypes.data, packed.ctypes.data, capacity, ctypes.byref(count), threads)
    assert rc == 0
    return int(count.value), ids, packed


def marked_tiles(heat, threshold=0.5):
    """NumPy restatement: tiles holding a value not <= tau, widened by one tile in every direction."""
    maps, H, W = heat.shape
    TY, TX = -(-H // TH), -(-W // TW)
    tau = np.float32(threshold) / np.float32(25.0) * (np.float32(1.0) - np.float32(1e-5))
    padded = np.zeros((maps, TY * TH, TX * TW), np.float32)
    padded[:, :H, :W] = heat
    active = ~(padded <= tau)
    active[:, H:, :] = False
    active[:, :, W:] = False
    raw = active.reshape(maps, TY, TH, TX, TW).any(axis=(2, 4))
    wide = np.zeros_like(raw)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            shifted = np.zeros_like(raw)
            ys = slice(max(dy, 0), TY + min(dy, 0)); yd = slice(max(-dy, 0), TY + min(-dy, 0))
            xs = slice(max(dx, 0), TX + min(dx, 0)); xd = slice(max(-dx, 0), TX + min(-dx, 0))
            shifted[:, yd, xd] = raw[:, ys, xs]
            wide |= shifted
    return wide, padded


def sparsified(heat, threshold=0.5):
    wide, padded = marked_tiles(heat, threshold)
    maps, H, W = heat.shape
    keep = np.repeat(np.repeat(wide, TH, axis=1), TW, axis=2)
    return np.where(keep, padded, np.float32(0.0))[:, :H, :W]


@pytest.mark.parametrize('shape', [(3, 64, 64), (2, 180, 320), (4, 37, 93), (1, 3, 5), (2, 8, 16)])
def test_packed_tiles_are_the_marked_tiles(shape):
    rng = np.random.default_rng(shape[1])
    heat = rng.uniform(0, 0.015, shape).astype(np.float32)
    maps, H, W = shape
    for m in range(maps):
        for _ in range(3):
            y, x = rng.integers(0, H), rng.integers(0, W)
            heat[m, y, x] = rng.uniform(0.03, 1.0)
    heat[0, H - 1, W - 1] = np.nan
    heat[maps - 1, 0, 0] = -5.0                                           # negative: not active by itself
    wide, padded = marked_tiles(heat)
    n, ids, packed = pack(heat)
    TY, TX = wide.shape[1:]
    want_ids = np.flatnonzero(wide.reshape(-1))
    assert n == len(want_ids)
    np.testing.assert_array_equal(ids[:n], want_ids)
    tiles = padded.reshape(maps, TY, TH, TX, TW).transpose(0, 1, 3, 2, 4).reshape(-1, TH * TW)
    np.testing.assert_array_equal(packed[:n].view(np.uint32), tiles[want_ids].view(np.uint32))
    assert (packed[n:] == 7.0).all() and (ids[n:] == -1).all()
    # too small a capacity: the count is still reported, nothing is written
    n2, ids2, packed2 = pack(heat, capacity=max(n - 1, 0))
    assert n2 == n and (packed2 == 7.0).all()


def test_zeroing_everything_outside_the_marked_tiles_does_not_change_the_decode():
    """The equivalence the sparse transfer rests on, checked with the C oracle on clean frames, the adversarial
    filter cases (ties, plateaus, borders, near-threshold sums) and maps with negative values."""
    from oracle import c_oracle
    from object_keypoints_b200 import synthetic
    from test_filter_bound import cases
    cfg = [1, 3]
    camera = synthetic.default_camera((64, 64))
    batch = synthetic.make_batch(24, cfg, (64, 64), seed=21, objects=(1, 3))
    heat = batch.heat.copy()
    rng = np.random.default_rng(5)
    for i, p in enumerate([p for p in cases().values() if p.shape == (64, 64)][:12]):
        heat[i, i % 3] = p
    heat[13, 1] += rng.uniform(-0.02, 0.0, (64, 64)).astype(np.float32)   # negative background around the blobs
    heat[14, 0, 20:24, 40:44] = 0.0199                                    # just below tau: sums stay below 0.5
    heat[15, 2, 30, 30] = 0.021                                           # just above tau, far from everything
    sparse = sparsified(heat.reshape(-1, 64, 64)).reshape(heat.shape)
    assert (sparse == 0).mean() > 0.4 and not np.array_equal(sparse, heat)
    want = c_oracle.decode(heat, batch.depth, batch.centers, cfg, camera, max_peaks=64)
    got = c_oracle.decode(sparse, batch.depth, batch.centers, cfg, camera, max_peaks=64)
    for key in want:
        np.testing.assert_array_equal(got[key].view(np.uint8), want[key].view(np.uint8), err_msg=key)
    assert want['n_objects'].sum() > 20
