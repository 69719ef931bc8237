"""CPU check of the arithmetic contract behind the strip kernel's candidate filter
(object_keypoints_b200/csrc/okp_peaks_strip.cuh): the kernel streams a separable float32 box sum
S~ and only runs the reference's exact raster-order sum S (perception/pipeline.py:70-71) where S~
cannot decide. This file restates the filter in NumPy float32, operation for operation, and
checks on adversarial inputs that (1) the bound |S - S~| <= 2.9e-6 * S~ the kernel relies on holds,
(2) no true peak is ever filtered out, and (3) filter + exact check reproduce the oracle's peak set.
The CUDA path itself is compared with the oracle in test_gpu_decode.py."""
import numpy as np
import pytest

from oracle import np_oracle

TIE = np.float32(1.00002)            # OKP_STRIP_TIE
SLACK = np.float32(1e-5)             # OKP_STRIP_THRESHOLD_SLACK
f32 = np.float32


def separable_sum(p):
    """S~ exactly as okp_strip_step computes it: horizontal 5-sums through the shared partial sums
    of a 4-pixel strip (strip s = pixels 4s-2 .. 4s+1, window columns 4s-4 .. 4s+3), then
    P(t-3) + P(t-1) + h(t) with P(t) = h(t-1) + h(t) down the rows. Zero outside the image."""
    H, W = p.shape
    strips = W // 4 + 1
    padded = np.zeros((H + 8, 4 * strips + 4), np.float32)          # rows -4.., columns -4..
    padded[4:4 + H, 4:4 + W] = p
    w = [padded[:, 4 * np.arange(strips) + k] for k in range(8)]     # w[k][row, strip]
    c34 = w[3] + w[4]
    t12 = w[1] + w[2]
    t56 = w[5] + w[6]
    tc = t12 + c34
    u = c34 + t56
    h = np.stack([w[0] + tc, tc + w[5], w[2] + u, u + w[7]], axis=-1)  # [row, strip, c]: pixel x = 4s + c - 2
    h = h.reshape(H + 8, 4 * strips)                                   # column index = x + 2
    rows = h.shape[0]
    P = np.zeros_like(h)
    P[1:] = h[:-1] + h[1:]
    out = np.zeros((H, W), np.float32)
    for y in range(H):
        t = y + 2 + 4                                                  # newest input row of box y, in padded rows
        assert t < rows
        v = (P[t - 3] + P[t - 1]) + h[t]
        out[y] = v[2:2 + W]
    return out


def filtered_peaks(p, threshold=0.5):
    """Candidate filter on S~ followed by the exact check, like the candidate warps do it."""
    H, W = p.shape
    approx = separable_sum(p)
    exact = np_oracle.box_sum(p)
    thr = f32(threshold)
    thr_lo = f32(thr - SLACK * np.abs(thr))
    peaks, candidates = [], 0
    for y in range(H):
        for x in range(W):
            sp = approx[y, x]
            if not sp > thr_lo:
                continue
            y0, y1, x0, x1 = max(y - 2, 0), min(y + 3, H), max(x - 2, 0), min(x + 3, W)
            block = approx[y0:y1, x0:x1]
            if (block > f32(sp * TIE)).any():
                continue
            candidates += 1
            s = exact[y, x]
            if not s > thr:
                continue
            ties = (f32(1) * block * TIE >= sp)
            if (exact[y0:y1, x0:x1][ties] > s).any():
                continue
            peaks.append((y, x))
    return peaks, candidates, approx, exact


def cases():
    rng = np.random.default_rng(42)
    out = {}
    out['dense_noise'] = rng.uniform(0, 0.2, (40, 64)).astype(np.float32)
    q = rng.uniform(0, 0.2, (40, 64)).astype(np.float32)
    out['quantised_ties'] = np.round(q * 8) / 8
    out['random_init_net'] = (0.5007 + rng.uniform(0, 0.0027, (64, 64))).astype(np.float32)
    plateau = rng.uniform(0, 0.01, (48, 48)).astype(np.float32)
    plateau[10:30, 12:33] = 1.0
    out['saturated_plateau'] = plateau
    yy, xx = np.mgrid[0:32, 0:44].astype(np.float32)
    blob = np.exp(-((yy - 15.5) ** 2 + (xx - 20.5) ** 2) / 4.0).astype(np.float32)              # centred on a half pixel
    out['half_pixel_blob'] = blob
    ulp = blob.copy()
    ulp[15, 20] = np.nextafter(ulp[15, 20], f32(2))                                           # 1-ulp asymmetry
    out['one_ulp_tie_break'] = ulp
    out['tiny_values'] = (rng.uniform(0, 1, (16, 20)) * 1e-38).astype(np.float32)             # denormal sums
    out['border_blobs'] = np.zeros((12, 16), np.float32)
    out['border_blobs'][0, 0] = out['border_blobs'][11, 15] = out['border_blobs'][0, 15] = 0.9
    return out


@pytest.mark.parametrize('name', sorted(cases()))
def test_filter_plus_exact_check_reproduces_the_oracle(name):
    p = cases()[name]
    threshold = 0.0 if name == 'tiny_values' else 0.5
    got, candidates, approx, exact = filtered_peaks(p, threshold)
    want = [tuple(v) for v in np_oracle.find_peaks(p, threshold=threshold)[0]]
    assert got == want
    # the bound itself, with its safety factor: gamma_24 on both sums
    assert (np.abs(exact.astype(np.float64) - approx.astype(np.float64)) <= 2.9e-6 * approx.astype(np.float64)).all()
    # the filter is worth having: on sparse maps candidates are (nearly) only the peaks
    if name in ('half_pixel_blob', 'one_ulp_tie_break', 'border_blobs'):
        assert candidates <= len(want) + 4


def test_separable_sum_is_a_reordering_of_the_same_25_terms():
    rng = np.random.default_rng(0)
    p = (rng.integers(0, 64, (20, 24)) / 64.0).astype(np.float32)     # exactly representable sums: any order agrees
    np.testing.assert_array_equal(separable_sum(p), np_oracle.box_sum(p))
