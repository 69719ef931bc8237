"""Target rasterisation: okp_rasterise_targets_f32 and its NumPy oracle against the
UNMODIFIED reference dataset code (perception/datasets/video.py:44-53,195-263; fixture
tests/golden/targets_64.npz from oracle/make_goldens.py::targets_case). Centre-vector and depth maps are
bit-exact; the heatmaps agree to 2e-7 (one float32 ulp: the reference's exp is numba's, not libm's)."""
import numpy as np
import pytest

from helpers import load_golden

TOL_HEAT = 2e-7


def test_oracle_matches_the_reference_dataset_code():
    from oracle import np_oracle
    g = load_golden('targets_64.npz')
    for n in range(g['keypoints'].shape[0]):
        heat, centers, depth = np_oracle.rasterise_targets(g['keypoints'][n], g['depths'][n], list(g['keypoint_config']),
                                                           tuple(g['size']))
        np.testing.assert_array_equal(centers, g['ref_centers'][n])
        np.testing.assert_array_equal(depth, g['ref_depth'][n])
        assert np.abs(heat - g['ref_heat'][n]).max() <= TOL_HEAT


@pytest.mark.gpu
def test_cuda_targets_match_reference_and_decode_back():
    import torch
    from oracle import np_oracle
    from object_keypoints_b200 import targets, KeypointDecoder, synthetic
    g = load_golden('targets_64.npz')
    cfg, size = list(g['keypoint_config']), tuple(int(v) for v in g['size'])
    heat, depth, centers = targets.rasterise_targets(g['keypoints'], g['depths'], cfg, size)
    np.testing.assert_array_equal(centers.cpu().numpy(), g['ref_centers'])
    np.testing.assert_array_equal(depth.cpu().numpy(), g['ref_depth'])
    assert np.abs(heat.cpu().numpy() - g['ref_heat']).max() <= TOL_HEAT
    # n_objects limits the objects drawn per frame; other shapes against the oracle
    rng = np.random.default_rng(3)
    for cfg, size, G in (([1, 1, 1], (64, 64), 4), ([2], (48, 80), 2), ([1, 3], (180, 320), 8)):
        Kp = 1 + sum(cfg)
        N = 3
        kp = rng.uniform(-4, 1, (N, G, Kp, 2)) + rng.uniform(0, 1, (N, G, 1, 2)) * np.array([size[1], size[0]])
        z = rng.uniform(0.4, 1.5, (N, G, Kp))
        counts = rng.integers(0, G + 1, N).astype(np.int32)
        heat, depth, centers = targets.rasterise_targets(kp, z, cfg, size, n_objects=counts)
        for n in range(N):
            want = np_oracle.rasterise_targets(kp[n, :counts[n]], z[n, :counts[n]], cfg, size)
            np.testing.assert_array_equal(centers[n].cpu().numpy(), want[1])
            np.testing.assert_array_equal(depth[n].cpu().numpy(), want[2])
            assert np.abs(heat[n].cpu().numpy() - want[0]).max() <= TOL_HEAT
    # round trip: targets of well separated objects decode back to their keypoints (config 1 style)
    cfg, size = [1, 3], (64, 64)
    kp = np.zeros((2, 2, 5, 2))
    for n in range(2):
        for obj, centre in enumerate(([18.0, 20.0], [46.0, 44.0])):
            kp[n, obj, 1:] = np.array(centre) + np.array([[0.3, -7.0], [-7.2, 0.4], [6.8, 0.9], [0.2, 7.1]]) + 0.37 * n
            kp[n, obj, 0] = kp[n, obj, 1:].mean(axis=0)
    z = np.full((2, 2, 5), 0.8)
    heat, depth, centers = targets.rasterise_targets(kp, z, cfg, size)
    decoder = KeypointDecoder(cfg, size, camera=synthetic.default_camera(size))
    t = decoder.decode_batch(heat, depth, centers).numpy()
    assert (t['n_objects'] == 2).all() and (t['kp_count'][:, :2] == np.array([1, 1, 3])).all()
    assert np.abs(t['kp_xy'][:, :2, 0, 0] - kp[:, :, 0]).max() < 0.75     # centroid in pixel-index coordinates
