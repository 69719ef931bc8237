"""The oracle is only worth something if it is pinned to the reference: these tests compare
both restatements (NumPy, C) with fixtures produced by the UNMODIFIED reference pipeline
(oracle/make_goldens.py) and with each other. CPU only."""
import numpy as np
import pytest

from helpers import (load_golden, golden_camera, reference_tables, assert_tables_match)
from oracle import np_oracle, c_oracle

DECODE_FIXTURES = ['valve_64.npz', 'cups_64.npz', 'valve_grid_180x320.npz', 'test_pipeline_180x320.npz',
                   'adversarial_64.npz', 'nan_64.npz']


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def test_box_sum_is_bitwise_torch_conv2d():
    """Summation order contract: 25 sequential float32 adds == torch CPU conv2d with ones(5,5)."""
    g = load_golden('boxsum.npz')
    for m, s in zip(g['maps'], g['sums']):
        assert np.array_equal(bits(np_oracle.box_sum(m)), bits(s))
    assert np.array_equal(bits(np_oracle.box_sum(g['big'][0])), bits(g['big_sums'][0]))


@pytest.mark.parametrize('name', DECODE_FIXTURES)
def test_numpy_oracle_matches_reference(name):
    g = load_golden(name)
    cam = np_oracle.camera_dict(golden_camera(g))
    got = np_oracle.decode(g['heat'], g['depth'], g['centers'], list(g['keypoint_config']), cam)
    assert_tables_match(got, reference_tables(g))


@pytest.mark.parametrize('name', DECODE_FIXTURES)
def test_c_oracle_matches_reference_and_numpy(name):
    g = load_golden(name)
    camera = golden_camera(g)
    cfg = list(g['keypoint_config'])
    got = c_oracle.decode(g['heat'], g['depth'], g['centers'], cfg, camera)
    assert_tables_match(got, reference_tables(g))
    again = np_oracle.decode(g['heat'], g['depth'], g['centers'], cfg, np_oracle.camera_dict(camera))
    for key, value in got.items():
        assert np.array_equal(bits(value), bits(again[key])), f"C and NumPy oracle differ on {key}"


def test_adversarial_flags_and_counts():
    g = load_golden('adversarial_64.npz')
    names = [str(n) for n in g['names']]
    got = c_oracle.decode(g['heat'], g['depth'], g['centers'], list(g['keypoint_config']), golden_camera(g))
    ref = reference_tables(g)
    by = {n: i for i, n in enumerate(names)}
    assert got['peak_count'][by['half_pixel_tie'], 0] == 2          # both tied pixels kept
    assert got['peak_count'][by['plateau'], 0] == 4
    assert got['peak_count'][by['constant_patch'], 0] == 16         # every pixel of the plateau
    assert got['flags'][by['no_centres']] & np_oracle.FLAG_NO_CENTERS
    assert got['flags'][by['all_zero']] & np_oracle.FLAG_NO_CENTERS
    assert got['flags'][by['argmax_resolution']] & np_oracle.FLAG_ARGMAX_RESOLVED
    # the reference prints one line per skipped vote (pipeline.py:123)
    skipped = ((got['peak_object'][:, 1:] == -1) &
               (np.arange(got['peak_object'].shape[2])[None, None] < got['peak_count'][:, 1:, None])).sum(axis=(1, 2))
    has_objects = got['n_objects'] > 0
    np.testing.assert_array_equal(skipped[has_objects], ref['n_skipped'][has_objects])
    assert (got['flags'][ref['n_skipped'] > 0] & np_oracle.FLAG_OUTLIER_SKIPPED).all()


def test_centroid_is_much_tighter_than_tolerance():
    g = load_golden('valve_64.npz')
    got = c_oracle.decode(g['heat'], g['depth'], g['centers'], list(g['keypoint_config']), golden_camera(g))
    ref = reference_tables(g)
    assert np.abs(got['peak_xy'] - ref['peak_xy']).max() < 5e-5


def test_geometry_against_opencv_goldens():
    from object_keypoints_b200 import camera_utils
    g = load_golden('geometry.npz')
    left = camera_utils.FisheyeCamera(g['K_left'], g['D_left'], [720, 1280])
    right = camera_utils.FisheyeCamera(g['K_right'], g['D_right'], [720, 1280])
    assert np.abs(c_oracle.project(g['X'], g['T_CW'], left) - g['project_left']).max() < 1e-9
    assert np.abs(c_oracle.project(g['X'], g['T_RL'] @ g['T_CW'], right) - g['project_right']).max() < 1e-9
    assert np.abs(left.project(g['X'], g['T_CW']) - g['project_left']).max() < 1e-9
    for tag in ['full', 'small', 'net']:
        cam = camera_utils.FisheyeCamera(g[f'K_{tag}'], g[f'D_{tag}'], g[f'image_size_{tag}'])
        assert np.abs(c_oracle.undistort(g[f'undistort_in_{tag}'], cam) - g[f'undistort_out_{tag}']).max() < 1e-9
        assert np.abs(cam.undistort(g[f'undistort_in_{tag}']) - g[f'undistort_out_{tag}']).max() < 1e-9
        out32 = c_oracle.undistort(g[f'undistort_in_{tag}'].astype(np.float32).astype(np.float64), cam, round_to_f32=True)
        assert np.array_equal(out32.astype(np.float32), g[f'undistort_out32_{tag}'])


def test_dlt_against_opencv_goldens():
    g = load_golden('geometry.npz')
    pts = np.stack([g['pairs_undistorted_left'], g['pairs_undistorted_right']], axis=1)
    proj = np.stack([g['P1'], g['P2']])
    want = g['pairs_plain_dlt']
    for got in (np_oracle.triangulate_dlt(pts, np.ones(pts.shape[:2], bool), proj),
                c_oracle.triangulate(pts, None, proj)):
        rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
        assert rel.max() < 1e-9


def test_golden_pixel_vectors_of_reference_test_suite():
    """test/test_pipeline.py:26-33,171-177: triangulating the printed pixel vectors gives the
    3D points back within 1e-3 m."""
    from object_keypoints_b200 import camera_utils
    g = load_golden('geometry.npz')
    left = camera_utils.FisheyeCamera(g['K_left'], g['D_left'], [720, 1280])
    right = camera_utils.FisheyeCamera(g['K_right'], g['D_right'], [720, 1280])
    uL = c_oracle.undistort(g['golden_left'], left)
    uR = c_oracle.undistort(g['golden_right'], right)
    X = c_oracle.triangulate(np.stack([uL, uR], axis=1), None, np.stack([g['P1'], g['P2']]))
    assert np.linalg.norm(X - g['golden_keypoints'], axis=1).max() < 1e-3
    # and the pixels themselves are the projections of those points (re-derived to all digits)
    assert np.abs(c_oracle.project(g['golden_keypoints'], np.eye(4), left) - g['golden_left']).max() < 1e-7
    assert np.abs(c_oracle.project(g['golden_keypoints'], g['T_RL'], right) - g['golden_right']).max() < 1e-7


def test_kmeans_stand_in_merges_double_detection():
    pts = np.array([[10.0, 10.0], [10.6, 10.2], [20.0, 10.0], [10.0, 22.0]], np.float32)
    cen = np_oracle.cluster_detections(pts, 3)
    want = np.array([[10.3, 10.1], [20.0, 10.0], [10.0, 22.0]])
    assert np.abs(np.sort(cen, axis=0) - np.sort(want, axis=0)).max() < 1e-5


def _multiview_scene(seed=1003, P=40, V=16):
    """BASELINE config 3 shape: P points seen from V poses, 0.3 px noise, 5 % gross outliers."""
    from object_keypoints_b200 import synthetic
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    camera = synthetic.default_camera((180, 320)).scale(4.0)
    X = np.stack([rng.uniform(-0.3, 0.3, P), rng.uniform(-0.2, 0.2, P), rng.uniform(-0.05, 0.05, P)], axis=1)
    poses = np.zeros((V, 4, 4))
    for v in range(V):
        R = Rotation.from_rotvec(rng.normal(0, 0.25, 3)).as_matrix()
        poses[v] = np.eye(4)
        poses[v][:3, :3] = R
        poses[v][:3, 3] = -R @ np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), -rng.uniform(0.6, 1.0)])
    clean = np.stack([camera.project(X, poses[v]) for v in range(V)], axis=1)
    obs = clean + rng.normal(0, 0.3, (P, V, 2))
    outliers = rng.uniform(size=(P, V)) < 0.05
    obs[outliers] += rng.normal(0, 20.0, (int(outliers.sum()), 2))
    return camera, X, poses, obs, clean, outliers


def test_robust_multiview_oracles_agree_and_recover_points():
    camera, X, poses, obs, clean, outliers = _multiview_scene()
    P, V = obs.shape[:2]
    Xc, valid_c, err_c, dropped_c = c_oracle.triangulate_robust(obs, None, poses, camera, 2.0, V)
    Xn, valid_n, err_n, dropped_n = np_oracle.triangulate_robust(obs, None, poses, np_oracle.camera_dict(camera),
                                                                 camera.K, 2.0)
    np.testing.assert_array_equal(valid_c, valid_n)
    np.testing.assert_array_equal(dropped_c, dropped_n)
    assert (np.linalg.norm(Xc - Xn, axis=1) <= 1e-7 * np.linalg.norm(Xn, axis=1) + 1e-10).all()
    np.testing.assert_allclose(err_c, err_n, rtol=0, atol=1e-6)
    assert np.isfinite(Xc).all() and np.linalg.norm(Xc - X, axis=1).max() < 5e-3
    gross = outliers & (np.linalg.norm(obs - clean, axis=2) > 6)
    assert gross.sum() > 0 and valid_c[gross].sum() == 0
    # nothing above the gate -> plain DLT (views outside the image are masked: their undistortion may not converge)
    quiet = clean + np.random.default_rng(1).normal(0, 0.02, clean.shape)
    inside = ((clean[..., 0] > 0) & (clean[..., 0] < 1280) & (clean[..., 1] > 0) & (clean[..., 1] < 720)).astype(np.uint8)
    assert (inside.sum(axis=1) >= 2).all()
    Xq, valid_q, _, dropped_q = c_oracle.triangulate_robust(quiet, inside, poses, camera, 2.0, V)
    assert dropped_q.sum() == 0 and np.array_equal(valid_q, inside)
    und = np.stack([camera.undistort(quiet[:, v]) for v in range(V)], axis=1)
    proj = np.stack([camera.K @ poses[v][:3] for v in range(V)])
    plain = c_oracle.triangulate(und, inside, proj)
    assert (np.linalg.norm(Xq - plain, axis=1) <= 1e-9 * np.linalg.norm(plain, axis=1) + 1e-12).all()


def _stereo_from_golden(g, scale=None):
    from object_keypoints_b200 import camera_utils
    left = camera_utils.FisheyeCamera(g['K_left'], g['D_left'], [720, 1280])
    right = camera_utils.FisheyeCamera(g['K_right'], g['D_right'], [720, 1280])
    if scale is not None:
        left, right = left.scale(scale), right.scale(scale)
    return camera_utils.StereoCamera(left, right, g['T_RL'])


def test_hartley_sturm_restatement_is_cv2_correct_matches():
    """A13 pinned: np_oracle.correct_matches against cv2.correctMatches outputs (rig F, forward
    motion, general pose) stored by oracle/make_goldens.py."""
    g = load_golden('geometry.npz')
    l, r = np_oracle.correct_matches(g['F'], g['pairs_undistorted_left'], g['pairs_undistorted_right'])
    assert np.abs(l - g['pairs_corrected_left']).max() < 1e-9
    assert np.abs(r - g['pairs_corrected_right']).max() < 1e-9
    for i in range(g['hs_F'].shape[0]):
        l, r = np_oracle.correct_matches(g['hs_F'][i], g['hs_left'][i], g['hs_right'][i])
        assert np.abs(l - g['hs_corrected_left'][i]).max() < 1e-9
        assert np.abs(r - g['hs_corrected_right'][i]).max() < 1e-9
        # the corrected pairs satisfy the epipolar constraint
        h = lambda x: np.concatenate([x, np.ones((len(x), 1))], axis=1)
        resid = np.einsum('ni,ij,nj->n', h(r), g['hs_F'][i], h(l)) / np.abs(g['hs_F'][i]).max()
        assert np.abs(resid).max() < 1e-6


def test_stereo_triangulate_restatement_is_the_reference():
    """StereoCamera.triangulate of the unmodified reference (float32 casts, correctMatches, DLT) on
    noisy pairs at full and at test scale; the plain DLT is measurably different."""
    g = load_golden('geometry.npz')
    for scale, left_px, right_px, want in ((None, g['pairs_left'], g['pairs_right'], g['pairs_stereo_triangulate']),
                                           (180 / 720, g['small_pairs_left'], g['small_pairs_right'],
                                            g['small_pairs_stereo_triangulate'])):
        stereo = _stereo_from_golden(g, scale)
        args = (left_px, right_px, np_oracle.camera_dict(stereo.left_camera), np_oracle.camera_dict(stereo.right_camera),
                stereo.left_camera.K, stereo.right_camera.K, stereo.T_RL, stereo.F)
        got = np_oracle.triangulate_stereo(*args)
        rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
        assert np.median(rel) < 1e-6 and rel.max() < 1e-4       # float32 roundings of undistort/correct outputs
        plain = np_oracle.triangulate_stereo(*args, optimal_correction=False)
        rel_plain = np.linalg.norm(plain - want, axis=1) / np.linalg.norm(want, axis=1)
        assert rel_plain.max() > 10 * rel.max()


def test_association_oracle_meets_reference_test_expectations():
    """test/test_pipeline.py:208-261 restated with consistent camera scales."""
    g = load_golden('geometry.npz')
    stereo = _stereo_from_golden(g, 0.25)
    keypoints_X = np.array([[0.0, 0.0, 1.0], [0.0, 0.25, 1.0], [0.0, -0.25, 1.0]])
    pl = stereo.left_camera.project(keypoints_X, np.eye(4))
    pr = stereo.right_camera.project(keypoints_X, stereo.T_RL)
    uL, uR = stereo.left_camera.undistort(pl), stereo.right_camera.undistort(pr)
    rng = np.random.default_rng(3)
    for _ in range(5):                                    # test_association_simple
        perm = rng.permutation(3)
        match, _ = np_oracle.associate(stereo.F, uL, uR[perm])
        assert (match != -1).all()
        np.testing.assert_array_equal(uR, uR[perm][match])
    # test_association_two_same: left 0 has no partner, the other two share one image column
    points_left = np.array([[160.251929, 92.04110211], [160.251929, 135.25386897], [160.251929, 48.82833525]])
    points_right = np.array([[149.9327, 139.14128], [149.93279695, 133.14128143], [149.88808034, 47.08818382]])
    match, _ = np_oracle.associate(stereo.F, stereo.left_camera.undistort(points_left),
                                   stereo.right_camera.undistort(points_right))
    np.testing.assert_array_equal(match, [-1, 1, 2])
