"""Camera geometry and triangulation kernels against OpenCV goldens and the C oracle."""
import numpy as np
import pytest

from helpers import load_golden, TOL_METRES_REL

pytestmark = pytest.mark.gpu


def cameras(g):
    from object_keypoints_b200 import camera_utils
    left = camera_utils.FisheyeCamera(g['K_left'], g['D_left'], [720, 1280])
    right = camera_utils.FisheyeCamera(g['K_right'], g['D_right'], [720, 1280])
    return left, right


def test_project_and_undistort_match_opencv():
    from object_keypoints_b200 import project_points, undistort_points, camera_utils
    g = load_golden('geometry.npz')
    left, right = cameras(g)
    assert np.abs(project_points(g['X'], g['T_CW'], left).cpu().numpy() - g['project_left']).max() < 1e-8
    assert np.abs(project_points(g['X'], g['T_RL'] @ g['T_CW'], right).cpu().numpy() - g['project_right']).max() < 1e-8
    for tag in ['full', 'small', 'net']:
        cam = camera_utils.FisheyeCamera(g[f'K_{tag}'], g[f'D_{tag}'], g[f'image_size_{tag}'])
        out = undistort_points(g[f'undistort_in_{tag}'], cam).cpu().numpy()
        assert np.abs(out - g[f'undistort_out_{tag}']).max() < 1e-8
        out32 = undistort_points(g[f'undistort_in_{tag}'].astype(np.float32).astype(np.float64), cam, round_to_f32=True)
        diff = np.abs(out32.cpu().numpy().astype(np.float32) - g[f'undistort_out32_{tag}'])
        assert diff.max() <= 1.5e-4 and (diff > 0).mean() < 0.01        # float32 ulp at ~1e3 px, rare


def test_two_view_dlt_matches_cv2_triangulate_points():
    from object_keypoints_b200 import triangulate
    g = load_golden('geometry.npz')
    pts = np.stack([g['pairs_undistorted_left'], g['pairs_undistorted_right']], axis=1)
    X = triangulate(pts, np.stack([g['P1'], g['P2']])).cpu().numpy()
    rel = np.linalg.norm(X - g['pairs_plain_dlt'], axis=1) / np.linalg.norm(g['pairs_plain_dlt'], axis=1)
    assert rel.max() < 1e-8


def test_golden_pixel_vectors_of_reference_tests():
    """test/test_pipeline.py:171-177 through the TriangulationComponent API."""
    from object_keypoints_b200 import TriangulationComponent, camera_utils
    g = load_golden('geometry.npz')
    left, right = cameras(g)
    stereo = camera_utils.StereoCamera(left, right, g['T_RL'])
    triangulation = TriangulationComponent()
    triangulation.reset(stereo)
    p_W = triangulation(g['golden_left'], g['golden_right'])
    np.testing.assert_array_less(np.linalg.norm(p_W - g['golden_keypoints'], axis=1), 1e-3)


def test_multiview_dlt_matches_oracle_and_recovers_points():
    """Config 3 shape: 40 points x 16 views, 0.3 px noise, 5% gross outliers, 2 px gate."""
    from object_keypoints_b200 import triangulate, triangulate_multiview, synthetic
    from oracle import c_oracle, np_oracle
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(1003)
    camera = synthetic.default_camera((180, 320)).scale(4.0)      # 720x1280 camera
    P, V = 40, 16
    X = np.stack([rng.uniform(-0.3, 0.3, P), rng.uniform(-0.2, 0.2, P), rng.uniform(-0.05, 0.05, P)], axis=1)
    poses = np.zeros((V, 4, 4))
    for v in range(V):
        R = Rotation.from_rotvec(rng.normal(0, 0.25, 3)).as_matrix()
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = -R @ np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), -rng.uniform(0.6, 1.0)])
        poses[v] = T
    obs = np.stack([camera.project(X, poses[v]) for v in range(V)], axis=1) + rng.normal(0, 0.3, (P, V, 2))
    outliers = rng.uniform(size=(P, V)) < 0.05
    obs[outliers] += rng.normal(0, 20.0, (int(outliers.sum()), 2))
    # plain V-view DLT: CUDA vs both oracles
    undist = np.stack([camera.undistort(obs[:, v]) for v in range(V)], axis=1)
    proj = np.stack([camera.K @ poses[v][:3] for v in range(V)])
    valid = np.ones((P, V), np.uint8)
    got = triangulate(undist, proj, valid).cpu().numpy()
    for want in (c_oracle.triangulate(undist, valid, proj), np_oracle.triangulate_dlt(undist, valid.astype(bool), proj)):
        assert (np.linalg.norm(got - want, axis=1) <= 1e-7 * np.linalg.norm(want, axis=1) + 1e-10).all()
    # robust form (one kernel): drop the worst view above the 2 px gate, solve again -- vs both oracles
    Xf, valid_f, err, dropped = triangulate_multiview(obs, None, poses, camera, max_error_px=2.0, return_dropped=True)
    Xf, valid_f, err, dropped = Xf.cpu().numpy(), valid_f.cpu().numpy(), err.cpu().numpy(), dropped.cpu().numpy()
    Xc, valid_c, err_c, dropped_c = c_oracle.triangulate_robust(obs, None, poses, camera, 2.0, V)
    np.testing.assert_array_equal(valid_f, valid_c)
    np.testing.assert_array_equal(dropped, dropped_c)
    assert (np.linalg.norm(Xf - Xc, axis=1) <= 1e-7 * np.linalg.norm(Xc, axis=1) + 1e-10).all()
    np.testing.assert_allclose(err, err_c, rtol=0, atol=1e-6)
    Xn, valid_n, _, _ = np_oracle.triangulate_robust(obs, None, poses, np_oracle.camera_dict(camera), camera.K, 2.0)
    np.testing.assert_array_equal(valid_f, valid_n)
    assert (np.linalg.norm(Xf - Xn, axis=1) <= 1e-7 * np.linalg.norm(Xn, axis=1) + 1e-10).all()
    assert np.isfinite(Xf).all()
    assert np.linalg.norm(Xf - X, axis=1).max() < 5e-3
    assert np.linalg.norm(Xf - X, axis=1).mean() < np.linalg.norm(got - X, axis=1).mean()
    true_err = np.linalg.norm(obs - np.stack([camera.project(X, poses[v]) for v in range(V)], axis=1), axis=2)
    assert valid_f[outliers & (true_err > 6)].sum() == 0
    assert (err[valid_f.astype(bool)] <= 2.0).all()
    # the plain gate (every view above the threshold at once) against the C oracle
    from object_keypoints_b200 import reprojection_filter
    v1, e1 = reprojection_filter(got, obs, valid, poses, camera, 2.0)
    v0, e0 = c_oracle.reprojection_filter(got, obs, valid, poses, camera, 2.0)
    np.testing.assert_array_equal(v1.cpu().numpy(), v0)
    np.testing.assert_allclose(e1.cpu().numpy(), e0, rtol=0, atol=1e-8)
    # per-point projections and masks
    per_point = np.broadcast_to(proj[None], (P, V, 3, 4)).copy()
    got2 = triangulate(undist, per_point, valid).cpu().numpy()
    np.testing.assert_allclose(got2, got, rtol=1e-12, atol=1e-14)
    few = valid.copy()
    few[0, 1:] = 0
    assert np.isnan(triangulate(undist, proj, few).cpu().numpy()[0]).all()


def test_stereo_pipeline_end_to_end_like_reference_test():
    """test/test_pipeline.py:179-206: extraction on both views + triangulation within 5e-2 m."""
    from object_keypoints_b200 import KeypointExtractionComponent, TriangulationComponent, camera_utils
    g = load_golden('test_pipeline_180x320.npz')
    geo = load_golden('geometry.npz')
    left, right = cameras(geo)
    stereo_small = camera_utils.StereoCamera(left.scale(180 / 720), right.scale(180 / 720), g['T_RL'])
    extraction = KeypointExtractionComponent({'keypoint_config': [1, 3]}, [180, 320])
    keypoints, _ = extraction(g['heat'])
    triangulation = TriangulationComponent()
    triangulation.reset(stereo_small)
    for c, count in enumerate([1, 1, 3]):
        L, R = np.stack(keypoints[0][c]), np.stack(keypoints[1][c])
        assert L.shape[0] == R.shape[0] == count
        X = triangulation(L, R)
        assert X.shape == (count, 3)
    X0 = triangulation(np.stack(keypoints[0][0]), np.stack(keypoints[1][0]))
    assert np.linalg.norm(X0[0] - g['keypoints_3d'][0]) < 5e-2


def test_correct_matches_is_cv2_and_the_oracle():
    """A13: okp_correct_matches_f64 against cv2.correctMatches goldens (rig, forward motion, general
    pose) and the NumPy restatement."""
    from object_keypoints_b200 import correct_matches
    from oracle import np_oracle
    g = load_golden('geometry.npz')
    cases = [(g['F'], g['pairs_undistorted_left'], g['pairs_undistorted_right'],
              g['pairs_corrected_left'], g['pairs_corrected_right'])]
    cases += [(g['hs_F'][i], g['hs_left'][i], g['hs_right'][i], g['hs_corrected_left'][i], g['hs_corrected_right'][i])
              for i in range(g['hs_F'].shape[0])]
    for F, l, r, want_l, want_r in cases:
        got_l, got_r = correct_matches(F, l, r)
        got_l, got_r = got_l.cpu().numpy(), got_r.cpu().numpy()
        assert np.abs(got_l - want_l).max() < 1e-7 and np.abs(got_r - want_r).max() < 1e-7
        ol, orr = np_oracle.correct_matches(F, l, r)
        assert np.abs(got_l - ol).max() < 1e-7 and np.abs(got_r - orr).max() < 1e-7
    # degenerate inputs: pairs already on their epipolar lines stay where they are
    got_l, got_r = correct_matches(g['F'], g['pairs_corrected_left'], g['pairs_corrected_right'])
    assert np.abs(got_l.cpu().numpy() - g['pairs_corrected_left']).max() < 1e-6
    empty_l, empty_r = correct_matches(g['F'], np.zeros((0, 2)), np.zeros((0, 2)))
    assert empty_l.shape == (0, 2) and empty_r.shape == (0, 2)


def test_stereo_triangulate_with_correction_matches_the_reference():
    """StereoCamera.triangulate (camera_utils.py:92-110) of the unmodified reference on noisy pairs,
    <= 1e-4 relative (north_star), at full and at test scale."""
    from object_keypoints_b200 import camera_utils
    g = load_golden('geometry.npz')
    left, right = cameras(g)
    for scale, lp, rp, want in ((1.0, g['pairs_left'], g['pairs_right'], g['pairs_stereo_triangulate']),
                                (180 / 720, g['small_pairs_left'], g['small_pairs_right'],
                                 g['small_pairs_stereo_triangulate'])):
        stereo = camera_utils.StereoCamera(left.scale(scale), right.scale(scale), g['T_RL']) if scale != 1.0 else \
            camera_utils.StereoCamera(left, right, g['T_RL'])
        got = stereo.triangulate(lp, rp)
        rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
        assert rel.max() <= TOL_METRES_REL
        plain = stereo.triangulate(lp, rp, optimal_correction=False)
        assert (np.linalg.norm(plain - want, axis=1) / np.linalg.norm(want, axis=1)).max() > rel.max()


def test_association_component_meets_reference_test_expectations():
    """test/test_pipeline.py:208-261 (AssociationComponent) and the oracle, batched and single."""
    from object_keypoints_b200 import AssociationComponent, associate, camera_utils, undistort_points
    from oracle import np_oracle
    g = load_golden('geometry.npz')
    left, right = cameras(g)
    stereo = camera_utils.StereoCamera(left.scale(0.25), right.scale(0.25), g['T_RL'])
    keypoints_X = np.array([[0.0, 0.0, 1.0], [0.0, 0.25, 1.0], [0.0, -0.25, 1.0]])
    points_left = stereo.left_camera.project(keypoints_X, np.eye(4))
    points_right = stereo.right_camera.project(keypoints_X, stereo.T_RL)
    association = AssociationComponent()
    association.reset(stereo)
    rng = np.random.default_rng(3)
    for _ in range(5):
        shuffled = points_right[rng.permutation(3)]
        associations = association(points_left, shuffled)
        assert (associations != -1).all()
        np.testing.assert_equal(points_right, shuffled[associations])
    two_left = np.array([[160.251929, 92.04110211], [160.251929, 135.25386897], [160.251929, 48.82833525]])
    two_right = np.array([[149.9327, 139.14128], [149.93279695, 133.14128143], [149.88808034, 47.08818382]])
    np.testing.assert_array_equal(association(two_left, two_right), [-1, 1, 2])
    # the 64x64 cameras of test_association_tricky: every left point gets its own right point
    K = np.array([[62.31692844, 0., 31.92640056], [0., 62.38274914, 32.92623658], [0., 0., 1.]])
    Kp = np.array([[62.07155716, 0., 31.79527486], [0., 62.14031698, 32.54056898], [0., 0., 1.]])
    D = np.array([-1.73678913e-01, 2.69084607e-02, -2.66312740e-04, -1.11094300e-04])
    Dp = np.array([-0.17596905, 0.02856535, -0.00036341, -0.00021308])
    tiny = camera_utils.StereoCamera(camera_utils.FisheyeCamera(K, D, [64, 64]), camera_utils.FisheyeCamera(Kp, Dp, [64, 64]),
                                     g['T_RL'])
    association.reset(tiny)
    tricky = association(np.array([[35.5, 25.5], [26.5, 39.5], [38.5, 39.5]]),
                         np.array([[29.5, 25.5], [20.5, 38.5], [33.5, 39.5]]))
    assert tricky.shape[0] == 3 and np.unique(tricky).size == 3 and (tricky >= 0).all()
    assert association(np.zeros((0, 2)), two_right).shape == (0,)
    np.testing.assert_array_equal(association(two_left, np.zeros((0, 2))), [-1, -1, -1])
    # batched, ragged, random: bit-equal to the oracle
    B, ML, MR = 37, 9, 7
    L = np.zeros((B, ML, 2))
    R = np.zeros((B, MR, 2))
    nl = rng.integers(0, ML + 1, B).astype(np.int32)
    nr = rng.integers(0, MR + 1, B).astype(np.int32)
    for b in range(B):
        X = np.stack([rng.uniform(-0.3, 0.3, ML), rng.uniform(-0.2, 0.2, ML), rng.uniform(0.5, 1.5, ML)], axis=1)
        L[b] = stereo.left_camera.undistort(stereo.left_camera.project(X, np.eye(4)) + rng.normal(0, 0.3, (ML, 2)))
        Rp = stereo.right_camera.undistort(stereo.right_camera.project(X, stereo.T_RL) + rng.normal(0, 0.3, (ML, 2)))
        R[b] = Rp[rng.permutation(ML)[:MR]]
    match, cost = associate(stereo.F, L, R, nl, nr, max_distance_px=1.0)
    match, cost = match.cpu().numpy(), cost.cpu().numpy()
    for b in range(B):
        want_m, want_c = np_oracle.associate(stereo.F, L[b, :nl[b]], R[b, :nr[b]], max_distance_px=1.0)
        np.testing.assert_array_equal(match[b, :nl[b]], want_m)
        assert (match[b, nl[b]:] == -1).all()
        np.testing.assert_allclose(cost[b, :nl[b]], want_c, rtol=0, atol=1e-9)
