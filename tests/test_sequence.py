"""BASELINE config 3: a sequence filmed by a moving camera -> per-pair epipolar association -> tracks -> robust V-view
triangulation (object_keypoints_b200/sequence.py). The reference stops at two views (scripts/label.py:285-305), so the
chain's statement of record is oracle/sequence_oracle.py (parity unpinned for V > 2; its pieces are pinned where a
reference exists). CPU: host logic + the oracle chain against ground truth. GPU: the CUDA chain against the oracle chain
on the same heatmaps -- matches, masks and drop counts bit-equal, points <= 1e-4 relative."""
import numpy as np
import pytest

from helpers import TOL_METRES_REL


def _small_sequence(n_frames):
    from object_keypoints_b200 import sequence, synthetic
    camera = synthetic.default_camera((180, 320))
    seq = sequence.synthetic_sequence(n_frames)
    keypoints, depths = sequence.project_sequence(seq, camera)
    return camera, seq, keypoints, depths


def test_view_schedule_and_pair_geometry():
    from object_keypoints_b200 import sequence
    from oracle import sequence_oracle
    frames = sequence.view_schedule(900, 16)
    assert frames.shape == (56, 16) and frames[0, 0] == 0 and frames[0, 1] == 56 and frames[55, 15] == 55 + 15 * 56
    assert len(np.unique(frames)) == frames.size
    camera, seq, keypoints, depths = _small_sequence(24)
    assert keypoints.shape == (24, 8, 5, 2) and (depths > 0.3).all()
    inside = (keypoints[..., 0] > 4) & (keypoints[..., 0] < 316) & (keypoints[..., 1] > 4) & (keypoints[..., 1] < 176)
    assert inside.mean() > 0.95
    # every object's first point is the mean of its keypoints (video.py:128)
    np.testing.assert_allclose(seq['scene'][:, 0], seq['scene'][:, 1:].mean(axis=1), atol=1e-12)
    schedule = sequence.view_schedule(24, 4)
    F = sequence.pair_fundamentals(camera, seq['T_CW'], schedule)
    K = np.asarray(camera.K)
    for a, v in ((0, 1), (3, 2), (5, 3)):
        T_RL = seq['T_CW'][schedule[a, v]] @ sequence_oracle.inv_transform(seq['T_CW'][schedule[a, 0]])
        np.testing.assert_allclose(F[a, v - 1], sequence_oracle.fundamental_matrix(T_RL, K, K), rtol=1e-12, atol=1e-18)
        # the epipolar constraint holds for the (undistorted = pinhole) projections of the scene points
        X = seq['scene'].reshape(-1, 3)
        def pinhole(T):
            Xc = X @ T[:3, :3].T + T[:3, 3]
            return (Xc / Xc[:, 2:3]) @ K.T
        left, right = pinhole(seq['T_CW'][schedule[a, 0]]), pinhole(seq['T_CW'][schedule[a, v]])
        residual = np.einsum('ni,ij,nj->n', right, F[a, v - 1], left)
        scale = np.linalg.norm(F[a, v - 1]) * np.linalg.norm(left, axis=1) * np.linalg.norm(right, axis=1)
        assert np.abs(residual / scale).max() < 1e-12


def test_oracle_chain_recovers_the_scene():
    from oracle import c_oracle, np_oracle, sequence_oracle
    camera, seq, keypoints, depths = _small_sequence(32)
    N = 32
    heat = np.zeros((N, 3, 180, 320), np.float32)
    depth = np.zeros_like(heat)
    centers = np.zeros((N, 2, 2, 180, 320), np.float32)
    for n in range(N):
        heat[n], centers[n], depth[n] = np_oracle.rasterise_targets(keypoints[n], depths[n], [1, 3], (180, 320))
    tables = c_oracle.decode(heat, depth, centers, [1, 3], camera)
    assert (tables['n_objects'] == 8).all()
    out = sequence_oracle.sequence_tracks(tables, seq['T_CW'], camera, views=4)
    ok = ~np.isnan(out['points'][..., 0])
    assert ok.sum() == 8 * 40                                     # 8 anchors x 40 keypoints
    truth = seq['scene'].reshape(-1, 3)
    error = np.linalg.norm(out['points'][ok][:, None] - truth[None], axis=2).min(axis=1)
    assert np.median(error) < 2e-3 and (error < 1e-2).mean() > 0.9
    assert (out['valid'].sum(axis=-1)[ok] >= 2).all()


@pytest.mark.gpu
def test_cuda_chain_matches_the_oracle_chain():
    import torch
    from object_keypoints_b200 import KeypointDecoder, sequence, targets
    from oracle import c_oracle, sequence_oracle
    camera, seq, keypoints, depths = _small_sequence(96)
    heat, depth, centers = targets.rasterise_targets(keypoints, depths, [1, 3], (180, 320))
    decoder = KeypointDecoder([1, 3], (180, 320), camera=camera)
    tables = decoder.decode_batch(heat, depth, centers)
    want_tables = c_oracle.decode(heat.cpu().numpy(), depth.cpu().numpy(), centers.cpu().numpy(), [1, 3], camera)
    got_tables = tables.numpy()
    np.testing.assert_array_equal(got_tables['peak_count'], want_tables['peak_count'])
    np.testing.assert_array_equal(got_tables['peak_xy'].view(np.uint32), want_tables['peak_xy'].view(np.uint32))
    for views, noise in ((6, 0.0), (8, 0.0)):
        chain = sequence.SequenceTriangulator(camera, views=views)
        got = {k: v.cpu().numpy() for k, v in chain(tables, chain.prepare(seq['T_CW'])).items()}
        want = sequence_oracle.sequence_tracks(want_tables, seq['T_CW'], camera, views=views)
        np.testing.assert_array_equal(got['match'], want['match'])
        np.testing.assert_array_equal(got['observed'], want['observed'])
        np.testing.assert_array_equal(got['observations'], want['observations'])
        np.testing.assert_array_equal(got['valid'], want['valid'])          # bit-equal masks after the reprojection filter
        np.testing.assert_array_equal(got['dropped'], want['dropped'])
        ok = ~np.isnan(want['points'][..., 0])
        assert (np.isnan(got['points'][..., 0]) == ~ok).all() and ok.sum() == (96 // views) * 40
        scale = np.linalg.norm(want['points'][ok], axis=-1)
        assert (np.linalg.norm(got['points'][ok] - want['points'][ok], axis=-1) <= TOL_METRES_REL * scale).all()
        used = got['valid'].astype(bool)
        assert np.abs(got['error'][used] - want['error'][used]).max() < 1e-6
        truth = seq['scene'].reshape(-1, 3)
        error = np.linalg.norm(got['points'][ok][:, None] - truth[None], axis=2).min(axis=1)
        assert np.median(error) < 2e-3
        assert (got['dropped'][ok] > 0).any() or views == 6            # wrong epipolar matches do occur and are filtered
