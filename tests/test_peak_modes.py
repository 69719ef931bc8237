"""K1 parameter modes beyond the reference pipeline's own (BASELINE.json: "3x3 max-pool NMS peak
extraction with per-type top-k and thresholding"): nms_size 3 / 5, NMS on the box sum or on the map itself, per-type
top-k. Pinned to CornerNet's unmodified _nms + _topk (perception/corner_net_lite/core/models/py_utils/utils.py:14-38,
fixture tests/golden/cornernet_topk.npz); the CUDA path is compared with the NumPy oracle for every combination."""
import numpy as np
import pytest

from helpers import load_golden


def test_oracle_top_k_is_cornernet_nms_plus_topk():
    from oracle import np_oracle
    g = load_golden('cornernet_topk.npz')
    k = int(g['k'])
    t = np_oracle.extract_peak_tables(g['heat'], threshold=float(g['threshold']), nms_size=3, use_box_sum=False, top_k=k,
                                      max_peaks=256)
    np.testing.assert_array_equal(t['peak_count'][:, 0], k)
    np.testing.assert_array_equal(t['peak_score'][:, 0, :k], g['ref_scores'])
    np.testing.assert_array_equal(t['peak_yx'][:, 0, :k, 0], g['ref_ys'])
    np.testing.assert_array_equal(t['peak_yx'][:, 0, :k, 1], g['ref_xs'])
    # and _nms itself: the suppressed map is the map where it equals its 3x3 maximum
    for m in range(g['heat'].shape[0]):
        yx, score = np_oracle.find_peaks(g['heat'][m, 0], threshold=0.0, nms_size=3, use_box_sum=False)
        keep = np.zeros(score.shape, bool)
        keep[yx[:, 0], yx[:, 1]] = True
        np.testing.assert_array_equal(np.where(keep, score, 0.0), g['ref_suppressed'][m, 0])


@pytest.mark.gpu
@pytest.mark.parametrize('nms_size,box_sum,top_k', [(3, False, 20), (3, False, 0), (5, False, 7), (3, True, 0), (3, True, 5),
                                                    (5, True, 3)])
def test_cuda_peak_modes_match_the_oracle(nms_size, box_sum, top_k):
    from oracle import np_oracle
    from object_keypoints_b200 import KeypointDecoder, synthetic
    g = load_golden('cornernet_topk.npz')
    threshold = float(g['threshold'])
    K = 256
    heat = np.concatenate([g['heat'][:3].reshape(1, 3, *g['heat'].shape[2:]), g['heat'][3:].reshape(1, 3, *g['heat'].shape[2:])])
    if box_sum:
        threshold, heat = 0.5, heat * np.float32(0.2)              # box sums of 25 pixels
    decoder = KeypointDecoder([1, 1], heat.shape[2:], max_peaks=K, threshold=threshold, nms_size=nms_size, box_sum=box_sum,
                              top_k=top_k)
    got = decoder.extract_peaks(heat).numpy()
    want = np_oracle.extract_peak_tables(heat, threshold=threshold, nms_size=nms_size, use_box_sum=box_sum, top_k=top_k, max_peaks=K)
    np.testing.assert_array_equal(got['peak_count'], want['peak_count'])
    np.testing.assert_array_equal(got['peak_yx'], want['peak_yx'])
    for key in ('peak_score', 'peak_xy', 'peak_conf'):
        np.testing.assert_array_equal(got[key].view(np.uint32), want[key].view(np.uint32), err_msg=key)
    assert want['peak_count'].min() > 0
    if top_k:
        assert (want['peak_count'] <= top_k).all() and (np.diff(want['peak_score'][0, 0, :top_k]) <= 0).all()
    # a full decode in top-k mode: the grouping consumes the score-ordered tables (config 2 style frames)
    if (nms_size, box_sum, top_k) == (5, True, 3):
        batch = synthetic.make_batch(8, [1, 1, 1], (64, 64), seed=9, objects=(1, 3))
        camera = synthetic.default_camera((64, 64))
        reference = KeypointDecoder([1, 1, 1], (64, 64), camera=camera).decode_batch(batch.heat, batch.depth, batch.centers).numpy()
        ranked = KeypointDecoder([1, 1, 1], (64, 64), camera=camera, top_k=3).decode_batch(batch.heat, batch.depth, batch.centers).numpy()
        np.testing.assert_array_equal(ranked['n_objects'], np.minimum(reference['n_objects'], 3))
        for n in range(8):                                       # same peaks when nothing is cut, ordered by score
            for c in range(4):
                k = int(reference['peak_count'][n, c])
                if k <= 3:
                    order = np.lexsort((np.arange(k), -reference['peak_score'][n, c, :k]))
                    np.testing.assert_array_equal(ranked['peak_yx'][n, c, :k], reference['peak_yx'][n, c, :k][order])
