"""Evaluation bookkeeping: the matching + error statistics of the reference's
scripts/eval_model.py Results.add / print_results (:137-232). tests/golden/evaluation.npz holds what the
UNMODIFIED reference class returned; the NumPy oracle is pinned to it here, the CUDA path (through the C
ABI) is compared with both in the gpu test."""
import numpy as np
import pytest

from helpers import load_golden, golden_camera


def sequences_from_status(out, kp_point, n_objects):
    """Status tables -> the reference's nested lists, as padded arrays: kept objects in order, per object the
    points in (type, slot) order without the ones dropped as 'not in view' (eval_model.py:159-187)."""
    N, O, C, S = out['status'].shape
    L = C * S
    kind = np.full((N, O, L), -1, np.int32)
    pred = np.zeros((N, O, L, 3))
    gt = np.zeros((N, O, L, 3))
    kept = np.zeros(N, np.int32)
    for n in range(N):
        k = 0
        for o in range(int(n_objects[n])):
            st = out['status'][n, o].reshape(-1)
            if (st == 3).any():
                assert ((st == 3) | (st == -1)).all()
                continue
            i = 0
            for j, code in enumerate(st):
                if code == 0:
                    kind[n, k, i] = 0
                    pred[n, k, i] = kp_point[n, o].reshape(-1, 3)[j]
                    gt[n, k, i] = out['gt_point'][n, o].reshape(-1, 3)[j]
                    i += 1
                elif code == 1:
                    kind[n, k, i] = 1
                    i += 1
            k += 1
        kept[n] = k
    return kind, pred, gt, kept


def check_against_reference(out, g, summary):
    kind, pred, gt, kept = sequences_from_status(out, g['kp_point'], g['n_objects'])
    L = g['ref_seq_kind'].shape[2]
    np.testing.assert_array_equal(kept, g['ref_kept_objects'])
    np.testing.assert_array_equal(kind[:, :, :L], g['ref_seq_kind'])
    assert (kind[:, :, L:] == -1).all()
    np.testing.assert_allclose(pred[:, :, :L], g['ref_seq_pred'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(gt[:, :, :L], g['ref_seq_gt'], rtol=0, atol=1e-9)
    for name, want in zip(g['ref_summary_columns'], g['ref_summary']):
        # 'missing' is printed with two decimals (eval_model.py:230)
        tol = 5e-3 if str(name) == 'missing_percentage' else 1e-9 * max(abs(want), 1.0)
        assert abs(summary[str(name)] - want) <= tol, (name, summary[str(name)], want)


def test_oracle_matches_the_reference_results_class():
    from oracle import np_oracle
    g = load_golden('evaluation.npz')
    camera = golden_camera(g)
    out = np_oracle.evaluation_match(g['kp_point'], g['kp_count'], g['n_objects'], g['T_WC'], g['scene_points'],
                                     np_oracle.camera_dict(camera), camera.image_size)
    assert set(np.unique(out["status"])) == {-1, 0, 1, 2, 3}, "fixture must exercise every status code"
    check_against_reference(out, g, np_oracle.evaluation_summary(out['status'], out['err'], out['err_xy']))


@pytest.mark.gpu
def test_cuda_evaluation_matches_reference_and_oracle():
    import torch
    from oracle import np_oracle
    from object_keypoints_b200 import evaluation
    g = load_golden('evaluation.npz')
    camera = golden_camera(g)
    results = evaluation.Results()
    results.set_calibration(camera)
    out = results.add_batch(torch.from_numpy(g['kp_point']).cuda(), torch.from_numpy(g['kp_count']).cuda(),
                            torch.from_numpy(g['n_objects']).cuda(), g['T_WC'], g['scene_points'])
    got = {k: v.cpu().numpy() for k, v in out.items()}
    want = np_oracle.evaluation_match(g['kp_point'], g['kp_count'], g['n_objects'], g['T_WC'], g['scene_points'],
                                      np_oracle.camera_dict(camera), camera.image_size)
    np.testing.assert_array_equal(got['status'], want['status'])
    np.testing.assert_array_equal(got['gt_object'], want['gt_object'])
    for key in ('gt_point', 'err', 'err_xy'):
        np.testing.assert_allclose(got[key], want[key], rtol=0, atol=1e-12, err_msg=key)
    check_against_reference(got, g, results.summary())
    # the reference's per-frame interface: objects as ObjectKeypointPipeline returns them
    single = evaluation.Results()
    single.set_calibration(camera)
    for n in range(g['kp_point'].shape[0]):
        objects = []
        for o in range(int(g['n_objects'][n])):
            p_C = [g['kp_point'][n, o, c, :g['kp_count'][n, o, c]] if g['kp_count'][n, o, c] else None
                   for c in range(g['kp_count'].shape[2])]
            objects.append({'p_C': p_C})
        single.add(g['T_WC'][n], objects, g['scene_points'])
    for name, value in results.summary().items():
        assert abs(single.summary()[name] - value) <= 1e-9 * max(abs(value), 1.0), name
