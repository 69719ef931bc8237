"""TEST INFRASTRUCTURE -- statement of record of BASELINE config 3's chain (object_keypoints_b200/sequence.py):
tracks of a moving camera. The reference stops at two views (scripts/label.py:285-305: T_RL = inv(T_WR) @ T_WL,
P1 = K [I|0], P2 = K [I|0] T_RL, undistort, cv2.triangulatePoints); north_star asks for V views with a reprojection
filter, so PARITY IS UNPINNED for V > 2 (no reference behaviour exists); the pieces are pinned where a reference exists:
undistortion and two-view DLT against OpenCV (tests/golden/geometry.npz), association against the scenarios of
test/test_pipeline.py:208-261.

Chain: undistort every peak -> per (anchor frame, view frame, map) epipolar association of the peak lists under the
pair's fundamental matrix -> tracks -> robust V-view triangulation per anchor with the anchor group's poses."""
import numpy as np

from oracle import c_oracle, np_oracle


def inv_transform(T):
    """SE(3) inverse [R^T | -R^T t] (perception/utils/linalg.py:4-10)."""
    out = np.eye(4)
    out[:3, :3] = T[:3, :3].T
    out[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return out


def fundamental_matrix(T_RL, K, Kp):
    """perception/utils/camera_utils.py:184-189: F = Kp^-T R K^T [K R^T t]x."""
    R, t = T_RL[:3, :3], T_RL[:3, 3]
    A = K @ R.T @ t
    C = np.array([[0.0, -A[2], A[1]], [A[2], 0.0, -A[0]], [-A[1], A[0], 0.0]])
    return np.linalg.inv(Kp).T @ R @ K.T @ C


def sequence_tracks(tables, T_CW, camera, views, max_distance_px=2.5, max_error_px=2.0, max_rounds=None):
    """tables: decode tables (NumPy) of the N frames; T_CW [N,4,4]; camera: object with K, D, Kinv, image_size.
    -> dict(points [A,C,K,3], valid / observed [A,C,K,V] uint8, error [A,C,K,V], dropped [A,C,K], match [A,V-1,C,K])."""
    peak_xy, peak_count = tables['peak_xy'], tables['peak_count']
    N, C, K = peak_xy.shape[:3]
    stride = N // views
    frames = np.arange(stride)[:, None] + stride * np.arange(views)[None, :]
    A, V = frames.shape
    Kmat = np.asarray(camera.K, dtype=np.float64)
    xy = peak_xy.astype(np.float64)
    und = c_oracle.undistort(xy.reshape(-1, 2), camera).reshape(N, C, K, 2)
    count = np.minimum(peak_count, K)
    match = np.full((A, V - 1, C, K), -1, np.int32)
    obs = np.zeros((A, C, K, V, 2))
    observed = np.zeros((A, C, K, V), np.uint8)
    for a in range(A):
        f0 = frames[a, 0]
        T0_inv = inv_transform(T_CW[f0])
        for c in range(C):
            n0 = int(count[f0, c])
            obs[a, c, :n0, 0] = xy[f0, c, :n0]
            observed[a, c, :n0, 0] = 1
        for v in range(1, V):
            f1 = frames[a, v]
            F = fundamental_matrix(T_CW[f1] @ T0_inv, Kmat, Kmat)
            for c in range(C):
                n0, n1 = int(count[f0, c]), int(count[f1, c])
                if n0 == 0 or n1 == 0:
                    continue
                m, _ = np_oracle.associate(F, und[f0, c, :n0], und[f1, c, :n1], max_distance_px)
                match[a, v - 1, c, :n0] = m
                for k in range(n0):
                    if m[k] >= 0:
                        obs[a, c, k, v] = xy[f1, c, m[k]]
                        observed[a, c, k, v] = 1
    rounds = V if max_rounds is None else int(max_rounds)
    points = np.zeros((A, C, K, 3))
    valid = observed.copy()
    error = np.zeros((A, C, K, V))
    dropped = np.zeros((A, C, K), np.int32)
    for a in range(A):
        X, va, err, dr = c_oracle.triangulate_robust(obs[a].reshape(C * K, V, 2), valid[a].reshape(C * K, V), T_CW[frames[a]],
                                                     camera, max_error_px, rounds)
        points[a], valid[a], error[a], dropped[a] = X.reshape(C, K, 3), va.reshape(C, K, V), err.reshape(C, K, V), dr.reshape(C, K)
    return {'points': points, 'valid': valid, 'observed': observed, 'error': error, 'dropped': dropped, 'match': match,
            'frames': frames, 'observations': obs}
