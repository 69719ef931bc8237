"""SE(3) helpers with the names and argument meaning of the reference's
``perception/utils/linalg.py:4-23`` (host side, float64 NumPy)."""
import numpy as np


def skew_matrix(v):
    """[v]x, the cross-product matrix (reference: perception/utils/linalg.py:4-7)."""
    v = np.asarray(v)
    out = np.zeros((3, 3), dtype=v.dtype)
    out[0, 1], out[0, 2] = -v[2], v[1]
    out[1, 0], out[1, 2] = v[2], -v[0]
    out[2, 0], out[2, 1] = -v[1], v[0]
    return out


def inv_transform(T):
    """Inverse of a rigid 4x4 transform, [R^T | -R^T t] (reference: linalg.py:9-13)."""
    T = np.asarray(T)
    Rt = T[:3, :3].T
    out = np.eye(4, dtype=T.dtype)
    out[:3, :3] = Rt
    out[:3, 3] = -Rt @ T[:3, 3]
    return out


def transform_points(T, points):
    """R p + t over an array of ... x 3 points (reference: linalg.py:15-20)."""
    points = np.asarray(points)
    return (T[:3, :3] @ points[..., None])[..., 0] + T[:3, 3]


def angle_between(R1, R2):
    """xyz Euler angles of R1^T R2 (reference: linalg.py:22-23)."""
    from scipy.spatial.transform import Rotation
    return Rotation.from_matrix(R1.T @ R2).as_euler('xyz', degrees=False)
