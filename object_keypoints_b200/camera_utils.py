"""Kalibr pinhole cameras (equidistant and radtan) with the class names, constructor
arguments and method meaning of the reference's ``perception/utils/camera_utils.py``.

The reference delegates the arithmetic to OpenCV; here it is restated in float64 NumPy
so that the camera objects do not need OpenCV and so that the very same formulas can be
read next to the CUDA implementation in ``csrc/okp_geometry.cuh``:

* equidistant projection   -- cv2.fisheye.projectPoints   (camera_utils.py:65-73)
* equidistant undistortion -- cv2.fisheye.undistortPoints (camera_utils.py:75-81)
* radtan projection / undistortion -- cv2.projectPoints / cv2.undistortPoints
  (camera_utils.py:45-62)

These host methods are calibration-time utilities (building synthetic scenes, loading
YAML). The pipeline itself never calls them: it hands K, D, Kinv and image_size to the
CUDA library (see ``pack_camera``).
"""
import numpy as np
import yaml

from . import linalg

_UNDISTORT_ITERS = 10          # cv::fisheye::undistortPoints default TermCriteria count
_UNDISTORT_EPS = 1e-8          # ... and epsilon
_FAR_AWAY = -1000000.0         # value OpenCV writes for points that do not converge


class PinholeCamera:
    """K (3x3), D (4,), image_size (height, width). Reference: camera_utils.py:7-43."""

    def __init__(self, K, D, image_size):
        self.K = np.asarray(K, dtype=np.float64)
        self.Kinv = np.linalg.inv(self.K)
        self.D = np.asarray(D, dtype=np.float64)
        self.image_size = np.array(image_size)
        # Same sanity check as the reference (principal point near the image centre).
        assert np.abs(self.K[0, 2] * 2.0 - self.image_size[1]) < 0.05 * self.image_size[1]

    def scale(self, scale):
        K = scale_camera_matrix(self.K, np.ones(2) * scale)
        return FisheyeCamera(K, self.D, self.image_size * scale)

    def cut(self, offset):
        offset = np.asarray(offset, dtype=np.float64)
        K = self.K.copy()
        K[0, 2] = self.K[0, 2] - offset[0]
        K[1, 2] = self.K[1, 2] - offset[1]
        return FisheyeCamera(K, self.D, self.image_size - 2.0 * offset[::-1])

    def unproject(self, xys, zs):
        """Pixels (already undistorted) and camera-frame depths -> N x 3 points."""
        xys = np.asarray(xys)
        homogeneous = np.concatenate([xys, np.ones((xys.shape[0], 1))], axis=1)
        rays = (self.Kinv @ homogeneous[:, :, None])[:, :, 0]
        return rays * np.asarray(zs)[:, None]

    def in_frame(self, x):
        x = np.asarray(x)
        outside = (x <= 0.0).any(axis=1) | (x >= self.image_size).any(axis=1)
        return ~outside


def _apply_extrinsics(X, T_CW):
    X = np.asarray(X, dtype=np.float64)
    T_CW = np.asarray(T_CW, dtype=np.float64)
    return X @ T_CW[:3, :3].T + T_CW[:3, 3]


class FisheyeCamera(PinholeCamera):
    """Kalibr ``pinhole`` + ``equidistant`` model (camera_utils.py:64-81)."""

    def project(self, X, T_CW=np.eye(4)):
        """N x 3 points in the frame T_CW maps from -> N x 2 distorted pixels (float64)."""
        Xc = _apply_extrinsics(X, T_CW)
        a = Xc[:, 0] / Xc[:, 2]
        b = Xc[:, 1] / Xc[:, 2]
        r = np.sqrt(a * a + b * b)
        theta = np.arctan(r)
        t2 = theta * theta
        k1, k2, k3, k4 = self.D[:4]
        theta_d = theta * (1.0 + k1 * t2 + k2 * t2 * t2 + k3 * t2 * t2 * t2 + k4 * t2 * t2 * t2 * t2)
        with np.errstate(divide='ignore', invalid='ignore'):
            s = np.where(r > 1e-8, theta_d / r, 1.0)
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        return np.stack([fx * (a * s) + cx, fy * (b * s) + cy], axis=1)

    def undistort(self, xy):
        """N x 2 distorted pixels -> N x 2 pixels of the ideal pinhole camera K.

        The result has the dtype of ``xy`` (float32 in, float32 out), like OpenCV."""
        xy = np.asarray(xy)
        out_dtype = xy.dtype if xy.dtype in (np.float32, np.float64) else np.float64
        out = undistort_equidistant(xy.astype(np.float64), self.K, self.D)
        return out.astype(out_dtype)


def undistort_equidistant(xy, K, D):
    """Float64 restatement of cv::fisheye::undistortPoints(xy, K, D, P=K)."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    k1, k2, k3, k4 = np.asarray(D, dtype=np.float64)[:4]
    out = np.empty((xy.shape[0], 2), dtype=np.float64)
    for n in range(xy.shape[0]):
        px = (xy[n, 0] - cx) / fx
        py = (xy[n, 1] - cy) / fy
        theta_d = np.sqrt(px * px + py * py)
        theta_d = min(max(-np.pi / 2.0, theta_d), np.pi / 2.0)
        converged = False
        theta = theta_d
        scale = 0.0
        if abs(theta_d) > _UNDISTORT_EPS:
            for _ in range(_UNDISTORT_ITERS):
                t2 = theta * theta
                t4 = t2 * t2
                t6 = t4 * t2
                t8 = t6 * t2
                k0t2, k1t4, k2t6, k3t8 = k1 * t2, k2 * t4, k3 * t6, k4 * t8
                fix = (theta * (1 + k0t2 + k1t4 + k2t6 + k3t8) - theta_d) / \
                      (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8)
                theta = theta - fix
                if abs(fix) < _UNDISTORT_EPS:
                    converged = True
                    break
            scale = np.tan(theta) / theta_d
        else:
            converged = True
        flipped = (theta_d < 0 and theta > 0) or (theta_d > 0 and theta < 0)
        if converged and not flipped:
            out[n, 0] = fx * (px * scale) + cx
            out[n, 1] = fy * (py * scale) + cy
        else:
            out[n, 0] = _FAR_AWAY
            out[n, 1] = _FAR_AWAY
    return out


class RadTanPinholeCamera(PinholeCamera):
    """Kalibr ``pinhole`` + ``radtan`` (k1, k2, p1, p2) model (camera_utils.py:45-62).

    Host-side only: the CUDA path covers the equidistant model the pipeline is run with
    (config/calibration.yaml)."""

    def _distort(self, a, b):
        k1, k2, p1, p2 = self.D[:4]
        r2 = a * a + b * b
        radial = 1.0 + k1 * r2 + k2 * r2 * r2
        ad = a * radial + 2.0 * p1 * a * b + p2 * (r2 + 2.0 * a * a)
        bd = b * radial + p1 * (r2 + 2.0 * b * b) + 2.0 * p2 * a * b
        return ad, bd

    def project(self, X, T_CW=np.eye(4)):
        Xc = _apply_extrinsics(X, T_CW)
        ad, bd = self._distort(Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2])
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        return np.stack([fx * ad + cx, fy * bd + cy], axis=1)

    def undistort(self, xy):
        xy = np.asarray(xy)
        out_dtype = xy.dtype if xy.dtype in (np.float32, np.float64) else np.float64
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        k1, k2, p1, p2 = self.D[:4]
        x0 = (xy[:, 0].astype(np.float64) - cx) / fx
        y0 = (xy[:, 1].astype(np.float64) - cy) / fy
        x, y = x0.copy(), y0.copy()
        for _ in range(5):                       # cv::undistortPoints default: 5 fixed-point steps
            r2 = x * x + y * y
            icdist = 1.0 / (1.0 + (k2 * r2 + k1) * r2)
            dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
            dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
            x = (x0 - dx) * icdist
            y = (y0 - dy) * icdist
        return np.stack([fx * x + cx, fy * y + cy], axis=1).astype(out_dtype)


class StereoCamera:
    """Left/right camera pair and T_RL (camera_utils.py:84-117)."""

    def __init__(self, left_camera, right_camera, T_RL):
        self.left_camera = left_camera
        self.right_camera = right_camera
        self.T_RL = np.asarray(T_RL, dtype=np.float64)
        self.T_LR = linalg.inv_transform(self.T_RL)
        self.F = fundamental_matrix(self.T_RL, left_camera.K, right_camera.K)

    def projection_matrices(self):
        """P1 = K [I|0], P2 = K' T_RL[:3] -- the pair camera_utils.py:103-104 builds."""
        P1 = self.left_camera.K @ np.eye(3, 4)
        P2 = self.right_camera.K @ self.T_RL[:3]
        return P1, P2

    def triangulate(self, left_keypoints, right_keypoints, optimal_correction=True):
        """N x 2 distorted pixel pairs -> N x 3 points in the left camera frame, on the GPU.

        Follows camera_utils.py:92-110: cast to float32, undistort both views, Hartley-Sturm
        correction with F (cv2.correctMatches), two-view DLT."""
        from .triangulation import triangulate_stereo
        return triangulate_stereo(self, left_keypoints, right_keypoints,
                                  optimal_correction=optimal_correction)

    @classmethod
    def from_file(cls, calibration_file):
        params = load_calibration_params(calibration_file)
        left = FisheyeCamera(params['K'], params['D'], params['image_size'])
        right = FisheyeCamera(params['Kp'], params['Dp'], params['image_size'])
        return cls(left, right, params['T_RL'])


def camera_matrix(intrinsics):
    fx, fy, cx, cy = intrinsics
    return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])


def projection_matrix(camera_matrix, T_CW):
    """3x4 projection K T_CW[:3] (camera_utils.py:125-130)."""
    return np.asarray(camera_matrix) @ np.asarray(T_CW)[:3, :]


def _read_yaml(calibration_file):
    with open(calibration_file, 'rt') as f:
        return yaml.load(f.read(), Loader=yaml.SafeLoader)


def from_calibration(calibration_file):
    """cam0 of a Kalibr YAML -> camera object (camera_utils.py:132-144)."""
    camera = _read_yaml(calibration_file)['cam0']
    K = camera_matrix(camera['intrinsics'])
    D = np.array(camera['distortion_coeffs'])
    model = (camera['camera_model'], camera['distortion_model'])
    if model == ('pinhole', 'equidistant'):
        return FisheyeCamera(K, D, camera['resolution'][::-1])
    if model == ('pinhole', 'radtan'):
        return RadTanPinholeCamera(K, D, camera['resolution'][::-1])
    raise ValueError(f"Unrecognized calibration type {camera['distortion_model']}.")


def load_calibration_params(calibration_file):
    """cam0/cam1 of a Kalibr stereo YAML -> dict (camera_utils.py:146-170)."""
    calibration = _read_yaml(calibration_file)
    left, right = calibration['cam0'], calibration['cam1']
    T_RL = np.array(right['T_cn_cnm1'])
    return {
        'K': camera_matrix(left['intrinsics']),
        'Kp': camera_matrix(right['intrinsics']),
        'D': np.array(left['distortion_coeffs']),
        'Dp': np.array(right['distortion_coeffs']),
        'T_LR': linalg.inv_transform(T_RL),
        'T_RL': T_RL,
        'image_size': right['resolution'][::-1],
    }


def scale_camera_matrix(K, scaling_factor):
    """Scale fx, cx by scaling_factor[0] and fy, cy by scaling_factor[1] (camera_utils.py:172-182)."""
    out = np.array(K, dtype=np.float64, copy=True)
    out[0, 0] = K[0, 0] * scaling_factor[0]
    out[1, 1] = K[1, 1] * scaling_factor[1]
    out[0, 2] = K[0, 2] * scaling_factor[0]
    out[1, 2] = K[1, 2] * scaling_factor[1]
    return out


def fundamental_matrix(T_RL, K, Kp):
    """F with x_R^T F x_L = 0 (camera_utils.py:184-189)."""
    R = T_RL[:3, :3]
    t = T_RL[:3, 3]
    return np.linalg.inv(Kp).T @ R @ K.T @ linalg.skew_matrix(K @ R.T @ t)
