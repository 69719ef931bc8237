"""Camera geometry and triangulation on the GPU (libokp.so), behind the reference's names.

* ``undistort_points`` / ``project_points`` -- FisheyeCamera.undistort / .project
  (perception/utils/camera_utils.py:65-81) for arrays of points;
* ``triangulate`` -- batched V-view DLT; at V = 2 it is cv2.triangulatePoints as used by
  StereoCamera.triangulate (camera_utils.py:103-108) and LabelingApp._triangulate
  (scripts/label.py:296-305);
* ``TriangulationComponent`` -- the ``reset(stereo)`` / ``__call__(left, right)`` component the
  reference's test-suite expects (test/test_pipeline.py:171-177), pairs matched by index;
* ``triangulate_multiview`` -- undistort, DLT over V views, reprojection-error gate, re-solve.
"""
import ctypes

import numpy as np
import torch

from . import _abi, _lib
from .pipeline import _device, _stream_handle


def _as_f64(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return x.to(device=device, dtype=torch.float64).contiguous()


def undistort_points(xy, camera, round_to_f32=False, device=None):
    """[n,2] distorted pixels -> [n,2] float64 CUDA tensor of pinhole pixels."""
    device = _device(device)
    xy = _as_f64(xy, device)
    out = torch.empty_like(xy)
    cam = _abi.pack_camera(camera)
    rc = _lib.lib().okp_fisheye_undistort_f64(xy.data_ptr(), xy.shape[0], ctypes.byref(cam), int(round_to_f32),
                                              out.data_ptr(), _stream_handle())
    _lib.check(rc, 'okp_fisheye_undistort_f64')
    return out


def project_points(X, T_CW, camera, device=None):
    """[n,3] points, 4x4 world->camera -> [n,2] float64 CUDA tensor of distorted pixels."""
    device = _device(device)
    X = _as_f64(X, device)
    out = torch.empty((X.shape[0], 2), dtype=torch.float64, device=device)
    cam = _abi.pack_camera(camera)
    T = np.ascontiguousarray(np.asarray(T_CW, dtype=np.float64).reshape(-1)[:16])
    rc = _lib.lib().okp_fisheye_project_f64(X.data_ptr(), X.shape[0], T.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                            ctypes.byref(cam), out.data_ptr(), _stream_handle())
    _lib.check(rc, 'okp_fisheye_project_f64')
    return out


def triangulate(points, projections, valid=None, device=None):
    """points [P,V,2] undistorted pixels, projections [V,3,4] (shared) or [P,V,3,4], valid [P,V]
    -> [P,3] float64 CUDA tensor (NaN where fewer than two valid views)."""
    device = _device(device)
    points = _as_f64(points, device)
    projections = _as_f64(projections, device)
    P, V = int(points.shape[0]), int(points.shape[1])
    per_point = int(projections.dim() == 4)
    valid_ptr = None
    if valid is not None:
        if isinstance(valid, np.ndarray):
            valid = torch.from_numpy(np.ascontiguousarray(valid))
        valid = valid.to(device=device, dtype=torch.uint8).contiguous()
        valid_ptr = valid.data_ptr()
    out = torch.empty((P, 3), dtype=torch.float64, device=device)
    rc = _lib.lib().okp_triangulate_f64(points.data_ptr(), valid_ptr, projections.data_ptr(), per_point, P, V,
                                        out.data_ptr(), _stream_handle())
    _lib.check(rc, 'okp_triangulate_f64')
    return out


def reprojection_filter(X, observations, valid, poses, camera, max_error_px, device=None):
    """X [P,3], observations [P,V,2] distorted pixels, valid [P,V], poses [V,4,4] world->camera ->
    (valid' [P,V] uint8, error [P,V] float64): views with error > max_error_px are cleared."""
    device = _device(device)
    X = _as_f64(X, device)
    observations = _as_f64(observations, device)
    poses = _as_f64(poses, device)
    P, V = int(observations.shape[0]), int(observations.shape[1])
    if isinstance(valid, np.ndarray):
        valid = torch.from_numpy(np.ascontiguousarray(valid))
    valid = valid.to(device=device, dtype=torch.uint8).contiguous().clone()
    err = torch.empty((P, V), dtype=torch.float64, device=device)
    cam = _abi.pack_camera(camera)
    rc = _lib.lib().okp_reprojection_filter_f64(X.data_ptr(), observations.data_ptr(), valid.data_ptr(), poses.data_ptr(),
                                                ctypes.byref(cam), P, V, float(max_error_px), err.data_ptr(),
                                                _stream_handle())
    _lib.check(rc, 'okp_reprojection_filter_f64')
    return valid, err


def triangulate_multiview(observations, valid, poses, camera, max_error_px=2.0, max_rounds=None, device=None,
                          return_dropped=False):
    """Config-3 path in ONE kernel (okp_triangulate_robust_f64): observations [P,V,2] distorted pixels
    seen from poses [V,4,4] (world->camera) with one equidistant camera. Per point: undistort, V-view
    DLT, reprojection error per view; while the worst valid view is farther than max_error_px (and
    more than two views remain, at most max_rounds drops) drop it and solve again.
    Returns (X [P,3], valid [P,V] uint8, error [P,V]) CUDA tensors."""
    device = _device(device)
    observations = _as_f64(observations, device)
    P, V = int(observations.shape[0]), int(observations.shape[1])
    poses_t = _as_f64(poses, device)
    if valid is None:
        valid = torch.ones((P, V), dtype=torch.uint8, device=device)
    else:
        if isinstance(valid, np.ndarray):
            valid = torch.from_numpy(np.ascontiguousarray(valid))
        valid = valid.to(device=device, dtype=torch.uint8).contiguous().clone()
    X = torch.empty((P, 3), dtype=torch.float64, device=device)
    err = torch.empty((P, V), dtype=torch.float64, device=device)
    dropped = torch.zeros((P,), dtype=torch.int32, device=device)
    cam = _abi.pack_camera(camera)
    rounds = V if max_rounds is None else int(max_rounds)
    rc = _lib.lib().okp_triangulate_robust_f64(observations.data_ptr(), valid.data_ptr(), poses_t.data_ptr(),
                                               ctypes.byref(cam), P, V, float(max_error_px), rounds, X.data_ptr(),
                                               err.data_ptr(), dropped.data_ptr(), _stream_handle())
    _lib.check(rc, 'okp_triangulate_robust_f64')
    if return_dropped:
        return X, valid, err, dropped
    return X, valid, err


def triangulate_stereo(stereo, left_keypoints, right_keypoints, optimal_correction=False):
    """StereoCamera.triangulate (camera_utils.py:92-110) on the GPU: float32 cast, undistort both
    views, [optional Hartley-Sturm correction], two-view DLT. Returns [N,3] float64 NumPy in the
    left camera frame."""
    left = np.asarray(left_keypoints).astype(np.float32).astype(np.float64)      # camera_utils.py:93-94
    right = np.asarray(right_keypoints).astype(np.float32).astype(np.float64)
    uL = undistort_points(left, stereo.left_camera, round_to_f32=True)
    uR = undistort_points(right, stereo.right_camera, round_to_f32=True)
    if optimal_correction:
        raise NotImplementedError("Hartley-Sturm correction (cv2.correctMatches) is not on the GPU path yet")
    P1, P2 = stereo.projection_matrices()
    points = torch.stack([uL, uR], dim=1)
    X = triangulate(points, np.stack([P1, P2]))
    return X.cpu().numpy()


class TriangulationComponent:
    """The component test/test_pipeline.py:171-177 uses: ``reset(stereo_camera)`` then
    ``__call__(left[N,2], right[N,2])`` -> [N,3] points in the left camera frame."""
    name = "triangulation"

    def __init__(self, optimal_correction=False):
        self.optimal_correction = optimal_correction
        self.stereo_camera = None

    def reset(self, stereo_camera):
        self.stereo_camera = stereo_camera

    def __call__(self, left_keypoints, right_keypoints):
        return triangulate_stereo(self.stereo_camera, left_keypoints, right_keypoints,
                                  optimal_correction=self.optimal_correction)
