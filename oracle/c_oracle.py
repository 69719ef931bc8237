"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/liboracle.so (okp_oracle.c).

Checker and CPU baseline only; never imported by the package. Tables are NumPy arrays with
the layout of include/okp.h, so they compare one to one with what the CUDA path returns.
"""
import ctypes
import os
import subprocess

import numpy as np

from object_keypoints_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(['make', '-C', HERE], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, 'liboracle.so')
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.okp_oracle_decode_f32.restype = ctypes.c_int
        _LIB.okp_oracle_triangulate_f64.restype = ctypes.c_int
        _LIB.okp_oracle_triangulate_robust_f64.restype = ctypes.c_int
        _LIB.okp_oracle_max_threads.restype = ctypes.c_int
    return _LIB


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def allocate_tables(N, C, keypoint_config, params):
    arrays = {name: np.zeros(shape, dtype=dtype)
              for name, dtype, shape in _abi.table_shapes(N, C, keypoint_config, params)}
    tables = _abi.OkpDecodeTables(**{name: a.ctypes.data for name, a in arrays.items()})
    return arrays, tables


def decode(heat, depth, centers, keypoint_config, camera, threads=0, tables=None, **params):
    """heat [N,C,H,W], depth [N,C,H,W], centers [N,T,2,H,W] float32 NumPy -> dict of tables.
    camera: object with K, D, Kinv, image_size (or None: no 3D points)."""
    heat = np.ascontiguousarray(heat, dtype=np.float32)
    depth = None if depth is None else np.ascontiguousarray(depth, dtype=np.float32)
    centers = None if centers is None else np.ascontiguousarray(centers, dtype=np.float32)
    N, C, H, W = heat.shape
    cfg = _abi.check_keypoint_config(keypoint_config)
    prm = _abi.make_params(**params)
    if tables is None:
        arrays, tab = allocate_tables(N, C, cfg, prm)
    else:
        arrays, tab = tables
    cfg_arr = (ctypes.c_int32 * max(len(cfg), 1))(*cfg)
    cam = _abi.pack_camera(camera) if camera is not None else None
    rc = lib().okp_oracle_decode_f32(_ptr(heat), _ptr(depth), _ptr(centers), N, C, H, W, cfg_arr,
                                     ctypes.byref(cam) if cam is not None else None, ctypes.byref(prm),
                                     ctypes.byref(tab), int(threads), 1)
    if rc != 0:
        raise RuntimeError(f"oracle decode failed: {_abi.ERRORS.get(rc, rc)}")
    return arrays


def max_threads():
    return int(lib().okp_oracle_max_threads())


def undistort(xy, camera, round_to_f32=False):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    out = np.empty_like(xy)
    cam = _abi.pack_camera(camera)
    lib().okp_oracle_undistort_f64(_ptr(xy), xy.shape[0], ctypes.byref(cam), int(round_to_f32), _ptr(out))
    return out


def project(X, T_CW, camera):
    X = np.ascontiguousarray(X, dtype=np.float64)
    T = np.ascontiguousarray(T_CW, dtype=np.float64)
    out = np.empty((X.shape[0], 2), dtype=np.float64)
    cam = _abi.pack_camera(camera)
    lib().okp_oracle_project_f64(_ptr(X), X.shape[0], _ptr(T), ctypes.byref(cam), _ptr(out))
    return out


def detection_to_point(xy, depth_map, camera, compat_clip_bug=True):
    xy = np.ascontiguousarray(xy, dtype=np.float32)
    depth_map = np.ascontiguousarray(depth_map, dtype=np.float32)
    out = np.empty((xy.shape[0], 3), dtype=np.float64)
    cam = _abi.pack_camera(camera)
    prm = _abi.make_params(compat_clip_bug=compat_clip_bug)
    lib().okp_oracle_detection_to_point_f32(_ptr(xy), xy.shape[0], _ptr(depth_map), depth_map.shape[0],
                                            depth_map.shape[1], ctypes.byref(cam), ctypes.byref(prm), _ptr(out))
    return out


def triangulate(points, valid, projections):
    points = np.ascontiguousarray(points, dtype=np.float64)
    P, V = points.shape[:2]
    projections = np.ascontiguousarray(projections, dtype=np.float64)
    per_point = int(projections.ndim == 4)
    valid_arr = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    out = np.empty((P, 3), dtype=np.float64)
    rc = lib().okp_oracle_triangulate_f64(_ptr(points), _ptr(valid_arr), _ptr(projections), per_point, P, V, _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle triangulate failed: {_abi.ERRORS.get(rc, rc)}")
    return out


def reprojection_filter(X, obs, valid, poses, camera, max_error_px):
    X = np.ascontiguousarray(X, dtype=np.float64)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    P, V = obs.shape[:2]
    valid = np.ascontiguousarray(valid, dtype=np.uint8).copy()
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    err = np.empty((P, V), dtype=np.float64)
    cam = _abi.pack_camera(camera)
    lib().okp_oracle_reprojection_filter_f64(_ptr(X), _ptr(obs), _ptr(valid), _ptr(poses), ctypes.byref(cam),
                                             P, V, ctypes.c_double(max_error_px), _ptr(err))
    return valid, err


def triangulate_robust(obs, valid, poses, camera, max_error_px, max_rounds):
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    P, V = obs.shape[:2]
    valid = np.ones((P, V), np.uint8) if valid is None else np.ascontiguousarray(valid, dtype=np.uint8).copy()
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    X = np.empty((P, 3), dtype=np.float64)
    err = np.empty((P, V), dtype=np.float64)
    dropped = np.zeros((P,), dtype=np.int32)
    cam = _abi.pack_camera(camera)
    rc = lib().okp_oracle_triangulate_robust_f64(_ptr(obs), _ptr(valid), _ptr(poses), ctypes.byref(cam), P, V,
                                                 ctypes.c_double(max_error_px), int(max_rounds), _ptr(X), _ptr(err),
                                                 _ptr(dropped))
    if rc != 0:
        raise RuntimeError(f"oracle triangulate_robust failed: {_abi.ERRORS.get(rc, rc)}")
    return X, valid, err, dropped
