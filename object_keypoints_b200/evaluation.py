"""Evaluation bookkeeping on the GPU: drop-in for ``Results`` of the reference's
``scripts/eval_model.py:137-232`` (``set_calibration`` / ``add`` / ``print_results``), plus the batched
form ``add_batch`` that scores a whole sequence straight from the decode tables (no host round trip).

The matching and the error statistics run in libokp.so (okp_eval_match_f64, okp_eval_summary_f64);
PyTorch owns the buffers and sorts the error vector for the two percentiles.
"""
import ctypes

import numpy as np
import torch

from . import _abi, _lib
from .pipeline import _device, _stream_handle

MATCHED, MISSING, POINT_NOT_IN_VIEW, OBJECT_NOT_IN_VIEW, EMPTY = 0, 1, 2, 3, -1
COLUMNS = ('mean', 'mean_xy', 'std', 'small', 'percentile25', 'percentile75', 'missing_percentage', 'points')


class Results:
    """eval_model.py:137-232. ``add(T_WC, objects, scene_points)`` takes what the reference's loop passes
    (:290): the frame's camera->world pose, the list of object dicts of ObjectKeypointPipeline.__call__
    and the scene's ground truth [G, Kp, 3] (row 0 of each object = its centre)."""
    MAX_COORDINATE = 2.0          # :171
    SMALL_ERROR = 0.03            # :210

    def __init__(self, device=None):
        self.camera = None
        self._packed = None
        self.device = None if device is None else torch.device(device)
        self._frames = []          # per add/add_batch call: dict of device tensors
        self._lib = _lib.lib()

    def set_calibration(self, camera):
        self.camera = camera
        self._packed = _abi.pack_camera(camera)
        size = np.asarray(camera.image_size, dtype=np.float64)
        self._limits = (float(size[0]), float(size[1]))     # in_frame compares (x, y) with image_size as stored

    # ---------------------------------------------------------------------------------------------
    def add_batch(self, kp_point, kp_count, n_objects, T_WC, scene_points, stream=None):
        """kp_point [N,O,C,S,3] float64, kp_count [N,O,C] int32, n_objects [N] int32: the decode tables
        (device tensors are used in place); T_WC [N,4,4]; scene_points [G,Kp,3]. Returns the per-slot device
        tensors status / gt_point / err / err_xy / gt_object of this batch. Asynchronous."""
        if self._packed is None:
            raise RuntimeError("call set_calibration(camera) first")
        device = _device(self.device if self.device is not None else
                         (kp_point.device if isinstance(kp_point, torch.Tensor) and kp_point.is_cuda else None))

        def dev(x, dtype):
            x = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
            return x.to(device=device, dtype=dtype).contiguous()
        kp_point, kp_count, n_objects = dev(kp_point, torch.float64), dev(kp_count, torch.int32), dev(n_objects, torch.int32)
        T_WC, scene = dev(T_WC, torch.float64).reshape(-1, 4, 4), dev(scene_points, torch.float64)
        N, O, C, S = kp_point.shape[:4]
        if T_WC.shape[0] != N or tuple(kp_count.shape) != (N, O, C) or scene.dim() != 3 or scene.shape[2] != 3:
            raise ValueError("inconsistent shapes")
        out = {
            'status': torch.empty((N, O, C, S), dtype=torch.int32, device=device),
            'gt_point': torch.empty((N, O, C, S, 3), dtype=torch.float64, device=device),
            'err': torch.empty((N, O, C, S), dtype=torch.float64, device=device),
            'err_xy': torch.empty((N, O, C, S), dtype=torch.float64, device=device),
            'gt_object': torch.empty((N, O), dtype=torch.int32, device=device),
        }
        stats = torch.empty((N, 8), dtype=torch.float64, device=device)
        rc = self._lib.okp_eval_match_f64(
            kp_point.data_ptr(), kp_count.data_ptr(), n_objects.data_ptr(), T_WC.data_ptr(), scene.data_ptr(),
            N, O, C, S, scene.shape[0], scene.shape[1], ctypes.byref(self._packed), self._limits[0], self._limits[1],
            self.MAX_COORDINATE, self.SMALL_ERROR, out['status'].data_ptr(), out['gt_point'].data_ptr(),
            out['err'].data_ptr(), out['err_xy'].data_ptr(), out['gt_object'].data_ptr(), stats.data_ptr(),
            _stream_handle(stream))
        _lib.check(rc, 'okp_eval_match_f64')
        self._frames.append(dict(out, stats=stats))
        return out

    def add_tables(self, tables, T_WC, scene_points, stream=None):
        """Scores the DecodeTables of KeypointDecoder.decode_batch in place."""
        return self.add_batch(tables['kp_point'], tables['kp_count'], tables['n_objects'], T_WC, scene_points, stream=stream)

    def add(self, T_WC, objects, scene_points):
        """The reference's per-frame call (eval_model.py:141): objects = list of dicts with 'p_C' = list over
        maps of (n, 3) arrays or None."""
        C = max([len(obj['p_C']) for obj in objects], default=1)
        S = max([len(p) for obj in objects for p in obj['p_C'] if p is not None], default=1)
        O = max(len(objects), 1)
        if O > _abi.OKP_MAX_OBJECTS or S > _abi.OKP_MAX_SLOTS:
            raise ValueError("too many objects / keypoints per type for one frame")
        kp_point = np.zeros((1, O, C, S, 3), np.float64)
        kp_count = np.zeros((1, O, C), np.int32)
        for o, obj in enumerate(objects):
            for c, points in enumerate(obj['p_C']):
                if points is not None and len(points):
                    kp_count[0, o, c] = len(points)
                    kp_point[0, o, c, :len(points)] = np.asarray(points, dtype=np.float64)
        return self.add_batch(kp_point, kp_count, np.array([len(objects)], np.int32),
                              np.asarray(T_WC, dtype=np.float64)[None], scene_points)

    # ---------------------------------------------------------------------------------------------
    def summary(self):
        """The row print_results prints (eval_model.py:192-232), errors in centimetres."""
        if not self._frames:
            raise RuntimeError("nothing was added")
        stats = torch.cat([f['stats'] for f in self._frames])
        totals = torch.empty(8, dtype=torch.float64, device=stats.device)
        _lib.check(self._lib.okp_eval_summary_f64(stats.data_ptr(), stats.shape[0], totals.data_ptr(), _stream_handle()),
                   'okp_eval_summary_f64')
        errors = torch.cat([f['err'][f['status'] == MATCHED] for f in self._frames])
        ordered = torch.sort(errors * 100.0).values

        def percentile(q):                      # np.percentile, linear interpolation
            position = (ordered.numel() - 1) * q / 100.0
            lo = int(np.floor(position))
            hi = min(lo + 1, ordered.numel() - 1)
            return float(ordered[lo] + (ordered[hi] - ordered[lo]) * (position - lo))
        matched, missing, small, mean, m2, sum_xy = [float(v) for v in totals[:6].cpu()]
        n_points = matched + missing
        return {
            'mean': mean * 100.0, 'mean_xy': sum_xy / matched * 100.0, 'std': (m2 / matched) ** 0.5 * 100.0,
            'small': small / n_points, 'percentile25': percentile(25), 'percentile75': percentile(75),
            'missing_percentage': missing / n_points * 100.0, 'points': int(n_points),
        }

    def print_results(self):
        row = self.summary()
        header = ["mean", "mean xy", "std", "< 3cm", "25th percentile", "75th percentile", "missing", "points"]
        cells = [f"{row['mean']}", f"{row['mean_xy']}", f"{row['std']}", f"{row['small']}", f"{row['percentile25']}",
                 f"{row['percentile75']}", f"{row['missing_percentage']:.02f}%", f"{row['points']}"]
        widths = [max(len(h), len(c)) for h, c in zip(header, cells)]
        print(" | ".join(h.ljust(w) for h, w in zip(header, widths)))
        print(" | ".join(c.ljust(w) for c, w in zip(cells, widths)))
        return row
