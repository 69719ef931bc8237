"""Ground-truth / training targets on the GPU: the heatmap, centre-vector and depth
maps the reference's dataset builds per frame on the CPU (``perception/datasets/video.py:44-53,195-263``),
for a whole batch in one kernel (okp_rasterise_targets_f32)."""
import ctypes

import numpy as np
import torch

from . import _abi, _lib
from .pipeline import _device, _stream_handle

KERNEL_SIZE = 8          # int(heatmap_size / 8)     video.py:19
LENGTH_SCALE = 2.0       # heatmap_size / 32         video.py:20
CENTER_RADIUS = 4.0      # heatmap_size / 16         video.py:18


def rasterise_targets(keypoints, depths, keypoint_config, size, n_objects=None, device=None, kernel_size=KERNEL_SIZE,
                      length_scale=LENGTH_SCALE, center_radius=CENTER_RADIUS, stream=None):
    """keypoints [N, G, Kp, 2] (x, y) in target pixels, Kp = 1 + sum(keypoint_config), each object's centre first;
    depths [N, G, Kp] camera-frame z; n_objects [N] (default: all G). -> (heat [N,C,H,W], depth [N,C,H,W],
    centers [N,C-1,2,H,W]) float32 CUDA tensors, the argument order ObjectKeypointPipeline.__call__ takes."""
    cfg = _abi.check_keypoint_config(keypoint_config)
    device = _device(device)

    def dev(x, dtype):
        x = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
        return x.to(device=device, dtype=dtype).contiguous()
    keypoints, depths = dev(keypoints, torch.float64), dev(depths, torch.float64)
    N, G, Kp = keypoints.shape[:3]
    C = 1 + len(cfg)
    if Kp != 1 + sum(cfg) or tuple(depths.shape) != (N, G, Kp) or keypoints.shape[3] != 2:
        raise ValueError(f"keypoints must be [N, G, {1 + sum(cfg)}, 2] and depths [N, G, {1 + sum(cfg)}]")
    H, W = int(size[0]), int(size[1])
    counts = None if n_objects is None else dev(n_objects, torch.int32)
    heat = torch.empty((N, C, H, W), dtype=torch.float32, device=device)
    depth = torch.empty((N, C, H, W), dtype=torch.float32, device=device)
    centers = torch.empty((N, C - 1, 2, H, W), dtype=torch.float32, device=device)
    cfg_array = (ctypes.c_int32 * max(len(cfg), 1))(*cfg)
    rc = _lib.lib().okp_rasterise_targets_f32(
        keypoints.data_ptr(), depths.data_ptr(), None if counts is None else counts.data_ptr(), N, G, C, H, W, cfg_array,
        int(kernel_size), float(length_scale), float(center_radius), heat.data_ptr(),
        centers.data_ptr() if C > 1 else None, depth.data_ptr(), _stream_handle(stream))
    _lib.check(rc, 'okp_rasterise_targets_f32')
    return heat, depth, centers
