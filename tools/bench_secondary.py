"""Timings of the kernels beside the decode path (triangulation, stereo correction / association, evaluation, target
rasterisation): CUDA events, inputs resident on the device, 10 repetitions after 3 warm-ups. One JSON object.
usage: python tools/bench_secondary.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from object_keypoints_b200 import (KeypointDecoder, synthetic, triangulate, triangulate_multiview, evaluation, targets,
                                   camera_utils)
from object_keypoints_b200.triangulation import correct_matches, associate


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / reps


rng = np.random.default_rng(0)
dev = torch.device('cuda')
camera = synthetic.default_camera((180, 320)).scale(4.0)
out = {}

# multi-view DLT (config 3: 16 views per point) and the two-view case
for V, P in ((2, 1 << 20), (16, 1 << 20)):
    X = np.stack([rng.uniform(-0.3, 0.3, 4096), rng.uniform(-0.2, 0.2, 4096), rng.uniform(0.6, 1.0, 4096)], axis=1)
    poses = np.tile(np.eye(4), (V, 1, 1))
    poses[:, 0, 3] = np.linspace(-0.2, 0.2, V)
    obs = np.stack([camera.project(X, poses[v]) for v in range(V)], axis=1)
    und = np.stack([camera.undistort(obs[:, v]) for v in range(V)], axis=1)
    points = torch.from_numpy(np.tile(und, (P // 4096, 1, 1))).to(dev)
    proj = torch.from_numpy(np.stack([camera.K @ poses[v][:3] for v in range(V)])).to(dev)
    ms = timed(lambda: triangulate(points, proj))
    out[f'triangulate_f64_V{V}'] = {'points': P, 'ms': ms, 'points_per_s': P / ms * 1e3, 'GB_per_s': P * (V * 16 + 24) / ms / 1e6}
    if V == 16:
        distorted = torch.from_numpy(np.tile(obs + rng.normal(0, 0.3, obs.shape), (P // 4096 // 4, 1, 1))).to(dev)
        poses_dev = torch.from_numpy(poses).to(dev)
        ms = timed(lambda: triangulate_multiview(distorted, None, poses_dev, camera, max_error_px=2.0))
        out['triangulate_robust_f64_V16'] = {'points': distorted.shape[0], 'ms': ms, 'points_per_s': distorted.shape[0] / ms * 1e3}

# Hartley-Sturm correction and stereo association
stereo = camera_utils.StereoCamera.from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config', 'calibration.yaml'))
n = 1 << 20
left = torch.from_numpy(rng.uniform(100, 1100, (n, 2))).to(dev)
right = left + torch.from_numpy(np.stack([rng.uniform(-60, -5, n), rng.normal(0, 0.4, n)], axis=1)).to(dev)
ms = timed(lambda: correct_matches(stereo.F, left, right))
out['correct_matches_f64'] = {'pairs': n, 'ms': ms, 'pairs_per_s': n / ms * 1e3}
B = 1 << 16
L8 = left[:B * 8].reshape(B, 8, 2).contiguous()
R8 = right[:B * 8].reshape(B, 8, 2).contiguous()
ms = timed(lambda: associate(stereo.F, L8, R8))
out['stereo_associate_f64_8x8'] = {'frame_pairs': B, 'ms': ms, 'frame_pairs_per_s': B / ms * 1e3}

# evaluation bookkeeping on the tables of a real decode, target rasterisation
cfg, size = [1, 3], (180, 320)
heat, depth, centers, _ = synthetic.torch_grid_batch(4096, cfg, size, seed=7, grid=(4, 2), device='cuda')
decoder = KeypointDecoder(cfg, size, camera=synthetic.default_camera(size))
tables = decoder.decode_batch(heat, depth, centers)
scene = torch.from_numpy(rng.uniform(-0.5, 0.5, (8, 5, 3)) + np.array([0, 0, 1.0])).to(dev)
T_WC = torch.eye(4, dtype=torch.float64, device=dev).repeat(4096, 1, 1)
results = evaluation.Results()
results.set_calibration(synthetic.default_camera(size))


def evaluate():
    results._frames.clear()
    results.add_tables(tables, T_WC, scene)


ms = timed(evaluate)
out['eval_match_f64'] = {'frames': 4096, 'ms': ms, 'frames_per_s': 4096 / ms * 1e3}
for size, frames, G in (((64, 64), 8192, 2), ((180, 320), 1024, 8)):
    kp = torch.from_numpy(rng.uniform(8, min(size) - 8, (frames, G, 5, 2))).to(dev)
    z = torch.from_numpy(rng.uniform(0.4, 1.5, (frames, G, 5))).to(dev)
    ms = timed(lambda: targets.rasterise_targets(kp, z, cfg, size))
    written = frames * (3 + 3 + 4) * size[0] * size[1] * 4
    out[f'rasterise_targets_{size[0]}x{size[1]}'] = {'frames': frames, 'ms': ms, 'frames_per_s': frames / ms * 1e3,
                                                     'GB_written_per_s': written / ms / 1e6}
print(json.dumps(out))
