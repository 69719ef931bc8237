"""Latency of the reference-style per-frame call: ObjectKeypointPipeline.__call__(heatmap, p_depth, p_centers) with
batch-1 CPU tensors in and the list of object dicts out (what scripts/eval_model.py:287-290 does per frame; the
unmodified reference takes ~3.1 ms per 64x64 frame and ~19.7 ms per 180x320 frame on this container's CPU, SURVEY.md
section 6). usage: python tools/bench_per_frame.py [frames]"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from object_keypoints_b200 import ObjectKeypointPipeline, synthetic

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}
for size, objects in (((64, 64), (1, 2)), ((180, 320), (4, 8))):
    cfg = [1, 3]
    batch = synthetic.make_batch(32, cfg, size, seed=3, objects=objects)
    pipeline = ObjectKeypointPipeline(size, None, {'keypoint_config': cfg})
    pipeline.reset(synthetic.default_camera(size))
    inputs = [(torch.from_numpy(batch.heat[i:i + 1]), torch.from_numpy(batch.depth[i:i + 1]),
               torch.from_numpy(batch.centers[i:i + 1])) for i in range(32)]
    resident = [tuple(t.cuda() for t in frame) for frame in inputs]
    for name, source in (('cpu_tensors', inputs), ('cuda_tensors', resident)):
        for i in range(20):
            pipeline(*source[i % 32])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        found = 0
        for i in range(frames):
            found += len(pipeline(*source[i % 32]))
        elapsed = time.perf_counter() - t0
        out[f"{size[0]}x{size[1]}_{name}"] = {'ms_per_frame': 1e3 * elapsed / frames, 'objects_per_frame': found / frames}
print(json.dumps(out))
