"""Sweeps the strip kernel's launch-plan knobs (environment variables read by okp_strip_plan at every
call) in one process. usage: python tools/sweep_k1.py [180x320|64x64] [frames] [reps]"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from object_keypoints_b200 import KeypointDecoder, synthetic

shape = sys.argv[1] if len(sys.argv) > 1 else '180x320'
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
H, W = [int(v) for v in shape.split('x')]
grid = (4, 2) if W >= 128 else (2, 1)
heat, depth, centers, _ = synthetic.torch_grid_batch(frames, [1, 3], (H, W), seed=7, grid=grid, device='cuda')
dec = KeypointDecoder([1, 3], (H, W), camera=synthetic.default_camera((H, W)))
tables = dec.tables(frames)
gb = frames * 3 * H * W * 4 / 1e9
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
rows = []
for stages, smem, service in itertools.product([2, 3, 4], [110, 220], [288, 448, 608]):
    os.environ['OKP_STRIP_STAGES'] = str(stages)
    os.environ['OKP_STRIP_SMEM_KB'] = str(smem)
    os.environ['OKP_STRIP_THREADS'] = str(service)
    try:
        for _ in range(2):
            dec.extract_peaks(heat, tables)
        torch.cuda.synchronize()
        total = 0.0
        for _ in range(reps):
            ev[0].record(); dec.extract_peaks(heat, tables); ev[1].record()
            torch.cuda.synchronize()
            total += ev[0].elapsed_time(ev[1])
        ms = total / reps
        rows.append((gb / (ms / 1e3), stages, smem, service, ms))
        print(f"{shape} stages={stages} smem_kb={smem} threads={service}: {ms * 1e3:.1f} us = {gb / (ms / 1e3):.0f} GB/s", flush=True)
    except Exception as e:
        print(f"{shape} stages={stages} smem_kb={smem} threads={service}: FAILED {e}", flush=True)
        break
rows.sort(reverse=True)
print("best:", rows[:3])
