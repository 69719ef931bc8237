"""Throughput of the host pass of the sparse transfer (okp_host_pack_tiles_f32) against the thread count, on one chunk
of the bench workload (256 valve frames of 180x320 = 177 MB). usage: python tools/bench_host_pack.py"""
import ctypes
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from object_keypoints_b200 import _lib, synthetic

L = _lib.lib()
frames, (H, W), C = 256, (180, 320), 3
device = 'cuda' if torch.cuda.is_available() else 'cpu'
heat = synthetic.torch_grid_batch(frames, [1, 3], (H, W), seed=1004, grid=(4, 2), device=device, chunk=64)[0].cpu()
if torch.cuda.is_available():
    heat = heat.pin_memory()
maps = frames * C
tiles = ((H + 3) // 4) * ((W + 15) // 16)
cap = maps * tiles
scratch = np.zeros(L.okp_host_pack_scratch_bytes(maps, H, W), np.uint8)
offsets = np.zeros(maps + 1, np.int64)
ids = torch.empty(cap, dtype=torch.int32)
packed = torch.empty((cap, 64), dtype=torch.float32)
if torch.cuda.is_available():
    ids, packed = ids.pin_memory(), packed.pin_memory()
count = ctypes.c_longlong()
cores = len(os.sched_getaffinity(0))
out = {'cores': cores, 'chunk_MB': heat.numel() * 4 / 1e6}
for threads in sorted({1, 2, 4, 8, 16, cores // 2, cores}):
    if threads < 1 or threads > cores:
        continue
    def run():
        L.okp_host_pack_tiles_f32(heat.data_ptr(), maps, H, W, ctypes.c_float(0.5), scratch.ctypes.data, offsets.ctypes.data,
                                  ids.data_ptr(), packed.data_ptr(), cap, ctypes.byref(count), threads)
    run()
    t0 = time.perf_counter()
    for _ in range(5):
        run()
    dt = (time.perf_counter() - t0) / 5
    out[f'threads_{threads}'] = {'ms': dt * 1e3, 'GB_per_s': heat.numel() * 4 / dt / 1e9, 'frames_per_s': frames / dt}
out['marked_fraction'] = count.value / cap
print(json.dumps(out))
