"""Debug helper: one extract_peaks call on a small random batch, compared with the C oracle.
usage: python tools/debug_extract.py H W N C K"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from object_keypoints_b200 import KeypointDecoder
from oracle import c_oracle
H, W, N, C, K = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (64, 64, 8, 3, 32)
rng = np.random.default_rng(0)
heat = rng.uniform(0, 0.2, (N, C, H, W)).astype(np.float32)
cfg = [1] * (C - 1)
dec = KeypointDecoder(cfg, (H, W), max_peaks=K)
t = dec.extract_peaks(heat)
torch.cuda.synchronize()
got = t.numpy()
want = c_oracle.decode(heat, np.zeros_like(heat), np.zeros((N, C - 1, 2, H, W), np.float32), cfg, None, max_peaks=K)
for key in ['peak_count', 'peak_yx', 'peak_score', 'peak_xy', 'peak_conf']:
    same = np.array_equal(got[key].view(np.uint32) if got[key].dtype == np.float32 else got[key],
                          want[key].view(np.uint32) if want[key].dtype == np.float32 else want[key])
    print(key, 'OK' if same else 'DIFFERS')
print(got['peak_count'].ravel()[:12], want['peak_count'].ravel()[:12])
