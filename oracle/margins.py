"""TEST INFRASTRUCTURE -- decide whether a synthetic frame is safe for BIT-EXACT comparison.

The reference's decisions hinge on exact float equality (NMS), on a threshold, on rounding
to the nearest pixel and on arg-min / arg-max. Two correct implementations that sum in a
different order may legitimately disagree when an input sits on such a knife edge, so the
bit-exact parity set only holds frames with a safety margin on every decision
("Rejection rules"); knife-edge inputs are tested separately.
"""
import numpy as np

from . import np_oracle as oracle


def _away_from_half(value, margin):
    frac = value - np.floor(value)
    return abs(frac - 0.5) >= margin


def frame_is_clean(heat, depth, centers, keypoint_config, cam, threshold=0.5,
                   score_margin=1e-4, threshold_margin=1e-3, rounding_margin=2e-3,
                   distance_margin=1e-3, outlier_distance=20.0):
    """heat [C,H,W], depth [C,H,W], centers [T,2,H,W] of ONE frame -> (bool, reason)."""
    C, H, W = heat.shape
    cfg = [1] + list(keypoint_config)
    peaks = []
    for c in range(C):
        score = oracle.box_sum(heat[c])
        # window maximum EXCLUDING the pixel itself
        padded = np.full((H + 4, W + 4), -np.inf, dtype=np.float32)
        padded[2:2 + H, 2:2 + W] = score
        others = np.full((H, W), -np.inf, dtype=np.float32)
        for dy in range(5):
            for dx in range(5):
                if dy == 2 and dx == 2:
                    continue
                others = np.maximum(others, padded[dy:dy + H, dx:dx + W])
        interesting = score > threshold - threshold_margin
        near_tie = interesting & (np.abs(score - others) < score_margin * np.abs(score))
        if near_tie.any():
            return False, f"map {c}: near-tie between neighbouring box sums"
        is_max = score > others
        if (is_max & (np.abs(score - threshold) < threshold_margin)).any():
            return False, f"map {c}: local maximum on the threshold"
        yx, _ = oracle.find_peaks(heat[c], threshold)
        per_map = []
        for y, x in yx:
            xy, conf = oracle.centroid(heat[c], int(y), int(x))
            per_map.append((xy, conf))
        peaks.append(per_map)
    if len(peaks[0]) == 0:
        return True, "no centres"
    center_xy = np.array([p[0] for p in peaks[0]], dtype=np.float64)
    counts = np.zeros((len(center_xy), C), dtype=int)
    for c in range(1, C):
        for xy, conf in peaks[c]:
            if not (_away_from_half(float(xy[0]), rounding_margin) and
                    _away_from_half(float(xy[1]), rounding_margin)):
                return False, "spoke centroid rounds on a half pixel"
            xi = min(max(int(np.rint(xy[0])), 0), W - 1)
            yi = min(max(int(np.rint(xy[1])), 0), H - 1)
            vote = np.array([xi + 0.5 + float(centers[c - 1, 0, yi, xi]),
                             yi + 0.5 + float(centers[c - 1, 1, yi, xi])])
            dist = np.sort(np.linalg.norm(center_xy - vote[None], axis=1))
            if abs(dist[0] - outlier_distance) < distance_margin:
                return False, "vote on the outlier radius"
            if len(dist) > 1 and dist[1] - dist[0] < distance_margin:
                return False, "vote equidistant to two centres"
            if dist[0] <= outlier_distance:
                o = int(np.linalg.norm(center_xy - vote[None], axis=1).argmin())
                counts[o, c] += 1
    for c in range(1, C):
        if (counts[:, c] > cfg[c]).any():
            return False, "more detections than keypoint_config allows"
    # depth lookup rounds the UNDISTORTED pixel
    for c in range(C):
        for xy, _ in peaks[c]:
            ux, uy = oracle.undistort_point(float(xy[0]), float(xy[1]), cam)
            if not (_away_from_half(ux, rounding_margin) and _away_from_half(uy, rounding_margin)):
                return False, "undistorted pixel rounds on a half pixel"
    return True, "ok"
