// okp_eval.cuh -- evaluation bookkeeping on the device.
//
// Replaces Results.add / Results.print_results of the reference's scripts/eval_model.py:137-232: every
// predicted 3D keypoint of every frame is matched to the scene's ground truth and scored, so that a whole
// test sequence is evaluated from the decode tables without leaving the GPU.
//
// One CTA per frame. The scene (G objects x Kp points, world frame) is moved into the frame's camera
// frame once and kept in shared memory; one thread per predicted object finds its ground-truth object,
// one thread per keypoint slot does the point matching; thread 0 folds the frame's errors into
// (count, mean, M2) so that the final reduction (okp_eval_summary_kernel, Chan's pairwise merge in a
// fixed order) is deterministic and as accurate as NumPy's two-pass std.
#pragma once
#include "okp_common.cuh"
#include "okp_geometry.cuh"

#define OKP_EVAL_EMPTY (-1)
#define OKP_EVAL_MATCHED 0
#define OKP_EVAL_MISSING 1
#define OKP_EVAL_POINT_NOT_IN_VIEW 2
#define OKP_EVAL_OBJECT_NOT_IN_VIEW 3
#define OKP_EVAL_STATS 8            // doubles per frame: matched, missing, small, mean, M2, sum_xy, 0, 0

struct OkpEvalDims { int N, O, C, S, G, Kp; };

// PinholeCamera.in_frame (camera_utils.py:36-43) for one pixel: limits are image_size AS GIVEN, (H, W)
// against (x, y) -- x is compared with the height, exactly like the reference.
__device__ __forceinline__ bool okp_in_frame(double u, double v, double limit_x, double limit_y) {
    return !(u <= 0.0 || v <= 0.0 || u >= limit_x || v >= limit_y);
}

__device__ __forceinline__ bool okp_point_in_view(const double* p, const OkpCamera& cam, double limit_x, double limit_y) {
    const double identity[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    double u, v;
    okp_project_point(p, identity, cam, &u, &v);
    return okp_in_frame(u, v, limit_x, limit_y);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
okp_eval_match_kernel(const double* __restrict__ kp_point, const int32_t* __restrict__ kp_count,
                      const int32_t* __restrict__ n_objects, const double* __restrict__ T_WC,
                      const double* __restrict__ scene, OkpEvalDims d, OkpCamera cam, double limit_x, double limit_y,
                      double max_coordinate, double small_error, int32_t* __restrict__ status,
                      double* __restrict__ gt_point, double* __restrict__ err, double* __restrict__ err_xy,
                      int32_t* __restrict__ gt_object, double* __restrict__ frame_stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* scene_C = reinterpret_cast<double*>(smem_raw);                    // [G][Kp][3]
    int* s_object = reinterpret_cast<int*>(scene_C + (size_t)d.G * d.Kp * 3);  // [O] ground-truth object or -1
    int* s_visible = s_object + d.O;                                          // [O]
    const int n = blockIdx.x;
    if (n >= d.N) return;
    const int slots = d.O * d.C * d.S;
    const size_t base = (size_t)n * slots;

    // T_CW = [R^T | -R^T t] (linalg.py:9-13), scene into the camera frame (linalg.py:15-20)
    const double* T = T_WC + (size_t)n * 16;
    double R[9], t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) R[3 * r + c] = T[4 * c + r];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = -(R[3 * r] * T[3] + R[3 * r + 1] * T[7] + R[3 * r + 2] * T[11]);
    for (int i = threadIdx.x; i < d.G * d.Kp; i += THREADS) {
        const double x = scene[3 * i], y = scene[3 * i + 1], z = scene[3 * i + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) scene_C[3 * i + r] = (R[3 * r] * x + R[3 * r + 1] * y + R[3 * r + 2] * z) + t[r];
    }
    __syncthreads();

    const int objects = okp_min(n_objects[n], d.O);
    for (int o = threadIdx.x; o < d.O; o += THREADS) {
        int g_best = -1, visible = 0;
        if (o < objects) {
            // nearest ground-truth centre in camera-frame XY, depth ignored (eval_model.py:153-155)
            const double cx = kp_point[(base + (size_t)o * d.C * d.S) * 3], cy = kp_point[(base + (size_t)o * d.C * d.S) * 3 + 1];
            double best = 0.0;
            for (int g = 0; g < d.G; ++g) {
                const double dx = scene_C[(size_t)g * d.Kp * 3] - cx, dy = scene_C[(size_t)g * d.Kp * 3 + 1] - cy;
                const double dist = sqrt(dx * dx + dy * dy);
                if (g == 0 || dist < best) { best = dist; g_best = g; }        // first minimum, like np.argmin
            }
            visible = okp_point_in_view(scene_C + (size_t)g_best * d.Kp * 3, cam, limit_x, limit_y) ? 1 : 0;   // :158-163
        }
        s_object[o] = g_best;
        s_visible[o] = visible;
        gt_object[(size_t)n * d.O + o] = g_best;
    }
    __syncthreads();

    for (int i = threadIdx.x; i < slots; i += THREADS) {
        const int o = i / (d.C * d.S), cs = i - o * d.C * d.S, c = cs / d.S, s = cs - c * d.S;
        int code = OKP_EVAL_EMPTY;
        double gx = 0.0, gy = 0.0, gz = 0.0, e = 0.0, exy = 0.0;
        if (o < objects && s < kp_count[((size_t)n * d.O + o) * d.C + c]) {
            const double px = kp_point[(base + i) * 3], py = kp_point[(base + i) * 3 + 1], pz = kp_point[(base + i) * 3 + 2];
            if (!s_visible[o]) {
                code = OKP_EVAL_OBJECT_NOT_IN_VIEW;
            } else if (!(px < max_coordinate && py < max_coordinate && pz < max_coordinate)) {
                code = OKP_EVAL_MISSING;                                        // :171, :183-185
            } else {
                const double* object_points = scene_C + (size_t)s_object[o] * d.Kp * 3;
                int k_best = 0;
                double best = 0.0;
                for (int k = 0; k < d.Kp; ++k) {                                // :172
                    const double dx = object_points[3 * k] - px, dy = object_points[3 * k + 1] - py, dz = object_points[3 * k + 2] - pz;
                    const double dist = sqrt((dx * dx + dy * dy) + dz * dz);
                    if (k == 0 || dist < best) { best = dist; k_best = k; }
                }
                const double* gt = object_points + 3 * k_best;
                if (!okp_point_in_view(gt, cam, limit_x, limit_y)) {
                    code = OKP_EVAL_POINT_NOT_IN_VIEW;                          // :174-178
                } else {
                    code = OKP_EVAL_MATCHED;
                    gx = gt[0]; gy = gt[1]; gz = gt[2];
                    const double dx = gx - px, dy = gy - py, dz = gz - pz;
                    e = sqrt((dx * dx + dy * dy) + dz * dz);                    // :206
                    exy = sqrt(dx * dx + dy * dy);                              // :207
                }
            }
        }
        status[base + i] = code;
        gt_point[(base + i) * 3] = gx; gt_point[(base + i) * 3 + 1] = gy; gt_point[(base + i) * 3 + 2] = gz;
        err[base + i] = e;
        err_xy[base + i] = exy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {          // the frame's (count, mean, M2) in slot order: Welford, deterministic
        double matched = 0.0, missing = 0.0, small = 0.0, mean = 0.0, m2 = 0.0, sum_xy = 0.0;
        for (int i = 0; i < slots; ++i) {
            const int code = status[base + i];
            if (code == OKP_EVAL_MISSING) missing += 1.0;
            if (code != OKP_EVAL_MATCHED) continue;
            const double e = err[base + i];
            matched += 1.0;
            const double delta = e - mean;
            mean += delta / matched;
            m2 += delta * (e - mean);
            sum_xy += err_xy[base + i];
            if (e < small_error) small += 1.0;                                  // :210 (< 3 cm)
        }
        double* out = frame_stats + (size_t)n * OKP_EVAL_STATS;
        out[0] = matched; out[1] = missing; out[2] = small; out[3] = mean; out[4] = m2; out[5] = sum_xy; out[6] = 0.0; out[7] = 0.0;
    }
}

// (count, mean, M2) of two disjoint sample sets (Chan et al.); the other fields add.
__device__ __forceinline__ void okp_eval_merge(double* a, const double* b) {
    const double na = a[0], nb = b[0], total = na + nb;
    if (nb > 0.0) {
        const double delta = b[3] - a[3];
        a[3] = total > 0.0 ? a[3] + delta * (nb / total) : 0.0;
        a[4] = a[4] + b[4] + delta * delta * (na * nb / total);
    }
    a[0] = total; a[1] += b[1]; a[2] += b[2]; a[5] += b[5];
}

// totals[8] = merge of all frames: matched, missing, small, mean error, M2, sum of xy errors (metres).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
okp_eval_summary_kernel(const double* __restrict__ frame_stats, int N, double* __restrict__ totals) {
    __shared__ double part[THREADS][OKP_EVAL_STATS];
    double acc[OKP_EVAL_STATS] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int n = threadIdx.x; n < N; n += THREADS) okp_eval_merge(acc, frame_stats + (size_t)n * OKP_EVAL_STATS);
#pragma unroll
    for (int i = 0; i < OKP_EVAL_STATS; ++i) part[threadIdx.x][i] = acc[i];
    __syncthreads();
    for (int stride = THREADS / 2; stride > 0; stride >>= 1) {
        if (threadIdx.x < stride) okp_eval_merge(part[threadIdx.x], part[threadIdx.x + stride]);
        __syncthreads();
    }
    if (threadIdx.x < OKP_EVAL_STATS) totals[threadIdx.x] = part[0][threadIdx.x];
}
