// okp_targets.cuh -- ground-truth / training targets on the device.
//
// Replaces, for a batch of frames, what the reference's dataset builds per frame on the CPU
// (perception/datasets/video.py): the Gaussian keypoint heatmaps (_set_keypoints :44-53 with the
// per-map normalisation of _extract_example :210-211), the centre-vector maps (_compute_centers
// :225-242) and the depth maps (_compute_depth :244-263). It makes the synthetic inputs of the
// benchmarks and tests GPU-resident instead of a host loop.
//
// One CTA per (frame, map). A pixel's value depends on the map's keypoints IN ORDER (the heat sum is
// accumulated in float32 keypoint by keypoint, later discs overwrite earlier ones), so every thread walks
// the frame's keypoint list in the reference's order for each of its pixels; the window / disc tests are
// integer or one float64 compare, the Gaussian is evaluated only inside its (2k+1)^2 window.
#pragma once
#include "okp_common.cuh"
#include "okp_group.cuh"      // OkpConfig

struct OkpTargetParams {
    int32_t kernel_size;      // 8  = int(heatmap_size / 8)        video.py:19
    double length_scale;      // 2  = heatmap_size / 32            video.py:20
    double center_radius;     // 4  = heatmap_size / 16            video.py:18
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
okp_rasterise_targets_kernel(const double* __restrict__ keypoints, const double* __restrict__ depths,
                             const int32_t* __restrict__ n_objects, int N, int G, int Kp, int C, int H, int W,
                             OkpConfig config, OkpTargetParams prm, float* __restrict__ heat,
                             float* __restrict__ centers, float* __restrict__ depth) {
    const int n = blockIdx.x / C, c = blockIdx.x - n * C;
    if (n >= N) return;
    extern __shared__ __align__(16) unsigned char target_smem[];
    double* s_xy = reinterpret_cast<double*>(target_smem);            // [G][Kp][2] this frame's keypoints
    double* s_z = s_xy + (size_t)G * Kp * 2;                          // [G][Kp]
    __shared__ float s_max[THREADS / 32];
    const int objects = okp_min(n_objects ? n_objects[n] : G, G);
    for (int i = threadIdx.x; i < G * Kp; i += THREADS) {
        s_xy[2 * i] = keypoints[((size_t)n * G * Kp + i) * 2];
        s_xy[2 * i + 1] = keypoints[((size_t)n * G * Kp + i) * 2 + 1];
        s_z[i] = depths[(size_t)n * G * Kp + i];
    }
    __syncthreads();
    int first = 0;                                                    // this map's keypoints inside an object: [first, first + count)
    for (int i = 0; i < c; ++i) first += config.cfg[i];
    const int count = config.cfg[c];
    const double inv_scale2 = prm.length_scale * prm.length_scale;
    const double radius2 = prm.center_radius * prm.center_radius * (1.0 + 1e-12);   // loose superset of the disc
    const size_t HW = (size_t)H * W;
    float* heat_map = heat + ((size_t)n * C + c) * HW;
    float* depth_map = depth + ((size_t)n * C + c) * HW;
    float* center_map = c > 0 ? centers + ((size_t)n * (C - 1) + (c - 1)) * 2 * HW : nullptr;

    float local_max = 0.0f;
    for (int p = threadIdx.x; p < H * W; p += THREADS) {
        const int i = p / W, j = p - i * W;
        // pixel centre as the reference holds it: float32 (j + 0.5, i + 0.5) (_pixel_indices, video.py:37-42)
        const double pxc = (double)((float)j + 0.5f), pyc = (double)((float)i + 0.5f);
        float h = 0.0f, z = 0.0f, vx = 0.0f, vy = 0.0f;
        for (int g = 0; g < objects; ++g) {
            const double* obj = s_xy + (size_t)g * Kp * 2;
            for (int k = first; k < first + count; ++k) {
                const double x = obj[2 * k], y = obj[2 * k + 1];
                // _set_keypoints: window around the truncated position, clipped to the map
                const int ix = (int)x, iy = (int)y;
                if (j >= ix - prm.kernel_size && j <= ix + prm.kernel_size && i >= iy - prm.kernel_size && i <= iy + prm.kernel_size) {
                    const double dx = x - (double)j, dy = y - (double)i;
                    h = (float)((double)h + exp(-(dx * dx + dy * dy) / inv_scale2));
                }
                // _compute_depth / _compute_centers: discs around the keypoint, later keypoints overwrite
                // sqrt only inside a slightly enlarged disc (d2 >= r^2 (1 + 1e-12) implies sqrt(d2) >= r); the reference's own
                // comparison decides the rest
                const double ddx = x - pxc, ddy = y - pyc, d2 = ddx * ddx + ddy * ddy;
                if (d2 < radius2 && sqrt(d2) < prm.center_radius) {
                    z = (float)s_z[g * Kp + k];
                    if (c > 0) { vx = (float)(obj[0] - pxc); vy = (float)(obj[1] - pyc); }
                }
            }
        }
        heat_map[p] = h;
        depth_map[p] = z;
        if (c > 0) { center_map[p] = vx; center_map[HW + p] = vy; }
        local_max = fmaxf(local_max, h);
    }
    // per-map normalisation: target / max(target.max(), 0.5), clipped to [0, 1] (video.py:210-211)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = local_max;
    __syncthreads();
    float peak = 0.5f;
    for (int w = 0; w < THREADS / 32; ++w) peak = fmaxf(peak, s_max[w]);
    for (int p = threadIdx.x; p < H * W; p += THREADS)                // every thread re-reads its own stores
        heat_map[p] = fminf(fmaxf(__fdiv_rn(heat_map[p], peak), 0.0f), 1.0f);
}
