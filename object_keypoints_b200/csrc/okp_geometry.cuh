// okp_geometry.cuh -- Kalibr pinhole-equidistant camera arithmetic in float64, per point in
// registers. Restates what the reference obtains from OpenCV:
//   okp_undistort_point   cv::fisheye::undistortPoints(xy, K, D, P=K)   camera_utils.py:75-81
//   okp_project_point     cv::fisheye::projectPoints                     camera_utils.py:65-73
//   okp_detection_to_point DetectionToPoint.__call__ + unproject          pipeline.py:164-171,
//                                                                          camera_utils.py:31-34
#pragma once
#include "okp_common.cuh"

__device__ __forceinline__ void okp_undistort_point(double u, double v, const OkpCamera& cam, double* ou, double* ov) {
    const double px = (u - cam.cx) / cam.fx;
    const double py = (v - cam.cy) / cam.fy;
    double theta_d = sqrt(px * px + py * py);
    const double half_pi = 3.14159265358979323846 / 2.0;
    theta_d = fmin(fmax(-half_pi, theta_d), half_pi);
    double theta = theta_d, scale = 0.0;
    bool converged = false;
    if (fabs(theta_d) > 1e-8) {
        for (int it = 0; it < 10; ++it) {           // Newton on theta (1 + k1 t^2 + ... + k4 t^8) = theta_d
            const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
            const double a = cam.k[0] * t2, b = cam.k[1] * t4, c = cam.k[2] * t6, d = cam.k[3] * t8;
            const double fix = (theta * (1 + a + b + c + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * c + 9 * d);
            theta = theta - fix;
            if (fabs(fix) < 1e-8) { converged = true; break; }
        }
        scale = tan(theta) / theta_d;
    } else {
        converged = true;
    }
    const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
    if (converged && !flipped) {
        *ou = cam.fx * (px * scale) + cam.cx;
        *ov = cam.fy * (py * scale) + cam.cy;
    } else {
        *ou = -1000000.0;                           // OpenCV's marker for "did not converge"
        *ov = -1000000.0;
    }
}

// T: row-major 3x4 (or the first 12 entries of a 4x4) world -> camera transform.
__device__ __forceinline__ void okp_project_point(const double* X, const double* T, const OkpCamera& cam,
                                                  double* ou, double* ov) {
    const double x = X[0], y = X[1], z = X[2];
    const double xc = T[0] * x + T[1] * y + T[2] * z + T[3];
    const double yc = T[4] * x + T[5] * y + T[6] * z + T[7];
    const double zc = T[8] * x + T[9] * y + T[10] * z + T[11];
    const double a = xc / zc, b = yc / zc;
    const double r = sqrt(a * a + b * b);
    const double th = atan(r), t2 = th * th;
    const double thd = th * (1.0 + cam.k[0] * t2 + cam.k[1] * t2 * t2 + cam.k[2] * t2 * t2 * t2 +
                             cam.k[3] * t2 * t2 * t2 * t2);
    const double s = r > 1e-8 ? thd / r : 1.0;
    *ou = cam.fx * (a * s) + cam.cx;
    *ov = cam.fy * (b * s) + cam.cy;
}

template <typename T>
__device__ __forceinline__ void okp_detection_to_point(float x, float y, const T* __restrict__ depth_map,
                                                       int H, int W, const OkpCamera& cam, int compat_clip_bug,
                                                       double* out) {
    double du, dv;
    okp_undistort_point((double)x, (double)y, cam, &du, &dv);
    const float ux = (float)du, uy = (float)dv;      // OpenCV returns float32 for float32 input
    int xi = __float2int_rn(ux), yi = __float2int_rn(uy);   // np.round: half to even
    if (compat_clip_bug) {                            // pipeline.py:162,169: (x, y) clipped with (H-1, W-1)
        xi = okp_clamp(xi, 0, cam.clip_x);
        yi = okp_clamp(yi, 0, cam.clip_y);
    }
    xi = okp_clamp(xi, 0, W - 1);                     // the reference would raise IndexError here
    yi = okp_clamp(yi, 0, H - 1);
    const double z = (double)okp_ld<T>(depth_map + (size_t)yi * W + xi);
    const double hx = (double)ux, hy = (double)uy;
#pragma unroll
    for (int r = 0; r < 3; ++r) out[r] = (cam.kinv[3 * r] * hx + cam.kinv[3 * r + 1] * hy + cam.kinv[3 * r + 2]) * z;
}

__global__ void okp_undistort_kernel(const double* __restrict__ xy, int n, OkpCamera cam, int round_to_f32,
                                     double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u, v;
    okp_undistort_point(xy[2 * i], xy[2 * i + 1], cam, &u, &v);
    if (round_to_f32) { u = (double)(float)u; v = (double)(float)v; }
    out[2 * i] = u;
    out[2 * i + 1] = v;
}

struct OkpPose { double m[12]; };

__global__ void okp_project_kernel(const double* __restrict__ X, int n, OkpPose T, OkpCamera cam,
                                   double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u, v;
    okp_project_point(X + 3 * (size_t)i, T.m, cam, &u, &v);
    out[2 * i] = u;
    out[2 * i + 1] = v;
}

__global__ void okp_detection_to_point_kernel(const float* __restrict__ xy, int n, const float* __restrict__ depth_map,
                                              int H, int W, OkpCamera cam, int compat_clip_bug,
                                              double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    okp_detection_to_point(xy[2 * i], xy[2 * i + 1], depth_map, H, W, cam, compat_clip_bug, out + 3 * (size_t)i);
}

// north_star "reprojection-error filtering": error of X[p] in view v against the distorted
// observation; views whose error exceeds max_error are dropped from the valid mask.
__global__ void okp_reprojection_filter_kernel(const double* __restrict__ X, const double* __restrict__ obs,
                                               uint8_t* __restrict__ valid, const double* __restrict__ poses,
                                               OkpCamera cam, int P, int V, double max_error,
                                               double* __restrict__ err) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)P * V) return;
    const int p = (int)(i / V), v = (int)(i - (long long)p * V);
    double uu, vv;
    okp_project_point(X + 3 * (size_t)p, poses + 16 * (size_t)v, cam, &uu, &vv);
    const double dx = uu - obs[2 * i], dy = vv - obs[2 * i + 1];
    const double e = sqrt(dx * dx + dy * dy);
    err[i] = e;
    if (!(e <= max_error)) valid[i] = 0;
}
