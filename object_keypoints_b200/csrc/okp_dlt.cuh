// okp_dlt.cuh -- K5: batched multi-view DLT triangulation, one thread per 3D point, float64,
// everything in registers.
//
// The reference triangulates two views with cv2.triangulatePoints (camera_utils.py:103-108,
// scripts/label.py:296-305): the homogeneous point is the right singular vector of the
// smallest singular value of A, whose rows are x P[2] - P[0] and y P[2] - P[1] per view. This
// kernel does the same for V >= 2 views without ever forming A^T A (which would square the
// condition number): the rows are streamed through Givens rotations into a 4x4 upper
// triangular R (A = Q R shares its right singular vectors with R), then a one-sided Jacobi SVD
// of R yields the vector.
#pragma once
#include "okp_common.cuh"
#include "okp_geometry.cuh"

__device__ __forceinline__ void okp_givens_append(double (&R)[4][4], double (&a)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (a[j] == 0.0) continue;
        const double r = hypot(R[j][j], a[j]);
        const double c = R[j][j] / r, s = a[j] / r;
#pragma unroll
        for (int k = j; k < 4; ++k) {
            const double rk = R[j][k], ak = a[k];
            R[j][k] = c * rk + s * ak;
            a[k] = c * ak - s * rk;
        }
    }
}

// Smallest right singular vector of the 4x4 matrix R (destroyed). One-sided (Hestenes) Jacobi.
__device__ __forceinline__ void okp_smallest_right_singular_vector(double (&R)[4][4], double (&h)[4]) {
    double Vm[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    // a column whose norm has fallen below 1e-15 of the matrix norm is rounding noise (exactly consistent observations:
    // the smallest singular value is 0); rotating against it changes nothing representable and the relative test below
    // would never be met, so such pairs count as converged (same statement in oracle/okp_oracle.c)
    double negligible = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) negligible += R[r][c] * R[r][c];
    negligible *= 1e-30;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    alpha += R[r][p] * R[r][p];
                    beta += R[r][q] * R[r][q];
                    gamma += R[r][p] * R[r][q];
                }
                if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 2.3e-16 * sqrt(alpha * beta)) continue;
                if (alpha <= negligible || beta <= negligible) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double ap = R[r][p], aq = R[r][q];
                    R[r][p] = cs * ap - sn * aq;
                    R[r][q] = sn * ap + cs * aq;
                    const double vp = Vm[r][p], vq = Vm[r][q];
                    Vm[r][p] = cs * vp - sn * vq;
                    Vm[r][q] = sn * vp + cs * vq;
                }
            }
        }
        if (!rotated) break;
    }
    int arg = 0;
    double smallest = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double nrm = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) nrm += R[r][c] * R[r][c];
        if (c == 0 || nrm < smallest) { smallest = nrm; arg = c; }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        // select column `arg` without dynamic register indexing
        h[r] = arg == 0 ? Vm[r][0] : (arg == 1 ? Vm[r][1] : (arg == 2 ? Vm[r][2] : Vm[r][3]));
    }
}

__global__ void __launch_bounds__(128)
okp_triangulate_kernel(const double* __restrict__ points, const uint8_t* __restrict__ valid,
                       const double* __restrict__ projections, int per_point, int P, int V,
                       double* __restrict__ out) {
    extern __shared__ double s_proj[];                 // [V][12] when the projections are shared
    if (!per_point) {
        for (int i = threadIdx.x; i < V * 12; i += blockDim.x) s_proj[i] = projections[i];
        __syncthreads();
    }
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double R[4][4] = {};
    int views = 0;
    for (int v = 0; v < V; ++v) {
        const size_t pv = (size_t)p * V + v;
        if (valid && !valid[pv]) continue;
        const double* M = per_point ? projections + pv * 12 : s_proj + v * 12;
        const double x = points[2 * pv], y = points[2 * pv + 1];
        double row[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) row[c] = x * M[8 + c] - M[c];
        okp_givens_append(R, row);
#pragma unroll
        for (int c = 0; c < 4; ++c) row[c] = y * M[8 + c] - M[4 + c];
        okp_givens_append(R, row);
        ++views;
    }
    if (views < 2) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        out[3 * (size_t)p] = nan; out[3 * (size_t)p + 1] = nan; out[3 * (size_t)p + 2] = nan;
        return;
    }
    double h[4];
    okp_smallest_right_singular_vector(R, h);
    out[3 * (size_t)p] = h[0] / h[3];
    out[3 * (size_t)p + 1] = h[1] / h[3];
    out[3 * (size_t)p + 2] = h[2] / h[3];
}

// ---------------------------------------------------------------------------------------------
// K5 + K6 fused: robust multi-view triangulation, one thread per 3D point.
//   repeat: V-view DLT over the valid views -> reprojection error of every view (distorted
//   pixels, okp_project_point) -> if the worst valid view is farther than max_error and more
//   than two views remain, drop it and solve again (at most max_rounds drops).
// The reference has no N-view code: the statement of record is
// oracle/np_oracle.py::triangulate_robust. The valid mask lives in one 64-bit register.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
okp_triangulate_robust_kernel(const double* __restrict__ obs, uint8_t* __restrict__ valid,
                              const double* __restrict__ poses, OkpCamera cam, int P, int V, double max_error,
                              int max_rounds, double* __restrict__ out, double* __restrict__ err,
                              int32_t* __restrict__ dropped, int group_points) {
    extern __shared__ double s_pose[];                 // [V][12] world -> camera, then [V][12] K * pose
    double* s_proj = s_pose + 12 * V;
    // group_points > 0: the points come in groups of that many, each with its own V poses (the 16 frames a track of a
    // moving camera was seen from); blockIdx.y is the group. 0: one pose set for every point.
    if (group_points > 0) poses += (size_t)blockIdx.y * V * 16;
    for (int i = threadIdx.x; i < V * 12; i += blockDim.x) s_pose[i] = poses[(i / 12) * 16 + (i % 12)];
    __syncthreads();
    for (int i = threadIdx.x; i < V * 12; i += blockDim.x) {          // camera_utils.py:125-130: K @ T[:3]
        const int v = i / 12, r = (i % 12) / 4, c = i % 4;
        const double* T = s_pose + 12 * v;
        s_proj[i] = r == 0 ? cam.fx * T[c] + cam.cx * T[8 + c] : (r == 1 ? cam.fy * T[4 + c] + cam.cy * T[8 + c] : T[8 + c]);
    }
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (group_points > 0) {
        if (p >= group_points) return;
        p += blockIdx.y * group_points;
    }
    if (p >= P) return;
    unsigned long long mask = 0;
    for (int v = 0; v < V; ++v)
        if (!valid || valid[(size_t)p * V + v]) mask |= 1ull << v;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double X[3] = {nan, nan, nan};
    int drops = 0;
    for (;;) {
        double R[4][4] = {};
        const int views = __popcll(mask);
        if (views < 2) { X[0] = X[1] = X[2] = nan; break; }
        for (int v = 0; v < V; ++v) {
            if (!((mask >> v) & 1ull)) continue;
            const size_t pv = (size_t)p * V + v;
            double x, y;
            okp_undistort_point(obs[2 * pv], obs[2 * pv + 1], cam, &x, &y);
            const double* M = s_proj + v * 12;
            double row[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) row[c] = x * M[8 + c] - M[c];
            okp_givens_append(R, row);
#pragma unroll
            for (int c = 0; c < 4; ++c) row[c] = y * M[8 + c] - M[4 + c];
            okp_givens_append(R, row);
        }
        double h[4];
        okp_smallest_right_singular_vector(R, h);
        X[0] = h[0] / h[3]; X[1] = h[1] / h[3]; X[2] = h[2] / h[3];
        int worst = -1;
        double worst_err = 0.0;
        for (int v = 0; v < V; ++v) {
            const size_t pv = (size_t)p * V + v;
            double uu, vv;
            okp_project_point(X, s_pose + 12 * v, cam, &uu, &vv);
            const double dx = uu - obs[2 * pv], dy = vv - obs[2 * pv + 1];
            const double e = sqrt(dx * dx + dy * dy);
            err[pv] = e;
            const double rank = e == e ? e : INFINITY;                 // a NaN error counts as the worst
            if (((mask >> v) & 1ull) && (worst < 0 || rank > worst_err)) { worst = v; worst_err = rank; }   // first maximum
        }
        if (!(worst_err > max_error) || views <= 2 || drops >= max_rounds) break;
        mask &= ~(1ull << worst);
        ++drops;
    }
    if (valid)
        for (int v = 0; v < V; ++v) valid[(size_t)p * V + v] = (uint8_t)((mask >> v) & 1ull);
    out[3 * (size_t)p] = X[0]; out[3 * (size_t)p + 1] = X[1]; out[3 * (size_t)p + 2] = X[2];
    if (dropped) dropped[p] = drops;
}
