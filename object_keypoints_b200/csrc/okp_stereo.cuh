// okp_stereo.cuh -- two-view stereo helpers, one thread per point pair, float64 in registers.
//
//   okp_correct_matches_kernel   cv2.correctMatches as StereoCamera.triangulate calls it
//                                (perception/utils/camera_utils.py:100-101): the Hartley-Sturm
//                                optimal correction (Hartley & Zisserman, Alg. 12.1). Moves a
//                                pair (x, x') by the smallest total squared distance onto a pair
//                                that satisfies x'^T F x = 0 exactly.
//   okp_associate_kernel         stereo association by epipolar distance, one-to-one, -1 for
//                                unmatched; the implementation is gone from the reference, its
//                                expectations survive in test/test_pipeline.py:208-261.
//
// The correction is geometric (a unique global minimum of a rational cost over one real
// parameter t), so any exact minimiser reproduces OpenCV's output. OpenCV finds the stationary
// points as the six roots of a degree-6 polynomial g(t) with a Durand-Kerner iteration; here the
// roots come from an Aberth-Ehrlich iteration started on the Newton polygon of the coefficients
// (so that roots of very different magnitude -- the epipole of a sideways stereo rig is ~1e6 px
// away -- all start at the right scale), the cost is evaluated at the real part of every root and
// at t = infinity like OpenCV does, and the winner is polished by Newton steps on g.
#pragma once
#include "okp_common.cuh"

struct OkpMat3 { double m[9]; };

// Unit right singular vector of the smallest singular value of the 3x3 matrix A (destroyed):
// one-sided (Hestenes) Jacobi, the same scheme as the 4x4 solve of okp_dlt.cuh.
__host__ __device__ __forceinline__ void okp_null_vector3(double (&A)[3][3], double (&h)[3]) {
    double Vm[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    alpha += A[r][p] * A[r][p];
                    beta += A[r][q] * A[r][q];
                    gamma += A[r][p] * A[r][q];
                }
                if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 2.3e-16 * sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double ap = A[r][p], aq = A[r][q];
                    A[r][p] = cs * ap - sn * aq;
                    A[r][q] = sn * ap + cs * aq;
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
    for (int c = 0; c < 3; ++c) {
        double nrm = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) nrm += A[r][c] * A[r][c];
        if (c == 0 || nrm < smallest) { smallest = nrm; arg = c; }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) h[r] = arg == 0 ? Vm[r][0] : (arg == 1 ? Vm[r][1] : Vm[r][2]);
}

struct OkpComplex { double re, im; };
__host__ __device__ __forceinline__ OkpComplex okp_cmul(OkpComplex a, OkpComplex b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ __forceinline__ OkpComplex okp_cdiv(OkpComplex a, OkpComplex b) {      // Smith's algorithm
    // branch-free (selects): the two cases of the textbook form differ by which component of b is the pivot; with
    // p = pivot, q = the other one, x / y = the matching components of a, both cases are ((x + y r) / den, +-(y - x r) / den)
    // -- the same operations on the same operands as the two-branch form, so the quotient is bit-identical, and the lanes
    // of a warp no longer split on |b.re| >= |b.im|
    const bool big = fabs(b.re) >= fabs(b.im);
    const double p = big ? b.re : b.im, q = big ? b.im : b.re;
    const double x = big ? a.re : a.im, y = big ? a.im : a.re;
    const double r = q / p, den = p + q * r;
    const double im = (y - x * r) / den;
    return {(x + y * r) / den, big ? im : -im};
}

// All complex roots of k[0] + k[1] t + ... + k[6] t^6 (degree n <= 6 after leading zeros are
// dropped). Returns n. Aberth-Ehrlich, Gauss-Seidel order, start values on the Newton polygon
// (Bini 1996): for every edge (i, j) of the upper convex hull of (i, log|k_i|), j - i roots on a
// circle of radius |k_i / k_j|^(1 / (j - i)).
__host__ __device__ __noinline__ int okp_poly6_roots(const double* k, OkpComplex* z) {
    int n = 6;
    while (n > 0 && k[n] == 0.0) --n;
    if (n == 0) return 0;
    double lg[7];
    for (int i = 0; i <= n; ++i) lg[i] = k[i] != 0.0 ? log(fabs(k[i])) : -1e300;
    int placed = 0, i = 0;
    while (i < n) {
        int best = i + 1;
        double slope = (lg[i + 1] - lg[i]);
        for (int j = i + 2; j <= n; ++j) {
            const double s = (lg[j] - lg[i]) / (double)(j - i);
            if (s >= slope) { slope = s; best = j; }
        }
        const int m = best - i;
        double radius = lg[i] <= -1e299 ? 0.0 : exp(-slope);
        if (!(radius < 1e150)) radius = 1e150;
        for (int r = 0; r < m; ++r) {
            const double ang = 6.283185307179586 * ((double)r / m + (double)placed / n) + 0.7;
            double sn, cs;
            sincos(ang, &sn, &cs);
            z[placed++] = {radius * cs, radius * sn};
        }
        i = best;
    }
    // Convergence: the largest relative correction of a sweep below 1e-14 -- or no longer contracting once it is small.
    // Aberth-Ehrlich converges cubically, so a sweep that is within 1e-9 of the roots and does not shrink the correction
    // fourfold has reached the rounding floor of an ill-conditioned root (the far roots of a sideways rig sit at |t| ~ 1e6
    // and stall at ~1e-15 .. 1e-13). Round 1 asked every root for 4e-16: one pair in five never got there and ran to the
    // cap of 200 sweeps, and with it the other 31 pairs of its warp (ncu r02m: 4 of 32 lanes active, 19.9 ms per 2^20
    // pairs). The winner is polished by Newton steps on g below, so the result does not move (<= 1e-12 px).
    double previous = 1e300;
    for (int it = 0; it < 200; ++it) {
        double worst = 0.0;
        for (int r = 0; r < n; ++r) {
            const OkpComplex x = z[r];
            OkpComplex p = {k[n], 0.0}, dp = {0.0, 0.0};
            for (int c = n - 1; c >= 0; --c) {
                dp = okp_cmul(dp, x);
                dp.re += p.re; dp.im += p.im;
                p = okp_cmul(p, x);
                p.re += k[c];
            }
            if (p.re == 0.0 && p.im == 0.0) continue;
            OkpComplex w;
            if (dp.re == 0.0 && dp.im == 0.0) {
                w = {1e-3 * (fabs(x.re) + 1e-3), 1e-3 * (fabs(x.im) + 1e-3)};     // stationary point: nudge off it
            } else {
                w = okp_cdiv(p, dp);
                OkpComplex s = {0.0, 0.0};
                for (int j = 0; j < n; ++j) {
                    if (j == r) continue;
                    const OkpComplex d = {x.re - z[j].re, x.im - z[j].im};
                    if (d.re == 0.0 && d.im == 0.0) continue;
                    const OkpComplex inv = okp_cdiv({1.0, 0.0}, d);
                    s.re += inv.re; s.im += inv.im;
                }
                const OkpComplex ws = okp_cmul(w, s);
                const OkpComplex den = {1.0 - ws.re, -ws.im};
                if (den.re != 0.0 || den.im != 0.0) w = okp_cdiv(w, den);
            }
            z[r] = {x.re - w.re, x.im - w.im};
            const double rel = (fabs(w.re) + fabs(w.im)) / (fabs(x.re) + fabs(x.im) + 1e-300);
            if (rel > worst) worst = rel;
        }
        if (worst < 1e-14 || (worst < 1e-9 && worst > 0.25 * previous)) break;
        previous = worst;
    }
    return n;
}

// The epipoles of F itself, once per call (host): right null vector (F e = 0) and left null vector (e^T F = 0).
struct OkpEpipoles { double right[3], left[3]; };
__host__ __device__ __forceinline__ OkpEpipoles okp_epipoles(const OkpMat3& Fm) {
    double B[3][3], Bt[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { B[r][c] = Fm.m[3 * r + c]; Bt[r][c] = Fm.m[3 * c + r]; }
    OkpEpipoles e;
    okp_null_vector3(B, e.right);
    okp_null_vector3(Bt, e.left);
    return e;
}

// One pair: Hartley-Sturm correction (__host__ too: tools/host_check_stereo.cu runs it on the CPU
// of the build container, which has no GPU). F row-major with x2^T F x1 = 0.
__host__ __device__ __forceinline__ void okp_correct_pair(const OkpMat3& Fm, const OkpEpipoles& ep, double x1, double y1,
                                                          double x2, double y2, double* o1, double* o2) {
    const double* F = Fm.m;
    // F~ = T2^T F T1 with T = [[1,0,x],[0,1,y],[0,0,1]]: both points move to the origin
    double A[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        A[0][c] = F[c];
        A[1][c] = F[3 + c];
        A[2][c] = x2 * F[c] + y2 * F[3 + c] + F[6 + c];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) A[r][2] = A[r][0] * x1 + A[r][1] * y1 + A[r][2];
    // epipoles: F~ e1 = 0, e2^T F~ = 0, scaled so that e_x^2 + e_y^2 = 1. F~ = T2^T F T1, so F (T1 e1) = 0 and
    // (T2 e2)^T F = 0: e = T^-1 (epipole of F) -- a translation of F's own epipoles (round 1 ran two 3x3 Jacobi SVDs of F~
    // per pair for them)
    const double e1[3] = {ep.right[0] - x1 * ep.right[2], ep.right[1] - y1 * ep.right[2], ep.right[2]};
    const double e2[3] = {ep.left[0] - x2 * ep.left[2], ep.left[1] - y2 * ep.left[2], ep.left[2]};
    const double s1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1]);
    const double s2 = sqrt(e2[0] * e2[0] + e2[1] * e2[1]);
    const double e1x = e1[0] / s1, e1y = e1[1] / s1, e1z = e1[2] / s1;
    const double e2x = e2[0] / s2, e2y = e2[1] / s2, e2z = e2[2] / s2;
    // F' = R2 F~ R1^T with R = [[ex, ey, 0], [-ey, ex, 0], [0, 0, 1]]: epipoles onto the x axes
    double G[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        G[0][c] = e2x * A[0][c] + e2y * A[1][c];
        G[1][c] = -e2y * A[0][c] + e2x * A[1][c];
        G[2][c] = A[2][c];
    }
    double Fp[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        Fp[r][0] = G[r][0] * e1x + G[r][1] * e1y;
        Fp[r][1] = -G[r][0] * e1y + G[r][1] * e1x;
        Fp[r][2] = G[r][2];
    }
    const double f1 = e1z, f2 = e2z, a = Fp[1][1], b = Fp[1][2], c = Fp[2][1], d = Fp[2][2];
    const double f1_2 = f1 * f1, f1_4 = f1_2 * f1_2, f2_2 = f2 * f2, f2_4 = f2_2 * f2_2;
    const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d;
    // g(t) = t ((a t + b)^2 + f2^2 (c t + d)^2)^2 - (a d - b c)(1 + f1^2 t^2)^2 (a t + b)(c t + d), expanded
    double k[7];
    k[6] = b * c2 * f1_4 * a - a2 * d * f1_4 * c;
    k[5] = f2_4 * c2 * c2 + 2 * a2 * f2_2 * c2 - a2 * d2 * f1_4 + b2 * c2 * f1_4 + a2 * a2;
    k[4] = 4 * a2 * a * b + 2 * b * c2 * f1_2 * a + 4 * f2_4 * c2 * c * d + 4 * a * b * f2_2 * c2 +
           4 * a2 * f2_2 * c * d - 2 * a2 * d * f1_2 * c - a * d2 * f1_4 * b + b2 * c * f1_4 * d;
    k[3] = 6 * a2 * b2 + 6 * f2_4 * c2 * d2 + 2 * b2 * f2_2 * c2 + 2 * a2 * f2_2 * d2 - 2 * a2 * d2 * f1_2 +
           2 * b2 * c2 * f1_2 + 8 * a * b * f2_2 * c * d;
    k[2] = 4 * a * b2 * b + 4 * b2 * f2_2 * c * d + 4 * f2_4 * c * d2 * d - a2 * d * c + b * c2 * a +
           4 * a * b * f2_2 * d2 - 2 * a * d2 * f1_2 * b + 2 * b2 * c * f1_2 * d;
    k[1] = f2_4 * d2 * d2 + b2 * b2 + 2 * b2 * f2_2 * d2 - a2 * d2 + b2 * c2;
    k[0] = -a * d2 * b + b2 * c * d;
    double big = 0.0;
#pragma unroll
    for (int i = 0; i < 7; ++i) big = fmax(big, fabs(k[i]));
    if (big > 0.0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) k[i] /= big;
    }
    OkpComplex z[6];
    const int n = okp_poly6_roots(k, z);
    // cost at t = infinity, then at the real part of every root (OpenCV's selection rule)
    double s_best = 1.0 / f1_2 + c2 / (a2 + f2_2 * c2);
    double t_best = 0.0;
    bool at_infinity = true;
    if (!(s_best == s_best)) s_best = INFINITY;
    for (int i = 0; i < n; ++i) {
        const double t = z[i].re;
        const double u = c * t + d, v = a * t + b;
        const double s = (t * t) / (1.0 + f1_2 * t * t) + (u * u) / (v * v + f2_2 * u * u);
        if (s < s_best) { s_best = s; t_best = t; at_infinity = false; }
    }
    double l1[3], l2[3];
    if (!at_infinity) {
        double t = t_best;
        for (int it = 0; it < 4; ++it) {              // polish: Newton on g, kept only while |g| shrinks
            double g = k[6], dg = 0.0;
            for (int i = 5; i >= 0; --i) { dg = dg * t + g; g = g * t + k[i]; }
            if (dg == 0.0 || g == 0.0) break;
            const double tn = t - g / dg;
            double gn = k[6];
            for (int i = 5; i >= 0; --i) gn = gn * tn + k[i];
            if (!(fabs(gn) < fabs(g))) break;
            t = tn;
        }
        const double u = c * t + d, v = a * t + b;
        // closest points to the origin on l1 = (t f1, 1, -t) and l2 = (-f2 u, v, u)
        l1[0] = t * t * f1; l1[1] = t; l1[2] = t * t * f1_2 + 1.0;
        l2[0] = f2 * u * u; l2[1] = -v * u; l2[2] = f2_2 * u * u + v * v;
    } else {                                           // t -> infinity: the limits of the expressions above
        l1[0] = f1; l1[1] = 0.0; l1[2] = f1_2;
        l2[0] = f2 * c2; l2[1] = -a * c; l2[2] = f2_2 * c2 + a2;
    }
    const double p1x = l1[0] / l1[2], p1y = l1[1] / l1[2];
    const double p2x = l2[0] / l2[2], p2y = l2[1] / l2[2];
    // back through R^T and the translations
    o1[0] = e1x * p1x - e1y * p1y + x1;
    o1[1] = e1y * p1x + e1x * p1y + y1;
    o2[0] = e2x * p2x - e2y * p2y + x2;
    o2[1] = e2y * p2x + e2x * p2y + y2;
}

__global__ void __launch_bounds__(64)
okp_correct_matches_kernel(OkpMat3 F, OkpEpipoles ep, const double* __restrict__ left, const double* __restrict__ right, int n,
                           int round_to_f32, double* __restrict__ left_out, double* __restrict__ right_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o1[2], o2[2];
    okp_correct_pair(F, ep, left[2 * i], left[2 * i + 1], right[2 * i], right[2 * i + 1], o1, o2);
    if (round_to_f32) {                               // OpenCV returns the input dtype (float32 in camera_utils.py:93-101)
        o1[0] = (double)(float)o1[0]; o1[1] = (double)(float)o1[1];
        o2[0] = (double)(float)o2[0]; o2[1] = (double)(float)o2[1];
    }
    left_out[2 * i] = o1[0]; left_out[2 * i + 1] = o1[1];
    right_out[2 * i] = o2[0]; right_out[2 * i + 1] = o2[1];
}

// ---------------------------------------------------------------------------------------------
// Association. One warp per frame pair. cost(i, j) = mean of the two point-to-epipolar-line
// distances (pixels) of UNDISTORTED points, x_R^T F x_L = 0. Greedy one-to-one: repeatedly take
// the globally smallest remaining cost (ties: smallest left index, then smallest right index)
// until it exceeds max_distance. Statement of record: oracle/np_oracle.py::associate.
// ---------------------------------------------------------------------------------------------
#define OKP_ASSOC_MAX 64

__global__ void __launch_bounds__(32)
okp_associate_kernel(OkpMat3 Fm, const double* __restrict__ F_pairs, const double* __restrict__ left,
                     const int32_t* __restrict__ n_left, const double* __restrict__ right,
                     const int32_t* __restrict__ n_right, int ML, int MR, double max_distance,
                     int32_t* __restrict__ match, double* __restrict__ match_cost) {
    extern __shared__ double s_cost[];               // [ML][MR]
    __shared__ unsigned long long s_used_left, s_used_right;
    const int b = blockIdx.x, lane = threadIdx.x;
    const int nl = okp_clamp(n_left[b], 0, ML), nr = okp_clamp(n_right[b], 0, MR);
    // one F for every pair (a stereo rig) or one per pair (two frames of a moving camera: scripts/label.py:285-305)
    const double* F = F_pairs ? F_pairs + 9 * (size_t)b : Fm.m;
    const double* L = left + (size_t)b * ML * 2;
    const double* R = right + (size_t)b * MR * 2;
    for (int e = lane; e < nl * nr; e += 32) {
        const int i = e / nr, j = e - i * nr;
        const double x = L[2 * i], y = L[2 * i + 1], xp = R[2 * j], yp = R[2 * j + 1];
        const double l0 = F[0] * x + F[1] * y + F[2], l1 = F[3] * x + F[4] * y + F[5], l2 = F[6] * x + F[7] * y + F[8];
        const double m0 = F[0] * xp + F[3] * yp + F[6], m1 = F[1] * xp + F[4] * yp + F[7];
        const double r = fabs(xp * l0 + yp * l1 + l2);
        s_cost[i * MR + j] = 0.5 * (r / sqrt(l0 * l0 + l1 * l1) + r / sqrt(m0 * m0 + m1 * m1));
    }
    for (int i = lane; i < ML; i += 32) {
        match[(size_t)b * ML + i] = -1;
        match_cost[(size_t)b * ML + i] = 0.0;
    }
    if (lane == 0) { s_used_left = 0ull; s_used_right = 0ull; }
    __syncwarp();
    const int rounds = nl < nr ? nl : nr;
    for (int round = 0; round < rounds; ++round) {
        const unsigned long long ul = s_used_left, ur = s_used_right;
        double best = INFINITY;
        int arg = 0x7fffffff;
        for (int e = lane; e < nl * nr; e += 32) {
            const int i = e / nr, j = e - i * nr;
            if (((ul >> i) & 1ull) || ((ur >> j) & 1ull)) continue;
            const double c = s_cost[i * MR + j];
            if (c < best || (c == best && e < arg)) { best = c; arg = e; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
            if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (!(best <= max_distance)) break;
        const int i = arg / nr, j = arg - i * nr;
        if (lane == 0) {
            match[(size_t)b * ML + i] = j;
            match_cost[(size_t)b * ML + i] = best;
            s_used_left = ul | (1ull << i);
            s_used_right = ur | (1ull << j);
        }
        __syncwarp();
    }
}
