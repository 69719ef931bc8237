/* TEST INFRASTRUCTURE -- plain C restatement of the reference's heatmap -> 3D keypoint path.
 *
 * This file is the CHECKER and the CPU baseline, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs load liboracle.so. It is pinned
 * against outputs of the unmodified reference (tests/golden/, written by oracle/make_goldens.py)
 * by tests/test_oracle.py. Paths cited below are relative to the reference root.
 *
 * Build: make -C oracle   (gcc -O3 -ffp-contract=off -fopenmp; no FMA contraction so that
 * float32 results are the separately rounded ones the contract in np_oracle.py states).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/okp.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------------------
 * A7: cv::fisheye::undistortPoints(xy, K, D, P=K)  -- camera_utils.py:75-81
 * ------------------------------------------------------------------------------------- */
static void undistort_point(double u, double v, const OkpCamera* cam, double* ou, double* ov) {
    const double px = (u - cam->cx) / cam->fx;
    const double py = (v - cam->cy) / cam->fy;
    double theta_d = sqrt(px * px + py * py);
    const double half_pi = 3.14159265358979323846 / 2.0;
    if (theta_d < -half_pi) theta_d = -half_pi;
    if (theta_d > half_pi) theta_d = half_pi;
    double theta = theta_d, scale = 0.0;
    int converged = 0;
    if (fabs(theta_d) > 1e-8) {
        for (int it = 0; it < 10; ++it) {
            const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
            const double a = cam->k[0] * t2, b = cam->k[1] * t4, c = cam->k[2] * t6, d = cam->k[3] * t8;
            const double fix = (theta * (1 + a + b + c + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * c + 9 * d);
            theta = theta - fix;
            if (fabs(fix) < 1e-8) { converged = 1; break; }
        }
        scale = tan(theta) / theta_d;
    } else {
        converged = 1;
    }
    const int flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
    if (converged && !flipped) {
        *ou = cam->fx * (px * scale) + cam->cx;
        *ov = cam->fy * (py * scale) + cam->cy;
    } else {
        *ou = -1000000.0;
        *ov = -1000000.0;
    }
}

void okp_oracle_undistort_f64(const double* xy, int n, const OkpCamera* cam, int round_to_f32, double* out) {
    for (int i = 0; i < n; ++i) {
        double u, v;
        undistort_point(xy[2 * i], xy[2 * i + 1], cam, &u, &v);
        if (round_to_f32) { u = (double)(float)u; v = (double)(float)v; }
        out[2 * i] = u;
        out[2 * i + 1] = v;
    }
}

/* A11: cv::fisheye::projectPoints -- camera_utils.py:65-73 */
void okp_oracle_project_f64(const double* X, int n, const double* T, const OkpCamera* cam, double* out) {
    for (int i = 0; i < n; ++i) {
        const double x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
        const double xc = T[0] * x + T[1] * y + T[2] * z + T[3];
        const double yc = T[4] * x + T[5] * y + T[6] * z + T[7];
        const double zc = T[8] * x + T[9] * y + T[10] * z + T[11];
        const double a = xc / zc, b = yc / zc;
        const double r = sqrt(a * a + b * b);
        const double th = atan(r), t2 = th * th;
        const double thd = th * (1.0 + cam->k[0] * t2 + cam->k[1] * t2 * t2 + cam->k[2] * t2 * t2 * t2 +
                                 cam->k[3] * t2 * t2 * t2 * t2);
        const double s = r > 1e-8 ? thd / r : 1.0;
        out[2 * i] = cam->fx * (a * s) + cam->cx;
        out[2 * i + 1] = cam->fy * (b * s) + cam->cy;
    }
}

/* A8: DetectionToPoint.__call__ (pipeline.py:164-171) + PinholeCamera.unproject (camera_utils.py:31-34) */
static void detection_to_point(const float* xy, const float* depth_map, int H, int W, const OkpCamera* cam,
                               int compat_clip_bug, double* out) {
    double du, dv;
    undistort_point((double)xy[0], (double)xy[1], cam, &du, &dv);
    const float ux = (float)du, uy = (float)dv;            /* OpenCV: float32 in, float32 out */
    long xi = lrintf(ux), yi = lrintf(uy);                 /* np.round: half to even */
    if (compat_clip_bug) {
        if (xi < 0) xi = 0; if (xi > cam->clip_x) xi = cam->clip_x;
        if (yi < 0) yi = 0; if (yi > cam->clip_y) yi = cam->clip_y;
    }
    if (xi < 0) xi = 0; if (xi > W - 1) xi = W - 1;
    if (yi < 0) yi = 0; if (yi > H - 1) yi = H - 1;
    const double z = (double)depth_map[yi * W + xi];
    const double hx = (double)ux, hy = (double)uy;
    for (int r = 0; r < 3; ++r)
        out[r] = (cam->kinv[3 * r] * hx + cam->kinv[3 * r + 1] * hy + cam->kinv[3 * r + 2]) * z;
}

void okp_oracle_detection_to_point_f32(const float* xy, int n, const float* depth_map, int H, int W,
                                       const OkpCamera* cam, const OkpDecodeParams* params, double* out) {
    for (int i = 0; i < n; ++i)
        detection_to_point(xy + 2 * i, depth_map, H, W, cam, params->compat_clip_bug, out + 3 * i);
}

/* ---------------------------------------------------------------------------------------
 * A2-A5: box sum, NMS, threshold, centroid -- pipeline.py:46-79, models.py:55-58
 * ------------------------------------------------------------------------------------- */
/* torch's max_pool2d update rule, `if (val > max || isnan(val)) max = val` (ATen MaxPoolKernel): NaN propagates, so no
 * pixel within reach of a NaN box sum passes `x == hmax` (perception/models.py:55-58). That form does not vectorise
 * (2.8x slower end to end), so it runs only on maps whose box sums hold a NaN (one vectorised look); every other map
 * takes the plain maximum, which gives the same result when no operand is NaN. */
#define MAXF_NAN(a, b) (((b) > (a) || (b) != (b)) ? (b) : (a))
#define MAXF_FAST(a, b) ((a) > (b) ? (a) : (b))
#define MAXF(a, b) (nan_aware ? MAXF_NAN(a, b) : MAXF_FAST(a, b))

typedef struct {
    float* padded;   /* (H+4) x (W+4), zero border */
    float* score;    /* (H+4) x (W+4), -inf border */
    float* rowmax;   /* (H+4) x W */
} MapScratch;

static inline __attribute__((always_inline)) void find_peaks_map_impl(const float* p, int H, int W, float threshold,
        MapScratch* s, int K, int32_t* count, int32_t* yx, float* score_out, float* xy, float* conf, const int nan_aware);

static void find_peaks_map(const float* p, int H, int W, float threshold, MapScratch* s,
                           int K, int32_t* count, int32_t* yx, float* score_out, float* xy, float* conf) {
    /* a box sum is NaN only if a NaN or an Inf enters it (x - x is 0 for every finite x): one vectorised look decides
     * which maximum runs */
    int has_nan = 0;
    for (size_t i = 0; i < (size_t)H * W; ++i) has_nan |= !(p[i] - p[i] == 0.0f);
    if (has_nan) find_peaks_map_impl(p, H, W, threshold, s, K, count, yx, score_out, xy, conf, 1);
    else find_peaks_map_impl(p, H, W, threshold, s, K, count, yx, score_out, xy, conf, 0);
}

static inline __attribute__((always_inline)) void find_peaks_map_impl(const float* p, int H, int W, float threshold,
        MapScratch* s, int K, int32_t* count, int32_t* yx, float* score_out, float* xy, float* conf, const int nan_aware) {
    const int PW = W + 4;
    memset(s->padded, 0, sizeof(float) * (size_t)(H + 4) * PW);
    for (int y = 0; y < H; ++y) memcpy(s->padded + (size_t)(y + 2) * PW + 2, p + (size_t)y * W, sizeof(float) * W);
    for (size_t i = 0; i < (size_t)(H + 4) * PW; ++i) s->score[i] = -INFINITY;
    /* box sum: 25 sequential float32 additions per pixel, raster tap order, starting from +0 */
    for (int y = 0; y < H; ++y) {
        float* acc = s->score + (size_t)(y + 2) * PW + 2;
        for (int x = 0; x < W; ++x) acc[x] = 0.0f;
        for (int dy = 0; dy < 5; ++dy) {
            const float* row = s->padded + (size_t)(y + dy) * PW;
            for (int dx = 0; dx < 5; ++dx) {
                const float* src = row + dx;
                for (int x = 0; x < W; ++x) acc[x] = acc[x] + src[x];
            }
        }
    }
    /* 5x5 maximum, separable (max is exact, so order does not matter) */
    for (int y = 0; y < H + 4; ++y) {
        const float* row = s->score + (size_t)y * PW;
        float* out = s->rowmax + (size_t)y * W;
        for (int x = 0; x < W; ++x) {
            float m = row[x];
            m = MAXF(m, row[x + 1]); m = MAXF(m, row[x + 2]); m = MAXF(m, row[x + 3]); m = MAXF(m, row[x + 4]);
            out[x] = m;
        }
    }
    int n = 0;
    for (int y = 0; y < H; ++y) {
        const float* r0 = s->rowmax + (size_t)y * W;
        const float* c = s->score + (size_t)(y + 2) * PW + 2;
        for (int x = 0; x < W; ++x) {
            float m = r0[x];
            m = MAXF(m, r0[x + W]); m = MAXF(m, r0[x + 2 * W]); m = MAXF(m, r0[x + 3 * W]); m = MAXF(m, r0[x + 4 * W]);
            const float v = c[x];
            if (v == m && v > threshold) {
                if (n < K) {
                    yx[2 * n] = y; yx[2 * n + 1] = x;
                    score_out[n] = v;
                    /* centroid over the clipped window, float32, raster order, no FMA */
                    float sy = 0.0f, sx = 0.0f, sp = 0.0f;
                    const int y0 = y - 2 < 0 ? 0 : y - 2, y1 = y + 3 > H ? H : y + 3;
                    const int x0 = x - 2 < 0 ? 0 : x - 2, x1 = x + 3 > W ? W : x + 3;
                    for (int i = y0; i < y1; ++i)
                        for (int j = x0; j < x1; ++j) {
                            const float q = p[(size_t)i * W + j];
                            sy = sy + q * (float)i;
                            sx = sx + q * (float)j;
                            sp = sp + q;
                        }
                    xy[2 * n] = sx / sp; xy[2 * n + 1] = sy / sp;
                    conf[n] = sp;
                }
                ++n;
            }
        }
    }
    *count = n;
}

/* ---------------------------------------------------------------------------------------
 * deterministic stand-in for KMeans(init='random') -- pipeline.py:146-148 (parity unpinned)
 * ------------------------------------------------------------------------------------- */
#define KMEANS_MAX_INITS 256
static void cluster_detections(const float* pts32, int n, int k, int iters, float* out) {
    double pts[OKP_MAX_PEAKS][2], cen[OKP_MAX_SLOTS][2], best_cen[OKP_MAX_SLOTS][2];
    int subset[OKP_MAX_SLOTS], assign[OKP_MAX_PEAKS], new_assign[OKP_MAX_PEAKS];
    double best = 0.0; int have_best = 0;
    for (int i = 0; i < n; ++i) { pts[i][0] = pts32[2 * i]; pts[i][1] = pts32[2 * i + 1]; }
    for (int c = 0; c < k; ++c) subset[c] = c;
    for (int init = 0; init < KMEANS_MAX_INITS; ++init) {
        for (int c = 0; c < k; ++c) { cen[c][0] = pts[subset[c]][0]; cen[c][1] = pts[subset[c]][1]; }
        int have_assign = 0;
        for (int it = 0; it < iters; ++it) {
            int same = have_assign;
            for (int i = 0; i < n; ++i) {
                int arg = 0; double dmin = 0.0;
                for (int c = 0; c < k; ++c) {
                    const double dx = pts[i][0] - cen[c][0], dy = pts[i][1] - cen[c][1];
                    const double d = dx * dx + dy * dy;
                    if (c == 0 || d < dmin) { dmin = d; arg = c; }
                }
                new_assign[i] = arg;
                if (have_assign && assign[i] != arg) same = 0;
            }
            if (same) break;
            memcpy(assign, new_assign, sizeof(int) * n);
            have_assign = 1;
            for (int c = 0; c < k; ++c) {
                double sx = 0.0, sy = 0.0; int m = 0;
                for (int i = 0; i < n; ++i) if (assign[i] == c) { sx += pts[i][0]; sy += pts[i][1]; ++m; }
                if (m) { cen[c][0] = sx / m; cen[c][1] = sy / m; }
            }
        }
        double inertia = 0.0;
        for (int i = 0; i < n; ++i) {
            double dmin = 0.0;
            for (int c = 0; c < k; ++c) {
                const double dx = pts[i][0] - cen[c][0], dy = pts[i][1] - cen[c][1];
                const double d = dx * dx + dy * dy;
                if (c == 0 || d < dmin) dmin = d;
            }
            inertia += dmin;
        }
        if (!have_best || inertia < best) { best = inertia; have_best = 1; memcpy(best_cen, cen, sizeof(cen)); }
        /* next k-subset in lexicographic order */
        int c = k - 1;
        while (c >= 0 && subset[c] == n - k + c) --c;
        if (c < 0) break;
        ++subset[c];
        for (int j = c + 1; j < k; ++j) subset[j] = subset[j - 1] + 1;
    }
    for (int c = 0; c < k; ++c) { out[2 * c] = (float)best_cen[c][0]; out[2 * c + 1] = (float)best_cen[c][1]; }
}

/* ---------------------------------------------------------------------------------------
 * A0/A6: whole decode of one frame -- pipeline.py:104-153, 182-200
 * ------------------------------------------------------------------------------------- */
static void decode_frame(const float* heat, const float* depth, const float* centers, int n, int C, int H, int W,
                         const int32_t* keypoint_config, const OkpCamera* cam, const OkpDecodeParams* prm,
                         const OkpDecodeTables* t, MapScratch* scratch) {
    const int K = prm->max_peaks, O = prm->max_objects, V = prm->max_votes, T = C - 1;
    int cfg[OKP_MAX_MAPS];
    cfg[0] = 1;
    int S = 1;
    for (int i = 0; i < T; ++i) { cfg[1 + i] = keypoint_config[i]; if (cfg[1 + i] > S) S = cfg[1 + i]; }
    const size_t HW = (size_t)H * W;
    uint32_t flags = 0;
    for (int c = 0; c < C; ++c) {
        const size_t m = (size_t)n * C + c;
        find_peaks_map(heat + m * HW, H, W, prm->threshold, scratch, K, t->peak_count + m, t->peak_yx + m * K * 2,
                       t->peak_score + m * K, t->peak_xy + m * K * 2, t->peak_conf + m * K);
        if (t->peak_count[m] > K) flags |= OKP_FLAG_PEAK_OVERFLOW;
    }
    const size_t m0 = (size_t)n * C;
    int n_center = t->peak_count[m0] < K ? t->peak_count[m0] : K;
    if (n_center == 0) { t->flags[n] = flags | OKP_FLAG_NO_CENTERS; t->n_objects[n] = 0; return; }
    if (n_center > O) flags |= OKP_FLAG_OBJECT_OVERFLOW;
    const int n_obj = n_center < O ? n_center : O;
    t->n_objects[n] = n_obj;
    const float* center_xy = t->peak_xy + m0 * K * 2;
    for (int o = 0; o < n_obj; ++o) t->peak_object[m0 * K + o] = o;
    /* spoke assignment */
    for (int c = 1; c < C; ++c) {
        const size_t m = (size_t)n * C + c;
        const int k_n = t->peak_count[m] < K ? t->peak_count[m] : K;
        const float* cmap = centers + ((size_t)n * T + (c - 1)) * 2 * HW;
        for (int k = 0; k < k_n; ++k) {
            const float px = t->peak_xy[(m * K + k) * 2], py = t->peak_xy[(m * K + k) * 2 + 1];
            long xi = lrintf(px), yi = lrintf(py);
            if (xi < 0) xi = 0; if (xi > W - 1) xi = W - 1;
            if (yi < 0) yi = 0; if (yi > H - 1) yi = H - 1;
            const double vx = ((double)xi + 0.5) + (double)cmap[(size_t)yi * W + xi];
            const double vy = ((double)yi + 0.5) + (double)cmap[HW + (size_t)yi * W + xi];
            t->peak_vote[(m * K + k) * 2] = vx;
            t->peak_vote[(m * K + k) * 2 + 1] = vy;
            int arg = 0; double dmin = 0.0;
            for (int o = 0; o < n_obj; ++o) {
                const double dx = (double)center_xy[2 * o] - vx, dy = (double)center_xy[2 * o + 1] - vy;
                const double d = sqrt(dx * dx + dy * dy);
                if (o == 0 || d < dmin) { dmin = d; arg = o; }
            }
            if (dmin > prm->outlier_distance) { flags |= OKP_FLAG_OUTLIER_SKIPPED; continue; }
            t->peak_object[m * K + k] = arg;
            const size_t ob = (size_t)n * O + arg;
            if (t->n_votes[ob] < V) {
                t->votes[(ob * V + t->n_votes[ob]) * 2] = vx;
                t->votes[(ob * V + t->n_votes[ob]) * 2 + 1] = vy;
            } else {
                flags |= OKP_FLAG_VOTE_OVERFLOW;
            }
            t->n_votes[ob] += 1;
        }
    }
    /* resolution + 3D */
    for (int o = 0; o < n_obj; ++o) {
        for (int c = 0; c < C; ++c) {
            const size_t m = (size_t)n * C + c;
            const size_t oc = ((size_t)n * O + o) * C + c;
            const int k_n = t->peak_count[m] < K ? t->peak_count[m] : K;
            int idx[OKP_MAX_PEAKS], cnt = 0;
            for (int k = 0; k < k_n; ++k) if (t->peak_object[m * K + k] == o) idx[cnt++] = k;
            t->kp_assigned[oc] = cnt;
            if (cnt == 0) continue;
            float pts[OKP_MAX_SLOTS][2];
            int ids[OKP_MAX_SLOTS], kept = cnt;
            if (cnt > cfg[c]) {
                if (cfg[c] == 1) {
                    int arg = 0;
                    for (int i = 1; i < cnt; ++i)
                        if (t->peak_conf[m * K + idx[i]] > t->peak_conf[m * K + idx[arg]]) arg = i;
                    ids[0] = idx[arg];
                    pts[0][0] = t->peak_xy[(m * K + ids[0]) * 2]; pts[0][1] = t->peak_xy[(m * K + ids[0]) * 2 + 1];
                    kept = 1;
                    flags |= OKP_FLAG_ARGMAX_RESOLVED;
                } else {
                    float gathered[OKP_MAX_PEAKS * 2];
                    for (int i = 0; i < cnt; ++i) {
                        gathered[2 * i] = t->peak_xy[(m * K + idx[i]) * 2];
                        gathered[2 * i + 1] = t->peak_xy[(m * K + idx[i]) * 2 + 1];
                    }
                    kept = cfg[c];
                    cluster_detections(gathered, cnt, kept, prm->kmeans_iterations, &pts[0][0]);
                    for (int i = 0; i < kept; ++i) ids[i] = -1;
                    flags |= OKP_FLAG_CLUSTERED;
                }
            } else {
                for (int i = 0; i < cnt; ++i) {
                    ids[i] = idx[i];
                    pts[i][0] = t->peak_xy[(m * K + idx[i]) * 2]; pts[i][1] = t->peak_xy[(m * K + idx[i]) * 2 + 1];
                }
            }
            t->kp_count[oc] = kept;
            for (int s = 0; s < kept; ++s) {
                t->kp_peak[oc * S + s] = ids[s];
                t->kp_xy[(oc * S + s) * 2] = pts[s][0];
                t->kp_xy[(oc * S + s) * 2 + 1] = pts[s][1];
                if (cam)
                    detection_to_point(pts[s], depth + m * HW, H, W, cam, prm->compat_clip_bug,
                                       t->kp_point + (oc * S + s) * 3);
            }
        }
    }
    t->flags[n] = flags;
}

static void reset_tables(int N, int C, const int32_t* keypoint_config, const OkpDecodeParams* prm,
                         const OkpDecodeTables* t) {
    const size_t K = prm->max_peaks, O = prm->max_objects, V = prm->max_votes;
    size_t S = 1;
    for (int i = 0; i < C - 1; ++i) if ((size_t)keypoint_config[i] > S) S = keypoint_config[i];
    const size_t NC = (size_t)N * C;
    memset(t->peak_count, 0, sizeof(int32_t) * NC);
    memset(t->peak_yx, 0xff, sizeof(int32_t) * NC * K * 2);
    memset(t->peak_score, 0, sizeof(float) * NC * K);
    memset(t->peak_xy, 0, sizeof(float) * NC * K * 2);
    memset(t->peak_conf, 0, sizeof(float) * NC * K);
    memset(t->peak_object, 0xff, sizeof(int32_t) * NC * K);
    memset(t->peak_vote, 0, sizeof(double) * NC * K * 2);
    memset(t->n_objects, 0, sizeof(int32_t) * N);
    memset(t->flags, 0, sizeof(uint32_t) * N);
    memset(t->kp_assigned, 0, sizeof(int32_t) * N * O * C);
    memset(t->kp_count, 0, sizeof(int32_t) * N * O * C);
    memset(t->kp_peak, 0xff, sizeof(int32_t) * N * O * C * S);
    memset(t->kp_xy, 0, sizeof(float) * N * O * C * S * 2);
    memset(t->kp_point, 0, sizeof(double) * N * O * C * S * 3);
    memset(t->n_votes, 0, sizeof(int32_t) * N * O);
    memset(t->votes, 0, sizeof(double) * N * O * V * 2);
}

/* All pointers are HOST pointers here. threads <= 0: use every core. Returns OKP_OK / OKP_E_*. */
int okp_oracle_decode_f32(const float* heat, const float* depth, const float* centers, int N, int C, int H, int W,
                          const int32_t* keypoint_config, const OkpCamera* cam, const OkpDecodeParams* prm,
                          const OkpDecodeTables* t, int threads, int reset) {
    if (!heat || !keypoint_config || !prm || !t) return OKP_E_NULL;
    if (N < 0 || C < 1 || C > OKP_MAX_MAPS || H < 1 || W < 1) return OKP_E_SHAPE;
    if (prm->max_peaks < 1 || prm->max_peaks > OKP_MAX_PEAKS || prm->max_objects < 1 ||
        prm->max_objects > OKP_MAX_OBJECTS || prm->max_votes < 1) return OKP_E_CAPACITY;
    if (prm->nms_size != 5 || !prm->box_sum) return OKP_E_UNSUPPORTED;
    if (reset) reset_tables(N, C, keypoint_config, prm, t);
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
#pragma omp parallel num_threads(threads)
    {
        MapScratch s;
        s.padded = (float*)malloc(sizeof(float) * (size_t)(H + 4) * (W + 4));
        s.score = (float*)malloc(sizeof(float) * (size_t)(H + 4) * (W + 4));
        s.rowmax = (float*)malloc(sizeof(float) * (size_t)(H + 4) * W);
#pragma omp for schedule(dynamic, 1)
        for (int n = 0; n < N; ++n)
            decode_frame(heat, depth, centers, n, C, H, W, keypoint_config, cam, prm, t, &s);
        free(s.padded); free(s.score); free(s.rowmax);
    }
    return OKP_OK;
}

int okp_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------
 * A12/A14: DLT -- camera_utils.py:103-108, scripts/label.py:296-305, any number of views.
 * One-sided Jacobi SVD of the stacked (x P[2] - P[0], y P[2] - P[1]) rows in float64.
 * ------------------------------------------------------------------------------------- */
static void smallest_right_singular_vector(double* A, int rows, double* h) {
    double Vm[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    /* a column whose norm has fallen below 1e-15 of the matrix norm is rounding noise (exactly consistent
     * observations: the smallest singular value is 0); rotating against it changes nothing representable and the
     * relative test below would never be met, so such pairs count as converged */
    double negligible = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < 4; ++c) negligible += A[r * 4 + c] * A[r * 4 + c];
    negligible *= 1e-30;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < rows; ++r) {
                    alpha += A[r * 4 + p] * A[r * 4 + p];
                    beta += A[r * 4 + q] * A[r * 4 + q];
                    gamma += A[r * 4 + p] * A[r * 4 + q];
                }
                if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 2.3e-16 * sqrt(alpha * beta)) continue;
                if (alpha <= negligible || beta <= negligible) continue;
                rotated = 1;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                for (int r = 0; r < rows; ++r) {
                    const double ap = A[r * 4 + p], aq = A[r * 4 + q];
                    A[r * 4 + p] = cs * ap - sn * aq;
                    A[r * 4 + q] = sn * ap + cs * aq;
                }
                for (int r = 0; r < 4; ++r) {
                    const double vp = Vm[r][p], vq = Vm[r][q];
                    Vm[r][p] = cs * vp - sn * vq;
                    Vm[r][q] = sn * vp + cs * vq;
                }
            }
        if (!rotated) break;
    }
    int arg = 0; double smallest = 0;
    for (int c = 0; c < 4; ++c) {
        double nrm = 0;
        for (int r = 0; r < rows; ++r) nrm += A[r * 4 + c] * A[r * 4 + c];
        if (c == 0 || nrm < smallest) { smallest = nrm; arg = c; }
    }
    for (int r = 0; r < 4; ++r) h[r] = Vm[r][arg];
}

int okp_oracle_triangulate_f64(const double* points, const uint8_t* valid, const double* projections,
                               int per_point, int P, int V, double* out) {
    if (!points || !projections || !out) return OKP_E_NULL;
    if (V < 1 || V > OKP_MAX_VIEWS || P < 0) return OKP_E_SHAPE;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; ++p) {
        double A[2 * OKP_MAX_VIEWS * 4];
        int rows = 0;
        for (int v = 0; v < V; ++v) {
            if (valid && !valid[(size_t)p * V + v]) continue;
            const double* M = projections + (per_point ? ((size_t)p * V + v) * 12 : (size_t)v * 12);
            const double x = points[((size_t)p * V + v) * 2], y = points[((size_t)p * V + v) * 2 + 1];
            for (int c = 0; c < 4; ++c) {
                A[rows * 4 + c] = x * M[8 + c] - M[c];
                A[(rows + 1) * 4 + c] = y * M[8 + c] - M[4 + c];
            }
            rows += 2;
        }
        if (rows < 4) { out[3 * p] = out[3 * p + 1] = out[3 * p + 2] = NAN; continue; }
        double h[4];
        smallest_right_singular_vector(A, rows, h);
        out[3 * p] = h[0] / h[3]; out[3 * p + 1] = h[1] / h[3]; out[3 * p + 2] = h[2] / h[3];
    }
    return OKP_OK;
}

/* reprojection error of X in every view + gating (north_star; no reference code) */
int okp_oracle_reprojection_filter_f64(const double* X, const double* obs, uint8_t* valid, const double* poses,
                                       const OkpCamera* cam, int P, int V, double max_error_px, double* err) {
    for (int p = 0; p < P; ++p)
        for (int v = 0; v < V; ++v) {
            double uv[2];
            okp_oracle_project_f64(X + 3 * p, 1, poses + (size_t)v * 16, cam, uv);
            const double dx = uv[0] - obs[((size_t)p * V + v) * 2], dy = uv[1] - obs[((size_t)p * V + v) * 2 + 1];
            const double e = sqrt(dx * dx + dy * dy);
            err[(size_t)p * V + v] = e;
            if (!(e <= max_error_px)) valid[(size_t)p * V + v] = 0;
        }
    return OKP_OK;
}

/* robust V-view triangulation: DLT, reprojection errors, drop the worst view above the gate, repeat
 * (north_star K5 + K6; no reference code -- same statement as np_oracle.triangulate_robust) */
int okp_oracle_triangulate_robust_f64(const double* obs, uint8_t* valid, const double* poses, const OkpCamera* cam,
                                      int P, int V, double max_error_px, int max_rounds, double* out, double* err,
                                      int32_t* dropped) {
    if (!obs || !poses || !cam || !out || !err) return OKP_E_NULL;
    if (V < 1 || V > OKP_MAX_VIEWS || P < 0) return OKP_E_SHAPE;
    double proj[OKP_MAX_VIEWS * 12];
    for (int v = 0; v < V; ++v) {
        const double* T = poses + (size_t)v * 16;
        for (int c = 0; c < 4; ++c) {
            proj[v * 12 + c] = cam->fx * T[c] + cam->cx * T[8 + c];
            proj[v * 12 + 4 + c] = cam->fy * T[4 + c] + cam->cy * T[8 + c];
            proj[v * 12 + 8 + c] = T[8 + c];
        }
    }
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; ++p) {
        uint8_t mask[OKP_MAX_VIEWS];
        double und[OKP_MAX_VIEWS * 2];
        for (int v = 0; v < V; ++v) {
            mask[v] = valid ? (valid[(size_t)p * V + v] != 0) : 1;
            okp_oracle_undistort_f64(obs + ((size_t)p * V + v) * 2, 1, cam, 0, und + 2 * v);
        }
        int drops = 0;
        double X[3] = {NAN, NAN, NAN};
        for (;;) {
            int views = 0;
            for (int v = 0; v < V; ++v) views += mask[v];
            if (views < 2) { X[0] = X[1] = X[2] = NAN; break; }
            okp_oracle_triangulate_f64(und, mask, proj, 0, 1, V, X);
            int worst = -1;
            double worst_err = 0.0;
            for (int v = 0; v < V; ++v) {
                double uv[2];
                okp_oracle_project_f64(X, 1, poses + (size_t)v * 16, cam, uv);
                const double dx = uv[0] - obs[((size_t)p * V + v) * 2], dy = uv[1] - obs[((size_t)p * V + v) * 2 + 1];
                const double e = sqrt(dx * dx + dy * dy);
                err[(size_t)p * V + v] = e;
                const double rank = e == e ? e : INFINITY;
                if (mask[v] && (worst < 0 || rank > worst_err)) { worst = v; worst_err = rank; }
            }
            if (!(worst_err > max_error_px) || views <= 2 || drops >= max_rounds) break;
            mask[worst] = 0;
            ++drops;
        }
        if (valid) for (int v = 0; v < V; ++v) valid[(size_t)p * V + v] = mask[v];
        out[3 * p] = X[0]; out[3 * p + 1] = X[1]; out[3 * p + 2] = X[2];
        if (dropped) dropped[p] = drops;
    }
    return OKP_OK;
}
