// okp_group.cuh -- K3 + K4: centre-vector voting into object instances, over-detection
// resolution and 3D points. One CTA per frame; everything it touches is a few hundred bytes of
// peak records plus 3 gathered floats per spoke peak (2 centre-vector components, 1 depth).
//
// Replaces ObjectExtraction.__call__ (perception/pipeline.py:104-153) and the DetectionToPoint
// loop of ObjectKeypointPipeline.__call__ (pipeline.py:189-199).
#pragma once
#include "okp_common.cuh"
#include "okp_geometry.cuh"

#define OKP_KMEANS_MAX_INITS 256

struct OkpConfig { int32_t cfg[OKP_MAX_MAPS]; };    // cfg[0] = 1 (centre map), then keypoint_config

// Deterministic stand-in for the reference's unseeded KMeans(init='random') (pipeline.py:146-148):
// Lloyd's algorithm from every k-subset of the detections (lexicographic, at most
// OKP_KMEANS_MAX_INITS), lowest inertia wins. Same statement as oracle/okp_oracle.c.
// xy: the map's peak_xy table; idx[0..n): the detections of this (object, type). Points are re-read
// from the table instead of being copied, to keep the per-thread stack small.
__device__ __noinline__ void okp_cluster_detections(const float* __restrict__ xy, const unsigned char* idx, int n, int k,
                                                    int iters, float* out) {
#define OKP_PT(i, d) ((double)xy[2 * (int)idx[(i)] + (d)])
    double cen[OKP_MAX_SLOTS][2], best_cen[OKP_MAX_SLOTS][2];
    int subset[OKP_MAX_SLOTS];
    unsigned char assign[OKP_MAX_PEAKS], new_assign[OKP_MAX_PEAKS];
    double best = 0.0;
    bool have_best = false;
    for (int c = 0; c < k; ++c) subset[c] = c;
    for (int init = 0; init < OKP_KMEANS_MAX_INITS; ++init) {
        for (int c = 0; c < k; ++c) { cen[c][0] = OKP_PT(subset[c], 0); cen[c][1] = OKP_PT(subset[c], 1); }
        bool have_assign = false;
        for (int it = 0; it < iters; ++it) {
            bool same = have_assign;
            for (int i = 0; i < n; ++i) {
                int arg = 0;
                double dmin = 0.0;
                for (int c = 0; c < k; ++c) {
                    const double dx = OKP_PT(i, 0) - cen[c][0], dy = OKP_PT(i, 1) - cen[c][1];
                    const double d = dx * dx + dy * dy;
                    if (c == 0 || d < dmin) { dmin = d; arg = c; }
                }
                new_assign[i] = (unsigned char)arg;
                if (have_assign && assign[i] != arg) same = false;
            }
            if (same) break;
            for (int i = 0; i < n; ++i) assign[i] = new_assign[i];
            have_assign = true;
            for (int c = 0; c < k; ++c) {
                double sx = 0.0, sy = 0.0;
                int m = 0;
                for (int i = 0; i < n; ++i)
                    if (assign[i] == c) { sx += OKP_PT(i, 0); sy += OKP_PT(i, 1); ++m; }
                if (m) { cen[c][0] = sx / m; cen[c][1] = sy / m; }
            }
        }
        double inertia = 0.0;
        for (int i = 0; i < n; ++i) {
            double dmin = 0.0;
            for (int c = 0; c < k; ++c) {
                const double dx = OKP_PT(i, 0) - cen[c][0], dy = OKP_PT(i, 1) - cen[c][1];
                const double d = dx * dx + dy * dy;
                if (c == 0 || d < dmin) dmin = d;
            }
            inertia += dmin;
        }
        if (!have_best || inertia < best) {
            best = inertia;
            have_best = true;
            for (int c = 0; c < k; ++c) { best_cen[c][0] = cen[c][0]; best_cen[c][1] = cen[c][1]; }
        }
        int c = k - 1;
        while (c >= 0 && subset[c] == n - k + c) --c;
        if (c < 0) break;
        ++subset[c];
        for (int j = c + 1; j < k; ++j) subset[j] = subset[j - 1] + 1;
    }
    for (int c = 0; c < k; ++c) { out[2 * c] = (float)best_cen[c][0]; out[2 * c + 1] = (float)best_cen[c][1]; }
#undef OKP_PT
}

// Shared-memory bytes ONE frame (= one warp) of okp_group_kernel needs; `stash` = the kept keypoints are also
// held in shared memory for the 3D lift (dropped when the worst-case capacities would not fit).
static inline size_t okp_group_smem_bytes(int C, int K, int O, int S, bool stash) {
    size_t bytes = (size_t)O * 2 * sizeof(double);                                                  // centres
    bytes += (size_t)C * K * (2 * sizeof(double) + 2 * sizeof(float) + sizeof(float) + sizeof(int));
    bytes += (size_t)O * C * sizeof(int);
    if (stash) bytes += (size_t)O * C * S * 2 * sizeof(float);
    bytes += (OKP_MAX_MAPS + 1) * sizeof(int);                                                      // counts, flags
    return (bytes + 15) / 16 * 16;
}

// Latency is what this kernel is made of (a frame is ~40 peaks, and every 3D lift is a serial float64
// Newton + tan chain of a few microseconds): ONE WARP PER FRAME, several frames per CTA, so that an SM holds
// 64 frames whose chains overlap (the CTA-per-frame form held 16 and took 4x longer). The frame's peak
// records are pulled into shared memory with one round trip, every later phase works on shared memory, and
// results leave as fire-and-forget stores. Global round trips on the critical path: counts -> records ->
// centre-vector gather -> depth gather. No block-level barrier anywhere: warps are independent.
template <typename E>
__global__ void __launch_bounds__(128)
okp_group_kernel(const E* __restrict__ depth, const E* __restrict__ centers, int N, int C, int H, int W,
                 OkpConfig config, OkpCamera cam, int have_camera, OkpDecodeParams prm, int S, int stash,
                 int frame_smem_bytes, OkpDecodeTables t) {
    constexpr int THREADS = 32;                            // a frame's team: the strides below
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    const int K = prm.max_peaks, O = prm.max_objects, V = prm.max_votes, T = C - 1;
    const size_t HW = (size_t)H * W;
    extern __shared__ __align__(16) unsigned char group_smem[];
    unsigned char* mine = group_smem + (size_t)(threadIdx.x >> 5) * frame_smem_bytes;
    double (*s_center)[2] = reinterpret_cast<double (*)[2]>(mine);          // [O][2]
    double* s_vote = reinterpret_cast<double*>(mine) + (size_t)O * 2;       // [C][K][2] predicted centre of a spoke peak
    float* s_xy = reinterpret_cast<float*>(s_vote + (size_t)C * K * 2);     // [C][K][2] centroid (x, y)
    float* s_conf = s_xy + (size_t)C * K * 2;                               // [C][K]
    int* s_obj = reinterpret_cast<int*>(s_conf + (size_t)C * K);            // [C][K]    object of a peak, -1 = none
    int* s_kept = s_obj + (size_t)C * K;                                    // [O][C]    keypoints kept per (object, map)
    int* s_counts = s_kept + (size_t)O * C;                                 // [OKP_MAX_MAPS]
    unsigned int* s_flags_ptr = reinterpret_cast<unsigned int*>(s_counts + OKP_MAX_MAPS);
    float* s_kept_xy = reinterpret_cast<float*>(s_flags_ptr + 1);           // [O][C][S][2] (only with stash)
#define s_flags (*s_flags_ptr)

    const size_t m0 = (size_t)n * C;
    if (lane == 0) s_flags = 0;
    __syncwarp();
    if (lane < C) {
        const int c = t.peak_count[m0 + lane];
        s_counts[lane] = c < K ? c : K;
        if (c > K) atomicOr(&s_flags, OKP_FLAG_PEAK_OVERFLOW);
    }
    // reset this frame's object tables (flat, coalesced; every region is contiguous per frame)
    {
        const int oc = O * C, ocs = oc * S;
        int32_t* assigned = t.kp_assigned + (size_t)n * oc;
        int32_t* count = t.kp_count + (size_t)n * oc;
        for (int i = lane; i < oc; i += THREADS) { assigned[i] = 0; count[i] = 0; }
        int32_t* peak = t.kp_peak + (size_t)n * ocs;
        for (int i = lane; i < ocs; i += THREADS) peak[i] = -1;
        float* xy = t.kp_xy + (size_t)n * ocs * 2;
        for (int i = lane; i < ocs * 2; i += THREADS) xy[i] = 0.0f;
        double* point = t.kp_point + (size_t)n * ocs * 3;
        for (int i = lane; i < ocs * 3; i += THREADS) point[i] = 0.0;
        int32_t* nv = t.n_votes + (size_t)n * O;
        for (int i = lane; i < O; i += THREADS) nv[i] = 0;
        double* votes = t.votes + (size_t)n * O * V * 2;
        for (int i = lane; i < O * V * 2; i += THREADS) votes[i] = 0.0;
    }
    __syncwarp();

    const int n_center = s_counts[0];
    if (n_center == 0) {                                   // pipeline.py:105-106
        if (lane == 0) { t.n_objects[n] = 0; t.flags[n] = s_flags | OKP_FLAG_NO_CENTERS; }
        return;
    }
    const int n_obj = n_center < O ? n_center : O;
    if (lane == 0 && n_center > O) atomicOr(&s_flags, OKP_FLAG_OBJECT_OVERFLOW);

    // ---- the frame's peak records into shared memory (one round trip); spoke peaks fetch their centre vector
    // in the same pass and vote as soon as the centres are known ----
    for (int i = lane; i < C * K; i += THREADS) {
        const int c = i / K, k = i - c * K;
        if (k >= s_counts[c]) continue;
        const size_t s = (m0 + c) * K + k;
        const float2 p = *reinterpret_cast<const float2*>(t.peak_xy + 2 * s);
        s_xy[2 * i] = p.x; s_xy[2 * i + 1] = p.y;
        s_conf[i] = t.peak_conf[s];
        if (c == 0) {                                      // pipeline.py:109-114: one object per centre peak
            const int o = k < n_obj ? k : -1;
            s_obj[i] = o;
            if (o >= 0) { s_center[o][0] = (double)p.x; s_center[o][1] = (double)p.y; t.peak_object[s] = o; }
        } else {
            const int xi = okp_clamp(__float2int_rn(p.x), 0, W - 1);      // np.round = half to even
            const int yi = okp_clamp(__float2int_rn(p.y), 0, H - 1);
            const E* cmap = centers + ((size_t)n * T + (c - 1)) * 2 * HW;
            const double vx = ((double)xi + 0.5) + (double)okp_ld<E>(cmap + (size_t)yi * W + xi);
            const double vy = ((double)yi + 0.5) + (double)okp_ld<E>(cmap + HW + (size_t)yi * W + xi);
            s_vote[2 * i] = vx; s_vote[2 * i + 1] = vy;
            t.peak_vote[2 * s] = vx;
            t.peak_vote[2 * s + 1] = vy;
        }
    }
    __syncwarp();

    // ---- spoke peaks vote for a centre (pipeline.py:115-128) ----
    for (int i = K + lane; i < C * K; i += THREADS) {
        const int c = i / K, k = i - c * K;
        if (k >= s_counts[c]) continue;
        const double vx = s_vote[2 * i], vy = s_vote[2 * i + 1];
        int arg = 0;
        double dmin = 0.0;
        for (int o = 0; o < n_obj; ++o) {
            const double dx = s_center[o][0] - vx, dy = s_center[o][1] - vy;
            const double d = sqrt(dx * dx + dy * dy);
            if (o == 0 || d < dmin) { dmin = d; arg = o; }             // first minimum, like np.argmin
        }
        if (dmin > prm.outlier_distance) {
            atomicOr(&s_flags, OKP_FLAG_OUTLIER_SKIPPED);              // the reference prints and skips
            arg = -1;
        }
        s_obj[i] = arg;
        t.peak_object[(m0 + c) * K + k] = arg;
    }
    __syncwarp();

    // ---- votes per object, in assignment order (type-major, raster order inside a type) ----
    for (int o = lane; o < n_obj; o += THREADS) {
        const size_t ob = (size_t)n * O + o;
        int nv = 0;
        for (int c = 1; c < C; ++c) {
            for (int k = 0; k < s_counts[c]; ++k) {
                if (s_obj[c * K + k] != o) continue;
                if (nv < V) {
                    t.votes[(ob * V + nv) * 2] = s_vote[(c * K + k) * 2];
                    t.votes[(ob * V + nv) * 2 + 1] = s_vote[(c * K + k) * 2 + 1];
                } else {
                    atomicOr(&s_flags, OKP_FLAG_VOTE_OVERFLOW);
                }
                ++nv;
            }
        }
        t.n_votes[ob] = nv;
    }

    // ---- per (object, map): resolve over-detection ----
    float* kept_xy = stash ? s_kept_xy : t.kp_xy + (size_t)n * O * C * S * 2;
    for (int i = lane; i < n_obj * C; i += THREADS) {
        const int o = i / C, c = i - o * C;
        const size_t oc = ((size_t)n * O + o) * C + c;
        const int limit = config.cfg[c];
        const int* obj = s_obj + c * K;
        const float* xy = s_xy + (size_t)c * K * 2;
        int cnt = 0;
        for (int k = 0; k < s_counts[c]; ++k) cnt += (obj[k] == o);
        t.kp_assigned[oc] = cnt;
        s_kept[i] = 0;
        if (cnt == 0) continue;                            // pipeline.py:150-152: empty array
        float pts[OKP_MAX_SLOTS][2];
        int ids[OKP_MAX_SLOTS];
        int kept = 0;
        if (cnt <= limit) {
            for (int k = 0; k < s_counts[c]; ++k)
                if (obj[k] == o) {
                    ids[kept] = k;
                    pts[kept][0] = xy[2 * k];
                    pts[kept][1] = xy[2 * k + 1];
                    ++kept;
                }
        } else if (limit == 1) {                           // pipeline.py:139-142: most confident detection
            int arg = -1;
            float best = 0.0f;
            for (int k = 0; k < s_counts[c]; ++k)
                if (obj[k] == o) {
                    const float conf = s_conf[c * K + k];
                    if (arg < 0 || conf > best) { best = conf; arg = k; }   // first maximum, like np.argmax
                }
            ids[0] = arg;
            pts[0][0] = xy[2 * arg];
            pts[0][1] = xy[2 * arg + 1];
            kept = 1;
            atomicOr(&s_flags, OKP_FLAG_ARGMAX_RESOLVED);
        } else {                                           // pipeline.py:143-148: cluster
            unsigned char members[OKP_MAX_PEAKS];
            int g = 0;
            for (int k = 0; k < s_counts[c]; ++k)
                if (obj[k] == o) members[g++] = (unsigned char)k;
            kept = limit;
            okp_cluster_detections(xy, members, g, kept, prm.kmeans_iterations, &pts[0][0]);
            for (int s = 0; s < kept; ++s) ids[s] = -1;
            atomicOr(&s_flags, OKP_FLAG_CLUSTERED);
        }
        t.kp_count[oc] = kept;
        s_kept[i] = kept;
        for (int s = 0; s < kept; ++s) {
            t.kp_peak[oc * S + s] = ids[s];
            t.kp_xy[(oc * S + s) * 2] = pts[s][0];
            t.kp_xy[(oc * S + s) * 2 + 1] = pts[s][1];
            if (stash) { s_kept_xy[((size_t)i * S + s) * 2] = pts[s][0]; s_kept_xy[((size_t)i * S + s) * 2 + 1] = pts[s][1]; }
        }
    }
    __syncwarp();                                       // also makes the kp_xy stores visible to the block (no stash)

    // ---- lift every kept keypoint to 3D, one thread each (pipeline.py:164-171, 189-199) ----
    if (have_camera) {
        for (int i = lane; i < n_obj * C * S; i += THREADS) {
            const int ocl = i / S, s = i - ocl * S;        // ocl = o * C + c inside the frame
            if (s >= s_kept[ocl]) continue;
            const int c = ocl % C;
            double p3[3];
            okp_detection_to_point(kept_xy[((size_t)ocl * S + s) * 2], kept_xy[((size_t)ocl * S + s) * 2 + 1],
                                   depth + (m0 + c) * HW, H, W, cam, prm.compat_clip_bug, p3);
            double* out = t.kp_point + (((size_t)n * O * C + ocl) * S + s) * 3;
            out[0] = p3[0]; out[1] = p3[1]; out[2] = p3[2];
        }
    }
    __syncwarp();
    if (lane == 0) { t.n_objects[n] = n_obj; t.flags[n] = s_flags; }
#undef s_flags
}
