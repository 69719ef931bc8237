// okp_peaks.cuh -- K1: 5x5 box sum + 5x5 max-pool NMS + threshold + ordered compaction +
// sub-pixel centroid (replaces perception/pipeline.py:46-79 and perception/models.py:55-58).
//
// Work decomposition: every map (frame n, channel c) is cut into tiles; one CTA processes one
// tile at a time (persistent grid-stride loop). A tile is staged in shared memory with a 4 px
// halo (2 for the box sum + 2 for the max pool of box sums). Peaks of a tile are written, sorted
// in raster order, to a per-tile slot list; okp_merge_peaks_kernel concatenates the tiles of a map
// into the final table (raster order is what the reference's boolean-mask indexing returns,
// pipeline.py:73).
#pragma once
#include "okp_common.cuh"

// ---------------------------------------------------------------------------------------------
// Generic tile kernel ("v0"): any H, W, any tile size. One thread per pixel, shared-memory taps.
// Slow but obviously right; kept as the in-library cross-check of the tuned kernels and as the
// path for shapes the tuned kernels do not cover.
// ---------------------------------------------------------------------------------------------
struct OkpTileGeometry {
    int H, W;            // map size
    int TH, TW;          // tile size (outputs)
    int tiles_y, tiles_x;
    int maps;            // N * C
    int radius;          // NMS window radius: 2 (5x5, the reference) or 1 (3x3)
    int box_sum;         // 1: the score is the 5x5 box sum (the reference); 0: the map value itself
};

// torch's max_pool2d update rule (`if (val > max || isnan(val)) max = val`): a NaN anywhere in the window makes the
// maximum NaN, so `x == hmax` (perception/models.py:55-58) fails for every pixel within reach of a NaN box sum.
__device__ __forceinline__ float okp_pool_max(float m, float v) { return (v > m || v != v) ? v : m; }

__device__ __forceinline__ void okp_centroid_from_smem(const float* raw, int pitch, int ry, int rx, int gy, int gx,
                                                       float* cx, float* cy, float* conf) {
    // raw(ry, rx) is the peak pixel inside the staged tile; (gy, gx) its image coordinates.
    // Pixels outside the image were staged as +0, which leaves every partial sum unchanged, so
    // summing the full 5x5 window in raster order equals the reference's border-clipped window.
    float sy = 0.0f, sx = 0.0f, sp = 0.0f;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const float q = raw[(ry + dy) * pitch + rx + dx];
            sy = __fadd_rn(sy, __fmul_rn(q, (float)(gy + dy)));
            sx = __fadd_rn(sx, __fmul_rn(q, (float)(gx + dx)));
            sp = __fadd_rn(sp, q);
        }
    }
    *cx = __fdiv_rn(sx, sp);
    *cy = __fdiv_rn(sy, sp);
    *conf = sp;
}

// Processes one (map, tile) work item with the whole CTA; shared memory as laid out by the kernels below.
template <int THREADS, typename T>
__device__ __forceinline__ void okp_generic_tile(const T* __restrict__ heat, const OkpTileGeometry& g, float threshold,
                                                 int K, long long work, long long out_index,
                                                 int32_t* __restrict__ tile_count,
                                                 OkpPeakRecord* __restrict__ tile_peaks, unsigned char* smem_raw,
                                                 int* s_count_ptr) {
    const int RP = g.TW + 8;                 // raw pitch
    const int SP = g.TW + 4;                 // score pitch
    float* raw = reinterpret_cast<float*>(smem_raw);
    float* score = raw + (g.TH + 8) * RP;
    int32_t* keys = reinterpret_cast<int32_t*>(score + (g.TH + 4) * SP);
    int32_t* sorted = keys + K;
    int& s_count = *s_count_ptr;
    const int tiles_per_map = g.tiles_y * g.tiles_x;
    {
        const int map = (int)(work / tiles_per_map);
        const int tile = (int)(work - (long long)map * tiles_per_map);
        const int ty0 = (tile / g.tiles_x) * g.TH;
        const int tx0 = (tile % g.tiles_x) * g.TW;
        const T* src = heat + (size_t)map * g.H * g.W;
        if (threadIdx.x == 0) s_count = 0;
        // ---- stage the tile + 4 px halo, zero outside the image (conv2d zero padding) ----
        for (int i = threadIdx.x; i < (g.TH + 8) * RP; i += THREADS) {
            const int ry = i / RP, rx = i - ry * RP;
            const int gy = ty0 - 4 + ry, gx = tx0 - 4 + rx;
            float v = 0.0f;
            if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) v = okp_ld<T>(src + (size_t)gy * g.W + gx);
            raw[i] = v;
        }
        __syncthreads();
        // ---- score on the tile + 2 px halo: box sum (25 sequential adds, raster tap order) or the value itself ----
        const int r = g.radius;
        for (int i = threadIdx.x; i < (g.TH + 4) * SP; i += THREADS) {
            const int sy = i / SP, sx = i - sy * SP;
            const int gy = ty0 - 2 + sy, gx = tx0 - 2 + sx;
            float acc = -INFINITY;           // max_pool2d pads with -inf
            if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) {
                if (g.box_sum) {
                    acc = 0.0f;
#pragma unroll
                    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 5; ++dx) acc = __fadd_rn(acc, raw[(sy + dy) * RP + sx + dx]);
                } else {
                    acc = raw[(sy + 2) * RP + sx + 2];
                }
            }
            score[i] = acc;
        }
        __syncthreads();
        // ---- NMS + threshold; unordered append ----
        for (int i = threadIdx.x; i < g.TH * g.TW; i += THREADS) {
            const int py = i / g.TW, px = i - py * g.TW;
            const int gy = ty0 + py, gx = tx0 + px;
            if (gy >= g.H || gx >= g.W) continue;
            const float v = score[(py + 2) * SP + px + 2];
            if (!(v > threshold)) continue;
            float m = v;
            for (int dy = 2 - r; dy <= 2 + r; ++dy)
                for (int dx = 2 - r; dx <= 2 + r; ++dx) m = okp_pool_max(m, score[(py + dy) * SP + px + dx]);
            if (v == m) {
                const int slot = atomicAdd(&s_count, 1);
                if (slot < K) keys[slot] = gy * g.W + gx;
            }
        }
        __syncthreads();
        const int total = s_count;
        if (total > K) {
            // Overflow: the table keeps the FIRST K peaks in raster order. Redo the scan in order
            // with one warp (rare path: only maps with more peaks than the table holds).
            __syncthreads();
            if (threadIdx.x < 32) {
                int filled = 0;
                for (int base = 0; base < g.TH * g.TW && filled < K; base += 32) {
                    const int i = base + threadIdx.x;
                    bool is_peak = false;
                    int key = 0;
                    if (i < g.TH * g.TW) {
                        const int py = i / g.TW, px = i - py * g.TW;
                        const int gy = ty0 + py, gx = tx0 + px;
                        if (gy < g.H && gx < g.W) {
                            const float v = score[(py + 2) * SP + px + 2];
                            if (v > threshold) {
                                float m = v;
                                for (int dy = 2 - r; dy <= 2 + r; ++dy)
                                    for (int dx = 2 - r; dx <= 2 + r; ++dx) m = okp_pool_max(m, score[(py + dy) * SP + px + dx]);
                                is_peak = (v == m);
                                key = gy * g.W + gx;
                            }
                        }
                    }
                    const unsigned ballot = __ballot_sync(0xffffffffu, is_peak);
                    const int slot = filled + __popc(ballot & ((1u << threadIdx.x) - 1u));
                    if (is_peak && slot < K) keys[slot] = key;
                    filled += __popc(ballot);
                }
            }
            __syncthreads();
        }
        const int n = total < K ? total : K;
        // ---- order by raster key (rank sort; n is tiny) ----
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const int key = keys[i];
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += (keys[j] < key);
            sorted[rank] = key;
        }
        __syncthreads();
        OkpPeakRecord* out = tile_peaks + (size_t)out_index * K;
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const int key = sorted[i];
            const int gy = key / g.W, gx = key - gy * g.W;
            const int ry = gy - ty0 + 4, rx = gx - tx0 + 4;
            OkpPeakRecord rec;
            rec.key = key;
            rec.score = score[(ry - 2) * SP + rx - 2];
            okp_centroid_from_smem(raw, RP, ry, rx, gy, gx, &rec.cx, &rec.cy, &rec.conf);
            rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
            out[i] = rec;
        }
        if (threadIdx.x == 0) tile_count[out_index] = total;
        __syncthreads();
    }
}

template <int THREADS, typename T>
__global__ void __launch_bounds__(THREADS)
okp_peaks_generic_kernel(const T* __restrict__ heat, OkpTileGeometry g, float threshold, int K,
                         int32_t* __restrict__ tile_count, OkpPeakRecord* __restrict__ tile_peaks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_count;
    const long long work_items = (long long)g.maps * g.tiles_y * g.tiles_x;
    for (long long work = blockIdx.x; work < work_items; work += gridDim.x)
        okp_generic_tile<THREADS, T>(heat, g, threshold, K, work, work, tile_count, tile_peaks, smem_raw, &s_count);
}

// ---------------------------------------------------------------------------------------------
// Merge the per-tile lists of a map into the final raster-ordered table. One warp per map.
// Also resets the assignment columns (peak_object, peak_vote) and clears unused slots, so the
// tables never need a separate memset.
// ---------------------------------------------------------------------------------------------
// m: map (row of the tables); lists: index of the map's first tile list in tile_count / tile_peaks.
__device__ __forceinline__ void okp_merge_map(const int32_t* tile_count, const OkpPeakRecord* tile_peaks, int m, size_t lists,
                                              int lane, int tiles_per_map, int W, int K, const OkpDecodeTables& t) {
    int total = 0;       // true number of peaks
    int kept = 0;        // candidates available (each tile contributes at most K)
    for (int tile = 0; tile < tiles_per_map; ++tile) {
        const int c = tile_count[lists + tile];
        total += c;
        kept += c < K ? c : K;
    }
    const int n_out = kept < K ? kept : K;
    // clear every slot first (lanes stride over K)
    for (int k = lane; k < K; k += 32) {
        const size_t s = (size_t)m * K + k;
        if (k >= n_out) {
            t.peak_yx[2 * s] = -1; t.peak_yx[2 * s + 1] = -1;
            t.peak_score[s] = 0.0f;
            t.peak_xy[2 * s] = 0.0f; t.peak_xy[2 * s + 1] = 0.0f;
            t.peak_conf[s] = 0.0f;
        }
        t.peak_object[s] = -1;
        t.peak_vote[2 * s] = 0.0; t.peak_vote[2 * s + 1] = 0.0;
    }
    // rank of every candidate among all candidates of the map
    for (int tile = 0; tile < tiles_per_map; ++tile) {
        const int c = okp_min(tile_count[lists + tile], K);
        const OkpPeakRecord* mine = tile_peaks + (lists + tile) * K;
        for (int i = lane; i < c; i += 32) {
            const OkpPeakRecord rec = mine[i];
            int rank = 0;
            if (tiles_per_map == 1) {
                rank = i;                            // already sorted inside the tile
            } else {
                for (int other = 0; other < tiles_per_map; ++other) {
                    const int oc = okp_min(tile_count[lists + other], K);
                    const OkpPeakRecord* theirs = tile_peaks + (lists + other) * K;
                    for (int j = 0; j < oc; ++j) rank += (theirs[j].key < rec.key);
                }
            }
            if (rank < K) {
                const size_t s = (size_t)m * K + rank;
                const int y = rec.key / W;
                t.peak_yx[2 * s] = y; t.peak_yx[2 * s + 1] = rec.key - y * W;
                t.peak_score[s] = rec.score;
                t.peak_xy[2 * s] = rec.cx; t.peak_xy[2 * s + 1] = rec.cy;
                t.peak_conf[s] = rec.conf;
            }
        }
    }
    if (lane == 0) t.peak_count[m] = total;
}

// C: maps per frame. The generic path announces itself per frame (OKP_FLAG_GENERIC_PATH; the grouping keeps the bit).
__global__ void __launch_bounds__(128)
okp_merge_peaks_kernel(const int32_t* __restrict__ tile_count, const OkpPeakRecord* __restrict__ tile_peaks,
                       int maps, int tiles_per_map, int W, int K, int C, OkpDecodeTables t) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= maps) return;
    okp_merge_map(tile_count, tile_peaks, warp, (size_t)warp * tiles_per_map, threadIdx.x & 31, tiles_per_map, W, K, t);
    if ((threadIdx.x & 31) == 0 && t.flags && warp % C == 0) t.flags[warp / C] = OKP_FLAG_GENERIC_PATH;
}

// ---------------------------------------------------------------------------------------------
// Overflow path of the stream kernel (okp_peaks_stream.cuh): maps with more peaks than the table
// holds need the FIRST K in raster order (and the true count), maps with negative / NaN values the exact
// arithmetic. Every CTA owns a contiguous slice of `per_cta` maps: one coalesced look at their peak_count;
// the maps of the slice that overflowed are redone tile by tile with the generic routine (tile lists in the
// CTA's own slice of the workspace) and merged by warp 0. With no overflow the kernel is that one look.
// The slices are small (maps / min(maps, OKP_OVERFLOW_CTAS)), so a batch in which EVERY map overflows -- an
// untrained network, BASELINE config 5 -- is spread over the whole GPU (round 1 used 256-map slices: 768 such
// maps ran on three CTAs).
// ---------------------------------------------------------------------------------------------
#define OKP_OVERFLOW_CTAS 592                 // 4 per SM

template <int THREADS, typename T>
__global__ void __launch_bounds__(THREADS)
okp_peaks_overflow_kernel(const T* __restrict__ heat, OkpTileGeometry g, float threshold, int K, int per_cta,
                          int32_t* __restrict__ tile_count, OkpPeakRecord* __restrict__ tile_peaks, OkpDecodeTables t) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_count;
    __shared__ int s_over[THREADS];
    okp_wait_for_predecessor();                           // launched as a programmatic dependent of the peak kernel
    const int tiles_per_map = g.tiles_y * g.tiles_x;
    const int first = blockIdx.x * per_cta;
    const int last = first + per_cta < g.maps ? first + per_cta : g.maps;
    for (int base = first; base < last; base += THREADS) {
        const int mine = base + threadIdx.x;
        const int over = (mine < last && t.peak_count[mine] > K) ? 1 : 0;
        s_over[threadIdx.x] = over;
        if (!__syncthreads_or(over)) continue;
        for (int i = 0; i < THREADS && base + i < last; ++i) {
            if (!s_over[i]) continue;                     // uniform: shared flag
            const int m = base + i;
            for (int tile = 0; tile < tiles_per_map; ++tile)
                okp_generic_tile<THREADS, T>(heat, g, threshold, K, (long long)m * tiles_per_map + tile,
                                          (long long)blockIdx.x * tiles_per_map + tile, tile_count, tile_peaks, smem_raw,
                                          &s_count);
            __threadfence_block();
            __syncthreads();
            if (threadIdx.x < 32)
                okp_merge_map(tile_count, tile_peaks, m, (size_t)blockIdx.x * tiles_per_map, threadIdx.x, tiles_per_map, g.W, K, t);
            __syncthreads();
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Per-type top-k (OkpDecodeParams.top_k > 0; BASELINE.json: "per-type top-k"; CornerNet's _topk,
// perception/corner_net_lite/core/models/py_utils/utils.py:27-38, applied to each map): the raster-ordered
// table of a map is re-ordered by score, descending (ties: raster order), and cut to k rows. One warp per
// map; the rows (at most K <= 256) are ranked by counting, written back, the tail cleared. A map whose table
// overflowed (peak_count > K) keeps its count so that OKP_FLAG_PEAK_OVERFLOW is still raised.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
okp_topk_kernel(int maps, int K, int top_k, OkpDecodeTables t) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= maps) return;
    __shared__ float s_score[4][OKP_MAX_PEAKS], s_conf[4][OKP_MAX_PEAKS], s_xy[4][OKP_MAX_PEAKS][2];
    __shared__ int s_yx[4][OKP_MAX_PEAKS][2];
    const int w = threadIdx.x >> 5;
    const int total = t.peak_count[warp];
    const int n = total < K ? total : K;
    const size_t base = (size_t)warp * K;
    for (int i = lane; i < n; i += 32) {
        s_score[w][i] = t.peak_score[base + i];
        s_conf[w][i] = t.peak_conf[base + i];
        s_xy[w][i][0] = t.peak_xy[2 * (base + i)]; s_xy[w][i][1] = t.peak_xy[2 * (base + i) + 1];
        s_yx[w][i][0] = t.peak_yx[2 * (base + i)]; s_yx[w][i][1] = t.peak_yx[2 * (base + i) + 1];
    }
    __syncwarp();
    const int keep = n < top_k ? n : top_k;
    for (int i = lane; i < n; i += 32) {
        const float mine = s_score[w][i];
        int rank = 0;                        // rows are in raster order: an equal score further left wins
        for (int j = 0; j < n; ++j) rank += (s_score[w][j] > mine) || (s_score[w][j] == mine && j < i);
        if (rank < keep) {
            t.peak_score[base + rank] = mine;
            t.peak_conf[base + rank] = s_conf[w][i];
            t.peak_xy[2 * (base + rank)] = s_xy[w][i][0]; t.peak_xy[2 * (base + rank) + 1] = s_xy[w][i][1];
            t.peak_yx[2 * (base + rank)] = s_yx[w][i][0]; t.peak_yx[2 * (base + rank) + 1] = s_yx[w][i][1];
        }
    }
    for (int i = keep + lane; i < n; i += 32) {
        t.peak_score[base + i] = 0.0f; t.peak_conf[base + i] = 0.0f;
        t.peak_xy[2 * (base + i)] = 0.0f; t.peak_xy[2 * (base + i) + 1] = 0.0f;
        t.peak_yx[2 * (base + i)] = -1; t.peak_yx[2 * (base + i) + 1] = -1;
    }
    if (lane == 0 && total <= K) t.peak_count[warp] = keep;
}
