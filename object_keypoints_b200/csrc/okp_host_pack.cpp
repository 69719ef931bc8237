// okp_host_pack.cpp -- HOST side of the sparse heatmap transfer (see csrc/okp_sparse.cuh for the argument why
// dropping empty regions leaves every table bit-identical). Plain C++ with OpenMP, compiled by g++ and linked into
// libokp.so: the pass has to run at memory speed (it reads every heatmap byte once), otherwise it would be slower than
// the PCIe copy it replaces. On x86-64 the marking loop has an AVX2 form that is selected at RUN time
// (__builtin_cpu_supports); every other host (aarch64 Grace, x86 without AVX2) runs the portable loop, which g++
// vectorises for the build target. The file is compiled without -mavx2 so nothing outside the guarded function uses it.
#if defined(__x86_64__) || defined(_M_X64)
#define OKP_HOST_X86 1
#include <immintrin.h>
#else
#define OKP_HOST_X86 0
#endif
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include "../../include/okp.h"

#define OKP_TILE_H 4
#define OKP_TILE_W 16
#define OKP_TILE_FLOATS (OKP_TILE_H * OKP_TILE_W)

static inline int tiles_y(int H) { return (H + OKP_TILE_H - 1) / OKP_TILE_H; }
static inline int tiles_x(int W) { return (W + OKP_TILE_W - 1) / OKP_TILE_W; }

// One full-height (4-row) band of 16-wide tiles: marks[tx] = 1 if the tile holds a value that is NOT <= tau (so NaN
// counts as active: it must reach the device unchanged).
static void mark_band_portable(const float* base, int W, int full, float tau, unsigned char* mrow) {
    for (int tx = 0; tx < full; ++tx) {
        unsigned char any = 0;
        for (int r = 0; r < OKP_TILE_H; ++r) {
            const float* row = base + (size_t)r * W + tx * OKP_TILE_W;
            for (int c = 0; c < OKP_TILE_W; ++c) any |= (unsigned char)!(row[c] <= tau);
        }
        mrow[tx] = any;
    }
}

#if OKP_HOST_X86
__attribute__((target("avx2")))
static void mark_band_avx2(const float* base, int W, int full, float tau, unsigned char* mrow) {
    const __m256 vtau = _mm256_set1_ps(tau);
    const float *r0 = base, *r1 = base + W, *r2 = base + 2 * (size_t)W, *r3 = base + 3 * (size_t)W;
    for (int tx = 0; tx < full; ++tx) {                    // four row streams, one mark store per tile
        const int x = tx * OKP_TILE_W;
        __m256 hit = _mm256_or_ps(_mm256_cmp_ps(_mm256_loadu_ps(r0 + x), vtau, _CMP_NLE_UQ),
                                  _mm256_cmp_ps(_mm256_loadu_ps(r0 + x + 8), vtau, _CMP_NLE_UQ));
        hit = _mm256_or_ps(hit, _mm256_or_ps(_mm256_cmp_ps(_mm256_loadu_ps(r1 + x), vtau, _CMP_NLE_UQ),
                                             _mm256_cmp_ps(_mm256_loadu_ps(r1 + x + 8), vtau, _CMP_NLE_UQ)));
        hit = _mm256_or_ps(hit, _mm256_or_ps(_mm256_cmp_ps(_mm256_loadu_ps(r2 + x), vtau, _CMP_NLE_UQ),
                                             _mm256_cmp_ps(_mm256_loadu_ps(r2 + x + 8), vtau, _CMP_NLE_UQ)));
        hit = _mm256_or_ps(hit, _mm256_or_ps(_mm256_cmp_ps(_mm256_loadu_ps(r3 + x), vtau, _CMP_NLE_UQ),
                                             _mm256_cmp_ps(_mm256_loadu_ps(r3 + x + 8), vtau, _CMP_NLE_UQ)));
        mrow[tx] = (unsigned char)(_mm256_movemask_ps(hit) != 0);
    }
}
#endif

typedef void (*MarkBandFn)(const float*, int, int, float, unsigned char*);

static MarkBandFn pick_mark_band() {
#if OKP_HOST_X86
    if (__builtin_cpu_supports("avx2")) return mark_band_avx2;
#endif
    return mark_band_portable;
}

// Row-wise: one pass over the map in memory order.
static void mark_active_tiles(const float* map, int H, int W, float tau, unsigned char* marks, int TX, MarkBandFn band) {
    const int full = W / OKP_TILE_W;                       // tile columns that are 16 wide
    const int TY = tiles_y(H);
    for (int ty = 0; ty < TY; ++ty) {
        const int y0 = ty * OKP_TILE_H, rows = y0 + OKP_TILE_H <= H ? OKP_TILE_H : H - y0;
        const float* base = map + (size_t)y0 * W;
        unsigned char* mrow = marks + (size_t)ty * TX;
        if (rows == OKP_TILE_H) {
            band(base, W, full, tau, mrow);
        } else {
            for (int tx = 0; tx < full; ++tx) {
                unsigned char any = 0;
                for (int r = 0; r < rows; ++r)
                    for (int c = 0; c < OKP_TILE_W; ++c) any |= (unsigned char)!(base[(size_t)r * W + tx * OKP_TILE_W + c] <= tau);
                mrow[tx] = any;
            }
        }
        if (full < TX) {
            unsigned char any = 0;
            for (int r = 0; r < rows; ++r)
                for (int x = full * OKP_TILE_W; x < W; ++x) any |= (unsigned char)!(base[(size_t)r * W + x] <= tau);
            mrow[full] = any;
        }
    }
}

extern "C" size_t okp_host_pack_scratch_bytes(int maps, int H, int W) {
    if (maps < 0 || H < 1 || W < 1) return 0;
    return (size_t)maps * 2 * tiles_y(H) * tiles_x(W) + 64;
}

extern "C" int okp_host_pack_tiles_f32(const float* heat_host, int maps, int H, int W, float threshold,
                                       unsigned char* scratch_host, long long* map_offsets_host,
                                       int32_t* tile_ids_host, float* packed_host, long long capacity_tiles,
                                       long long* n_tiles_out, int threads) {
    if (maps < 0 || H < 1 || W < 1) return OKP_E_SHAPE;
    if (!n_tiles_out) return OKP_E_NULL;
    *n_tiles_out = 0;
    if (maps == 0) return OKP_OK;
    if (!heat_host || !scratch_host || !map_offsets_host || !tile_ids_host || !packed_host) return OKP_E_NULL;
    const int TY = tiles_y(H), TX = tiles_x(W), tiles = TY * TX;
    if ((long long)maps * tiles > 0x7fffffffLL) return OKP_E_SHAPE;
    // a box sum above the threshold needs a value above threshold / 25; the slack covers the 24 float32 roundings
    const float tau = threshold > 0.0f ? threshold / 25.0f * (1.0f - 1e-5f) : -INFINITY;
    if (threads <= 0) threads = omp_get_max_threads();
    const MarkBandFn band = pick_mark_band();
    // pass 1: per map, activity marks -> marks widened by one tile in every direction, and their count
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int m = 0; m < maps; ++m) {
        unsigned char* raw = scratch_host + (size_t)m * 2 * tiles;
        unsigned char* wide = raw + tiles;
        mark_active_tiles(heat_host + (size_t)m * H * W, H, W, tau, raw, TX, band);
        long long count = 0;
        for (int ty = 0; ty < TY; ++ty)
            for (int tx = 0; tx < TX; ++tx) {
                unsigned char any = 0;
                for (int y = ty > 0 ? ty - 1 : 0; y <= ty + 1 && y < TY; ++y)
                    for (int x = tx > 0 ? tx - 1 : 0; x <= tx + 1 && x < TX; ++x) any |= raw[y * TX + x];
                wide[ty * TX + tx] = any;
                count += any;
            }
        map_offsets_host[m + 1] = count;
    }
    map_offsets_host[0] = 0;
    for (int m = 0; m < maps; ++m) map_offsets_host[m + 1] += map_offsets_host[m];
    const long long total = map_offsets_host[maps];
    *n_tiles_out = total;
    if (total > capacity_tiles) return OKP_OK;            // the caller sees n_tiles > capacity and sends the maps densely
    // pass 2: pack the marked tiles (positions beyond the map's edge are filled with +0)
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int m = 0; m < maps; ++m) {
        const unsigned char* wide = scratch_host + (size_t)m * 2 * tiles + tiles;
        const float* map = heat_host + (size_t)m * H * W;
        long long at = map_offsets_host[m];
        for (int t = 0; t < tiles; ++t) {
            if (!wide[t]) continue;
            const int ty = t / TX, tx = t - ty * TX;
            const int y0 = ty * OKP_TILE_H, x0 = tx * OKP_TILE_W;
            float* dst = packed_host + at * OKP_TILE_FLOATS;
            if (y0 + OKP_TILE_H <= H && x0 + OKP_TILE_W <= W) {
                for (int r = 0; r < OKP_TILE_H; ++r)                 // 64 bytes per row: the compiler emits vector moves
                    memcpy(dst + r * OKP_TILE_W, map + (size_t)(y0 + r) * W + x0, OKP_TILE_W * sizeof(float));
            } else {
                for (int r = 0; r < OKP_TILE_H; ++r)
                    for (int c = 0; c < OKP_TILE_W; ++c)
                        dst[r * OKP_TILE_W + c] = (y0 + r < H && x0 + c < W) ? map[(size_t)(y0 + r) * W + x0 + c] : 0.0f;
            }
            tile_ids_host[at] = m * tiles + t;
            ++at;
        }
    }
    return OKP_OK;
}
