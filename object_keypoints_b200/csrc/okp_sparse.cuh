// okp_sparse.cuh -- sparse host -> device transfer of heatmaps for the end-to-end (HOST buffer) form of the
// decode: okp_host_pack_tiles_f32 (host, OpenMP + AVX2, csrc/okp_host_pack.cpp) + okp_scatter_tiles_kernel (device).
//
// The end-to-end path is bound by PCIe: every float32 heatmap byte has to cross the bus (691 KB per 180x320 valve
// frame, 54 GB/s -> 78 k frames/s) while K1 could take 7 M frames/s. A network's heatmaps are almost empty, and
// empty regions cannot influence the result:
//   * a pixel whose 5x5 box sum exceeds the threshold has a value above threshold / 25 in its window, so every
//     pixel that can be a peak -- or can beat one in the NMS comparison -- lies within 2 px of an ACTIVE pixel
//     (value > tau = threshold / 25, minus a rounding slack), and its window within 4 px of it;
//   * a pixel with no active pixel in its window has a box sum <= 25 tau (1 + gamma) < threshold whatever the
//     inactive values are (also when they are replaced by zero, also when they are negative), so it neither is a
//     peak nor beats one.
// Hence a map in which everything farther than 4 px (Chebyshev) from all active pixels is replaced by +0 decodes
// to bit-identical tables. The host pass marks tiles of 4 x 16 pixels (one 64-byte line per row) that hold an
// active pixel, dilates the marks by one tile in every direction (>= 4 px), and packs the marked tiles; only those
// cross PCIe; the device scatters them into a zeroed dense map and the unchanged kernels run on it.
// Dense inputs (an untrained network: every pixel ~0.5) mark every tile: the caller then copies the map as it is.
#pragma once
#include "okp_common.cuh"

#define OKP_TILE_H 4
#define OKP_TILE_W 16
#define OKP_TILE_FLOATS (OKP_TILE_H * OKP_TILE_W)

static inline int okp_tiles_y(int H) { return (H + OKP_TILE_H - 1) / OKP_TILE_H; }
static inline int okp_tiles_x(int W) { return (W + OKP_TILE_W - 1) / OKP_TILE_W; }

// One thread per float4 of a packed tile: 16 threads per tile, 4 consecutive columns of one tile row each.
__global__ void __launch_bounds__(256)
okp_scatter_tiles_kernel(const float* __restrict__ packed, const int32_t* __restrict__ tile_ids, long long n_tiles,
                         int H, int W, int TX, int tiles, float* __restrict__ heat) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * 16) return;
    const long long tile = i >> 4;
    const int part = (int)(i & 15), r = part >> 2, c4 = (part & 3) * 4;
    const int id = tile_ids[tile];
    const int m = id / tiles, t = id - m * tiles, ty = t / TX, tx = t - ty * TX;
    const int y = ty * OKP_TILE_H + r, x = tx * OKP_TILE_W + c4;
    if (y >= H || x >= W) return;
    const float4 v = *reinterpret_cast<const float4*>(packed + tile * OKP_TILE_FLOATS + r * OKP_TILE_W + c4);
    float* dst = heat + ((size_t)m * H + y) * W + x;
    if (x + 3 < W && ((uintptr_t)dst & 15u) == 0) {
        *reinterpret_cast<float4*>(dst) = v;
    } else {
        dst[0] = v.x;
        if (x + 1 < W) dst[1] = v.y;
        if (x + 2 < W) dst[2] = v.z;
        if (x + 3 < W) dst[3] = v.w;
    }
}
