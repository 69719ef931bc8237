// okp_peaks_strip.cuh -- tuned K1 (placeholder until the strip kernel lands: the plan function
// declines every shape, so the generic tile kernel runs).
#pragma once
#include "okp_common.cuh"
#include "okp_peaks.cuh"

static inline bool okp_strip_plan(int maps, int H, int W, int K, OkpTileGeometry* geo, size_t* smem_bytes) {
    (void)maps; (void)H; (void)W; (void)K; (void)geo; (void)smem_bytes;
    return false;
}

static inline int okp_strip_launch(const float* heat, OkpTileGeometry geo, float threshold, int K, int32_t* tile_count,
                                   OkpPeakRecord* tile_peaks, size_t smem_bytes, cudaStream_t stream) {
    (void)heat; (void)geo; (void)threshold; (void)K; (void)tile_count; (void)tile_peaks; (void)smem_bytes; (void)stream;
    return OKP_E_UNSUPPORTED;
}
