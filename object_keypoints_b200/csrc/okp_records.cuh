// okp_records.cuh -- the multi-GPU exchange step: frames shard independently, the
// only cross-GPU traffic is the gather of the per-frame 3D keypoint records.
//
// A record is R = 2 + O*C + O*C*S*3 float64 per frame: n_objects, flags, kp_count[O,C], kp_point[O,C,S,3]
// (the part of OkpDecodeTables a consumer of the pipeline reads; the reference returns it as the list of
// object dicts of ObjectKeypointPipeline.__call__, perception/pipeline.py:195-199).
//
// One kernel packs the rank's records and stores them at the rank's rows of EVERY destination buffer it is
// given. With one destination (the local send buffer) it is the pack step in front of an NCCL all_gather;
// with `world` destinations that are peer-mapped buffers of the other GPUs it IS the all_gather: the stores
// travel over NVLink / NVSwitch while the block keeps packing (no staging copy, no second kernel).
#pragma once
#include "okp_common.cuh"

#define OKP_MAX_PEERS 16

struct OkpPeerBuffers { double* dst[OKP_MAX_PEERS]; };

// 128 threads x <= 32 registers and no shared memory: a block fits into the 4096 registers per SM that two resident
// CTAs of the persistent K1 (okp_peaks_stream_kernel) leave free, so the exchange of step k really runs beside the
// decode of step k + 1 instead of waiting for an SM to drain.
__global__ void __launch_bounds__(128, 16)
okp_pack_records_kernel(const int32_t* __restrict__ n_objects, const uint32_t* __restrict__ flags,
                        const int32_t* __restrict__ kp_count, const double* __restrict__ kp_point, int N, int n_count,
                        int n_point, long long first_row, int n_dst, OkpPeerBuffers peers) {
    const int R = 2 + n_count + n_point;
    const long long total = (long long)N * R;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / R), j = (int)(i - (long long)n * R);
        double v;
        if (j == 0) v = (double)n_objects[n];
        else if (j == 1) v = (double)flags[n];
        else if (j < 2 + n_count) v = (double)kp_count[(size_t)n * n_count + (j - 2)];
        else v = kp_point[(size_t)n * n_point + (j - 2 - n_count)];
        const long long at = first_row * R + i;
#pragma unroll
        for (int d = 0; d < OKP_MAX_PEERS; ++d)          // constant indices: the pointers stay in the parameter bank
            if (d < n_dst) peers.dst[d][at] = v;
    }
}
