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
//
// Compact records (round 2). The float64 record above is 3.9 KB per valve frame, three quarters of it padding, and
// packing it re-reads kp_point from HBM in a second kernel that competes with the next step's decode. The grouping
// (okp_group_frame) therefore emits a COMPACT record itself, the moment a value is known, straight into the sink
// buffers it is given -- local memory, or the peer-mapped buffer of the gathering rank, in which case the stores
// travel over NVLink while the kernel works and there is no separate exchange kernel at all:
//     int32  n_objects; uint32 flags;
//     int32  kp_count[O][C]                      (rows o < n_objects are written)
//     double point[O][P][3], P = 1 + sum(keypoint_config): object o's points in (map, slot) order, i.e. without the
//                                                [C][S] padding of the table (slots s < kp_count[o][c] are written)
// = 8 + 4 O C (rounded up to 8) + 24 O P bytes of STRIDE (2120 for valve with O = 16), of which a frame with 8 complete
// valve objects writes 8 + 96 + 960 = 1064. Unwritten parts keep whatever the buffer held: readers go by the counts
// (sharding.unpack_compact_records).
#pragma once
#include "okp_common.cuh"

#define OKP_MAX_PEERS 16

struct OkpRecordSinks {
    unsigned char* base[OKP_MAX_PEERS];   // [rows, stride] byte buffers
    long long first_row;                  // row of frame 0 of this call
    int n;                                // sinks in use (0: records are not emitted)
    int stride;                           // bytes per record, >= okp_record_bytes()
    int points_offset;                    // 8 + round_up(4 O C, 8)
};

static inline int okp_compact_points_offset(int O, int C) { return 8 + (4 * O * C + 7) / 8 * 8; }
static inline int okp_compact_record_bytes(int O, int C, int P) { return okp_compact_points_offset(O, C) + 24 * O * P; }

// The sink loops run over constant indices so that the pointers stay in the parameter bank.
__device__ __forceinline__ void okp_record_header(const OkpRecordSinks& k, int n, int n_objects, unsigned int flags) {
    const size_t at = (size_t)(k.first_row + n) * k.stride;
#pragma unroll
    for (int d = 0; d < OKP_MAX_PEERS; ++d)
        if (d < k.n) *reinterpret_cast<int2*>(k.base[d] + at) = make_int2(n_objects, (int)flags);
}
__device__ __forceinline__ void okp_record_count(const OkpRecordSinks& k, int n, int O, int C, int o, int c, int count) {
    const size_t at = (size_t)(k.first_row + n) * k.stride + 8 + 4 * (size_t)(o * C + c);
#pragma unroll
    for (int d = 0; d < OKP_MAX_PEERS; ++d)
        if (d < k.n) *reinterpret_cast<int32_t*>(k.base[d] + at) = count;
}
__device__ __forceinline__ void okp_record_point(const OkpRecordSinks& k, int n, int O, int C, int P, const OkpConfig& config,
                                                 int o, int c, int s, const double* p3) {
    if (k.n == 0) return;
    int slot = s;
    for (int i = 0; i < c; ++i) slot += config.cfg[i];
    const size_t at = (size_t)(k.first_row + n) * k.stride + k.points_offset + 24 * (size_t)(o * P + slot);
#pragma unroll
    for (int d = 0; d < OKP_MAX_PEERS; ++d)
        if (d < k.n) {
            double* dst = reinterpret_cast<double*>(k.base[d] + at);
            dst[0] = p3[0]; dst[1] = p3[1]; dst[2] = p3[2];
        }
}

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
