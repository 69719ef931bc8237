// okp_peaks_strip.cuh -- tuned K1 for sm_100a: TMA-fed, register-window box sum.
//
// Replaces perception/pipeline.py:46-79 + perception/models.py:55-58 for every map of a batch.
//
// Why this shape. The box sum has to be the 25 float32 additions of the reference in raster tap
// order (torch's CPU conv2d is bitwise that), so it cannot be made separable
// or reassociated: 24 dependent-order FADDs per pixel. At B200's measured FADD rate
// (profiles/r01_fp32_issue_microbench.txt: ~36.7e12 adds/s) that is 1.53e12 pixel/s = 6.1 TB/s of
// heatmap bytes, i.e. the kernel sits just under the HBM roofline and is bound by FADD *issue*.
// Everything else is therefore organised to cost as few issue slots as possible:
//
//   * heatmap rows arrive in shared memory by TMA (cp.async.bulk.tensor, one elected thread, no
//     per-thread load/store instructions); the tensor map's out-of-bounds zero fill IS conv2d's zero
//     padding (rows -2,-1,H,H+1 and columns -2,-1,W,W+1 cost nothing). TMA wants the innermost start
//     coordinate 16-byte aligned (tools/microbench/tma_probe.cu: column -2 is an illegal instruction),
//     so the box starts at column -4 and the strips are shifted instead: strip s produces the pixels
//     x = 4s-2 .. 4s+1, whose window is the columns 4s-4 .. 4s+3 = two aligned LDS.128 (the price is
//     one extra strip per row: 81 instead of 80 at W = 320);
//   * a thread owns a 4-pixel-wide column strip of one map and slides down it, keeping the 5x8
//     window in registers: per row step 2 LDS.128, 96 FADD, 1 STS.128 (box sums to a ring), 3 FMNMX
//     and one compare -- no halo is recomputed, no row is read twice;
//   * pixels whose box sum exceeds the threshold are rare: they set a bit in a row-indexed bitmap
//     ring; after each 5-row batch (one __syncthreads) the set bits are tested against their 5x5
//     neighbourhood in the box-sum ring (nearest neighbours first, early exit), peaks get their
//     centroid from L2 and are appended to the map's list, which is sorted by raster key and
//     written to the final tables at the end (no separate merge pass, no workspace traffic).
//
// A CTA owns M whole maps (M x W/4 threads). Maps whose peak count exceeds the table capacity K
// need "the first K in raster order": they are redone by the generic kernels (okp_peaks.cuh),
// which skip every other map.
#pragma once
#include <cuda.h>

#include "okp_common.cuh"
#include "okp_peaks.cuh"

#define OKP_STRIP_RB 5            // rows per batch (= register window depth, so slots are static)
#define OKP_STRIP_NS 2            // TMA stages
#define OKP_STRIP_SR 16           // ring depth (rows) of the box-sum and bitmap rings, >= 2 RB + 4

struct OkpStripPlan {
    int H, W, maps;
    int SW;                       // pitch of the box-sum ring: W + 4, column index = x + 2
    int strips;                   // W / 4 + 1 (strip s = pixels 4s-2 .. 4s+1)
    int half_strips;              // strips served by TMA box 0 (all of them when halves == 1)
    int halves;                   // 1 or 2 TMA boxes per row (box width <= 256 elements)
    int BW;                       // box width in floats
    int M;                        // maps per CTA
    int nb;                       // batches
    int K;                        // table capacity per map
    int wpr;                      // bitmap words per row (bit index = x + 2)
    int threads;
    int half_bytes;               // bytes of one TMA box: M * RB * BW * 4
    int half_stride;              // half_bytes rounded up to 128 (TMA destinations are 128-byte aligned)
    int stage_bytes;              // halves * half_stride
    int off_score, off_bitmap, off_list, off_count, off_mbar;
    int smem_bytes;
    int grid;
};

struct OkpStripPeak { int32_t key; float score, cx, cy, conf; };

__device__ __forceinline__ uint32_t okp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void okp_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(okp_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void okp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(okp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void okp_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OKP_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OKP_DONE_%=;\n"
        "bra OKP_WAIT_%=;\n"
        "OKP_DONE_%=:\n"
        "}\n" ::"r"(okp_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void okp_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(okp_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(okp_smem_u32(bar)) : "memory");
}

// 24 neighbours of the 5x5 window, nearest first (most non-peaks are rejected by the first few).
__constant__ signed char okp_nms_order[24][2] = {
    {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1},
    {0, -2}, {0, 2}, {-2, 0}, {2, 0}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2},
    {-2, -1}, {-2, 1}, {2, -1}, {2, 1}, {-2, -2}, {-2, 2}, {2, -2}, {2, 2}};

// One row step of the sliding window. I = step inside the batch = register slot of the new row.
// Column c of the strip is pixel x = xs + c - 2 (xs = 4 * strip); vmask has bit c set when that pixel
// is inside the image.
template <int I>
__device__ __forceinline__ void okp_strip_step(float (&w)[5][8], const unsigned char* raw_row, int y, int H, int SW,
                                               float threshold, float* score_map, uint32_t* bitmap_map, int wpr, int xs,
                                               uint32_t vmask) {
    const float4* rp = reinterpret_cast<const float4*>(raw_row);
    const float4 lo = rp[0], hi = rp[1];
    w[I][0] = lo.x; w[I][1] = lo.y; w[I][2] = lo.z; w[I][3] = lo.w;
    w[I][4] = hi.x; w[I][5] = hi.y; w[I][6] = hi.z; w[I][7] = hi.w;
    if (y < 0 || y >= H) return;                                    // uniform over the CTA
    constexpr int R0 = (I + 1) % 5, R1 = (I + 2) % 5, R2 = (I + 3) % 5, R3 = (I + 4) % 5, R4 = I;
    float a[4];
    // raster tap order: row y-2 first (0 + a00 is a00), then rows y-1 .. y+2, left to right
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] = w[R0][c];
#pragma unroll
    for (int d = 1; d < 5; ++d)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = __fadd_rn(a[c], w[R0][c + d]);
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = __fadd_rn(a[c], w[R1][c + d]);
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = __fadd_rn(a[c], w[R2][c + d]);
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = __fadd_rn(a[c], w[R3][c + d]);
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = __fadd_rn(a[c], w[R4][c + d]);
    const int ring = y & (OKP_STRIP_SR - 1);
    *reinterpret_cast<float4*>(score_map + (size_t)ring * SW + xs) = make_float4(a[0], a[1], a[2], a[3]);
    if (fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3])) > threshold) {  // rare
        uint32_t bits = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            bool cand = ((vmask >> c) & 1u) && a[c] > threshold;
#pragma unroll
            for (int o = 0; o < 4; ++o)                              // same-row neighbours this thread already knows
                if (o != c && (o - c) <= 2 && (c - o) <= 2) cand = cand && (!((vmask >> o) & 1u) || a[c] >= a[o]);
            bits |= cand ? (1u << c) : 0u;
        }
        if (bits) atomicOr(bitmap_map + ring * wpr + (xs >> 5), bits << (xs & 31));
    }
}

__global__ void __launch_bounds__(512)
okp_peaks_strip_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ heat, OkpStripPlan p,
                       float threshold, OkpDecodeTables t) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RB = OKP_STRIP_RB, NS = OKP_STRIP_NS, SR = OKP_STRIP_SR;
    float* score = reinterpret_cast<float*>(smem + p.off_score);             // [M][SR][SW], column x + 2
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem + p.off_bitmap);     // [M][SR][wpr]
    OkpStripPeak* list = reinterpret_cast<OkpStripPeak*>(smem + p.off_list); // [M][K]
    int* count = reinterpret_cast<int*>(smem + p.off_count);                 // [M]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_mbar);         // [NS]

    const int tid = threadIdx.x;
    const int mm = tid / p.strips;                       // map slot inside the CTA
    const int s = tid - mm * p.strips;                   // strip inside the map
    const int half = s >= p.half_strips ? 1 : 0;
    const int xs = 4 * s;                                // pixels xs-2 .. xs+1
    const int H = p.H, W = p.W, SW = p.SW;
    uint32_t vmask = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) vmask |= (xs + c - 2 >= 0 && xs + c - 2 < W) ? (1u << c) : 0u;
    const int first_map = blockIdx.x * p.M;

    for (int i = tid; i < p.M * SR * p.wpr; i += blockDim.x) bitmap[i] = 0;
    if (tid < p.M) count[tid] = 0;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) okp_mbar_init(full + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const CUtensorMap* tmap_ptr = &tmap;                 // address of the __grid_constant__ parameter itself
    auto issue = [=](int b) {                            // one thread: TMA the new rows of batch b
        uint64_t* bar = full + (b % NS);
        unsigned char* dst = smem + (size_t)(b % NS) * p.stage_bytes;
        okp_mbar_expect_tx(bar, (uint32_t)(p.halves * p.half_bytes));
        okp_tma_load_3d(dst, tmap_ptr, -4, b * RB - 2, first_map, bar);
        if (p.halves == 2) okp_tma_load_3d(dst + p.half_stride, tmap_ptr, 4 * p.half_strips - 4, b * RB - 2, first_map, bar);
    };
    if (tid == 0) {
        for (int b = 0; b < NS && b < p.nb; ++b) issue(b);
    }

    // this thread's window row inside a stage: box [M][RB][BW], first column 4 * (s - half * half_strips)
    const int thread_raw = half * p.half_stride + (mm * RB * p.BW + 4 * (s - half * p.half_strips)) * 4;
    const int row_pitch = p.BW * 4;
    float* score_map = score + (size_t)mm * SR * SW;
    uint32_t* bitmap_map = bitmap + (size_t)mm * SR * p.wpr;

    float w[5][8];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) w[i][j] = 0.0f;

    for (int b = 0; b < p.nb; ++b) {
        okp_mbar_wait(full + (b % NS), (uint32_t)((b / NS) & 1));
        const unsigned char* raw = smem + (size_t)(b % NS) * p.stage_bytes + thread_raw;
        const int y0 = b * RB - 4;                       // new row of step i is y0 + i + 2
        okp_strip_step<0>(w, raw + 0 * row_pitch, y0 + 0, H, SW, threshold, score_map, bitmap_map, p.wpr, xs, vmask);
        okp_strip_step<1>(w, raw + 1 * row_pitch, y0 + 1, H, SW, threshold, score_map, bitmap_map, p.wpr, xs, vmask);
        okp_strip_step<2>(w, raw + 2 * row_pitch, y0 + 2, H, SW, threshold, score_map, bitmap_map, p.wpr, xs, vmask);
        okp_strip_step<3>(w, raw + 3 * row_pitch, y0 + 3, H, SW, threshold, score_map, bitmap_map, p.wpr, xs, vmask);
        okp_strip_step<4>(w, raw + 4 * row_pitch, y0 + 4, H, SW, threshold, score_map, bitmap_map, p.wpr, xs, vmask);
        __syncthreads();                                 // batch b's box sums and bits are visible; its stage is free
        if (tid == 0 && b + NS < p.nb) issue(b + NS);

        // ---- NMS of the rows whose 5x5 neighbourhood is now complete: [y0 - 2, y0 + 3) ----
        const int words = p.M * RB * p.wpr;
        for (int idx = tid; idx < words; idx += blockDim.x) {
            const int m2 = idx / (RB * p.wpr);
            const int rem = idx - m2 * (RB * p.wpr);
            const int rr = rem / p.wpr, wi = rem - rr * p.wpr;
            const int r = y0 - 2 + rr;
            if (r < 0 || r >= H) continue;
            uint32_t* word = bitmap + ((size_t)m2 * SR + (r & (SR - 1))) * p.wpr + wi;
            uint32_t bits = *word;
            if (!bits) continue;
            *word = 0;
            const float* sm = score + (size_t)m2 * SR * SW + 2;     // sm[row * SW + x]
            while (bits) {
                const int c = __ffs(bits) - 1;
                bits &= bits - 1;
                const int x = wi * 32 + c - 2;
                const float v = sm[(r & (SR - 1)) * SW + x];
                bool peak = true;
#pragma unroll 1
                for (int k = 0; k < 24; ++k) {
                    const int ny = r + okp_nms_order[k][0], nx = x + okp_nms_order[k][1];
                    if (ny < 0 || ny >= H || nx < 0 || nx >= W) continue;        // max_pool2d pads with -inf
                    if (sm[(ny & (SR - 1)) * SW + nx] > v) { peak = false; break; }
                }
                if (!peak) continue;
                const int slot = atomicAdd(count + m2, 1);
                if (slot >= p.K) continue;               // overflow: the map is redone by the generic kernels
                // centroid over the border-clipped window, raster order (pipeline.py:46-62); the rows were
                // just streamed through L2
                const float* src = heat + (size_t)(first_map + m2) * H * W;
                float sy = 0.0f, sx = 0.0f, sp = 0.0f;
                for (int i = okp_max(r - 2, 0); i < okp_min(r + 3, H); ++i)
                    for (int j = okp_max(x - 2, 0); j < okp_min(x + 3, W); ++j) {
                        const float q = __ldg(src + (size_t)i * W + j);
                        sy = __fadd_rn(sy, __fmul_rn(q, (float)i));
                        sx = __fadd_rn(sx, __fmul_rn(q, (float)j));
                        sp = __fadd_rn(sp, q);
                    }
                OkpStripPeak pk;
                pk.key = r * W + x;
                pk.score = v;
                pk.cx = __fdiv_rn(sx, sp);
                pk.cy = __fdiv_rn(sy, sp);
                pk.conf = sp;
                list[(size_t)m2 * p.K + slot] = pk;
            }
        }
    }
    __syncthreads();

    // ---- epilogue: raster order (rank by key), final tables, unused slots cleared ----
    const int map = first_map + mm;
    if (map >= p.maps) return;
    const int total = count[mm];
    if (s == 0) t.peak_count[map] = total;
    if (total > p.K) return;                             // tables of this map are written by the overflow path
    const OkpStripPeak* mine = list + (size_t)mm * p.K;
    for (int i = s; i < p.K; i += p.strips) {
        const size_t slot_i = (size_t)map * p.K + i;
        t.peak_object[slot_i] = -1;
        t.peak_vote[2 * slot_i] = 0.0; t.peak_vote[2 * slot_i + 1] = 0.0;
        if (i >= total) {
            t.peak_yx[2 * slot_i] = -1; t.peak_yx[2 * slot_i + 1] = -1;
            t.peak_score[slot_i] = 0.0f;
            t.peak_xy[2 * slot_i] = 0.0f; t.peak_xy[2 * slot_i + 1] = 0.0f;
            t.peak_conf[slot_i] = 0.0f;
        } else {
            const OkpStripPeak pk = mine[i];
            int rank = 0;
            for (int j = 0; j < total; ++j) rank += (mine[j].key < pk.key);
            const size_t dst = (size_t)map * p.K + rank;
            const int y = pk.key / W;
            t.peak_yx[2 * dst] = y; t.peak_yx[2 * dst + 1] = pk.key - y * W;
            t.peak_score[dst] = pk.score;
            t.peak_xy[2 * dst] = pk.cx; t.peak_xy[2 * dst + 1] = pk.cy;
            t.peak_conf[dst] = pk.conf;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: plan + tensor map + launch
// ---------------------------------------------------------------------------------------------
static inline int okp_round_up_int(int v, int m) { return (v + m - 1) / m * m; }

static inline bool okp_strip_plan(int maps, int H, int W, int K, OkpStripPlan* out) {
    if (maps < 1 || H < 1 || W < 4 || (W % 4) != 0 || W > 500) return false;
    OkpStripPlan p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.maps = maps; p.K = K;
    p.SW = W + 4;
    p.strips = W / 4 + 1;
    if (4 * p.strips + 4 <= 256) {
        p.halves = 1; p.half_strips = p.strips;
    } else {
        p.halves = 2; p.half_strips = (p.strips + 1) / 2;
    }
    p.BW = 4 * p.half_strips + 4;                         // columns 4s-4 .. 4s+3 of the box's strips
    p.wpr = (W + 4 + 31) / 32;
    p.nb = (H + 6 + OKP_STRIP_RB - 1) / OKP_STRIP_RB;
    const int per_map = OKP_STRIP_NS * OKP_STRIP_RB * p.BW * p.halves * 4 + OKP_STRIP_SR * p.SW * 4 +
                        OKP_STRIP_SR * p.wpr * 4 + K * (int)sizeof(OkpStripPeak) + 4;
    const int budget = 110 * 1024;                        // two CTAs per SM
    int M = budget / per_map;
    if (M > 512 / p.strips) M = 512 / p.strips;
    if (M > maps) M = maps;
    if (M > 256) M = 256;
    if (M < 1) {
        M = 1;
        if (per_map + 1024 > 220 * 1024 || p.strips > 512) return false;
    }
    p.M = M;
    p.threads = M * p.strips;
    p.half_bytes = M * OKP_STRIP_RB * p.BW * 4;
    p.half_stride = okp_round_up_int(p.half_bytes, 128);
    p.stage_bytes = p.halves * p.half_stride;
    int off = OKP_STRIP_NS * p.stage_bytes;
    p.off_score = off; off += M * OKP_STRIP_SR * p.SW * 4;
    p.off_bitmap = off; off += M * OKP_STRIP_SR * p.wpr * 4;
    off = okp_round_up_int(off, 8);
    p.off_list = off; off += M * K * (int)sizeof(OkpStripPeak);
    p.off_count = off; off += M * 4;
    off = okp_round_up_int(off, 8);
    p.off_mbar = off; off += OKP_STRIP_NS * 8;
    p.smem_bytes = off;
    p.grid = (maps + M - 1) / M;
    *out = p;
    return true;
}

typedef CUresult (*OkpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline OkpEncodeTiledFn okp_encode_tiled_fn() {
    static OkpEncodeTiledFn fn = nullptr;                 // resolved once; a function pointer is not library state
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) != cudaSuccess ||
            status != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (OkpEncodeTiledFn)ptr;
    }
    return fn;
}

static inline int okp_strip_launch(const float* heat, const OkpStripPlan& p, float threshold,
                                   const OkpDecodeTables& tables, cudaStream_t stream) {
    OkpEncodeTiledFn encode = okp_encode_tiled_fn();
    if (!encode) return OKP_E_CUDA;
    if (((uintptr_t)heat & 15u) != 0) return OKP_E_UNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.maps};
    const cuuint64_t strides[2] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.W * p.H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)p.BW, (cuuint32_t)OKP_STRIP_RB, (cuuint32_t)p.M};
    const cuuint32_t elem[3] = {1, 1, 1};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)heat, dims, strides, box, elem,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OKP_E_CUDA;
    OKP_CUDA_CHECK(cudaFuncSetAttribute(okp_peaks_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    okp_peaks_strip_kernel<<<p.grid, p.threads, p.smem_bytes, stream>>>(tmap, heat, p, threshold, tables);
    OKP_CUDA_CHECK(cudaGetLastError());
    return OKP_OK;
}
