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
#include <stdlib.h>

#include "okp_common.cuh"
#include "okp_peaks.cuh"

#define OKP_STRIP_RB 5            // rows per batch (= register window depth, so slots are static)
#define OKP_STRIP_NS 2            // TMA stages
#define OKP_STRIP_MAX_LAG 8        // swept[] barriers

struct OkpStripPlan {
    int H, W, maps;
    int SW;                       // pitch of the box-sum ring: W + 4, column index = x + 2
    int strips;                   // W / 4 + 1 (strip s = pixels 4s-2 .. 4s+1)
    int half_strips;              // strips served by TMA box 0 (all of them when halves == 1)
    int halves;                   // 1 or 2 TMA boxes per row (box width <= 256 elements)
    int BW;                       // box width in floats
    int M;                        // maps per CTA
    int service_warps;            // warps that issue TMA and run NMS (the rest slide windows)
    int SR;                       // ring depth (rows, power of two) of the box-sum and bitmap rings
    int lag;                      // compute batch b may start once the NMS of batch b - lag is done
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

struct OkpStripPeak { int32_t key; float score; };

__device__ __forceinline__ uint32_t okp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void okp_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(okp_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void okp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(okp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool okp_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(okp_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Waits for the phase with the given parity. A wait that lasts longer than ~2 s of SM clocks can only
// be a protocol bug: trap (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void okp_mbar_wait(uint64_t* bar, uint32_t parity) {
    if (okp_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!okp_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void okp_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(okp_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(okp_smem_u32(bar)) : "memory");
}

// One row step of the sliding window. I = step inside the batch = register slot of the new row.
// Column c of the strip is pixel x = xs + c - 2 (xs = 4 * strip); vmask has bit c set when that pixel
// is inside the image. sw keeps the strip's last five rows of box sums (-inf outside the image), so
// that the candidates of row y - 2 can be narrowed down to the strip's own 5 x 4 block maximum before
// anything is written to the bitmap: one or two rows per blob instead of every row above threshold.
template <int I>
__device__ __forceinline__ void okp_strip_step(float (&w)[5][8], float (&sw)[5][4], const unsigned char* raw_row, int y,
                                               int H, int SW, float threshold, float* score_map, uint32_t* bitmap_map,
                                               int bitmap_pitch, int xs, uint32_t vmask, int ring_mask) {
    const float4* rp = reinterpret_cast<const float4*>(raw_row);
    const float4 lo = rp[0], hi = rp[1];
    w[I][0] = lo.x; w[I][1] = lo.y; w[I][2] = lo.z; w[I][3] = lo.w;
    w[I][4] = hi.x; w[I][5] = hi.y; w[I][6] = hi.z; w[I][7] = hi.w;
    constexpr int R0 = (I + 1) % 5, R1 = (I + 2) % 5, R2 = (I + 3) % 5, R3 = (I + 4) % 5, R4 = I;
    if (y >= 0 && y < H) {                                          // uniform over the CTA
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
        *reinterpret_cast<float4*>(score_map + (size_t)(y & ring_mask) * SW + xs) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) sw[I][c] = a[c];
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) sw[I][c] = -INFINITY;           // max_pool2d pads with -inf
    }
    // ---- candidates of row y - 2 (slot R2), whose five rows of box sums are now known ----
    const int yc = y - 2;
    if (yc < 0 || yc >= H) return;                                  // uniform
    const float (&b)[4] = sw[R2];
    if (fmaxf(fmaxf(fmaxf(b[0], b[1]), b[2]), b[3]) > threshold) {  // rare; runs warp-wide, keep it short
        const float ninf = -INFINITY;
        float cm[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float m = fmaxf(fmaxf(fmaxf(sw[0][c], sw[1][c]), fmaxf(sw[2][c], sw[3][c])), sw[4][c]);
            cm[c] = ((vmask >> c) & 1u) ? m : ninf;                  // columns outside the image never win
        }
        const float n012 = fmaxf(fmaxf(cm[0], cm[1]), cm[2]), n123 = fmaxf(fmaxf(cm[1], cm[2]), cm[3]);
        const float nall = fmaxf(n012, cm[3]);
        uint32_t bits = (b[0] > threshold && b[0] == n012) ? 1u : 0u;
        bits |= (b[1] > threshold && b[1] == nall) ? 2u : 0u;
        bits |= (b[2] > threshold && b[2] == nall) ? 4u : 0u;
        bits |= (b[3] > threshold && b[3] == n123) ? 8u : 0u;
        bits &= vmask;
        if (bits) atomicOr(bitmap_map + (yc & ring_mask) * bitmap_pitch + (xs >> 5), bits << (xs & 31));
    }
}

__device__ __forceinline__ void okp_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(okp_smem_u32(bar)) : "memory");
}

// Roles. Warps [0, CW) are compute warps (thread = one strip of one map); they never meet a CTA-wide
// barrier inside the row loop: they wait for TMA data (full[]), slide down RB rows, and arrive on
// done[]. The last warps are service warps: they wait on done[b], re-arm the freed stage with the
// TMA of batch b + NS, then run NMS on the rows batch b completed. The only back-pressure on the
// compute warps is the ring guard swept[] (the service warps may lag at most one batch, because the
// box-sum ring holds SR = 16 rows).
__global__ void __launch_bounds__(768, 1)
okp_peaks_strip_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ heat, OkpStripPlan p,
                       float threshold, OkpDecodeTables t) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RB = OKP_STRIP_RB;
    const int SR = p.SR;
    constexpr int NS = OKP_STRIP_NS;
    const int NSW = p.service_warps;
    float* score = reinterpret_cast<float*>(smem + p.off_score);             // [M][SR][SW], column x + 2
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem + p.off_bitmap);     // [SR][M][wpr], bit = x + 2
    OkpStripPeak* list = reinterpret_cast<OkpStripPeak*>(smem + p.off_list); // [M][K]
    int* count = reinterpret_cast<int*>(smem + p.off_count);                 // [M]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_mbar);         // [NS] TMA landed
    uint64_t* done = full + OKP_STRIP_NS;                                // [NS] compute warps finished the batch
    uint64_t* ready = done + OKP_STRIP_NS;                               // [lag] same event, for the NMS warps (they may lag)
    uint64_t* swept = ready + OKP_STRIP_MAX_LAG;                         // [lag] NMS warps finished the batch

    const int tid = threadIdx.x;
    const int H = p.H, W = p.W, SW = p.SW;
    const int first_map = blockIdx.x * p.M;
    const int compute_warps = (p.threads + 31) >> 5;
    const bool service = (tid >> 5) >= compute_warps;

    for (int i = tid; i < p.M * SR * p.wpr; i += blockDim.x) bitmap[i] = 0;
    if (tid < p.M) count[tid] = 0;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { okp_mbar_init(full + i, 1); okp_mbar_init(done + i, compute_warps); }
        for (int i = 0; i < p.lag; ++i) { okp_mbar_init(ready + i, compute_warps); okp_mbar_init(swept + i, NSW - 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (service) {
        const int lane = tid & 31;
        const int sw = (tid >> 5) - compute_warps;        // service warp index
        const CUtensorMap* tmap_ptr = &tmap;             // address of the __grid_constant__ parameter itself
        auto issue = [=](int b) {                        // one thread: TMA the new rows of batch b
            uint64_t* bar = full + (b % NS);
            unsigned char* dst = smem + (size_t)(b % NS) * p.stage_bytes;
            okp_mbar_expect_tx(bar, (uint32_t)(p.halves * p.half_bytes));
            okp_tma_load_3d(dst, tmap_ptr, -4, b * RB - 2, first_map, bar);
            if (p.halves == 2) okp_tma_load_3d(dst + p.half_stride, tmap_ptr, 4 * p.half_strips - 4, b * RB - 2, first_map, bar);
        };
        if (sw == 0) {
            // producer warp: keeps NS batches of rows in flight; nothing else, so that a burst of NMS work
            // never delays the next TMA
            if (lane == 0) {
                for (int b = 0; b < NS && b < p.nb; ++b) issue(b);
                for (int b = 0; b + NS < p.nb; ++b) {
                    okp_mbar_wait(done + (b % NS), (uint32_t)((b / NS) & 1));   // every compute warp has left stage b % NS
                    issue(b + NS);
                }
            }
        } else {
        const int nw = NSW - 1, nwi = sw - 1;             // NMS warps
        const int MW = p.M * p.wpr;                       // bitmap words per ring row
        const int chunks_per_row = (MW + 31) >> 5;
        const float ninf = -INFINITY;
        int slot = 0;                                     // b % lag, ((b / lag) & 1) without dividing
        uint32_t parity = 0;
        for (int b = 0; b < p.nb; ++b) {
            okp_mbar_wait(ready + slot, parity);          // batch b: box sums + candidate bits are visible
            // ---- NMS of the rows whose 5x5 neighbourhood is now complete: [y0 - 2, y0 + 3) ----
            // The row's bitmap is cut into chunks of 32 words, dealt round-robin to the service warps. A
            // non-zero word (32 pixels of one row, a few candidates) is tested by the whole warp at once:
            // lane i owns pixel x0 + i, takes the vertical maximum of its column over the 5 rows, lanes
            // 0..3 do the same for the 4 halo columns, and the horizontal 5-window comes from shuffles.
            const int y0 = b * RB - 4;
            int rr = 0, cc = nwi;                         // chunk = rr * chunks_per_row + cc, without dividing
            for (;; cc += nw) {
                while (cc >= chunks_per_row) { cc -= chunks_per_row; ++rr; }
                if (rr >= RB) break;
                const int r = y0 - 2 + rr;
                if (r < 0 || r >= H) continue;
                const int j = (cc << 5) + lane;
                uint32_t* row_words = bitmap + (size_t)(r & (SR - 1)) * MW;
                uint32_t mine = 0;
                if (j < MW) {
                    mine = row_words[j];
                    if (mine) row_words[j] = 0;
                }
                uint32_t nonzero = __ballot_sync(0xffffffffu, mine != 0);
                while (nonzero) {
                    const int src = __ffs(nonzero) - 1;
                    nonzero &= nonzero - 1;
                    const uint32_t bits = __shfl_sync(0xffffffffu, mine, src);
                    const int jj = j - lane + src;
                    const int m2 = jj / p.wpr, wi = jj - m2 * p.wpr;
                    const float* sm = score + (size_t)m2 * SR * SW + 2;     // sm[ring row * SW + x]
                    const int x0 = wi * 32 - 2;                              // pixel of bit 0
                    const int x = x0 + lane;
                    // halo columns x0-2, x0-1, x0+32, x0+33 on lanes 0..3
                    const int xh = lane < 2 ? x0 - 2 + lane : x0 + 30 + lane;
                    const bool x_in = x >= 0 && x < W;
                    const bool xh_in = lane < 4 && xh >= 0 && xh < W;
                    float vm = ninf, hv = ninf, v = ninf;
#pragma unroll
                    for (int d = -2; d <= 2; ++d) {
                        const int ry = r + d;
                        const bool row_in = ry >= 0 && ry < H;             // max_pool2d pads with -inf
                        const float* row = sm + ((ry & (SR - 1)) * SW);
                        const float u = (row_in && x_in) ? row[x] : ninf;
                        const float uh = (row_in && xh_in) ? row[xh] : ninf;
                        if (d == 0) v = u;
                        vm = fmaxf(vm, u);
                        hv = fmaxf(hv, uh);
                    }
                    float l1 = __shfl_up_sync(0xffffffffu, vm, 1), l2 = __shfl_up_sync(0xffffffffu, vm, 2);
                    float r1 = __shfl_down_sync(0xffffffffu, vm, 1), r2 = __shfl_down_sync(0xffffffffu, vm, 2);
                    const float h0 = __shfl_sync(0xffffffffu, hv, 0), h1 = __shfl_sync(0xffffffffu, hv, 1);
                    const float h2 = __shfl_sync(0xffffffffu, hv, 2), h3 = __shfl_sync(0xffffffffu, hv, 3);
                    if (lane == 0) { l1 = h1; l2 = h0; }
                    if (lane == 1) l2 = h1;
                    if (lane == 31) { r1 = h2; r2 = h3; }
                    if (lane == 30) r2 = h2;
                    const float hm = fmaxf(fmaxf(fmaxf(l1, l2), fmaxf(r1, r2)), vm);
                    if (((bits >> lane) & 1u) && v == hm) {                // a peak: box sum equals the 5x5 maximum
                        const int entry = atomicAdd(count + m2, 1);
                        if (entry < p.K) {                                    // overflow: redone by the overflow kernel
                            OkpStripPeak pk;
                            pk.key = r * W + x;
                            pk.score = v;
                            list[(size_t)m2 * p.K + entry] = pk;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) okp_mbar_arrive(swept + slot);
            if (++slot == p.lag) { slot = 0; parity ^= 1u; }
        }
        }
    } else {
        const bool active = tid < p.threads;              // the last compute warp may be partly idle
        const int ct = active ? tid : p.threads - 1;      // idle lanes shadow a real strip (their stores are masked)
        const int mm = ct / p.strips;                     // map slot inside the CTA
        const int s = ct - mm * p.strips;                 // strip inside the map
        const int half = s >= p.half_strips ? 1 : 0;
        const int xs = 4 * s;                             // pixels xs-2 .. xs+1
        uint32_t vmask = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) vmask |= (xs + c - 2 >= 0 && xs + c - 2 < W) ? (1u << c) : 0u;
        if (!active) vmask = 0;
        // this thread's window row inside a stage: box [M][RB][BW], first column 4 * (s - half * half_strips)
        const int thread_raw = half * p.half_stride + (mm * RB * p.BW + 4 * (s - half * p.half_strips)) * 4;
        const int row_pitch = p.BW * 4;
        float* score_map = score + (size_t)mm * SR * SW;
        uint32_t* bitmap_map = bitmap + (size_t)mm * p.wpr;   // + ring row * M * wpr
        const int bitmap_pitch = p.M * p.wpr;

        float w[5][8], sw[5][4];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) w[i][j] = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) sw[i][j] = -INFINITY;
        }

        const int lag = p.lag;
        int lag_slot = 0, ready_slot = 0;                 // b % lag and ((b - lag) / lag) & 1 without dividing
        uint32_t lag_parity = 0;
        for (int b = 0; b < p.nb; ++b) {
            // ring guard: batch b overwrites ring rows the service warp reads until it has finished batch b - 2
            if (b >= lag) {
                okp_mbar_wait(swept + lag_slot, lag_parity);
            }
            okp_mbar_wait(full + (b % NS), (uint32_t)((b / NS) & 1));
            const unsigned char* raw = smem + (size_t)(b % NS) * p.stage_bytes + thread_raw;
            const int y0 = b * RB - 4;                    // new row of step i is y0 + i + 2
            okp_strip_step<0>(w, sw, raw + 0 * row_pitch, y0 + 0, H, SW, threshold, score_map, bitmap_map, bitmap_pitch, xs, vmask, SR - 1);
            okp_strip_step<1>(w, sw, raw + 1 * row_pitch, y0 + 1, H, SW, threshold, score_map, bitmap_map, bitmap_pitch, xs, vmask, SR - 1);
            okp_strip_step<2>(w, sw, raw + 2 * row_pitch, y0 + 2, H, SW, threshold, score_map, bitmap_map, bitmap_pitch, xs, vmask, SR - 1);
            okp_strip_step<3>(w, sw, raw + 3 * row_pitch, y0 + 3, H, SW, threshold, score_map, bitmap_map, bitmap_pitch, xs, vmask, SR - 1);
            okp_strip_step<4>(w, sw, raw + 4 * row_pitch, y0 + 4, H, SW, threshold, score_map, bitmap_map, bitmap_pitch, xs, vmask, SR - 1);
            __syncwarp();
            if ((tid & 31) == 0) { okp_mbar_arrive(done + (b % NS)); okp_mbar_arrive(ready + ready_slot); }
            if (++ready_slot == lag) ready_slot = 0;
            if (b >= lag && ++lag_slot == lag) { lag_slot = 0; lag_parity ^= 1u; }
        }
    }
    __syncthreads();

    // ---- epilogue: raster order (rank by key), centroids, final tables, unused slots cleared ----
    if (tid >= p.threads) return;
    const int mm = tid / p.strips;
    const int s = tid - mm * p.strips;
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
            const int y = pk.key / W, x = pk.key - y * W;
            // centroid over the border-clipped window, raster order (pipeline.py:46-62). The 25 loads are
            // independent (clamped address, zero weight outside the image: adding +0 changes nothing), so
            // the whole map's peaks cost one L2 round trip here instead of one per tap inside the row loop.
            const float* src = heat + (size_t)map * H * W;
            float q[25];
#pragma unroll
            for (int k = 0; k < 25; ++k) {
                const int i2 = y + k / 5 - 2, j2 = x + k % 5 - 2;
                const bool in = i2 >= 0 && i2 < H && j2 >= 0 && j2 < W;
                q[k] = in ? __ldg(src + (size_t)i2 * W + j2) : 0.0f;
            }
            float sy = 0.0f, sx = 0.0f, sp = 0.0f;
#pragma unroll
            for (int k = 0; k < 25; ++k) {
                const int i2 = y + k / 5 - 2, j2 = x + k % 5 - 2;
                const bool in = i2 >= 0 && i2 < H && j2 >= 0 && j2 < W;
                if (in) {
                    sy = __fadd_rn(sy, __fmul_rn(q[k], (float)i2));
                    sx = __fadd_rn(sx, __fmul_rn(q[k], (float)j2));
                    sp = __fadd_rn(sp, q[k]);
                }
            }
            t.peak_yx[2 * dst] = y; t.peak_yx[2 * dst + 1] = x;
            t.peak_score[dst] = pk.score;
            t.peak_xy[2 * dst] = __fdiv_rn(sx, sp); t.peak_xy[2 * dst + 1] = __fdiv_rn(sy, sp);
            t.peak_conf[dst] = sp;
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
    p.service_warps = 3;                                  // 1 TMA producer + 2 NMS warps
    if (const char* e = getenv("OKP_STRIP_SERVICE_WARPS")) p.service_warps = atoi(e) >= 2 && atoi(e) <= 8 ? atoi(e) : 3;   // tuning aid
    p.SR = 16;
    if (const char* e = getenv("OKP_STRIP_RING")) p.SR = atoi(e) == 32 ? 32 : 16;                                 // tuning aid
    p.lag = (p.SR - 8) / OKP_STRIP_RB + 1;                // rows [5b'-8, ..) of a pending NMS must not alias rows <= 5b
    const int per_map = OKP_STRIP_NS * OKP_STRIP_RB * p.BW * p.halves * 4 + p.SR * p.SW * 4 +
                        p.SR * p.wpr * 4 + K * (int)sizeof(OkpStripPeak) + 4;
    int budget = 110 * 1024;                              // two CTAs per SM
    if (const char* e = getenv("OKP_STRIP_SMEM_KB")) budget = atoi(e) * 1024;                                     // tuning aid
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
    p.off_score = off; off += M * p.SR * p.SW * 4;
    p.off_bitmap = off; off += M * p.SR * p.wpr * 4;
    off = okp_round_up_int(off, 8);
    p.off_list = off; off += M * K * (int)sizeof(OkpStripPeak);
    p.off_count = off; off += M * 4;
    off = okp_round_up_int(off, 8);
    p.off_mbar = off; off += (2 * OKP_STRIP_NS + 2 * OKP_STRIP_MAX_LAG) * 8;
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
    const int block = (p.threads + 31) / 32 * 32 + 32 * p.service_warps;    // compute warps + service warps
    okp_peaks_strip_kernel<<<p.grid, block, p.smem_bytes, stream>>>(tmap, heat, p, threshold, tables);
    OKP_CUDA_CHECK(cudaGetLastError());
    return OKP_OK;
}
