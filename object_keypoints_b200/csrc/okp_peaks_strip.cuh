// okp_peaks_strip.cuh -- tuned K1 for sm_100a: TMA-fed strip kernel, bounded filter + exact check.
//
// Replaces perception/pipeline.py:46-79 + perception/models.py:55-58 for every map of a batch.
//
// Why this shape. A peak is a pixel whose 5x5 box sum S (25 float32 additions in raster tap order,
// torch's CPU conv2d is bitwise that) exceeds the threshold and equals the
// maximum of S over its 5x5 neighbourhood. Computing S exactly for every pixel costs 24
// dependent-order FADDs per pixel and made the first version of this kernel issue-bound at 0.31 of
// the HBM roofline (profiles/r01a_k1_strip_ncu.md). Peaks are rare, so the stream now computes a
// cheap BOUND instead and the exact arithmetic runs only where the bound cannot decide:
//
//   * every pixel gets the separable sum S~ (horizontal 5-sums shared between the 4 pixels of a
//     strip, vertical 5-sum from running pair sums: 5.25 FADD per pixel). For non-negative inputs
//     any summation order of the same 25 terms is within gamma_24 * sum|x| of the true sum, so
//     |S - S~| <= 2.9e-6 * S~. With tie = 1 + 2e-5: S~(p) <= threshold * (1 - 1e-5) proves
//     S(p) <= threshold, and S~(q) > tie * S~(p) proves S(q) > S(p). Pixels that survive both tests
//     against their whole neighbourhood are CANDIDATES (true peaks plus near ties, a few per blob);
//   * a candidate's S is then computed exactly (25 loads from L2, raster-order __fadd_rn), compared
//     with the threshold, and with the exact S of those neighbours whose S~ is within the tie band
//     (normally none). The peak set, the scores and the raster order are therefore bit-identical to
//     the exact kernel's; ties keep every tied pixel like `x == hmax` does;
//   * maps holding a negative value (sign bit seen by a 2-LOP3-per-row check) void the bound: they
//     are handed to the exact generic kernels through the overflow path, like maps with more than K
//     peaks ("the first K in raster order").
//
// Data movement: heatmap rows arrive in shared memory by TMA (cp.async.bulk.tensor, one elected
// thread); the tensor map's out-of-bounds zero fill IS conv2d's zero padding. TMA wants the
// innermost start coordinate 16-byte aligned (tools/microbench/tma_probe.cu), so the box starts at
// column -4 and strip s produces the pixels x = 4s-2 .. 4s+1 from the columns 4s-4 .. 4s+3 = two
// aligned LDS.128. A thread owns one strip of one map and slides down it; per row step: 2 LDS.128,
// 21 FADD, 1 STS.128 (S~ to a ring for the neighbourhood test), 3 FMNMX + 1 compare. Pixels above
// the threshold that top their strip's own 5x4 block set a bit in a row-indexed bitmap ring; the
// service warps test those bits against the ring and queue the survivors. The epilogue runs the
// exact check for all of them at once (one thread each, so the CTA pays one L2 round trip; doing it
// in the service warps serialised the round trips and stalled the stream on the ring guard), ranks
// the confirmed peaks by raster key and writes them, with centroids, to the tables.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "okp_common.cuh"
#include "okp_peaks.cuh"

#define OKP_STRIP_RB 5            // rows per batch (= period of the register rings, so slots are static)
#define OKP_STRIP_MAX_NS 8        // TMA stages (upper bound; the plan picks NS)
#define OKP_STRIP_MAX_LAG 8       // swept[] barriers
#define OKP_STRIP_TIE 1.00002f    // S~(q) > tie * S~(p) proves S(q) > S(p)   (see the header)
#define OKP_STRIP_THRESHOLD_SLACK 1e-5f

struct OkpStripPlan {
    int H, W, maps;
    int SW;                       // pitch of the S~ ring: W + 4, column index = x + 2
    int strips;                   // W / 4 + 1 (strip s = pixels 4s-2 .. 4s+1)
    int half_strips;              // strips served by TMA box 0 (all of them when halves == 1)
    int halves;                   // 1 or 2 TMA boxes per row (box width <= 256 elements)
    int BW;                       // box width in floats
    int M;                        // maps per CTA
    int NS;                       // TMA stages
    int service_warps;            // warps that issue TMA and test candidates (the rest slide windows)
    int SR;                       // ring depth (rows, power of two) of the S~ and bitmap rings
    int lag;                      // compute batch b may start once the candidates of batch b - lag are done
    int nb;                       // batches
    int K;                        // table capacity per map
    int PK;                       // candidate slots per map (2 K)
    int wpr;                      // bitmap words per row (bit index = x + 2)
    int threads;
    int half_bytes;               // bytes of one TMA box: M * RB * BW * 4
    int half_stride;              // half_bytes rounded up to 128 (TMA destinations are 128-byte aligned)
    int stage_bytes;              // halves * half_stride
    int off_score, off_bitmap, off_list, off_pending, off_count, off_mbar;
    int smem_bytes;
    int grid;
};

struct OkpStripPeak { int32_t key; float score, cx, cy, conf; };   // a confirmed peak, ready for the tables
struct OkpStripCandidate { int32_t key; uint32_t ties; };          // survived the S~ test; ties = neighbours in the tie band

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
__device__ __forceinline__ void okp_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(okp_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void okp_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(okp_smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(okp_smem_u32(bar)) : "memory");
}

// The reference's box sum at pixel (y, x): 25 additions in raster tap order starting from +0, zero
// outside the image (perception/pipeline.py:70-71). The loads are independent (one L2 round trip).
__device__ __forceinline__ float okp_exact_box_sum(const float* __restrict__ src, int H, int W, int y, int x) {
    float q[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) {
        const int i2 = y + k / 5 - 2, j2 = x + k % 5 - 2;
        const bool in = i2 >= 0 && i2 < H && j2 >= 0 && j2 < W;
        q[k] = in ? __ldg(src + (size_t)i2 * W + j2) : 0.0f;
    }
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 25; ++k) acc = __fadd_rn(acc, q[k]);
    return acc;
}

// One row step of the sliding window. I = step inside the batch = slot of the register rings.
// (lo, hi) are the columns 4s-4 .. 4s+3 of the new input row y + 2; the step produces S~ of row y for
// the strip's pixels x = xs + c - 2 (c = 0..3, xs = 4 * strip). vmask has bit c set when that pixel is
// inside the image. pr keeps the running pair sums h(t-1) + h(t) of the horizontal sums, sv the
// strip's last five rows of S~, so that the candidates of row y - 2 can be
// narrowed down to the strip's own 5 x 4 block before anything is written to the bitmap: one or two
// rows per blob instead of every row above threshold.
template <int I>
__device__ __forceinline__ void okp_strip_step(const float4 lo, const float4 hi, float (&pr)[5][4], float (&hp)[4],
                                               float (&sv)[5][4], uint32_t& sign, int y, int H, int SW, float thr_lo,
                                               float* score_map, uint32_t* bitmap_map, int bitmap_pitch, int xs,
                                               uint32_t vmask, int ring_mask) {
    // horizontal 5-sums of the four windows w[c .. c+4] (w = lo.xyzw, hi.xyzw): 9 adds
    const float c34 = lo.w + hi.x;
    const float t12 = lo.y + lo.z;
    const float t56 = hi.y + hi.z;
    const float tc = t12 + c34;
    const float u = c34 + t56;
    float h[4];
    h[0] = lo.x + tc; h[1] = tc + hi.y; h[2] = lo.z + u; h[3] = u + hi.w;
    // the strip's own pixels are w[2..5]: a set sign bit anywhere voids the bound (see the header)
    sign |= __float_as_uint(lo.z) | __float_as_uint(lo.w);
    sign |= __float_as_uint(hi.x) | __float_as_uint(hi.y);
    // vertical: S~(t) = P(t-3) + P(t-1) + h(t), P(t) = h(t-1) + h(t): 12 adds
    constexpr int R2 = (I + 3) % 5;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        v[c] = (pr[(I + 2) % 5][c] + pr[(I + 4) % 5][c]) + h[c];
        pr[I][c] = hp[c] + h[c];
        hp[c] = h[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) sv[I][c] = v[c];
    if (y >= 0 && y < H)                                            // uniform over the CTA
        *reinterpret_cast<float4*>(score_map + (size_t)(y & ring_mask) * SW + xs) = make_float4(v[0], v[1], v[2], v[3]);
    // ---- candidates of row y - 2 (slot R2), whose five rows of S~ are now known ----
    const int yc = y - 2;
    if (yc < 0 || yc >= H) return;                                  // uniform
    const float (&b)[4] = sv[R2];
    // warp-uniform branch (taken by about a quarter of the warp rows of a busy map): the column maxima of
    // the neighbouring strips come from the neighbouring lanes, so that a bit is only set for a pixel that
    // tops (within the tie band) its whole 5 x 5 neighbourhood as far as this warp can see it
    if (__any_sync(0xffffffffu, fmaxf(fmaxf(b[0], b[1]), fmaxf(b[2], b[3])) > thr_lo)) {
        const float ninf = -INFINITY;
        // rows y-4, y-3 (slots I+1, I+2) and y-1, y (slots I+4, I) may lie outside the image: max_pool2d pads with -inf
        const bool up2 = yc >= 2, up1 = yc >= 1, dn1 = yc + 1 < H, dn2 = yc + 2 < H;
        float cm[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float m = b[c];
            m = fmaxf(m, up2 ? sv[(I + 1) % 5][c] : ninf);
            m = fmaxf(m, up1 ? sv[(I + 2) % 5][c] : ninf);
            m = fmaxf(m, dn1 ? sv[(I + 4) % 5][c] : ninf);
            m = fmaxf(m, dn2 ? sv[I][c] : ninf);
            cm[c] = ((vmask >> c) & 1u) ? m : ninf;                  // columns outside the image never win
        }
        // columns -2, -1 from the strip to the left, 4, 5 from the strip to the right. vmask bits 4 / 5 say that
        // lane - 1 / lane + 1 holds that strip; where it does not (warp edge: unknown, image edge: nothing
        // there) -inf keeps the test conservative, the candidate warps see the whole neighbourhood anyway
        float l2 = __shfl_up_sync(0xffffffffu, cm[2], 1), l3 = __shfl_up_sync(0xffffffffu, cm[3], 1);
        float r0 = __shfl_down_sync(0xffffffffu, cm[0], 1), r1 = __shfl_down_sync(0xffffffffu, cm[1], 1);
        if (!(vmask & 16u)) { l2 = ninf; l3 = ninf; }
        if (!(vmask & 32u)) { r0 = ninf; r1 = ninf; }
        const float m12 = fmaxf(cm[1], cm[2]);
        const float n0 = fmaxf(fmaxf(l2, l3), fmaxf(cm[0], m12));
        const float n1 = fmaxf(fmaxf(l3, cm[0]), fmaxf(m12, cm[3]));
        const float n2 = fmaxf(fmaxf(cm[0], m12), fmaxf(cm[3], r0));
        const float n3 = fmaxf(fmaxf(m12, cm[3]), fmaxf(r0, r1));
        const float tie = OKP_STRIP_TIE;
        uint32_t bits = (b[0] > thr_lo && b[0] * tie >= n0) ? 1u : 0u;
        bits |= (b[1] > thr_lo && b[1] * tie >= n1) ? 2u : 0u;
        bits |= (b[2] > thr_lo && b[2] * tie >= n2) ? 4u : 0u;
        bits |= (b[3] > thr_lo && b[3] * tie >= n3) ? 8u : 0u;
        bits &= vmask;
        if (bits) atomicOr(bitmap_map + (yc & ring_mask) * bitmap_pitch + (xs >> 5), bits << (xs & 31));
    }
}

// Roles. Warps [0, CW) are compute warps (thread = one strip of one map); they never meet a CTA-wide
// barrier inside the row loop: they wait for TMA data (full[]), slide down RB rows, and arrive on
// done[]. The last warps are service warps: warp 0 of them keeps NS batches of rows in flight, the
// others wait on ready[b] and test the candidates of the rows batch b completed. The only
// back-pressure on the compute warps is the ring guard swept[] (the S~ ring holds SR = 16 rows).
__global__ void __launch_bounds__(576, 1)
okp_peaks_strip_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ heat, OkpStripPlan p,
                       float threshold, float thr_lo, OkpDecodeTables t) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RB = OKP_STRIP_RB;
    const int SR = p.SR;
    const int NS = p.NS;
    const int NSW = p.service_warps;
    float* score = reinterpret_cast<float*>(smem + p.off_score);             // [M][SR][SW], column x + 2
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem + p.off_bitmap);     // [SR][M][wpr], bit = x + 2
    OkpStripPeak* list = reinterpret_cast<OkpStripPeak*>(smem + p.off_list); // [M][K]
    OkpStripCandidate* pending = reinterpret_cast<OkpStripCandidate*>(smem + p.off_pending);   // [M][PK]
    int* count = reinterpret_cast<int*>(smem + p.off_count);                 // [M] peaks, [M] negative-input flags, [M] candidates
    int* negative = count + p.M;
    int* n_pending = count + 2 * p.M;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_mbar);         // [NS] TMA landed
    uint64_t* done = full + OKP_STRIP_MAX_NS;                            // [NS] compute warps finished the batch
    uint64_t* ready = done + OKP_STRIP_MAX_NS;                           // [lag] same event, for the candidate warps (they may lag)
    uint64_t* swept = ready + OKP_STRIP_MAX_LAG;                         // [lag] candidate warps finished the batch

    const int tid = threadIdx.x;
    const int H = p.H, W = p.W, SW = p.SW;
    const int first_map = blockIdx.x * p.M;
    const int compute_warps = (p.threads + 31) >> 5;
    const bool service = (tid >> 5) >= compute_warps;

    for (int i = tid; i < p.M * SR * p.wpr; i += blockDim.x) bitmap[i] = 0;
    for (int i = tid; i < 3 * p.M; i += blockDim.x) count[i] = 0;
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
        if (sw == 0) {
            // producer warp: keeps NS batches of rows in flight; nothing else, so that a burst of
            // candidate work never delays the next TMA
            if (lane == 0) {
                const CUtensorMap* tmap_ptr = &tmap;      // address of the __grid_constant__ parameter itself
                int stage = 0;
                uint32_t parity = 0;
                for (int b = 0; b < p.nb; ++b) {
                    if (b >= NS) okp_mbar_wait(done + stage, parity);        // every compute warp has left the stage
                    uint64_t* bar = full + stage;
                    unsigned char* dst = smem + (size_t)stage * p.stage_bytes;
                    okp_mbar_expect_tx(bar, (uint32_t)(p.halves * p.half_bytes));
                    okp_tma_load_3d(dst, tmap_ptr, -4, b * RB - 2, first_map, bar);
                    if (p.halves == 2) okp_tma_load_3d(dst + p.half_stride, tmap_ptr, 4 * p.half_strips - 4, b * RB - 2, first_map, bar);
                    if (++stage == NS) { stage = 0; if (b >= NS) parity ^= 1u; }
                }
            }
        } else {
            const int nw = NSW - 1, nwi = sw - 1;         // candidate warps
            const int MW = p.M * p.wpr;                   // bitmap words per ring row
            const int chunks_per_row = (MW + 31) >> 5;
            const float tie = OKP_STRIP_TIE;
            int slot = 0;                                 // b % lag, ((b / lag) & 1) without dividing
            uint32_t parity = 0;
            for (int b = 0; b < p.nb; ++b) {
                okp_mbar_wait(ready + slot, parity);      // batch b: S~ rows + candidate bits are visible
                // Rows whose 5x5 neighbourhood is now complete: [y0 - 2, y0 + 3). A row's bitmap is cut into
                // chunks of 32 words, dealt round-robin to the candidate warps; a lane owns one word and
                // walks its set bits (a few per blob): neighbourhood test on S~; survivors are queued for the epilogue.
                const int y0 = b * RB - 4;
                int rr = 0, cc = nwi;                     // chunk = rr * chunks_per_row + cc, without dividing
                for (;; cc += nw) {
                    while (cc >= chunks_per_row) { cc -= chunks_per_row; ++rr; }
                    if (rr >= RB) break;
                    const int r = y0 - 2 + rr;
                    if (r < 0 || r >= H) continue;
                    const int j = (cc << 5) + lane;
                    if (j >= MW) continue;
                    uint32_t* word = bitmap + (size_t)(r & (SR - 1)) * MW + j;
                    uint32_t mine = *word;
                    if (!mine) continue;
                    *word = 0;
                    const int m2 = j / p.wpr, wi = j - m2 * p.wpr;
                    const float* sm = score + (size_t)m2 * SR * SW + 2;      // sm[ring row * SW + x]
                    do {
                        const int bit = __ffs(mine) - 1;
                        mine &= mine - 1;
                        const int x = wi * 32 - 2 + bit;                      // inside the image (vmask)
                        const float sp = sm[(r & (SR - 1)) * SW + x];
                        const float lim = sp * tie;
                        bool alive = true;
                        uint32_t ties = 0;                                    // neighbours the bound cannot order
#pragma unroll
                        for (int k = 0; k < 25; ++k) {
                            if (k == 12) continue;
                            const int ry = r + k / 5 - 2, rx = x + k % 5 - 2;
                            if (ry >= 0 && ry < H && rx >= 0 && rx < W) {    // max_pool2d pads with -inf
                                const float v = sm[(ry & (SR - 1)) * SW + rx];
                                if (v > lim) alive = false;
                                else if (v * tie >= sp) ties |= 1u << k;
                            }
                        }
                        if (!alive) continue;
                        // survivor: the exact check needs global loads, which must not hold up the ring; it
                        // runs in the epilogue for all survivors of the CTA at once
                        const int entry = atomicAdd(n_pending + m2, 1);
                        if (entry < p.PK) {
                            OkpStripCandidate cd;
                            cd.key = r * W + x;
                            cd.ties = ties;
                            pending[(size_t)m2 * p.PK + entry] = cd;
                        }
                    } while (mine);
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
        if (active && (tid & 31) > 0 && s > 0) vmask |= 16u;                   // lane - 1 holds strip s - 1 of this map
        if (active && (tid & 31) < 31 && s + 1 < p.strips && tid + 1 < p.threads) vmask |= 32u;   // lane + 1 holds strip s + 1
        // this thread's window row inside a stage: box [M][RB][BW], first column 4 * (s - half * half_strips)
        const int thread_raw = half * p.half_stride + (mm * RB * p.BW + 4 * (s - half * p.half_strips)) * 4;
        const int row_pitch = p.BW * 4;
        float* score_map = score + (size_t)mm * SR * SW;
        uint32_t* bitmap_map = bitmap + (size_t)mm * p.wpr;   // + ring row * M * wpr
        const int bitmap_pitch = p.M * p.wpr;

        float pr[5][4], hp[4], sv[5][4];
        uint32_t sign = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { pr[i][j] = 0.0f; sv[i][j] = -INFINITY; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) hp[j] = 0.0f;

        const int lag = p.lag;
        int lag_slot = 0, ready_slot = 0;                 // b % lag and ((b - lag) / lag) & 1 without dividing
        uint32_t lag_parity = 0;
        int stage = 0;
        uint32_t full_parity = 0;
#define OKP_STRIP_ARGS H, SW, thr_lo, score_map, bitmap_map, bitmap_pitch, xs, vmask, SR - 1
        for (int b = 0; b < p.nb; ++b) {
            // ring guard: batch b overwrites ring rows the candidate warps read until they have finished batch b - lag
            if (b >= lag) okp_mbar_wait(swept + lag_slot, lag_parity);
            okp_mbar_wait(full + stage, full_parity);
            const unsigned char* raw = smem + (size_t)stage * p.stage_bytes + thread_raw;
            const int y0 = b * RB - 4;                    // new row of step i is y0 + i + 2
            // rows are fetched one step ahead of their use (two register pairs, ping-pong)
            float4 a0 = reinterpret_cast<const float4*>(raw)[0], a1 = reinterpret_cast<const float4*>(raw)[1];
            float4 b0 = reinterpret_cast<const float4*>(raw + row_pitch)[0], b1 = reinterpret_cast<const float4*>(raw + row_pitch)[1];
            okp_strip_step<0>(a0, a1, pr, hp, sv, sign, y0 + 0, OKP_STRIP_ARGS);
            a0 = reinterpret_cast<const float4*>(raw + 2 * row_pitch)[0]; a1 = reinterpret_cast<const float4*>(raw + 2 * row_pitch)[1];
            okp_strip_step<1>(b0, b1, pr, hp, sv, sign, y0 + 1, OKP_STRIP_ARGS);
            b0 = reinterpret_cast<const float4*>(raw + 3 * row_pitch)[0]; b1 = reinterpret_cast<const float4*>(raw + 3 * row_pitch)[1];
            okp_strip_step<2>(a0, a1, pr, hp, sv, sign, y0 + 2, OKP_STRIP_ARGS);
            a0 = reinterpret_cast<const float4*>(raw + 4 * row_pitch)[0]; a1 = reinterpret_cast<const float4*>(raw + 4 * row_pitch)[1];
            okp_strip_step<3>(b0, b1, pr, hp, sv, sign, y0 + 3, OKP_STRIP_ARGS);
            okp_strip_step<4>(a0, a1, pr, hp, sv, sign, y0 + 4, OKP_STRIP_ARGS);
            __syncwarp();
            if ((tid & 31) == 0) { okp_mbar_arrive(done + stage); okp_mbar_arrive(ready + ready_slot); }
            if (++stage == NS) { stage = 0; full_parity ^= 1u; }
            if (++ready_slot == lag) ready_slot = 0;
            if (b >= lag && ++lag_slot == lag) { lag_slot = 0; lag_parity ^= 1u; }
        }
#undef OKP_STRIP_ARGS
        if (active && (sign >> 31)) negative[mm] = 1;     // benign race: every writer stores 1
    }
    __syncthreads();

    // ---- epilogue A: exact check of every surviving candidate, one thread each (all loads of the CTA in
    // flight together: one L2 round trip). The 25 window values give the reference's box sum (raster-order
    // adds) and, for a confirmed peak, its centroid (pipeline.py:46-62) without loading anything twice. ----
    for (int i = tid; i < p.M * p.PK; i += blockDim.x) {
        const int mm = i / p.PK, j = i - mm * p.PK;
        if (first_map + mm >= p.maps || j >= n_pending[mm]) continue;
        const OkpStripCandidate cd = pending[i];
        const int y = cd.key / W, x = cd.key - y * W;
        const float* src = heat + (size_t)(first_map + mm) * H * W;
        float q[25];
#pragma unroll
        for (int k = 0; k < 25; ++k) {
            const int i2 = y + k / 5 - 2, j2 = x + k % 5 - 2;
            const bool in = i2 >= 0 && i2 < H && j2 >= 0 && j2 < W;
            q[k] = in ? __ldg(src + (size_t)i2 * W + j2) : 0.0f;       // +0 outside: conv2d's zero padding
        }
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 25; ++k) sum = __fadd_rn(sum, q[k]);
        if (!(sum > threshold)) continue;
        bool alive = true;
        uint32_t ties = cd.ties;
        while (ties) {                                           // neighbours the bound could not order (rare)
            const int k = __ffs(ties) - 1;
            ties &= ties - 1;
            if (okp_exact_box_sum(src, H, W, y + k / 5 - 2, x + k % 5 - 2) > sum) { alive = false; break; }
        }
        if (!alive) continue;
        const int entry = atomicAdd(count + mm, 1);               // a peak: S equals the 5x5 maximum
        if (entry >= p.K) continue;                               // overflow: redone by the overflow kernel
        float sy = 0.0f, sx = 0.0f, sp = 0.0f;
#pragma unroll
        for (int k = 0; k < 25; ++k) {
            const int i2 = y + k / 5 - 2, j2 = x + k % 5 - 2;
            const bool in = i2 >= 0 && i2 < H && j2 >= 0 && j2 < W;
            if (in) {                                             // border-clipped window, raster order
                sy = __fadd_rn(sy, __fmul_rn(q[k], (float)i2));
                sx = __fadd_rn(sx, __fmul_rn(q[k], (float)j2));
                sp = __fadd_rn(sp, q[k]);
            }
        }
        OkpStripPeak pk;
        pk.key = cd.key;
        pk.score = sum;
        pk.cx = __fdiv_rn(sx, sp);
        pk.cy = __fdiv_rn(sy, sp);
        pk.conf = sp;
        list[(size_t)mm * p.K + entry] = pk;
    }
    __syncthreads();

    // ---- epilogue B: raster order (rank by key), final tables, unused slots cleared ----
    if (tid >= p.threads) return;
    const int mm = tid / p.strips;
    const int s = tid - mm * p.strips;
    const int map = first_map + mm;
    if (map >= p.maps) return;
    // a map with negative values, or with more candidates than slots, is reported as overflowing, which
    // hands it to the exact generic path
    const int total = (negative[mm] || n_pending[mm] > p.PK) ? p.K + 1 : count[mm];
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
    p.service_warps = 4;                                  // 1 TMA producer + 3 candidate warps
    if (const char* e = getenv("OKP_STRIP_SERVICE_WARPS")) p.service_warps = atoi(e) >= 2 && atoi(e) <= 4 ? atoi(e) : 4;   // tuning aid
    p.SR = 16;
    if (const char* e = getenv("OKP_STRIP_RING")) p.SR = atoi(e) == 32 ? 32 : 16;                                 // tuning aid
    p.NS = 4;                                             // ~3 batches in flight per CTA cover the HBM latency
    if (const char* e = getenv("OKP_STRIP_STAGES")) p.NS = atoi(e) >= 2 && atoi(e) <= OKP_STRIP_MAX_NS ? atoi(e) : 4;      // tuning aid
    p.lag = (p.SR - 8) / OKP_STRIP_RB + 1;                // rows [5b'-8, ..) of a pending sweep must not alias rows <= 5b
    const int per_map = p.NS * OKP_STRIP_RB * p.BW * p.halves * 4 + p.SR * p.SW * 4 +
                        p.SR * p.wpr * 4 + K * (int)sizeof(OkpStripPeak) + 2 * K * (int)sizeof(OkpStripCandidate) + 12;
    int budget = 110 * 1024;                              // two CTAs per SM
    if (const char* e = getenv("OKP_STRIP_SMEM_KB")) budget = atoi(e) * 1024;                                     // tuning aid
    int M = (budget - 1024) / per_map;
    if (M > 448 / p.strips) M = 448 / p.strips;           // 14 compute warps + service warps <= 576 threads
    if (M > maps) M = maps;
    if (M > 256) M = 256;
    if (M < 1) {
        M = 1;
        if (per_map + 1024 > 220 * 1024 || p.strips > 448) return false;
    }
    p.M = M;
    p.threads = M * p.strips;
    p.half_bytes = M * OKP_STRIP_RB * p.BW * 4;
    p.half_stride = okp_round_up_int(p.half_bytes, 128);
    p.stage_bytes = p.halves * p.half_stride;
    int off = p.NS * p.stage_bytes;
    p.off_score = off; off += M * p.SR * p.SW * 4;
    p.off_bitmap = off; off += M * p.SR * p.wpr * 4;
    off = okp_round_up_int(off, 8);
    p.PK = 2 * K;
    p.off_list = off; off += M * K * (int)sizeof(OkpStripPeak);
    off = okp_round_up_int(off, 8);
    p.off_pending = off; off += M * p.PK * (int)sizeof(OkpStripCandidate);
    p.off_count = off; off += 3 * M * 4;
    off = okp_round_up_int(off, 8);
    p.off_mbar = off; off += (2 * OKP_STRIP_MAX_NS + 2 * OKP_STRIP_MAX_LAG) * 8;
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
    const float thr_lo = threshold - OKP_STRIP_THRESHOLD_SLACK * fabsf(threshold);
    okp_peaks_strip_kernel<<<p.grid, block, p.smem_bytes, stream>>>(tmap, heat, p, threshold, thr_lo, tables);
    OKP_CUDA_CHECK(cudaGetLastError());
    return OKP_OK;
}
