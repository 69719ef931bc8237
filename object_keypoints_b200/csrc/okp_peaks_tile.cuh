// okp_peaks_tile.cuh -- K1, third form (round 2): sparse tile kernel. EXPERIMENT, compiled into the tuning build only
// (-DOKP_TUNING_KNOBS, OKP_PEAKS_TILE=1): same results as okp_peaks_stream.cuh (and the same epilogue), fewer thread
// instructions on the maps a trained network writes -- and 3.6x slower, see profiles/r02e_tile_kernel.md for why.
//
// Replaces perception/pipeline.py:46-79 + perception/models.py:55-58 for every map of a batch.
//
// Why. The stream kernel computes the bounded box sum S~ for EVERY pixel (21 FADD + the gates per 4-pixel strip row):
// 17-21 thread instructions per pixel against a budget of 23 per pixel at 100 % of the HBM peak (float32; 11.5 for
// bfloat16 maps) -- issue-bound at 0.53 of the roofline at the network's 64x64 and at 0.17 on bfloat16 maps
// (profiles/r01m_k1_64x64_ncu.md). A heatmap is almost empty, and empty regions cannot matter:
//   * with tau = threshold / 25 * (1 - 4e-5), a pixel whose 5x5 window holds no value above tau has a box sum
//     <= 25 tau (1 + gamma_24) < threshold (1 - 1e-5): it is no peak, and it cannot beat a peak in the NMS comparison
//     (a peak's box sum exceeds the threshold). Call a value ACTIVE when its bit pattern, as an unsigned integer, exceeds
//     that of tau: values above tau, and every negative value, NaN and Inf (which void the bound and are caught below);
//   * so only pixels within 2 px of an active value can be peaks, their NMS neighbours that matter lie within 4 px of it,
//     and the box sums of those need the map within 6 px of it. Everything else is looked at once and never added up.
//
// Work decomposition. A map is cut into TILES of TW x TH outputs; a tile arrives in shared memory by TMA with its 4-px
// halo ([TH + 8] x [TW + 8] elements; the tensor map's zero fill outside the image is conv2d's zero padding), so a tile
// is a self-contained problem: no state is carried between tiles and no neighbour is ever "held by another warp". A CTA
// is persistent and warp-specialised like the stream kernel (producer lane / compute warps / epilogue warps, candidate
// buffers handed over per GROUP of M maps); per tile its CW compute warps run four short phases between named barriers:
//   P1  scan   one LDS.128 + two 3-input integer maxima per 4 pixels: which 4-pixel QUADS of which 4-row groups are active;
//   L   list   8-row BLOCKS of quad columns with an active value in reach become work items (a 64-bit mask per block row);
//   P2  slide  one lane per item: down the block's rows with the running sums of the stream kernel (9 + 12 FADD per quad
//              row after 4 priming rows), S~ goes to a shared-memory plane, quads with S~ above the threshold to a hot list;
//   N   NMS    one lane per hot quad: vertical gate from the plane, then the full 5x5 test; survivors are the candidates of
//              the stream kernel's epilogue (exact raster-order box sum from L2, exact check of the neighbours inside the
//              tie band, raster ranks, centroids, table rows -- okp_stream_epilogue).
// On the bench workload that is ~3 thread instructions per pixel at 180x320 and ~6 at 64x64 (P1 1.5, the rest in
// proportion to the blobs) instead of 17-21; a dense map (every quad active) costs about what the stream kernel costs.
//
// Shapes: W % 4 == 0 (float32) / W % 8 == 0 (bfloat16), threshold > 0; any H and any W (no 500-column limit).
#pragma once
#include "okp_peaks_stream.cuh"

#define OKP_TILE_MAX_MS 8                 // maps per stage
#define OKP_TILE_HOT_CAP 1024             // hot quads per stage before a map is handed to the overflow path

struct OkpTilePlan {
    OkpStreamPlan sp;             // epilogue layout (candidate buffers ...) and sp.s.{H, W, maps, M, K, PK, IC, NS, esize}
    int TW, TH;                   // outputs per tile
    int tiles_x, tiles_y;
    int MS;                       // maps per stage: the same tile of MS consecutive maps
    int BW, SR;                   // staged tile: SR = TH + 8 rows of BW elements
    int lead;                     // elements in front of tile column -4 (TMA boxes start on 16 bytes: 0 float32, 4 bfloat16)
    int QW;                       // quads per tile row, halo quads included: TW / 4 + 2   (<= 64)
    int NBX;                      // 8-row blocks of the extended output rows -2 .. TH + 1
    int NSUB;                     // 4-row groups of staged rows: SR / 4
    int NW16;                     // 16-quad windows per row: ceil(QW / 16)
    int tile_bytes;               // SR * BW * esize
    int stage_bytes;              // MS * tile_bytes rounded up to 128
    int plane_floats;             // S~ plane of one map of the stage: SR * 4 * QW floats
    int off_plane, off_blk, off_list, off_hot, off_ctr;
    int CW;                       // compute warps
    int stages_per_group;         // (M / MS) * tiles_x * tiles_y
    int list_cap;                 // MS * NBX * QW
    uint32_t tau_bits;            // activity threshold (bit pattern; the upper half is the bfloat16 pattern)
};

// One quad (4 consecutive elements at a 16- or 8-byte aligned address) as float32 and the OR of its raw sign / >= 2.0 bits.
template <typename T> struct OkpQuad;
template <> struct OkpQuad<float> {
    static __device__ __forceinline__ float4 load(const unsigned char* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ float2 load_hi(const unsigned char* p) { return reinterpret_cast<const float2*>(p)[1]; }
    static __device__ __forceinline__ float2 load_lo(const unsigned char* p) { return reinterpret_cast<const float2*>(p)[0]; }
    static __device__ __forceinline__ uint32_t bits(const float4 v) {
        return (__float_as_uint(v.x) | __float_as_uint(v.y)) | (__float_as_uint(v.z) | __float_as_uint(v.w));
    }
};
template <> struct OkpQuad<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const unsigned char* p) {
        const uint2 a = *reinterpret_cast<const uint2*>(p);
        return make_float4(__uint_as_float(a.x << 16), __uint_as_float(a.x & 0xFFFF0000u), __uint_as_float(a.y << 16),
                           __uint_as_float(a.y & 0xFFFF0000u));
    }
    static __device__ __forceinline__ float2 load_hi(const unsigned char* p) {
        const uint32_t a = reinterpret_cast<const uint32_t*>(p)[1];
        return make_float2(__uint_as_float(a << 16), __uint_as_float(a & 0xFFFF0000u));
    }
    static __device__ __forceinline__ float2 load_lo(const unsigned char* p) {
        const uint32_t a = reinterpret_cast<const uint32_t*>(p)[0];
        return make_float2(__uint_as_float(a << 16), __uint_as_float(a & 0xFFFF0000u));
    }
    static __device__ __forceinline__ uint32_t bits(const float4 v) {
        return (__float_as_uint(v.x) | __float_as_uint(v.y)) | (__float_as_uint(v.z) | __float_as_uint(v.w));
    }
};

// P1: the unsigned maximum of the bit patterns of the quad at p and the quad two rows below it.
template <typename T> __device__ __forceinline__ uint32_t okp_tile_scan2(const unsigned char* p, int two_rows);
template <> __device__ __forceinline__ uint32_t okp_tile_scan2<float>(const unsigned char* p, int two_rows) {
    const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + two_rows);
    return max(max(max(a.x, a.y), max(a.z, a.w)), max(max(b.x, b.y), max(b.z, b.w)));
}
template <> __device__ __forceinline__ uint32_t okp_tile_scan2<__nv_bfloat16>(const unsigned char* p, int two_rows) {
    const uint2 a = *reinterpret_cast<const uint2*>(p), b = *reinterpret_cast<const uint2*>(p + two_rows);
    const uint32_t m = __vmaxu2(__vmaxu2(a.x, a.y), __vmaxu2(b.x, b.y));      // per 16-bit half
    return max(m << 16, m & 0xFFFF0000u);                                      // compared with tau_bits' upper half
}

template <typename T, bool FUSED>
__global__ void __maxnreg__(96)
okp_peaks_tile_kernel(const __grid_constant__ CUtensorMap tmap, const T* __restrict__ heat, const __grid_constant__ OkpTilePlan tp,
                      float threshold, float thr_lo, const __grid_constant__ OkpDecodeTables t,
                      const __grid_constant__ OkpGroupArgs ga) {
    extern __shared__ __align__(128) unsigned char smem[];
    const OkpStreamPlan& sp = tp.sp;
    const OkpStripPlan& p = sp.s;
    const int NS = p.NS;
    int* n_peaks = reinterpret_cast<int*>(smem + sp.off_misc);                                 // [M] (+ n_items) (epilogue)
    int* pend = n_peaks + p.M + 1 + p.M + 1;                                                   // [M] (see okp_stream_epilogue)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + sp.off_mbar);       // [NS] TMA landed
    uint64_t* done = full + OKP_STRIP_MAX_NS;                               // [NS] the compute warps have left the staged tile
    uint64_t* cand_full = done + OKP_STRIP_MAX_NS;                          // [2] candidate buffer complete
    uint64_t* cand_free = cand_full + 2;                                    // [2] epilogue finished with the buffer

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int CW = tp.CW;
    const int H = p.H, W = p.W;

    float* plane = reinterpret_cast<float*>(smem + tp.off_plane);                              // [MS][SR][4 QW] S~
    uint32_t* blk = reinterpret_cast<uint32_t*>(smem + tp.off_blk);                            // [MS][NBX][2] active quads in reach of a block
    uint16_t* list = reinterpret_cast<uint16_t*>(smem + tp.off_list);                          // work items (m, block, quad)
    uint32_t* hot = reinterpret_cast<uint32_t*>(smem + tp.off_hot);                            // (m, extended row, quad)
    int* ctr = reinterpret_cast<int*>(smem + tp.off_ctr);                                      // [0] items, [1] hot quads

    for (int b = 0; b < 2; ++b) {
        int* count = reinterpret_cast<int*>(smem + sp.off_count[b]);
        for (int i = tid; i < 2 * p.M; i += blockDim.x) count[i] = 0;
    }
    for (int i = tid; i < p.M + 1; i += blockDim.x) n_peaks[i] = 0;
    for (int i = tid; i < p.M; i += blockDim.x) pend[i] = 0;
    for (int i = tid; i < tp.MS * tp.NBX * 2; i += blockDim.x) blk[i] = 0;
    if (tid < 2) ctr[tid] = 0;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { okp_mbar_init(full + i, 1); okp_mbar_init(done + i, CW); }
        for (int i = 0; i < 2; ++i) { okp_mbar_init(cand_full + i, CW); okp_mbar_init(cand_free + i, sp.EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int my_groups = blockIdx.x < sp.groups ? (sp.groups - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int tiles_per_map = tp.tiles_x * tp.tiles_y;

    if (warp == CW) {
        // ------------------------------- producer: one lane, the tiles of consecutive groups back to back ---
        if (lane == 0) {
            const CUtensorMap* tmap_ptr = &tmap;
            int stage = 0;
            uint32_t parity = 0;
            long long q = 0;
            for (int it = 0; it < my_groups; ++it) {
                const int first_map = (blockIdx.x + it * gridDim.x) * p.M;
                for (int ms = 0; ms < p.M; ms += tp.MS) {
                    for (int ty = 0; ty < tp.tiles_y; ++ty) {
                        for (int tx = 0; tx < tp.tiles_x; ++tx, ++q) {
                            if (q >= NS) {
                                while (!okp_mbar_try_wait_suspend(done + stage, parity)) {}
                            }
                            uint64_t* bar = full + stage;
                            okp_mbar_expect_tx(bar, (uint32_t)(tp.MS * tp.tile_bytes));
                            okp_tma_load_3d(smem + (size_t)stage * tp.stage_bytes, tmap_ptr, tx * tp.TW - 4 - tp.lead, ty * tp.TH - 4,
                                            first_map + ms, bar);
                            if (++stage == NS) { stage = 0; if (q >= NS) parity ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp < CW) {
        // ------------------------------- compute warps: scan, list, slide, NMS per staged tile -------------
        const int cthreads = CW * 32;
        const int esize = (int)sizeof(T);
        const int row_bytes = tp.BW * esize;
        const int QW = tp.QW, SR = tp.SR, NBX = tp.NBX;
        const int plane_pitch = 4 * QW;                                // floats per plane row
        const unsigned lt = (1u << lane) - 1u;
        const float tie = OKP_STRIP_TIE;
        int stage = 0;
        uint32_t full_parity = 0;
        for (int it = 0; it < my_groups; ++it) {
            const int buf = it & 1;
            if (it >= 2) okp_mbar_wait(cand_free + buf, (uint32_t)(((it >> 1) - 1) & 1));   // the epilogue released the buffer
            int* count = reinterpret_cast<int*>(smem + sp.off_count[buf]);                  // [M] candidates, [M] redo
            OkpStripCandidate* pending = reinterpret_cast<OkpStripCandidate*>(smem + sp.off_pending[buf]);
            for (int ms = 0; ms < p.M; ms += tp.MS) {
                for (int tile = 0; tile < tiles_per_map; ++tile) {
                    const int ty = tile / tp.tiles_x, tx = tile - ty * tp.tiles_x;
                    const int ty0 = ty * tp.TH, tx0 = tx * tp.TW;
                    okp_mbar_wait(full + stage, full_parity);
                    const unsigned char* raw = smem + (size_t)stage * tp.stage_bytes;

                    // ---- P1: active quads per 4-row group. A warp looks at 16 quads x 4 rows at a time: lane = (quad, row
                    // parity), one LDS.128 for each of its two rows; the 16-bit vote goes to the blocks within reach. Halo
                    // quads and rows outside the image hold the zero fill and are skipped ----
                    {
                        const int q16 = lane & 15, rp = lane >> 4;
                        const int qs = tx0 == 0 ? 1 : 0, qe = tx0 + tp.TW >= W ? QW - 1 : QW;
                        const int ss = ty0 == 0 ? 1 : 0, se = okp_min(tp.NSUB, (H - ty0 + 7) >> 2);
                        const int myq = qs + q16;
                        for (int m = 0; m < tp.MS; ++m) {
                            for (int s = ss + warp; s < se; s += CW) {
                                const unsigned char* src = raw + (size_t)m * tp.tile_bytes + (size_t)(4 * s + rp) * row_bytes +
                                                           (tp.lead + 4 * myq) * esize;
                                for (int q0 = myq; q0 - q16 < qe; q0 += 16, src += 64 * esize) {
                                    const uint32_t mx = okp_tile_scan2<T>(src, 2 * row_bytes);
                                    const unsigned b = __ballot_sync(0xffffffffu, q0 < qe && mx > tp.tau_bits);
                                    if (b != 0u && lane == 0) {
                                        // bit i of the vote = quad q0 + i of this window (lane 0: q0 = qs + 16 w)
                                        const uint64_t bits = (uint64_t)((b | (b >> 16)) & 0xFFFFu) << q0;
                                        // raw rows 4s .. 4s+3 are within reach of block bx (raw rows 8bx-2 .. 8bx+13) for 2bx-1 <= s <= 2bx+3
                                        const int lo = s >= 2 ? (s - 2) >> 1 : 0, hi = okp_min(NBX - 1, (s + 1) >> 1);
                                        for (int bx = lo; bx <= hi; ++bx) {
                                            uint32_t* dst = blk + (m * NBX + bx) * 2;
                                            if ((uint32_t)bits) atomicOr(dst, (uint32_t)bits);
                                            if ((uint32_t)(bits >> 32)) atomicOr(dst + 1, (uint32_t)(bits >> 32));
                                        }
                                    }
                                }
                            }
                        }
                    }
                    okp_named_barrier(2, cthreads);
                    if (tid == 0) ctr[1] = 0;                                      // every thread has finished the previous tile's N

                    // ---- L: items = (map of the stage, block, quad) whose quad or a neighbouring quad is active in reach ----
                    for (int mb = warp; mb < tp.MS * NBX; mb += CW) {
                        const uint64_t d = (uint64_t)blk[2 * mb] | ((uint64_t)blk[2 * mb + 1] << 32);
                        if (d == 0) continue;                                     // uniform
                        uint64_t dil = d | (d << 1) | (d >> 1);
                        if (QW < 64) dil &= (1ull << QW) - 1ull;
                        const int m = mb / NBX, bx = mb - m * NBX;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const uint32_t wbits = (uint32_t)(dil >> (32 * half));
                            if (wbits == 0u) continue;                            // uniform
                            int base = 0;
                            if (lane == 0) base = atomicAdd(ctr, __popc(wbits));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if ((wbits >> lane) & 1u)
                                list[base + __popc(wbits & lt)] = (uint16_t)((32 * half + lane) | (bx << 7) | (m << 12));
                        }
                    }
                    okp_named_barrier(2, cthreads);
                    for (int i = tid; i < tp.MS * NBX * 2; i += cthreads) blk[i] = 0;      // read above; next written behind the barrier below

                    // ---- P2: one lane per item slides down its block: running sums like okp_strip_step, S~ into the plane ----
                    const int n_list = ctr[0];
                    for (int i = tid; i < n_list; i += cthreads) {
                        const int item = list[i];
                        const int qi = item & 127, bx = (item >> 7) & 31, m = item >> 12;
                        const int ql = qi > 0 ? qi - 1 : 0, qr = qi + 1 < QW ? qi + 1 : QW - 1;
                        const unsigned char* row = raw + (size_t)m * tp.tile_bytes + (size_t)(8 * bx) * row_bytes + tp.lead * esize;
                        float* out = plane + (size_t)m * tp.plane_floats + (size_t)(8 * bx + 2) * plane_pitch + 4 * qi;
                        const int cl = 4 * ql * esize, cc = 4 * qi * esize, cr = 4 * qr * esize;
                        // outputs: extended rows e = 8 bx + k, k = 0..7 (image row ty0 + e - 2), from the staged rows e .. e + 4
                        const int rows = okp_min(12, SR - 8 * bx);                // staged rows this item may read
                        const bool real_quad = qi >= 1 && qi <= QW - 2 && tx0 + 4 * (qi - 1) < W;
                        float pa[4], pb[4], pc[4], pd[4], hp[4];                  // pair sums P(t-3) .. P(t), h(t-1)
#pragma unroll
                        for (int c = 0; c < 4; ++c) { pa[c] = pb[c] = pc[c] = pd[c] = 0.0f; hp[c] = 0.0f; }
                        uint32_t or_bits = 0;
#pragma unroll
                        for (int k = 0; k < 12; ++k) {
                            if (k < rows) {
                                const float2 l = OkpQuad<T>::load_hi(row + cl);
                                const float4 cq = OkpQuad<T>::load(row + cc);
                                const float2 r = OkpQuad<T>::load_lo(row + cr);
                                row += row_bytes;
                                or_bits |= OkpQuad<T>::bits(cq);
                                // horizontal 5-sums of the windows a[c .. c+4], a = (l.x, l.y, cq.xyzw, r.x, r.y): 9 adds
                                const float c34 = cq.y + cq.z;
                                const float t12 = l.y + cq.x;
                                const float t56 = cq.w + r.x;
                                const float tc = t12 + c34;
                                const float uu = c34 + t56;
                                float h[4];
                                h[0] = l.x + tc; h[1] = tc + cq.w; h[2] = cq.x + uu; h[3] = uu + r.y;
                                float sv[4];
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    sv[c] = (pb[c] + pd[c]) + h[c];               // S~(t) = P(t-3) + P(t-1) + h(t)
                                    pa[c] = pb[c]; pb[c] = pc[c]; pc[c] = pd[c];
                                    pd[c] = hp[c] + h[c];
                                    hp[c] = h[c];
                                }
                                if (k >= 4) {
                                    *reinterpret_cast<float4*>(out) = make_float4(sv[0], sv[1], sv[2], sv[3]);
                                    out += plane_pitch;
                                    const int e = 8 * bx + k - 4;
                                    const int gy = ty0 + e - 2;
                                    if (real_quad && e >= 2 && e < tp.TH + 2 && gy < H &&
                                        fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3])) > thr_lo) {
                                        const int slot = atomicAdd(ctr + 1, 1);
                                        if (slot < OKP_TILE_HOT_CAP) hot[slot] = (uint32_t)qi | ((uint32_t)e << 7) | ((uint32_t)m << 15);
                                        else count[p.M + ms + m] = 1;             // too many: the map goes to the overflow path
                                    }
                                }
                            }
                        }
                        // a negative value, a NaN, an Inf or a value >= 2.0 in reach voids the bound: the map is redone exactly
                        if (or_bits >> 30) count[p.M + ms + m] = 1;                // benign race: every writer stores 1
                    }
                    okp_named_barrier(2, cthreads);
                    if (lane == 0) okp_mbar_arrive(done + stage);                  // the staged tile may be overwritten
                    if (tid == 0) ctr[0] = 0;                                      // read by everyone before the barrier above

                    // ---- N: one lane per hot quad: vertical gate, then the 5x5 neighbourhood from the plane ----
                    const int n_hot = okp_min(ctr[1], OKP_TILE_HOT_CAP);
                    for (int i = tid; i < n_hot; i += cthreads) {
                        const uint32_t entry = hot[i];
                        const int qi = entry & 127, e = (entry >> 7) & 255, m = entry >> 15;
                        const int gy = ty0 + e - 2, gx0 = tx0 + 4 * (qi - 1);
                        const float* centre = plane + (size_t)m * tp.plane_floats + (size_t)(e + 2) * plane_pitch + 4 * qi;
                        const float ninf = -INFINITY;
                        const float4 b4 = *reinterpret_cast<const float4*>(centre);
                        const float4 u4 = *reinterpret_cast<const float4*>(centre - plane_pitch);
                        const float4 d4 = *reinterpret_cast<const float4*>(centre + plane_pitch);
                        const bool up1 = gy >= 1, dn1 = gy + 1 < H;
                        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
                        const float uv[4] = {u4.x, u4.y, u4.z, u4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
                        int possible = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float v = fmaxf(up1 ? uv[c] : ninf, dn1 ? dv[c] : ninf);
                            possible |= ((int)(b[c] > thr_lo) & (int)(b[c] * tie >= v)) << c;
                        }
                        if (!possible) continue;
                        // columns gx0 - 2 .. gx0 + 5 of the rows gy - 2 .. gy + 2; -inf outside the image (max_pool2d's padding)
                        float v[5][8];
#pragma unroll
                        for (int r = 0; r < 5; ++r) {
                            const float* src = centre + (r - 2) * plane_pitch;
                            const float2 l = reinterpret_cast<const float2*>(src - 4)[1];
                            const float4 cq = *reinterpret_cast<const float4*>(src);
                            const float2 rr = reinterpret_cast<const float2*>(src + 4)[0];
                            const bool row_in = gy + r - 2 >= 0 && gy + r - 2 < H;
                            v[r][0] = l.x; v[r][1] = l.y; v[r][2] = cq.x; v[r][3] = cq.y; v[r][4] = cq.z; v[r][5] = cq.w; v[r][6] = rr.x; v[r][7] = rr.y;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const bool in = row_in && gx0 + j - 2 >= 0 && gx0 + j - 2 < W;
                                v[r][j] = in ? v[r][j] : ninf;
                            }
                        }
                        float own[8], cm[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            own[j] = fmaxf(fmaxf(v[0][j], v[1][j]), fmaxf(v[3][j], v[4][j]));
                            cm[j] = fmaxf(own[j], v[2][j]);
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (!((possible >> c) & 1)) continue;
                            const int j = c + 2;
                            const float others = fmaxf(fmaxf(fmaxf(cm[j - 2], cm[j - 1]), fmaxf(cm[j + 1], cm[j + 2])), own[j]);
                            if (!(b[c] * tie >= others)) continue;               // a provably larger neighbour
                            uint32_t ties = 0;                                    // neighbours inside the tie band: exact check in the epilogue
#pragma unroll
                            for (int r = 0; r < 5; ++r) {
#pragma unroll
                                for (int dx = 0; dx < 5; ++dx) {
                                    if (r == 2 && dx == 2) continue;
                                    if (v[r][j + dx - 2] * tie >= b[c]) ties |= 1u << (r * 5 + dx);
                                }
                            }
                            const int mm = ms + m;
                            const int slot = atomicAdd(count + mm, 1);
                            if (slot < p.PK) {
                                OkpStripCandidate cd;
                                cd.key = gy * W + gx0 + c;
                                cd.ties = ties;
                                pending[(size_t)mm * p.PK + slot] = cd;
                            }
                        }
                    }
                    if (++stage == NS) { stage = 0; full_parity ^= 1u; }
                }
            }
            __syncwarp();
            if (lane == 0) okp_mbar_arrive(cand_full + buf);
        }
    } else {
        // ------------------------------- epilogue warps: one finished candidate buffer at a time ---------
        okp_stream_epilogue<T, FUSED>(smem, sp, heat, threshold, t, ga, tid - (CW + 1) * 32, my_groups, false);
    }
}

// tau: the largest float32 (bfloat16 for 2-byte maps) below threshold / 25 * (1 - 4e-5), as a bit pattern.
static inline uint32_t okp_tile_tau_bits(float threshold, int esize) {
    const double tau = (double)threshold / 25.0 * (1.0 - 4e-5);
    float f = (float)tau;
    if ((double)f > tau) f = nextafterf(f, 0.0f);
    uint32_t bits;
    memcpy(&bits, &f, 4);
    if (esize == 2) bits &= 0xFFFF0000u;                  // the largest bfloat16 <= tau (truncation rounds a positive value down)
    return bits;
}

// group_frame_bytes > 0 asks for the fused form (see okp_stream_plan). Returns false when the shape or the threshold is
// outside this kernel (the caller then uses the stream kernel).
static inline bool okp_tile_plan(int maps, int C, int H, int W, int K, int esize, float threshold, int group_frame_bytes,
                                 int lean, OkpTilePlan* out) {
    const int align = 16 / esize;
    if (maps < 1 || H < 1 || W < 4 || (W % align) != 0 || !(threshold > 1e-30f)) return false;
    OkpTilePlan tp;
    memset(&tp, 0, sizeof(tp));
    OkpStreamPlan& sp = tp.sp;
    OkpStripPlan& p = sp.s;
    p.H = H; p.W = W; p.maps = maps; p.K = K; p.esize = esize;
    p.PK = 2 * K;
    sp.C = C; sp.lean = lean;
    tp.tau_bits = okp_tile_tau_bits(threshold, esize);
    if (tp.tau_bits == 0) return false;
    const bool fused = group_frame_bytes > 0;
    // tile: the whole map when it is small, else ~160 x 36 outputs (a quarter more staged bytes than outputs)
    const int max_tw = okp_env_int("OKP_TILE_TW", 16, 240, 160), max_th = okp_env_int("OKP_TILE_TH", 8, 240, 36);
    tp.tiles_x = (W + max_tw - 1) / max_tw;
    tp.TW = okp_round_up_int((W + tp.tiles_x - 1) / tp.tiles_x, align > 4 ? align : 4);
    tp.tiles_x = (W + tp.TW - 1) / tp.TW;
    if (H <= 72) { tp.tiles_y = 1; tp.TH = okp_round_up_int(H, 4); }
    else { tp.tiles_y = (H + max_th - 1) / max_th; tp.TH = okp_round_up_int((H + tp.tiles_y - 1) / tp.tiles_y, 4); tp.tiles_y = (H + tp.TH - 1) / tp.TH; }
    tp.lead = esize == 2 ? 4 : 0;
    tp.BW = okp_round_up_int(tp.TW + 8 + tp.lead, align);
    tp.SR = tp.TH + 8;
    tp.QW = tp.TW / 4 + 2;
    if (tp.QW > 64 || tp.BW > 256 || tp.SR > 256) return false;
    tp.NBX = (tp.TH + 4 + 7) / 8;
    if (tp.NBX > 32) return false;
    tp.NSUB = tp.SR / 4;
    tp.NW16 = (tp.QW + 15) / 16;
    tp.tile_bytes = tp.SR * tp.BW * esize;
    tp.plane_floats = tp.SR * 4 * tp.QW;
    tp.CW = okp_env_int("OKP_TILE_COMPUTE_WARPS", 1, 8, 4);
    p.NS = okp_env_int("OKP_TILE_STAGES", 2, OKP_STRIP_MAX_NS, 2);
    sp.EW = okp_env_int("OKP_STREAM_EPILOGUE_WARPS", 0, 4, 0);
    if (sp.EW == 0) sp.EW = 2;
    const int budget = okp_env_int("OKP_TILE_SMEM_KB", 16, 224, H * W <= 72 * 72 ? 72 : 110) * 1024;
    // maps per stage: small maps are staged several at a time so that a pass has enough items for the CTA's lanes
    int MS = okp_env_int("OKP_TILE_MS", 0, OKP_TILE_MAX_MS, 0);
    if (MS == 0) MS = 1;
    // M: maps per group (candidate buffers are per group); a multiple of MS and, fused, of C
    int unit = MS;
    if (fused) { unit = C; while (unit % MS) unit += C; }
    int M = okp_env_int("OKP_TILE_GROUP_MAPS", 0, 256, 0);
    if (M == 0) M = unit * ((tp.tiles_x * tp.tiles_y > 1 ? 1 : (fused ? 2 : 6)));
    M = (M + unit - 1) / unit * unit;
    for (;;) {
        tp.MS = MS;
        p.M = M;
        p.IC = M * 64;
        tp.stage_bytes = okp_round_up_int(MS * tp.tile_bytes, 128);
        tp.list_cap = MS * tp.NBX * 64;
        int off = p.NS * tp.stage_bytes;
        tp.off_plane = off; off += MS * tp.plane_floats * 4;
        tp.off_blk = off; off += MS * tp.NBX * 2 * 4;
        tp.off_list = off; off += okp_round_up_int(tp.list_cap * 2, 16);
        tp.off_hot = off; off += OKP_TILE_HOT_CAP * 4;
        tp.off_ctr = off; off += 16;
        for (int b = 0; b < 2; ++b) { sp.off_pending[b] = off; off += M * p.PK * (int)sizeof(OkpStripCandidate); }
        sp.off_peaks = off; off += M * p.PK * (int)sizeof(OkpStripPeak);
        sp.off_items = off; off += p.IC * 4;
        for (int b = 0; b < 2; ++b) { sp.off_count[b] = off; off += 2 * M * 4; }
        sp.off_misc = off; off += (3 * M + 2) * 4;
        off = okp_round_up_int(off, 16);
        sp.off_group = off;
        if (fused) off += (M / C) * group_frame_bytes;
        sp.off_mbar = off; off += (2 * OKP_STRIP_MAX_NS + 4) * 8 + 32;
        sp.smem_bytes = off;
        if (off <= budget) break;
        if (M > unit) { M -= unit; continue; }
        if (p.NS > 2) { --p.NS; continue; }
        if (off <= 224 * 1024) break;
        return false;
    }
    sp.F = fused ? p.M / C : 0;
    sp.groups = (maps + p.M - 1) / p.M;
    sp.threads = (tp.CW + 1 + sp.EW) * 32;
    tp.stages_per_group = (p.M / tp.MS) * tp.tiles_x * tp.tiles_y;
    *out = tp;
    return true;
}

// ga: grouping arguments (fused form, tp.sp.F > 0) or NULL (peaks only).
template <typename T>
static inline int okp_tile_launch(const T* heat, const OkpTilePlan& tp, float threshold, const OkpDecodeTables& tables,
                                  const OkpGroupArgs* ga, cudaStream_t stream) {
    const OkpStreamPlan& sp = tp.sp;
    const OkpStripPlan& p = sp.s;
    OkpEncodeTiledFn encode = okp_encode_tiled_fn();
    if (!encode) return OKP_E_CUDA;
    if (((uintptr_t)heat & 15u) != 0) return OKP_E_UNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.maps};
    const cuuint64_t strides[2] = {(cuuint64_t)p.W * sizeof(T), (cuuint64_t)p.W * p.H * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)tp.BW, (cuuint32_t)tp.SR, (cuuint32_t)tp.MS};
    const cuuint32_t elem[3] = {1, 1, 1};
    const CUtensorMapDataType dtype = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUresult r = encode(&tmap, dtype, 3, (void*)heat, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OKP_E_CUDA;
    const bool fused = ga != nullptr && sp.F > 0;
    auto kernel = fused ? okp_peaks_tile_kernel<T, true> : okp_peaks_tile_kernel<T, false>;
    OkpGroupArgs none;
    if (!fused) memset(&none, 0, sizeof(none));
    OKP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.smem_bytes));
    int per_sm = 0;
    OKP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, sp.threads, sp.smem_bytes));
    if (per_sm < 1) return OKP_E_UNSUPPORTED;
    int device = 0, sms = 148;
    OKP_CUDA_CHECK(cudaGetDevice(&device));
    OKP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    long long grid = (long long)per_sm * sms;               // persistent: every CTA resident, groups dealt round-robin
    if (grid > sp.groups) grid = sp.groups;
    const float thr_lo = threshold - OKP_STRIP_THRESHOLD_SLACK * fabsf(threshold);
    kernel<<<(unsigned)grid, sp.threads, sp.smem_bytes, stream>>>(tmap, heat, tp, threshold, thr_lo, tables, fused ? *ga : none);
    OKP_CUDA_CHECK(cudaGetLastError());
    return OKP_OK;
}
