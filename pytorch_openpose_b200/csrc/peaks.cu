// Heat-map peaks: upsample/average -> smoothing -> 4-neighbour NMS -> ordered peak list (src/body.py:54-94), fused.
//
// The reference materialises heatmap_avg (H, W, 19) at frame resolution, smooths each of the 18 part maps with
// scipy.ndimage.gaussian_filter(sigma=3) in float64 (25 taps, 'reflect' border, axis 0 then axis 1) and keeps pixels
// that are >= their four neighbours (zero outside the image) and > thre1.  Here the full-resolution map is never
// written: one CTA owns a 64x32 tile of one part map of one frame and
//   1. computes the averaged map on the tile + halo straight from the low-resolution net outputs (CompositeMap:
//      x pass into shared memory, register-blocked y pass over 8-row strips -- the same fmaf chains as the
//      materialising kernels of prepost.cu, so the values are bit-identical to the planes opb_body_maps returns);
//      tiles whose low-resolution footprint already bounds the map below the threshold stop before that, tiles whose
//      window maximum is below it stop after it;
//   2. runs a float32 SCREEN of the Gaussian on deviations from the window mid-range: with
//         eps = 4e-6 * max|raw - mid|   (>= gamma_53 * max|raw - mid|)
//      the exact smoothed value G of scipy's float64 arithmetic satisfies |G - (mid + g)| <= eps, so a pixel with
//      g <= thre - mid - eps, or g < g_neighbour - 2 eps for one of its neighbours, cannot be a peak;
//   3. for the 32x16 sub-tiles that still hold a possible peak, repeats the smoothing in scipy's exact arithmetic --
//      float64, centre tap first, then the symmetric pairs from the far tap inwards, pair summed before the
//      multiply, no FMA contraction (this file is compiled with --fmad=false and uses explicit _rn intrinsics) --
//      and applies the reference's >= / > comparisons to those values.  Only step 3 emits peaks, so the result is
//      what the exact filter alone would give (ties and plateaus included).
// Mode 1 (Batch_body, srcmx/utilmx.py:230-263) replaces steps 2-3 by the 5x5 float32 blur in fixed tap order
// (== oracle blur5_fixed_order bit for bit), NMS and scoring on the blurred value.
//
// Peaks are appended unordered with warp-aggregated atomics and then rank-sorted by (part, y, x), which is the
// reference's order (part-major, np.nonzero row-major); the rank is the candidate id.
#include "composite.cuh"
#include <algorithm>

namespace opb {
namespace {

constexpr int FT_W = 64, FT_H = 32;
constexpr int kThreads = 256;
constexpr int kLoCap = 2048;           // low-resolution footprint values of all scales of one tile
constexpr int kXsRows = 28;            // x-pass rows of one scale of one tile

template <int MODE>
struct Geo {
    static constexpr int RAD = MODE == 0 ? kGaussRadius : 2;
    static constexpr int HALO = RAD + 1;                          // filter radius + the NMS ring
    static constexpr int WIN_W = FT_W + 2 * HALO;                 // 90 / 70
    static constexpr int WIN_H = FT_H + 2 * HALO;                 // 58 / 38
    static constexpr int ROW_OFF = (HALO + 7) / 8 * 8;            // raw rows start at the 8-row strip boundary y0 - ROW_OFF
    static constexpr int RAW_H = ROW_OFF + (FT_H + HALO + 7) / 8 * 8;   // 64 / 48
    static constexpr int RAW_W = WIN_W;
    static constexpr int STRIPS = RAW_H / 8;
    static constexpr int OUT_H = FT_H + 2, OUT_W = FT_W + 2;      // tile + NMS ring
    static constexpr int SLOTS = (STRIPS * RAW_W + kThreads - 1) / kThreads;
};

// exact pass geometry (one 32x16 sub-tile, 128 threads): same as the round-1 kernel
constexpr int ST_W = 32, ST_H = 16, R = kGaussRadius;
constexpr int EX_RAW_W = ST_W + 2 + 2 * R;   // 58
constexpr int EX_SM_W = ST_W + 2;            // 34
constexpr int EX_SM_H = ST_H + 2;            // 18

template <int MODE>
struct Smem {
    using G = Geo<MODE>;
    static constexpr size_t raw = 0;                                                    // float [RAW_H][RAW_W]
    static constexpr size_t maps = raw + sizeof(float) * G::RAW_H * G::RAW_W;           // int rowmap[WIN_H], colmap[WIN_W]
    static constexpr size_t misc = maps + sizeof(int) * (G::WIN_H + G::WIN_W);          // reductions, flags, per-scale footprints
    static constexpr size_t misc_bytes = 64 * 4 + kMaxScales * 5 * 4;
    static constexpr size_t uni = (misc + misc_bytes + 15) / 16 * 16;
    // phase 1: lo + xs ; phase 2: ver + out (float) ; phase 3: 2 x (ver64 + sm64)
    static constexpr size_t p1 = sizeof(float) * (kLoCap + kXsRows * G::RAW_W);
    static constexpr size_t p2 = sizeof(float) * (G::OUT_H * G::RAW_W + G::OUT_H * G::OUT_W);
    static constexpr size_t p3 = MODE == 0 ? 2 * sizeof(double) * (EX_SM_H * EX_RAW_W + EX_SM_H * EX_SM_W) : 0;
    static constexpr size_t uni_bytes = p1 > p2 ? (p1 > p3 ? p1 : p3) : (p2 > p3 ? p2 : p3);
    static constexpr size_t total = uni + uni_bytes;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // scipy 'reflect' (d c b a | a b c d | d c b a), valid for any offset
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i >= n ? period - 1 - i : i;
}
__device__ __forceinline__ int reflect101(int i, int n) {          // torch 'reflect': no edge repeat (pad < n)
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

struct PeakParams {
    MapSource src;
    const FramePost* frames;
    const unsigned* tile_mask;   // [frames][tiles_y][tiles_x] parts that can hold a peak (composite sources), or null: all
    int H, W, parts, n_frames;
    double thre;
    GaussTaps taps;
    float taps_f[kGaussRadius + 1];
    float blur[25];
    double* smoothed_out;
    int preblurred;          // mode 1 on a plane that already holds the blurred map (stage-level entry point)
};

// ---- tile bounds: which parts can exceed the threshold anywhere in a tile? -----------------------------------------
// The window of tile (tx, ty) reads, at scale s, the low-resolution rows [r0_s(ty), r1_s(ty)] x columns [c0_s(tx),
// c1_s(tx)] -- separable ranges.  One CTA per (tile column, frame): per scale it reduces min / max over the column
// range for every row (all 24 channel slots at once: the NHWC pixels are read as whole 96-byte rows), then over the row
// range of every tile of the column, and adds that scale's bound
//     max(mid * sum_min, mid * sum_max) + l1 * halfwidth      (+ float32 rounding slack)
// to the tile's per-channel bound.  Smoothing cannot raise a maximum (positive taps, sum <= 1 + 1e-15), so a part whose
// bound stays below the threshold has no peak in the tile and its bit in the tile's mask stays clear.
struct BoundParams {
    CompositeMap comp;
    int frame_base, H, W, parts, halo, tiles_x, tiles_y, max_ho;
    double cut;
    unsigned* mask;
};
__global__ void __launch_bounds__(256) tile_bounds_kernel(const __grid_constant__ BoundParams p) {
    extern __shared__ __align__(16) float4 bsm[];
    float4* colmin = bsm;                                 // [max_ho][6]
    float4* colmax = colmin + (size_t)p.max_ho * 6;       // [max_ho][6]
    float* bound = (float*)(colmax + (size_t)p.max_ho * 6);   // [tiles_y][24]
    const int tx = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    const int x0 = tx * FT_W;
    const int xlo = max(0, x0 - p.halo), xhi = min(p.W - 1, x0 + FT_W + p.halo - 1);
    for (int i = tid; i < p.tiles_y * 24; i += blockDim.x) bound[i] = 0.f;
    for (int s = 0; s < p.comp.n_scales; ++s) {
        const CompositeScale& c = p.comp.sc[s];
        const int c0 = __ldg(c.xf + xlo), c1 = min(__ldg(c.xf + xhi) + kUpTaps - 1, c.wo - 1);
        const float* base = c.src + (size_t)(frame + p.frame_base) * c.frame_stride;
        __syncthreads();
        for (int i = tid; i < c.ho * 6; i += blockDim.x) {
            const int r = i / 6, q = i - r * 6;
            const float4* px = (const float4*)(base + ((size_t)r * c.wo + c0) * c.cstride) + q;
            float4 lo = make_float4(INFINITY, INFINITY, INFINITY, INFINITY), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            for (int cc = c0; cc <= c1; ++cc, px += c.cstride / 4) {
                const float4 v = __ldg(px);
                // NaN poisons the maximum on purpose (comparisons below are written so that NaN bounds never skip)
                lo.x = fminf(lo.x, v.x); lo.y = fminf(lo.y, v.y); lo.z = fminf(lo.z, v.z); lo.w = fminf(lo.w, v.w);
                hi.x = v.x == v.x ? fmaxf(hi.x, v.x) : INFINITY;
                hi.y = v.y == v.y ? fmaxf(hi.y, v.y) : INFINITY;
                hi.z = v.z == v.z ? fmaxf(hi.z, v.z) : INFINITY;
                hi.w = v.w == v.w ? fmaxf(hi.w, v.w) : INFINITY;
            }
            colmin[i] = lo;
            colmax[i] = hi;
        }
        __syncthreads();
        for (int i = tid; i < p.tiles_y * 6; i += blockDim.x) {
            const int ty = i / 6, q = i - ty * 6;
            const int y0 = ty * FT_H;
            const int ylo = max(0, y0 - p.halo), yhi = min(p.H - 1, y0 + FT_H + p.halo - 1);
            const int r0 = __ldg(c.ybf + (ylo >> 3));
            const int r1 = __ldg(c.ybf + (yhi >> 3)) + __ldg(c.ybr + (yhi >> 3)) - 1;
            float4 lo = colmin[r0 * 6 + q], hi = colmax[r0 * 6 + q];
            for (int r = r0 + 1; r <= r1; ++r) {
                const float4 a = colmin[r * 6 + q], b = colmax[r * 6 + q];
                lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z); lo.w = fminf(lo.w, a.w);
                hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z); hi.w = fmaxf(hi.w, b.w);
            }
            const float los[4] = {lo.x, lo.y, lo.z, lo.w}, his[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float mid = 0.5f * (los[j] + his[j]), hw = 0.5f * (his[j] - los[j]);
                const float amax = fmaxf(fabsf(los[j]), fabsf(his[j]));
                bound[ty * 24 + q * 4 + j] += fmaxf(mid * c.sum_min, mid * c.sum_max) + c.l1 * hw * 1.0001f + 1e-5f * amax;
            }
        }
    }
    __syncthreads();
    for (int ty = tid; ty < p.tiles_y; ty += blockDim.x) {
        unsigned m = 0;
        for (int part = 0; part < p.parts; ++part)
            if (!((double)bound[ty * 24 + part] < p.cut)) m |= 1u << part;       // NaN bound: keep
        p.mask[((size_t)frame * p.tiles_y + ty) * p.tiles_x + tx] = m;
    }
}

__device__ __forceinline__ void block_minmax(float& vmin, float& vmax, float* red /*[16]*/) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                                   // red may still be read from a previous reduction
    if (lane == 0) {
        red[warp] = vmin;
        red[8 + warp] = vmax;
    }
    __syncthreads();
    vmin = red[0];
    vmax = red[8];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) {
        vmin = fminf(vmin, red[w]);
        vmax = fmaxf(vmax, red[8 + w]);
    }
}

__device__ __forceinline__ void emit_peak(bool peak, const PeakBuffers& pb, int part, int y, int x, float score) {
    const unsigned ballot = __ballot_sync(0xffffffffu, peak);
    if (ballot) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(pb.count, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (peak) {
            const int slot = base + __popc(ballot & ((1u << lane) - 1));
            if (slot < pb.capacity) {
                pb.keys[slot] = ((unsigned long long)part << 40) | ((unsigned long long)y << 20) | (unsigned)x;
                pb.scores[slot] = score;
            }
        }
    }
}

// shared-memory views of one CTA
struct TileSmem {
    float* raw;
    int *rowmap, *colmap;
    float* red;
    int* flags;
    int *fp_r0, *fp_nr, *fp_c0, *fp_nc, *fp_off;
    uint8_t* uni;
};

// one part map of one tile: window -> (screen) -> exact smoothing / blur -> NMS -> peaks.  Every early return is taken
// by the whole CTA.
template <int MODE>
__device__ __forceinline__ void tile_part(const PeakParams& p, const TileSmem& sm, const PeakBuffers& pb, int frame, int part,
                                          int x0, int y0, int xlo, int xhi, int ylo, int yhi) {
    using G = Geo<MODE>;
    const int tid = threadIdx.x;
    const int H = p.H, W = p.W;
    float* raw = sm.raw;
    const int* rowmap = sm.rowmap;
    const int* colmap = sm.colmap;
    int* flags = sm.flags;
    uint8_t* uni = sm.uni;
    const bool debug_all = MODE == 0 && p.smoothed_out != nullptr;
    if (tid < 4) flags[tid] = 0;

    float vmin = INFINITY, vmax = -INFINITY;
    if (p.src.planar != nullptr) {
        // ---- window from a materialised plane (stage-level entry points, fallback for unusual resize ratios)
        const float* map = p.src.planar + ((size_t)(frame + p.src.frame_base) * p.src.planes_per_frame + part) * H * W;
        constexpr int PER = (G::RAW_H * G::RAW_W + kThreads - 1) / kThreads;
        float vals[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = tid + j * kThreads;
            const int ry = i / G::RAW_W, cx = i - ry * G::RAW_W;
            const int iy = y0 - G::ROW_OFF + ry, ix = x0 - G::HALO + cx;
            const bool ok = i < G::RAW_H * G::RAW_W && iy >= ylo && iy <= yhi && ix >= xlo && ix <= xhi;
            vals[j] = ok ? __ldg(map + (size_t)iy * W + ix) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = tid + j * kThreads;
            if (i < G::RAW_H * G::RAW_W) {
                raw[i] = vals[j];
                const int ry = i / G::RAW_W, cx = i - ry * G::RAW_W;
                const int iy = y0 - G::ROW_OFF + ry, ix = x0 - G::HALO + cx;
                if (iy >= ylo && iy <= yhi && ix >= xlo && ix <= xhi) {
                    if (vals[j] == vals[j]) {
                        vmin = fminf(vmin, vals[j]);
                        vmax = fmaxf(vmax, vals[j]);
                    } else {
                        vmax = INFINITY;                           // a NaN in the map: never skip
                    }
                }
            }
        }
    } else {
        // ---- window from the low-resolution net outputs
        const CompositeMap& cm = p.src.comp;
        float* lo = (float*)uni;
        float* xs = lo + kLoCap;
        const int sb_first = ylo >> 3, sb_last = yhi >> 3;
        for (int s = 0; s < cm.n_scales; ++s) {
            const CompositeScale& c = cm.sc[s];
            const int nr = sm.fp_nr[s], nc = sm.fp_nc[s], r0 = sm.fp_r0[s], c0 = sm.fp_c0[s];
            const float* base = c.src + (size_t)(frame + p.src.frame_base) * c.frame_stride + part;
            float* dst = lo + sm.fp_off[s];
            for (int i = tid; i < nr * nc; i += kThreads) {
                const int r = i / nc, cc = i - r * nc;
                dst[i] = __ldg(base + ((size_t)(r0 + r) * c.wo + c0 + cc) * c.cstride);
            }
        }
        __syncthreads();
        float acc[G::SLOTS][8];
#pragma unroll
        for (int j = 0; j < G::SLOTS; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
        const int sb_base = (y0 - G::ROW_OFF) >> 3;       // arithmetic shift: may be negative
        for (int s = 0; s < cm.n_scales; ++s) {
            const CompositeScale& c = cm.sc[s];
            const int nr = sm.fp_nr[s], nc = sm.fp_nc[s], r0 = sm.fp_r0[s], c0 = sm.fp_c0[s];
            const float* los = lo + sm.fp_off[s];
            // x pass: rows of the footprint x window columns.  A lane owns the columns lane, lane + 32, lane + 64 and keeps
            // their tap tables in registers; the warps stride over the rows.
            {
                constexpr int CPL = (G::RAW_W + 31) / 32;
                const int lane = tid & 31, warp = tid >> 5;
                int fxr[CPL];
                float wxr[CPL][kUpTaps];
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int cx = lane + 32 * q;
                    const int ix = x0 - G::HALO + cx;
                    const bool ok = cx < G::RAW_W && ix >= xlo && ix <= xhi;
                    fxr[q] = ok ? __ldg(c.xf + ix) - c0 : -1;
#pragma unroll
                    for (int k = 0; k < kUpTaps; ++k) wxr[q][k] = ok ? __ldg(c.xw + (size_t)ix * kUpTaps + k) : 0.f;
                }
                const int last = c.wo - 1 - c0;
                for (int r = warp; r < nr; r += kThreads / 32) {
                    const float* row = los + r * nc;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const int cx = lane + 32 * q;
                        if (cx < G::RAW_W) {
                            float t = 0.f;
                            if (fxr[q] >= 0) {
#pragma unroll
                                for (int k = 0; k < kUpTaps; ++k) t = fmaf(wxr[q][k], row[min(fxr[q] + k, last)], t);
                            }
                            xs[r * G::RAW_W + cx] = t;
                        }
                    }
                }
            }
            __syncthreads();
            // y pass: (strip, column) items, 8 output rows each
#pragma unroll
            for (int j = 0; j < G::SLOTS; ++j) {
                const int item = tid + j * kThreads;
                const int b = item / G::RAW_W, cx = item - b * G::RAW_W;
                const int sb = sb_base + b;
                const int ix = x0 - G::HALO + cx;
                if (item < G::STRIPS * G::RAW_W && sb >= sb_first && sb <= sb_last && ix >= xlo && ix <= xhi) {
                    const int f0 = __ldg(c.ybf + sb), Rr = __ldg(c.ybr + sb);
                    const float4* wt = (const float4*)(c.ybw + (size_t)sb * c.yb_rs * 8);
                    const float* col = xs + (f0 - r0) * G::RAW_W + cx;
                    for (int r = 0; r < Rr; ++r) {
                        const float v = col[r * G::RAW_W];
                        const float4 wa = __ldg(wt + 2 * r), wb = __ldg(wt + 2 * r + 1);
                        acc[j][0] = fmaf(wa.x, v, acc[j][0]);
                        acc[j][1] = fmaf(wa.y, v, acc[j][1]);
                        acc[j][2] = fmaf(wa.z, v, acc[j][2]);
                        acc[j][3] = fmaf(wa.w, v, acc[j][3]);
                        acc[j][4] = fmaf(wb.x, v, acc[j][4]);
                        acc[j][5] = fmaf(wb.y, v, acc[j][5]);
                        acc[j][6] = fmaf(wb.z, v, acc[j][6]);
                        acc[j][7] = fmaf(wb.w, v, acc[j][7]);
                    }
                }
            }
            __syncthreads();
        }
        bool bad = false;
#pragma unroll
        for (int j = 0; j < G::SLOTS; ++j) {
            const int item = tid + j * kThreads;
            const int b = item / G::RAW_W, cx = item - b * G::RAW_W;
            const int ix = x0 - G::HALO + cx;
            if (item < G::STRIPS * G::RAW_W) {
                const int iy0 = y0 - G::ROW_OFF + b * 8;
                const bool col_ok = ix >= xlo && ix <= xhi;
                const int i_lo = col_ok ? ylo - iy0 : 8, i_hi = yhi - iy0;      // rows of the strip inside the window
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float v = acc[j][i];
                    raw[(b * 8 + i) * G::RAW_W + cx] = v;
                    if (i >= i_lo && i <= i_hi) {
                        vmin = fminf(vmin, v);                       // fminf / fmaxf drop a NaN operand ...
                        vmax = fmaxf(vmax, v);
                        bad |= v != v;                               // ... so it is tracked separately: never skip on NaN
                    }
                }
            }
        }
        if (bad) vmax = INFINITY;
    }
    block_minmax(vmin, vmax, sm.red);                    // also the barrier that publishes raw / flags

    if (MODE == 1) {
        // ---- 5x5 float32 blur (reflect-101 border), NMS and score on the blurred value
        const float thre = (float)p.thre;
        if (thre > 0.f && !(vmax > thre)) return;        // positive taps summing to < 1: blurred <= max(raw) <= thre
        float* outf = (float*)uni;
        for (int i = tid; i < G::OUT_H * G::OUT_W; i += kThreads) {
            const int r = i / G::OUT_W, c = i - r * G::OUT_W;
            const int y = y0 - 1 + r, x = x0 - 1 + c;
            float acc = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W && p.preblurred) {
                acc = raw[rowmap[r + 2] * G::RAW_W + colmap[c + 2]];
            } else if (y >= 0 && y < H && x >= 0 && x < W) {
#pragma unroll
                for (int dy = 0; dy < 5; ++dy) {
                    const float* row = raw + rowmap[r + dy] * G::RAW_W;
#pragma unroll
                    for (int dx = 0; dx < 5; ++dx) acc = __fadd_rn(acc, __fmul_rn(p.blur[dy * 5 + dx], row[colmap[c + dx]]));
                }
            }
            outf[i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < FT_H * FT_W; i += kThreads) {
            const int ty = i / FT_W, tx = i - ty * FT_W;
            const int y = y0 + ty, x = x0 + tx;
            bool peak = false;
            float v = 0.f;
            if (y < H && x < W) {
                const float* o = outf + (ty + 1) * G::OUT_W + tx + 1;
                v = o[0];
                peak = v > thre && v >= o[-1] && v >= o[1] && v >= o[-G::OUT_W] && v >= o[G::OUT_W];
            }
            emit_peak(peak, pb, part, y, x, v);
        }
        return;
    }

    // ---- mode 0 ----
    // The taps are positive and sum to one, so a smoothed value cannot exceed the raw maximum of its window (up to
    // ~1e-15 relative rounding): a tile whose whole window stays 1e-6 below thre has no peak.
    const double skip_below = p.thre > 0 ? p.thre * 0.999999 : p.thre * 1.000001;
    if (!debug_all && (double)vmax < skip_below) return;

    if (!debug_all) {
        // float32 screen on deviations from the mid-range
        const float mid = 0.5f * (vmin + vmax);
        const float dmax = fmaxf(vmax - mid, mid - vmin);
        const bool finite = dmax < INFINITY && mid == mid;
        if (finite) {
            float* ver = (float*)uni;                              // [OUT_H][RAW_W]
            float* outf = ver + G::OUT_H * G::RAW_W;               // [OUT_H][OUT_W]
            // vertical: column c, half of the 34 rows
            constexpr int VRUN = G::OUT_H / 2;                     // 17
            if (tid < 2 * G::RAW_W) {
                const int c = tid % G::RAW_W, r0 = (tid / G::RAW_W) * VRUN;
                const int cb = colmap[c];
                float v[VRUN + 2 * R];
#pragma unroll
                for (int i = 0; i < VRUN + 2 * R; ++i) v[i] = raw[rowmap[r0 + i] * G::RAW_W + cb] - mid;
#pragma unroll
                for (int o = 0; o < VRUN; ++o) {
                    float a = p.taps_f[0] * v[o + R];
#pragma unroll
                    for (int d = 1; d <= R; ++d) a = fmaf(p.taps_f[d], v[o + R - d] + v[o + R + d], a);
                    ver[(r0 + o) * G::RAW_W + c] = a;
                }
            }
            __syncthreads();
            // horizontal: row r, run of 11 of the 66 columns; positions outside the image hold the zero border (-mid)
            constexpr int HRUN = 11, NSEG = G::OUT_W / HRUN;       // 6
            if (tid < G::OUT_H * NSEG) {
                const int r = tid / NSEG, c0 = (tid - r * NSEG) * HRUN;
                const int y = y0 - 1 + r;
                float v[HRUN + 2 * R];
#pragma unroll
                for (int i = 0; i < HRUN + 2 * R; ++i) v[i] = ver[r * G::RAW_W + c0 + i];
#pragma unroll
                for (int o = 0; o < HRUN; ++o) {
                    const int x = x0 - 1 + c0 + o;
                    float a = -mid;
                    if (y >= 0 && y < H && x >= 0 && x < W) {
                        a = p.taps_f[0] * v[o + R];
#pragma unroll
                        for (int d = 1; d <= R; ++d) a = fmaf(p.taps_f[d], v[o + R - d] + v[o + R + d], a);
                    }
                    outf[r * G::OUT_W + c0 + o] = a;
                }
            }
            __syncthreads();
            const double eps = (double)dmax * 4e-6 + (double)fmaxf(fabsf(vmin), fabsf(vmax)) * 1e-13 + 1e-300;
            const double cut = p.thre - (double)mid - eps;
            for (int i = tid; i < FT_H * FT_W; i += kThreads) {
                const int ty = i / FT_W, tx = i - ty * FT_W;       // a warp covers 32 consecutive columns of one row
                const int y = y0 + ty, x = x0 + tx;
                bool cand = false;
                if (y < H && x < W) {
                    const float* o = outf + (ty + 1) * G::OUT_W + tx + 1;
                    const double v = (double)o[0] + 2.0 * eps;
                    cand = (double)o[0] > cut && v >= (double)o[-1] && v >= (double)o[1] && v >= (double)o[-G::OUT_W] &&
                           v >= (double)o[G::OUT_W];
                }
                if (__any_sync(0xffffffffu, cand) && (tid & 31) == 0) flags[(ty >> 4) * 2 + (tx >> 5)] = 1;
            }
        } else if (tid < 4) {
            flags[tid] = 1;                                        // NaN / Inf in the window: exact arithmetic decides
        }
        __syncthreads();
    } else {
        if (tid < 4) flags[tid] = 1;
        __syncthreads();
    }

    // ---- exact float64 pass over the flagged 32x16 sub-tiles: two at a time, 128 threads each
    const int half = tid >> 7, t2 = tid & 127;
    double* ver64 = (double*)uni + half * (EX_SM_H * EX_RAW_W + EX_SM_H * EX_SM_W);
    double* sm64 = ver64 + EX_SM_H * EX_RAW_W;
    for (int pass = 0; pass < 2; ++pass) {
        if (!(flags[pass * 2] | flags[pass * 2 + 1])) continue;        // uniform over the CTA
        const int sub = pass * 2 + half;
        const int ox = (sub & 1) * ST_W, oy = (sub >> 1) * ST_H;
        const bool active = flags[sub] != 0 && x0 + ox < W && y0 + oy < H;
        // sub-tile window (ry, rx) <-> tile window (oy + ry, ox + rx)
        constexpr int RUN = 9;
        if (active && t2 < 2 * EX_RAW_W) {
            const int c = t2 % EX_RAW_W, r0 = (t2 / EX_RAW_W) * RUN;
            const int cb = colmap[ox + c];
            double v[RUN + 2 * R];
#pragma unroll
            for (int i = 0; i < RUN + 2 * R; ++i) v[i] = (double)raw[rowmap[oy + r0 + i] * G::RAW_W + cb];
#pragma unroll
            for (int o = 0; o < RUN; ++o) {
                double acc = __dmul_rn(v[o + R], p.taps.w[0]);
#pragma unroll
                for (int d = R; d >= 1; --d) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[o + R - d], v[o + R + d]), p.taps.w[d]));
                ver64[(r0 + o) * EX_RAW_W + c] = acc;
            }
        }
        __syncthreads();
        constexpr int RUNH = 5, NSEG = 7;                              // 7 runs of 5 columns cover the 34 columns
        if (active && t2 < EX_SM_H * NSEG) {
            const int r = t2 / NSEG, seg = t2 - r * NSEG;
            const int c0 = seg * RUNH;
            const int y = y0 + oy - 1 + r;
            double v[RUNH + 2 * R];
#pragma unroll
            for (int i = 0; i < RUNH + 2 * R; ++i) v[i] = c0 + i < EX_RAW_W ? ver64[r * EX_RAW_W + c0 + i] : 0.0;
#pragma unroll
            for (int o = 0; o < RUNH; ++o) {
                const int c = c0 + o;
                if (c >= EX_SM_W) break;
                const int x = x0 + ox - 1 + c;
                double acc = 0.0;
                if (y >= 0 && y < H && x >= 0 && x < W) {
                    acc = __dmul_rn(v[o + R], p.taps.w[0]);
#pragma unroll
                    for (int d = R; d >= 1; --d)
                        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[o + R - d], v[o + R + d]), p.taps.w[d]));
                }
                sm64[r * EX_SM_W + c] = acc;
            }
        }
        __syncthreads();
        for (int i = t2; i < ST_H * ST_W; i += 128) {
            const int ty = i / ST_W, tx = i - ty * ST_W;
            const int y = y0 + oy + ty, x = x0 + ox + tx;
            bool peak = false;
            float score = 0.f;
            if (active && y < H && x < W) {
                const double* o = sm64 + (ty + 1) * EX_SM_W + tx + 1;
                const double v = o[0];
                if (p.smoothed_out) p.smoothed_out[((size_t)part * H + y) * W + x] = v;
                peak = v >= o[-EX_SM_W] && v >= o[EX_SM_W] && v >= o[-1] && v >= o[1] && v > p.thre;
                score = raw[(G::ROW_OFF + oy + ty) * G::RAW_W + G::HALO + ox + tx];      // RAW score, src/body.py:89
            }
            emit_peak(peak, pb, part, y, x, score);
        }
        __syncthreads();
    }
}

// one CTA = one 64x32 tile of one frame; it walks the parts its mask names
template <int MODE>
__global__ void __launch_bounds__(kThreads, 3) find_peaks_kernel(const __grid_constant__ PeakParams p) {
    using G = Geo<MODE>;
    using S = Smem<MODE>;
    extern __shared__ __align__(16) uint8_t smem[];
    // one CTA per (tile, frame) walks the parts its mask names (composite sources: usually none or one; materialised
    // planes have no mask and walk all parts -- measured faster than one CTA per part: the tile set-up is shared)
    const int frame = blockIdx.z;
    unsigned mask = p.parts >= 32 ? 0xffffffffu : (1u << p.parts) - 1;
    if (p.tile_mask) mask &= __ldg(p.tile_mask + ((size_t)frame * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    if (mask == 0) return;

    TileSmem sm;
    sm.raw = (float*)(smem + S::raw);
    sm.rowmap = (int*)(smem + S::maps);
    sm.colmap = sm.rowmap + G::WIN_H;
    sm.red = (float*)(smem + S::misc);                 // [16]
    sm.flags = (int*)(sm.red + 16);                    // [4] sub-tile flags
    sm.fp_r0 = sm.flags + 16;                          // per scale: first row, rows, first col, cols, lo offset
    sm.fp_nr = sm.fp_r0 + kMaxScales;
    sm.fp_c0 = sm.fp_nr + kMaxScales;
    sm.fp_nc = sm.fp_c0 + kMaxScales;
    sm.fp_off = sm.fp_nc + kMaxScales;
    sm.uni = smem + S::uni;

    const int tid = threadIdx.x;
    const int H = p.H, W = p.W;
    const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
    const int ylo = max(0, y0 - G::HALO), yhi = min(H - 1, y0 + FT_H + G::HALO - 1);
    const int xlo = max(0, x0 - G::HALO), xhi = min(W - 1, x0 + FT_W + G::HALO - 1);

    // window row / column -> raw buffer row / column through the border reflection.  Rows and columns of a border tile
    // that lie beyond the image reflect to positions outside the window; they only feed outputs outside the image
    // (which are the zero border, never computed), so they are clamped into the buffer.
    for (int i = tid; i < G::WIN_H + G::WIN_W; i += kThreads) {
        if (i < G::WIN_H) {
            const int iy = y0 - G::HALO + i;
            const int r = (MODE == 0 ? reflect_idx(iy, H) : reflect101(iy, H)) - (y0 - G::ROW_OFF);
            sm.rowmap[i] = min(max(r, 0), G::RAW_H - 1);
        } else {
            const int ix = x0 - G::HALO + (i - G::WIN_H);
            const int c = (MODE == 0 ? reflect_idx(ix, W) : reflect101(ix, W)) - (x0 - G::HALO);
            sm.colmap[i - G::WIN_H] = min(max(c, 0), G::RAW_W - 1);
        }
    }
    if (p.src.planar == nullptr && tid == 0) {
        const CompositeMap& cm = p.src.comp;
        int off = 0;
        for (int s = 0; s < cm.n_scales; ++s) {
            const CompositeScale& c = cm.sc[s];
            const int r0 = c.ybf[ylo >> 3];
            const int r1 = c.ybf[yhi >> 3] + c.ybr[yhi >> 3] - 1;
            const int c0 = c.xf[xlo];
            const int c1 = min(c.xf[xhi] + kUpTaps - 1, c.wo - 1);
            sm.fp_r0[s] = r0;
            sm.fp_nr[s] = r1 - r0 + 1;
            sm.fp_c0[s] = c0;
            sm.fp_nc[s] = c1 - c0 + 1;
            sm.fp_off[s] = off;
            off += (r1 - r0 + 1) * (c1 - c0 + 1);
        }
    }
    const PeakBuffers pb = p.frames[frame].pb;
    for (int part = 0; part < p.parts; ++part) {
        if (!((mask >> part) & 1)) continue;
        __syncthreads();                                 // tables visible; previous part's shared memory is free
        tile_part<MODE>(p, sm, pb, frame, part, x0, y0, xlo, xhi, ylo, yhi);
    }
}

// rank sort: position = number of smaller keys (keys are unique); the last block to finish turns the per-part counts
// into part_begin
__global__ void __launch_bounds__(256) order_peaks_kernel(const FramePost* __restrict__ frames, int parts) {
    __shared__ unsigned long long sk[256];
    __shared__ int s_last;
    const PeakBuffers pb = frames[blockIdx.y].pb;
    const int n = min(*pb.count, pb.capacity);
    const int active_blocks = max(1, (n + 255) / 256);
    if ((int)blockIdx.x >= active_blocks) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long key = i < n ? pb.keys[i] : ~0ull;
    int rank = 0;
    for (int base = 0; base < n; base += 256) {
        const int j = base + threadIdx.x;
        sk[threadIdx.x] = j < n ? pb.keys[j] : ~0ull;
        __syncthreads();
        const int m = min(256, n - base);
        for (int t = 0; t < m; ++t) rank += sk[t] < key;
        __syncthreads();
    }
    if (i < n) {
        double* c = pb.candidates + (size_t)rank * 4;
        c[0] = (double)(unsigned)(key & 0xFFFFF);
        c[1] = (double)(unsigned)((key >> 20) & 0xFFFFF);
        c[2] = (double)pb.scores[i];
        c[3] = (double)rank;                                            // cumulative peak id, src/body.py:90
        atomicAdd(&pb.part_count[(int)(key >> 40)], 1);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(pb.ticket, 1) == active_blocks - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        int acc = 0;
        for (int q = 0; q < parts; ++q) {
            pb.part_begin[q] = acc;
            acc += ((volatile int*)pb.part_count)[q];
        }
        pb.part_begin[parts] = acc;
    }
}

// utilmx.GaussianBlurConv (utilmx.py:243-263) on materialised planes (opb_batch_maps, Batch_hand): depthwise 5x5 on a
// reflect-padded map, fixed operation order (taps in row-major order, product and sum rounded separately) == the
// oracle's blur5_fixed_order bit for bit; torch adds the same 25 products in an unspecified order.
struct Blur5 { float w[25]; };
constexpr int BW = 32, BH = 8;
__global__ void __launch_bounds__(BW * BH) blur5_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                        const Blur5 k) {
    __shared__ float tile[BH + 4][BW + 4];
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH;
    const int tid = threadIdx.y * BW + threadIdx.x;
    for (int i = tid; i < (BH + 4) * (BW + 4); i += BW * BH) {
        const int ry = i / (BW + 4), rx = i - ry * (BW + 4);
        tile[ry][rx] = __ldg(in + plane + (size_t)reflect101(y0 - 2 + ry, H) * W + reflect101(x0 - 2 + rx, W));
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) acc = __fadd_rn(acc, __fmul_rn(k.w[dy * 5 + dx], tile[threadIdx.y + dy][threadIdx.x + dx]));
    out[plane + (size_t)y * W + x] = acc;
}

const float kBlur5[25] = {0.00078633f, 0.00655965f, 0.01330373f, 0.00655965f, 0.00078633f,      // utilmx.py:248-252
                          0.00655965f, 0.05472157f, 0.11098164f, 0.05472157f, 0.00655965f,
                          0.01330373f, 0.11098164f, 0.22508352f, 0.11098164f, 0.01330373f,
                          0.00655965f, 0.05472157f, 0.11098164f, 0.05472157f, 0.00655965f,
                          0.00078633f, 0.00655965f, 0.01330373f, 0.00655965f, 0.00078633f};

template <int MODE>
void launch_find_peaks(const PeakParams& p, dim3 grid, cudaStream_t stream) {
    static bool attr[64] = {};
    if (first_use_on_device(attr))
        OPB_CUDA(cudaFuncSetAttribute(find_peaks_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<MODE>::total));
    find_peaks_kernel<MODE><<<grid, kThreads, Smem<MODE>::total, stream>>>(p);
}

}  // namespace

// Does every tile's footprint of this composite map fit the fused kernel's shared-memory budget?  (Always for the
// reference's resize ratios; tiny frames at large scales have ratios near 1 and fall back to materialised planes.)
bool composite_fits_fused(const std::vector<std::vector<int>>& xf, const std::vector<std::vector<int>>& ybf,
                          const std::vector<std::vector<int>>& ybr, const std::vector<int>& wo, const std::vector<int>& yb_rs,
                          int H, int W) {
    const int S = (int)xf.size();
    for (int s = 0; s < S; ++s) {
        if (yb_rs[s] > 24) return false;
        for (size_t i = 1; i < xf[s].size(); ++i)
            if (xf[s][i] < xf[s][i - 1]) return false;
        for (size_t i = 1; i < ybf[s].size(); ++i)
            if (ybf[s][i] < ybf[s][i - 1] || ybf[s][i] + ybr[s][i] < ybf[s][i - 1] + ybr[s][i - 1]) return false;
    }
    for (int mode = 0; mode < 2; ++mode) {
        const int halo = mode == 0 ? kGaussRadius + 1 : 3;
        for (int y0 = 0; y0 < H; y0 += FT_H)
            for (int x0 = 0; x0 < W; x0 += FT_W) {
                const int ylo = std::max(0, y0 - halo), yhi = std::min(H - 1, y0 + FT_H + halo - 1);
                const int xlo = std::max(0, x0 - halo), xhi = std::min(W - 1, x0 + FT_W + halo - 1);
                int total = 0;
                for (int s = 0; s < S; ++s) {
                    const int r0 = ybf[s][ylo >> 3], r1 = ybf[s][yhi >> 3] + ybr[s][yhi >> 3] - 1;
                    const int c0 = xf[s][xlo], c1 = std::min(xf[s][xhi] + kUpTaps - 1, wo[s] - 1);
                    if (r1 - r0 + 1 > kXsRows) return false;
                    total += (r1 - r0 + 1) * (c1 - c0 + 1);
                }
                if (total > kLoCap) return false;
            }
    }
    return true;
}

size_t find_peaks_mask_words(int n_frames, int H, int W) {
    return (size_t)n_frames * cdiv(H, FT_H) * cdiv(W, FT_W);
}

void find_peaks_launch(const MapSource& src, int n_frames, int H, int W, int parts, int mode, double thre,
                       const FramePost* frames_dev, double* smoothed_out, unsigned* tile_mask_scratch, cudaStream_t stream) {
    OPB_REQUIRE(H < (1 << 20) && W < (1 << 20), "find_peaks: image too large for the key packing");
    OPB_REQUIRE(src.planar != nullptr || src.comp.fused_ok, "find_peaks: composite map does not fit the fused kernel");
    OPB_REQUIRE(n_frames >= 1 && n_frames * parts <= 65535 && parts >= 1 && parts <= 24, "find_peaks: bad frame / part count");
    dim3 grid(cdiv(W, FT_W), cdiv(H, FT_H), n_frames);
    PeakParams p;
    p.src = src;
    p.frames = frames_dev;
    p.tile_mask = nullptr;
    p.H = H; p.W = W; p.parts = parts; p.n_frames = n_frames;
    p.thre = thre;
    p.taps = gauss_taps_sigma3();
    for (int d = 0; d <= kGaussRadius; ++d) p.taps_f[d] = (float)p.taps.w[d];
    memcpy(p.blur, kBlur5, sizeof(p.blur));
    p.smoothed_out = smoothed_out;
    p.preblurred = mode == 2;
    if (src.planar == nullptr) {
        OPB_REQUIRE(tile_mask_scratch != nullptr, "find_peaks: composite sources need the tile-mask scratch");
        BoundParams b;
        b.comp = src.comp;
        b.frame_base = src.frame_base;
        b.H = H; b.W = W; b.parts = parts;
        b.halo = mode == 0 ? kGaussRadius + 1 : 3;
        b.tiles_x = grid.x; b.tiles_y = grid.y;
        b.max_ho = 1;
        for (int s = 0; s < src.comp.n_scales; ++s) b.max_ho = std::max(b.max_ho, src.comp.sc[s].ho);
        // mode 0: smoothed <= window maximum * (1 + 1e-15) -> 1e-6 relative margin; mode 1: blurred < window maximum
        b.cut = mode == 0 ? (thre > 0 ? thre * 0.999999 : thre * 1.000001) : (double)(float)thre;
        b.mask = tile_mask_scratch;
        const size_t smem = (size_t)b.max_ho * 12 * sizeof(float4) + (size_t)b.tiles_y * 24 * sizeof(float);
        OPB_REQUIRE(smem <= 200 * 1024, "find_peaks: net output too tall for the bounds kernel");
        static bool attr[64] = {};
        if (first_use_on_device(attr))
            OPB_CUDA(cudaFuncSetAttribute(tile_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        tile_bounds_kernel<<<dim3(grid.x, n_frames), 256, smem, stream>>>(b);
        OPB_CUDA(cudaGetLastError());
        p.tile_mask = tile_mask_scratch;
    }
    if (mode == 0) launch_find_peaks<0>(p, grid, stream);
    else launch_find_peaks<1>(p, grid, stream);
    OPB_CUDA(cudaGetLastError());
}

void order_peaks_launch(const FramePost* frames_dev, int n_frames, int max_capacity, int parts, cudaStream_t stream) {
    dim3 grid(cdiv(max_capacity, 256), n_frames);
    order_peaks_kernel<<<grid, 256, 0, stream>>>(frames_dev, parts);
    OPB_CUDA(cudaGetLastError());
}

void blur5_planar_launch(const float* maps_planar, float* out_planar, int n_maps, int H, int W, cudaStream_t stream) {
    Blur5 k;
    memcpy(k.w, kBlur5, sizeof(k.w));
    OPB_REQUIRE(n_maps <= 65535, "blur5: too many maps in one launch");
    dim3 grid(cdiv(W, BW), cdiv(H, BH), n_maps), block(BW, BH);
    blur5_kernel<<<grid, block, 0, stream>>>(maps_planar, out_planar, H, W, k);
    OPB_CUDA(cudaGetLastError());
}

}  // namespace opb
