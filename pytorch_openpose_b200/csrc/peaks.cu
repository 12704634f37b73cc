// Gaussian smoothing + 4-neighbour NMS + ordered peak list (src/body.py:70-94).
//
// The reference smooths each of the 18 part maps with scipy.ndimage.gaussian_filter(sigma=3) in float64
// (25 taps, 'reflect' border, axis 0 then axis 1) and keeps pixels that are >= their four neighbours (zero
// outside the image) and > thre1.  Ties and plateaus make the >= comparisons sensitive to the last bit, so the
// device filter reproduces scipy's arithmetic exactly: float64, centre tap first, then the symmetric pairs
// from the far tap inwards, pair summed before the multiply, no FMA contraction (this file is compiled with
// --fmad=false and uses explicit _rn intrinsics).  One CTA smooths a 32x16 tile from a shared-memory halo
// tile; tiles whose raw maximum is <= thre1 are skipped (the weights are positive and sum to one, so the
// smoothed value cannot exceed the raw maximum of its window).
//
// Peaks are appended unordered with warp-aggregated atomics and then rank-sorted by (part, y, x), which is
// the reference's order (part-major, np.nonzero row-major); the rank is the candidate id.
#include "opb_common.cuh"

namespace opb {
namespace {

constexpr int TW = 32, TH = 16, R = kGaussRadius;
constexpr int RAW_W = TW + 2 + 2 * R;   // 58: tile + 1-pixel NMS ring + filter halo
constexpr int RAW_H = TH + 2 + 2 * R;   // 42
constexpr int SM_W = TW + 2;            // 34
constexpr int SM_H = TH + 2;            // 18

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // scipy 'reflect' (d c b a | a b c d | d c b a), valid for any offset
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i >= n ? period - 1 - i : i;
}

__global__ void __launch_bounds__(128, 5) smooth_nms_kernel(const float* __restrict__ heat, int H, int W,
                                                         const GaussTaps taps, double thre, PeakBuffers pb,
                                                         double* __restrict__ smoothed_out) {
    __shared__ double raw[RAW_H][RAW_W];
    __shared__ double ver[SM_H][RAW_W];
    __shared__ double sm[SM_H][SM_W];
    __shared__ int any_above;

    const int part = blockIdx.z;
    const float* map = heat + (size_t)part * H * W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x;

    // reflected source row / column of every halo line, computed once per CTA (the modulo is expensive)
    __shared__ int s_row[RAW_H], s_col[RAW_W];
    if (tid < RAW_H) s_row[tid] = reflect_idx(y0 - 1 - R + tid, H);
    else if (tid < RAW_H + RAW_W) s_col[tid - RAW_H] = reflect_idx(x0 - 1 - R + (tid - RAW_H), W);
    if (tid == 0) any_above = 0;
    __syncthreads();
    // The taps are positive and sum to one, so a smoothed value cannot exceed the raw maximum of its window
    // (up to ~1e-15 relative rounding): a tile whose whole halo window stays 1e-6 below thre has no peak.
    const double skip_below = thre > 0 ? thre * 0.999999 : thre * 1.000001;
    bool above = false;
    {
        // batch the halo loads (all issued before the first use) -- one-at-a-time loads left the kernel waiting on
        // global-memory latency (ncu: long-scoreboard stalls, fp64 pipe 3.5 % busy)
        constexpr int PER = (RAW_H * RAW_W + 127) / 128;       // 20 elements per thread at 128 threads
        float vals[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = tid + j * 128;
            const int ry = i / RAW_W, rx = i - ry * RAW_W;
            vals[j] = i < RAW_H * RAW_W ? __ldg(map + (size_t)s_row[ry] * W + s_col[rx]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = tid + j * 128;
            if (i < RAW_H * RAW_W) {
                const int ry = i / RAW_W, rx = i - ry * RAW_W;
                raw[ry][rx] = (double)vals[j];
                above |= !((double)vals[j] < skip_below);          // NaN counts as "above": never skipped
            }
        }
    }
    if (above) any_above = 1;
    __syncthreads();
    if (!any_above && smoothed_out == nullptr) return;

    // Both passes are register-blocked: a thread pulls a run of 9+24 inputs into registers once and produces 9
    // outputs from it (the naive form re-reads two shared-memory doubles per tap and is LDS-bound, not fp64-bound).
    // vertical pass (scipy axis 0): SM_H = 18 rows x RAW_W = 58 columns; thread = (column, half of the rows)
    constexpr int RUN = 9;
    if (tid < 2 * RAW_W) {
        const int c = tid % RAW_W, r0 = (tid / RAW_W) * RUN;
        double v[RUN + 2 * R];
#pragma unroll
        for (int i = 0; i < RUN + 2 * R; ++i) v[i] = raw[r0 + i][c];
#pragma unroll
        for (int o = 0; o < RUN; ++o) {
            double acc = __dmul_rn(v[o + R], taps.w[0]);
#pragma unroll
            for (int d = R; d >= 1; --d) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[o + R - d], v[o + R + d]), taps.w[d]));
            ver[r0 + o][c] = acc;
        }
    }
    __syncthreads();
    // horizontal pass (scipy axis 1): 18 rows x 34 columns; thread = (row, run of <= 5 columns); positions outside the
    // image are the NMS zero border
    constexpr int RUNH = 5, NSEG = 7;                              // 7 runs of 5 columns cover the 34 columns
    if (tid < SM_H * NSEG) {
        const int r = tid / NSEG, seg = tid - r * NSEG;
        const int c0 = seg * RUNH;
        const int y = y0 - 1 + r;
        double v[RUNH + 2 * R];
#pragma unroll
        for (int i = 0; i < RUNH + 2 * R; ++i) v[i] = c0 + i < RAW_W ? ver[r][c0 + i] : 0.0;
#pragma unroll
        for (int o = 0; o < RUNH; ++o) {
            const int c = c0 + o;
            if (c >= SM_W) break;
            const int x = x0 - 1 + c;
            double acc = 0.0;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                acc = __dmul_rn(v[o + R], taps.w[0]);
#pragma unroll
                for (int d = R; d >= 1; --d)
                    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[o + R - d], v[o + R + d]), taps.w[d]));
            }
            sm[r][c] = acc;
        }
    }
    __syncthreads();

    for (int i = tid; i < TH * TW; i += blockDim.x) {
        const int ty = i / TW, tx = i - ty * TW;
        const int y = y0 + ty, x = x0 + tx;
        bool peak = false;
        if (y < H && x < W) {
            const double v = sm[ty + 1][tx + 1];
            if (smoothed_out) smoothed_out[((size_t)part * H + y) * W + x] = v;
            peak = v >= sm[ty][tx + 1] && v >= sm[ty + 2][tx + 1] && v >= sm[ty + 1][tx] && v >= sm[ty + 1][tx + 2] &&
                   v > thre;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, peak);
        if (ballot) {
            const int lane = tid & 31;
            int base = 0;
            if (lane == 0) base = atomicAdd(pb.count, __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (peak) {
                const int slot = base + __popc(ballot & ((1u << lane) - 1));
                if (slot < pb.capacity) {
                    pb.keys[slot] = ((unsigned long long)part << 40) | ((unsigned long long)y << 20) | (unsigned)x;
                    pb.scores[slot] = map[(size_t)y * W + x];           // RAW score, src/body.py:89
                }
            }
        }
    }
}

// rank sort: position = number of smaller keys (keys are unique); also counts peaks per part
__global__ void __launch_bounds__(256) sort_peaks_kernel(PeakBuffers pb, int parts, int* __restrict__ part_count) {
    __shared__ unsigned long long sk[256];
    const int n = min(*pb.count, pb.capacity);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
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
        atomicAdd(&part_count[(int)(key >> 40)], 1);
    }
}

__global__ void part_prefix_kernel(const int* __restrict__ part_count, int* __restrict__ part_begin, int parts) {
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int p = 0; p < parts; ++p) {
            part_begin[p] = acc;
            acc += part_count[p];
        }
        part_begin[parts] = acc;
    }
}

// utilmx.findpeaks_torch (utilmx.py:230-241) on an already blurred float32 map: > thre (compared in float32, like torch
// does for a Python scalar), >= the four zero-padded neighbours; the score is the BLURRED value (Batch_model.py:194).
__global__ void __launch_bounds__(256) nms_f32_kernel(const float* __restrict__ maps, int H, int W, float thre, PeakBuffers pb) {
    const int part = blockIdx.z;
    const float* map = maps + (size_t)part * H * W;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    bool peak = false;
    float v = 0.f;
    if (x < W && y < H) {
        v = __ldg(map + (size_t)y * W + x);
        if (v > thre) {
            const float l = x > 0 ? __ldg(map + (size_t)y * W + x - 1) : 0.f;
            const float r = x + 1 < W ? __ldg(map + (size_t)y * W + x + 1) : 0.f;
            const float u = y > 0 ? __ldg(map + (size_t)(y - 1) * W + x) : 0.f;
            const float d = y + 1 < H ? __ldg(map + (size_t)(y + 1) * W + x) : 0.f;
            peak = v >= l && v >= r && v >= u && v >= d;
        }
    }
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
                pb.scores[slot] = v;
            }
        }
    }
}

}  // namespace

void nms_f32_launch(const float* blurred_planar, int H, int W, int parts, float thre, PeakBuffers pb, cudaStream_t stream) {
    OPB_REQUIRE(H < (1 << 20) && W < (1 << 20), "nms: image too large for the key packing");
    OPB_CUDA(cudaMemsetAsync(pb.count, 0, sizeof(int), stream));
    dim3 grid(cdiv(W, 32), cdiv(H, 8), parts);
    nms_f32_kernel<<<grid, 256, 0, stream>>>(blurred_planar, H, W, thre, pb);
    OPB_CUDA(cudaGetLastError());
}

void smooth_nms_launch(const float* heat_planar, int H, int W, int parts, double thre, PeakBuffers pb,
                       double* smoothed_out, cudaStream_t stream) {
    OPB_REQUIRE(H < (1 << 20) && W < (1 << 20), "smooth_nms: image too large for the key packing");
    OPB_CUDA(cudaMemsetAsync(pb.count, 0, sizeof(int), stream));
    dim3 grid(cdiv(W, TW), cdiv(H, TH), parts);
    smooth_nms_kernel<<<grid, 128, 0, stream>>>(heat_planar, H, W, gauss_taps_sigma3(), thre, pb, smoothed_out);
    OPB_CUDA(cudaGetLastError());
}

// part_count_scratch: device int[parts]
void sort_peaks_launch2(PeakBuffers pb, int parts, int* part_count_scratch, cudaStream_t stream) {
    OPB_CUDA(cudaMemsetAsync(part_count_scratch, 0, sizeof(int) * parts, stream));
    sort_peaks_kernel<<<cdiv(pb.capacity, 256), 256, 0, stream>>>(pb, parts, part_count_scratch);
    OPB_CUDA(cudaGetLastError());
    part_prefix_kernel<<<1, 32, 0, stream>>>(part_count_scratch, pb.part_begin, parts);
    OPB_CUDA(cudaGetLastError());
}

}  // namespace opb
