// On-demand evaluation of the averaged full-resolution maps (src/body.py:54-68) from the low-resolution net outputs.
//
// The x8 cubic upsample, the crop and the cubic resize to the frame size are linear and separable: per axis they
// collapse into one banded operator with <= 6 taps per output index (host tables, net.cu::composite_taps).  The
// materialising kernels of prepost.cu apply them as an x pass (fmaf chain over the 6 column taps, starting from 0)
// and a y pass (one fmaf chain over all scales and their row taps, 1/n_scales folded into the row weights).  The
// functions below perform the SAME chains for a single position, so a value sampled here is bit-identical to the
// element of the materialised plane (zero-weight taps of the strip tables are exact no-ops for finite inputs).
#pragma once
#include "opb_common.cuh"

namespace opb {

// one channel at one position
__device__ __forceinline__ float composite_at(const CompositeMap& m, int frame, int ch, int y, int x) {
    float acc = 0.f;
    for (int s = 0; s < m.n_scales; ++s) {
        const CompositeScale& c = m.sc[s];
        const float* base = c.src + (size_t)frame * c.frame_stride + ch;
        const int fy = __ldg(c.yf + y), fx = __ldg(c.xf + x);
        float wx[kUpTaps];
        int col[kUpTaps];
#pragma unroll
        for (int j = 0; j < kUpTaps; ++j) {
            wx[j] = __ldg(c.xw + (size_t)x * kUpTaps + j);
            col[j] = min(fx + j, c.wo - 1) * c.cstride;
        }
#pragma unroll
        for (int k = 0; k < kUpTaps; ++k) {
            const float* row = base + (size_t)min(fy + k, c.ho - 1) * c.wo * c.cstride;
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < kUpTaps; ++j) t = fmaf(wx[j], __ldg(row + col[j]), t);
            acc = fmaf(__ldg(c.yw + (size_t)y * kUpTaps + k), t, acc);
        }
    }
    return acc;
}

// two adjacent channels (ch even: the x / y components of one limb's PAF) at one position
__device__ __forceinline__ float2 composite_at2(const CompositeMap& m, int frame, int ch, int y, int x) {
    float2 acc = make_float2(0.f, 0.f);
    for (int s = 0; s < m.n_scales; ++s) {
        const CompositeScale& c = m.sc[s];
        const float* base = c.src + (size_t)frame * c.frame_stride + ch;
        const int fy = __ldg(c.yf + y), fx = __ldg(c.xf + x);
        float wx[kUpTaps];
        int col[kUpTaps];
#pragma unroll
        for (int j = 0; j < kUpTaps; ++j) {
            wx[j] = __ldg(c.xw + (size_t)x * kUpTaps + j);
            col[j] = min(fx + j, c.wo - 1) * c.cstride;
        }
#pragma unroll
        for (int k = 0; k < kUpTaps; ++k) {
            const float* row = base + (size_t)min(fy + k, c.ho - 1) * c.wo * c.cstride;
            float2 t = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kUpTaps; ++j) {
                const float2 v = __ldg((const float2*)(row + col[j]));
                t.x = fmaf(wx[j], v.x, t.x);
                t.y = fmaf(wx[j], v.y, t.y);
            }
            const float wy = __ldg(c.yw + (size_t)y * kUpTaps + k);
            acc.x = fmaf(wy, t.x, acc.x);
            acc.y = fmaf(wy, t.y, acc.y);
        }
    }
    return acc;
}

}  // namespace opb
