// CUDA-core kernels around the tensor-core convolution:
//   * conv_first   : conv1_1 (src/model.py:35 / :145), Cin = 3 so K = 27 -- bandwidth bound, not GEMM
//                    shaped; reads the uint8 padded image, applies the reference's x/256-0.5
//                    normalisation (src/body.py:40) on the fly, writes bf16 NHWC.
//   * maxpool2     : nn.MaxPool2d(2,2,0) (src/model.py:11-12) on bf16 NHWC (used when a pool is not fused)
//   * conv_direct  : scalar reference convolution with the tensor-core kernel's exact I/O contract; only
//                    the cross-check entry point opb_conv2d(impl=1) reaches it.
#include "opb_common.cuh"

namespace opb {
namespace {

// ---- conv1_1: one thread per output pixel, 64 output channels, weights broadcast from smem -------
__global__ void __launch_bounds__(128) conv_first_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                         const float* __restrict__ w /*[27][64]*/,
                                                         const float* __restrict__ bias, int N, int H, int W,
                                                         int out_cstride) {
    __shared__ float sw[27 * 64];
    __shared__ float sb[64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const size_t total = (size_t)N * H * W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const size_t img = pix / ((size_t)W * H);
    float v[27];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int yy = y + dy - 1, xx = x + dx - 1;
            const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
            const uint8_t* p = in + ((img * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[(dy * 3 + dx) * 3 + c] = ok ? ((float)p[c] * (1.0f / 256.0f) - 0.5f) : 0.0f;   // exact in fp32
        }
    }
    __nv_bfloat16* o = out + pix * out_cstride;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = sb[c0 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[k], sw[k * 64 + c0 + j], acc[j]);
        }
        __nv_bfloat162 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
        *(uint4*)(o + c0) = *(uint4*)h;
    }
}

// ---- 2x2 max-pool, 8 channels (16 bytes) per thread -------------------------------------------------
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
    uint4 r;
    __nv_bfloat162* pa = (__nv_bfloat162*)&a;
    __nv_bfloat162* pb = (__nv_bfloat162*)&b;
    __nv_bfloat162* pr = (__nv_bfloat162*)&r;
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}
__global__ void maxpool2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int Ho,
                                int Wo, int C, int in_cstride, int out_cstride) {
    const int cv = C / 8;
    const size_t total = (size_t)N * Ho * Wo * cv;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % cv) * 8;
    size_t r = i / cv;
    const int x = (int)(r % Wo);
    r /= Wo;
    const int y = (int)(r % Ho);
    const size_t n = r / Ho;
    const int Hi = Ho * 2, Wi = Wo * 2;
    const __nv_bfloat16* p = in + ((n * Hi + 2 * y) * Wi + 2 * x) * in_cstride + c;
    uint4 a = *(const uint4*)p;
    uint4 b = *(const uint4*)(p + in_cstride);
    uint4 cc = *(const uint4*)(p + (size_t)Wi * in_cstride);
    uint4 d = *(const uint4*)(p + (size_t)Wi * in_cstride + in_cstride);
    *(uint4*)(out + ((n * Ho + y) * Wo + x) * out_cstride + c) = bf16x8_max(bf16x8_max(a, b), bf16x8_max(cc, d));
}

// ---- scalar reference convolution (cross-check only) ------------------------------------------------
__global__ void conv_direct_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ w,
                                   const float* __restrict__ bias, void* __restrict__ out, int N, int H, int W,
                                   int cin, int in_cstride, int cout_store, int out_cstride, int ks, int relu,
                                   int out_f32) {
    const size_t total = (size_t)N * H * W * cout_store;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int co = (int)(i % cout_store);
    size_t r = i / cout_store;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const size_t n = r / H;
    const int pad = ks / 2;
    const size_t K = (size_t)ks * ks * cin;
    float acc = 0.f;
    for (int dy = 0; dy < ks; ++dy) {
        const int yy = y + dy - pad;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < ks; ++dx) {
            const int xx = x + dx - pad;
            if (xx < 0 || xx >= W) continue;
            const __nv_bfloat16* a = in + ((n * H + yy) * W + xx) * in_cstride;
            const __nv_bfloat16* b = w + (size_t)co * K + (size_t)(dy * ks + dx) * cin;
            for (int c = 0; c < cin; ++c) acc = fmaf(__bfloat162float(a[c]), __bfloat162float(b[c]), acc);
        }
    }
    acc += bias[co];
    if (relu) acc = fmaxf(acc, 0.f);
    const size_t o = ((n * H + y) * W + x) * out_cstride + co;
    if (out_f32)
        ((float*)out)[o] = acc;
    else
        ((__nv_bfloat16*)out)[o] = __float2bfloat16_rn(acc);
}

}  // namespace

void conv_first_launch(const TensorView& in_u8, const TensorView& out, const float* w27x64, const float* bias,
                       cudaStream_t stream) {
    OPB_REQUIRE(in_u8.elem == 1 && in_u8.c == 3 && in_u8.cstride == 3, "conv_first: input must be dense u8 HWC3");
    OPB_REQUIRE(out.elem == 2 && out.c == 64 && out.cstride % 8 == 0 && out.coff == 0, "conv_first: output bf16 64ch");
    const size_t total = in_u8.pixels();
    conv_first_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(
        (const uint8_t*)in_u8.base, (__nv_bfloat16*)out.base, w27x64, bias, in_u8.n, in_u8.h, in_u8.w, out.cstride);
    OPB_CUDA(cudaGetLastError());
}

void maxpool2_launch(const TensorView& in, const TensorView& out, cudaStream_t stream) {
    OPB_REQUIRE(in.elem == 2 && out.elem == 2 && in.c == out.c && in.c % 8 == 0, "maxpool2: bf16, C % 8 == 0");
    OPB_REQUIRE(in.h == out.h * 2 && in.w == out.w * 2 && in.n == out.n, "maxpool2: dims");
    OPB_REQUIRE(in.cstride % 8 == 0 && out.cstride % 8 == 0 && in.coff % 8 == 0 && out.coff % 8 == 0, "maxpool2: align");
    const size_t total = out.pixels() * (in.c / 8);
    maxpool2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)in.ptr(), (__nv_bfloat16*)out.ptr(), in.n, out.h, out.w, in.c, in.cstride, out.cstride);
    OPB_CUDA(cudaGetLastError());
}

void conv_direct_launch(const ConvOp& op, cudaStream_t stream) {
    OPB_REQUIRE(!op.pool, "conv_direct: no fused pool");
    const size_t total = op.in.pixels() * op.cout_store;
    conv_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)op.in.ptr(), op.w, op.bias, op.out.ptr(), op.in.n, op.in.h, op.in.w, op.in.c,
        op.in.cstride, op.cout_store, op.out.cstride, op.ks, op.relu ? 1 : 0, op.out.elem == 4 ? 1 : 0);
    OPB_CUDA(cudaGetLastError());
}

}  // namespace opb
