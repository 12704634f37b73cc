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

// ---- conv1_1: one thread = 4 horizontally adjacent output pixels x 64 output channels (16 at a time).  Register
// blocking over pixels divides the shared-memory weight reads per FMA by four (the 1-pixel form was LDS-bound: ncu
// l1tex 86 %, FMA pipe 41 %).  W is a multiple of 8 (padded image), so strips never straddle rows. ---------------
__global__ void __launch_bounds__(128) conv_first_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                         const float* __restrict__ w /*[27][64]*/,
                                                         const float* __restrict__ bias, int N, int H, int W,
                                                         int out_cstride) {
    __shared__ float sw[27 * 64];
    __shared__ float sb[64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int strips_per_row = W >> 2;
    const size_t total = (size_t)N * H * strips_per_row;
    const size_t sidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= total) return;
    const int x0 = (int)(sidx % strips_per_row) * 4;
    const int y = (int)((sidx / strips_per_row) % H);
    const size_t img = sidx / ((size_t)strips_per_row * H);
    // 3 rows x 6 columns x 3 channels of normalised input (x/256 - 0.5 is exact in fp32; zero outside the image)
    float v[3][6][3];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        const bool rowok = (yy >= 0) && (yy < H);
        const uint8_t* row = in + (img * H + (rowok ? yy : 0)) * (size_t)W * 3;
#pragma unroll
        for (int dx = 0; dx < 6; ++dx) {
            const int xx = x0 + dx - 1;
            const bool ok = rowok && (xx >= 0) && (xx < W);
            const uint8_t* p = row + (ok ? xx : 0) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[dy][dx][c] = ok ? ((float)p[c] * (1.0f / 256.0f) - 0.5f) : 0.0f;
        }
    }
    __nv_bfloat16* o = out + ((img * H + y) * (size_t)W + x0) * out_cstride;
#pragma unroll 1
    for (int c16 = 0; c16 < 4; ++c16) {
        float acc[4][16];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[p][j] = sb[c16 * 16 + j];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* wk = &sw[((dy * 3 + dx) * 3 + c) * 64 + c16 * 16];
                    float wv[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 t = *(const float4*)(wk + q * 4);
                        wv[q * 4] = t.x; wv[q * 4 + 1] = t.y; wv[q * 4 + 2] = t.z; wv[q * 4 + 3] = t.w;
                    }
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(v[dy][dx + p][c], wv[j], acc[p][j]);
                }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
                __nv_bfloat162 h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    h[j] = __floats2bfloat162_rn(fmaxf(acc[p][h8 * 8 + 2 * j], 0.f), fmaxf(acc[p][h8 * 8 + 2 * j + 1], 0.f));
                *(uint4*)(o + (size_t)p * out_cstride + c16 * 16 + h8 * 8) = *(uint4*)h;
            }
        }
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
    OPB_REQUIRE(in_u8.w % 4 == 0, "conv_first: padded width must be a multiple of 4");
    const size_t total = in_u8.pixels() / 4;                   // one thread per 4-pixel strip
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
