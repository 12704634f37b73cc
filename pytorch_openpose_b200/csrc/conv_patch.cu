// Patch-resident tcgen05 implicit-GEMM convolution for the 3x3 and 7x7 layers (sm_100a).
//
// conv_tc.cu fetches one 128-pixel x 64-channel A tile per filter tap, i.e. every input pixel crosses L2->SM once
// per tap (49x for 7x7).  ncu showed that kernel pinned at the L2 ("LTS") operand-bandwidth cap (~11.6 TB/s) with
// the tensor pipe only ~48 % busy.  This kernel keeps a HALO PATCH of the input resident in shared memory and
// issues every tap as a UMMA whose A descriptor simply starts at a shifted row of the patch:
//
//   super-tile   16 x 16 output pixels = two 8-wide halves (M = 2 x 128); both halves reuse each B (weight) stage,
//                which halves the weight traffic per FLOP; accumulators: 2 halves x BLOCK_N columns x 2 stages of TMEM.
//   MODE 0       one patch per 64-channel chunk: TMA box {64 ch, 24 cols, 16+ks-1 rows}; tap (dy,dx) of half h reads
//                rows starting at patch row dy*24 + dx + 8h (128 B per row, SWIZZLE_128B), 8-row groups 24 rows
//                (3072 B) apart.  The swizzle XOR is a function of the absolute smem address, so the unaligned start
//                rows need no descriptor base offset (verified on hardware).
//   MODE 1       one patch per (chunk, dx): box {64, 16 cols, 16+ks-1 rows} at column offset dx, so only row shifts
//                by whole 8-row swizzle atoms occur (dy*16 rows = dy*2048 B); costs ks x more
//                patch traffic than MODE 0 (still ~5x less than per-tap tiles).
//   Zero padding is the TMA out-of-bounds fill, as in conv_tc.cu.  Epilogue as in conv_tc.cu (bias, ReLU, fused 2x2
//   max-pool, bf16/fp32 stores at a channel offset/stride).
//
// Warp roles (224 threads): warp 0 = weight (B) TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 =
// epilogue, warp 6 = patch (A) TMA producer (separate so that patch prefetch is not throttled by the B ring).
#include "opb_common.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>

namespace opb {
namespace {

using namespace tc;

constexpr int kThreads = 224;
constexpr int kAccStages = 2;
constexpr int kHalves = 2;
constexpr int kTile = 16;                          // super-tile is 16 x 16 pixels
constexpr int FLAG_RELU = 1, FLAG_F32 = 2, FLAG_POOL = 4;

struct PProb {
    void* out;
    const float* bias;
    int H, W, N;
    int tiles_x, tiles_y;
    int out_cstride, cout_store;
    int n_tiles_n, cin_chunks;
    int tile_begin, flags;
};

struct alignas(64) PatchParams {
    CUtensorMap tmA[kConvMaxProblems];
    CUtensorMap tmW[kConvMaxProblems];
    PProb prob[kConvMaxProblems];
    int nprob, total_tiles, ks;
};
static_assert(sizeof(PatchParams) <= 4000, "kernel parameter space");

template <int BLOCK_N, int MODE>
struct PCfg {
    static constexpr int kPitch = MODE == 0 ? 24 : 16;                 // patch columns (rows of 128 B per patch line)
    static constexpr int kPatchBytesMax = kPitch * (kTile + 6) * 128;  // ks = 7
    static constexpr int kNumPatch = 2;
    static constexpr int kBBytes = BLOCK_N * 128;
    static constexpr int kBStages = MODE == 0 ? (BLOCK_N == 128 ? 5 : 8) : (BLOCK_N == 128 ? 8 : 12);
    static constexpr int kTmemCols = kAccStages * kHalves * BLOCK_N;   // 512 / 256
    static constexpr int kNumBars = 2 * kNumPatch + 2 * kBStages + 2 * kAccStages;
    static constexpr int kBarBytes = kNumBars * 8 + 16;
    static constexpr int kBiasBytes = kAccStages * BLOCK_N * 4;
    static constexpr int kSmemBytes = 1024 + kNumPatch * kPatchBytesMax + kBStages * kBBytes + kBarBytes + kBiasBytes;
    static_assert(kPatchBytesMax % 1024 == 0 && kBBytes % 1024 == 0, "swizzle atoms must stay 1024-B aligned");
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// K-major SWIZZLE_128B descriptor with explicit 8-row-group stride and swizzle base offset
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

struct TileCoord {
    int pi, img, x0, y0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const PatchParams& p, int t, int block_n) {
    int pi = 0;
    while (pi + 1 < p.nprob && t >= p.prob[pi + 1].tile_begin) ++pi;
    const PProb& q = p.prob[pi];
    const int local = t - q.tile_begin;
    const int nt = local % q.n_tiles_n;
    const int mt = local / q.n_tiles_n;
    const int per_img = q.tiles_x * q.tiles_y;
    const int img = mt / per_img;
    const int r = mt - img * per_img;
    const int tyi = r / q.tiles_x;
    const int txi = r - tyi * q.tiles_x;
    TileCoord c;
    c.pi = pi;
    c.img = img;
    c.x0 = txi * kTile;
    c.y0 = tyi * kTile;
    c.n0 = nt * block_n;
    return c;
}

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_patch_kernel(const __grid_constant__ PatchParams p) {
    using C = PCfg<BLOCK_N, MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* patches = smem;
    uint8_t* bstages = smem + C::kNumPatch * C::kPatchBytesMax;
    uint64_t* pfull = (uint64_t*)(bstages + C::kBStages * C::kBBytes);
    uint64_t* pempty = pfull + C::kNumPatch;
    uint64_t* bfull = pempty + C::kNumPatch;
    uint64_t* bempty = bfull + C::kBStages;
    uint64_t* tfull = bempty + C::kBStages;
    uint64_t* tempty = tfull + kAccStages;
    uint32_t* tmem_slot = (uint32_t*)(tempty + kAccStages);
    float* sbias = (float*)((uint8_t*)pfull + C::kBarBytes);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const int ks = p.ks;
    const int pad = ks >> 1;
    const int patch_rows = kTile + ks - 1;
    const uint32_t patch_bytes = (uint32_t)(C::kPitch * patch_rows * 128);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nprob; ++i) {
            prefetch_tensormap(&p.tmA[i]);
            prefetch_tensormap(&p.tmW[i]);
        }
        for (int s = 0; s < C::kNumPatch; ++s) {
            mbar_init(&pfull[s], 1);
            mbar_init(&pempty[s], 1);
        }
        for (int s = 0; s < C::kBStages; ++s) {
            mbar_init(&bfull[s], 1);
            mbar_init(&bempty[s], 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // broadcast so the compiler knows the TMEM base is warp-uniform (keeps UTCHMMA operands in uniform registers)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 6) {
        // ================= patch (A) producer =================
        {   // all 32 lanes walk the loop (warp-uniform); one elected lane issues each async op
            int pb = 0;
            uint32_t pphase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t, BLOCK_N);
                const int cin_chunks = p.prob[tc.pi].cin_chunks;
                const CUtensorMap* tmA = &p.tmA[tc.pi];
                for (int cc = 0; cc < cin_chunks; ++cc) {
                    const int nshift = MODE == 0 ? 1 : ks;
                    for (int dx = 0; dx < nshift; ++dx) {
                        mbar_wait(&pempty[pb], pphase ^ 1, 10);
                        mbar_arrive_expect_tx_elect(&pfull[pb], patch_bytes);
                        tma_load_4d_elect(patches + pb * C::kPatchBytesMax, tmA, &pfull[pb], cc * 64, tc.x0 - pad + dx,
                                    tc.y0 - pad, tc.img);
                        if (++pb == C::kNumPatch) { pb = 0; pphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 0) {
        // ================= weight (B) producer =================
        {   // all 32 lanes walk the loop (warp-uniform); one elected lane issues each async op
            int bs = 0;
            uint32_t bphase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t, BLOCK_N);
                const int cin_chunks = p.prob[tc.pi].cin_chunks;
                const CUtensorMap* tmW = &p.tmW[tc.pi];
                for (int cc = 0; cc < cin_chunks; ++cc) {
                    for (int i = 0; i < ks * ks; ++i) {
                        // MODE 0 consumes taps row-major (dy, dx); MODE 1 consumes column-major (dx outer, dy inner)
                        const int tap = MODE == 0 ? i : (i % ks) * ks + (i / ks);
                        mbar_wait(&bempty[bs], bphase ^ 1, 11);
                        mbar_arrive_expect_tx_elect(&bfull[bs], C::kBBytes);
                        tma_load_2d_elect(bstages + bs * C::kBBytes, tmW, &bfull[bs], (tap * cin_chunks + cc) * 64, tc.n0);
                        if (++bs == C::kBStages) { bs = 0; bphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        {   // all 32 lanes walk the loop (warp-uniform); one elected lane issues each async op
            constexpr uint32_t idesc = make_idesc(BLOCK_N);
            int pb = 0, bs = 0, acc = 0;
            uint32_t pphase = 0, bphase = 0, acc_phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t, BLOCK_N);
                const PProb& q = p.prob[tc.pi];
                const bool half_valid[kHalves] = {true, tc.x0 + 8 < q.W};
                uint32_t accum[kHalves] = {0, 0};
                mbar_wait(&tempty[acc], acc_phase ^ 1, 12);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (kHalves * BLOCK_N);
                for (int cc = 0; cc < q.cin_chunks; ++cc) {
                    const int nshift = MODE == 0 ? 1 : ks;
                    for (int sh = 0; sh < nshift; ++sh) {
                        mbar_wait(&pfull[pb], pphase, 13);
                        tc_fence_after();
                        const uint32_t patch_addr = smem_u32(patches + pb * C::kPatchBytesMax);
                        const int ntap = MODE == 0 ? ks * ks : ks;
                        for (int i = 0; i < ntap; ++i) {
                            const int dy = MODE == 0 ? i / ks : i;
                            const int dx = MODE == 0 ? i - dy * ks : 0;        // MODE 1: dx is baked into the patch
                            mbar_wait(&bfull[bs], bphase, 14);
                            tc_fence_after();
                            const uint64_t bdesc = make_desc(smem_u32(bstages + bs * C::kBBytes), 1024, 0);
#pragma unroll
                            for (int h = 0; h < kHalves; ++h) {
                                if (!half_valid[h]) continue;
                                const uint32_t a_addr = patch_addr + (uint32_t)((dy * C::kPitch + dx + h * 8) * 128);
                                // Measured on B200: the 128-B swizzle is applied to the absolute shared-memory
                                // address, so a start row that is not 1024-B aligned needs NO base offset (setting
                                // (addr >> 7) & 7 there gives wrong results; tests/test_gpu_conv.py covers it).
                                const uint64_t adesc = make_desc(a_addr, C::kPitch * 128, 0);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma_bf16_elect(d_tmem + h * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, accum[h]);
                                    accum[h] = 1;
                                }
                            }
                            umma_commit_elect(&bempty[bs]);
                            if (++bs == C::kBStages) { bs = 0; bphase ^= 1; }
                        }
                        umma_commit_elect(&pempty[pb]);
                        if (++pb == C::kNumPatch) { pb = 0; pphase ^= 1; }
                    }
                }
                umma_commit_elect(&tfull[acc]);
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ================= epilogue =================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int ep_tid = threadIdx.x - 64;
        const int tx = row & 7, ty = row >> 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(p, t, BLOCK_N);
            const PProb& q = p.prob[tc.pi];
            float* bias_s = sbias + acc * BLOCK_N;
            if (ep_tid < BLOCK_N) bias_s[ep_tid] = __ldg(q.bias + tc.n0 + ep_tid);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const bool relu = q.flags & FLAG_RELU;
            const bool pool = q.flags & FLAG_POOL;
            const bool f32 = q.flags & FLAG_F32;
            const int n_valid = q.cout_store - tc.n0;
            const int y = tc.y0 + ty;

            mbar_wait(&tfull[acc], acc_phase, 15);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < kHalves; ++h) {
                if (tc.x0 + 8 * h >= q.W) break;                     // half entirely outside the image (uniform)
                const int x = tc.x0 + 8 * h + tx;
                const bool inside = (x < q.W) && (y < q.H);
                size_t pix;
                bool writer;
                if (pool) {
                    pix = ((size_t)tc.img * (q.H >> 1) + (y >> 1)) * (q.W >> 1) + (x >> 1);
                    writer = inside && !(tx & 1) && !(ty & 1);
                } else {
                    pix = ((size_t)tc.img * q.H + y) * q.W + x;
                    writer = inside;
                }
                const size_t out_off = pix * q.out_cstride + tc.n0;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (kHalves * BLOCK_N) + h * BLOCK_N;
#pragma unroll 1
                for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                    if (c0 >= n_valid) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float a = __uint_as_float(v[j]) + bias_s[c0 + j];
                        f[j] = relu ? fmaxf(a, 0.f) : a;
                    }
                    if (pool) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float m = fmaxf(f[j], __shfl_xor_sync(0xffffffffu, f[j], 1));
                            f[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                        }
                    }
                    if (writer) {
                        if (f32) {
                            float* o = (float*)q.out + out_off + c0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (c0 + j < n_valid) *(float4*)(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        } else {
                            __nv_bfloat16* o = (__nv_bfloat16*)q.out + out_off + c0;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                if (c0 + j < n_valid) {
                                    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[j], f[j + 1]);
                                    __nv_bfloat162 h1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
                                    __nv_bfloat162 h3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                                    uint4 u;
                                    u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = *(uint32_t*)&h2; u.w = *(uint32_t*)&h3;
                                    *(uint4*)(o + j) = u;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

struct PatchLaunch : ConvLaunch {
    PatchParams params;
    int grid = 0, block_n = 128, mode = 0;
    void run(cudaStream_t stream) const override;
};

template <int BN, int MODE>
void launch_one(const PatchLaunch& L, cudaStream_t stream) {
    static bool attr[64] = {};
    if (first_use_on_device(attr)) {
        OPB_CUDA(cudaFuncSetAttribute(conv_patch_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PCfg<BN, MODE>::kSmemBytes));
    }
    conv_patch_kernel<BN, MODE><<<L.grid, kThreads, PCfg<BN, MODE>::kSmemBytes, stream>>>(L.params);
}

void PatchLaunch::run(cudaStream_t stream) const {
    if (block_n == 128) {
        if (mode == 0) launch_one<128, 0>(*this, stream);
        else launch_one<128, 1>(*this, stream);
    } else {
        if (mode == 0) launch_one<64, 0>(*this, stream);
        else launch_one<64, 1>(*this, stream);
    }
    OPB_CUDA(cudaGetLastError());
}

}  // namespace

ConvLaunch* conv_patch_plan(const std::vector<ConvOp>& ops, int block_n, int num_sms, int mode) {
    OPB_REQUIRE(!ops.empty() && (int)ops.size() <= kConvMaxProblems, "conv_patch: 1..8 problems per launch");
    OPB_REQUIRE(block_n == 64 || block_n == 128, "conv_patch: block_n must be 64 or 128");
    OPB_REQUIRE(mode == 0 || mode == 1, "conv_patch: mode 0 or 1");
    auto L = std::make_unique<PatchLaunch>();
    PatchParams& P = L->params;
    memset(&P, 0, sizeof(P));
    P.nprob = (int)ops.size();
    P.ks = ops[0].ks;
    OPB_REQUIRE(P.ks == 3 || P.ks == 7, "conv_patch: kernel size 3 or 7");
    const int pitch = mode == 0 ? 24 : 16;
    int tile = 0;
    for (int i = 0; i < P.nprob; ++i) {
        const ConvOp& op = ops[i];
        OPB_REQUIRE(op.ks == P.ks, "conv_patch: grouped problems must share the kernel size");
        OPB_REQUIRE(op.in.elem == 2 && op.in.c % 64 == 0, "conv_patch: input must be bf16 with C % 64 == 0");
        OPB_REQUIRE(op.in.cstride % 8 == 0 && op.in.coff % 8 == 0, "conv_patch: input slice must be 16-byte aligned");
        OPB_REQUIRE(op.cout_pad % block_n == 0 && op.cout_store % 8 == 0 && op.cout_store <= op.cout_pad,
                    "conv_patch: bad output channel padding");
        OPB_REQUIRE(op.out.elem == 2 || op.out.elem == 4, "conv_patch: output must be bf16 or fp32");
        OPB_REQUIRE((op.out.coff * op.out.elem) % 16 == 0 && (op.out.cstride * op.out.elem) % 16 == 0,
                    "conv_patch: output slice must be 16-byte aligned");
        const int H = op.in.h, W = op.in.w, N = op.in.n;
        if (op.pool) {
            OPB_REQUIRE(H % 2 == 0 && W % 2 == 0 && op.relu, "conv_patch: fused pool needs even dims and ReLU");
            OPB_REQUIRE(op.out.h == H / 2 && op.out.w == W / 2 && op.out.n == N, "conv_patch: pooled output dims");
        } else {
            OPB_REQUIRE(op.out.h == H && op.out.w == W && op.out.n == N, "conv_patch: output dims");
        }
        PProb& q = P.prob[i];
        q.out = op.out.ptr();
        q.bias = op.bias;
        q.H = H; q.W = W; q.N = N;
        q.tiles_x = cdiv(W, kTile);
        q.tiles_y = cdiv(H, kTile);
        q.out_cstride = op.out.cstride;
        q.cout_store = op.cout_store;
        q.n_tiles_n = op.cout_pad / block_n;
        q.cin_chunks = op.in.c / 64;
        q.tile_begin = tile;
        q.flags = (op.relu ? FLAG_RELU : 0) | (op.out.elem == 4 ? FLAG_F32 : 0) | (op.pool ? FLAG_POOL : 0);
        tile += q.tiles_x * q.tiles_y * N * q.n_tiles_n;

        cuuint64_t adims[4] = {(cuuint64_t)op.in.c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t astr[3] = {(cuuint64_t)op.in.cstride * 2, (cuuint64_t)op.in.cstride * 2 * W,
                              (cuuint64_t)op.in.cstride * 2 * W * H};
        cuuint32_t abox[4] = {64, (cuuint32_t)pitch, (cuuint32_t)(kTile + P.ks - 1), 1};
        tensor_map_encode_bf16(&P.tmA[i], op.in.ptr(), 4, adims, astr, abox);
        const cuuint64_t K = (cuuint64_t)op.ks * op.ks * op.in.c;
        cuuint64_t wdims[2] = {K, (cuuint64_t)op.cout_pad};
        cuuint64_t wstr[1] = {K * 2};
        cuuint32_t wbox[2] = {64, (cuuint32_t)block_n};
        tensor_map_encode_bf16(&P.tmW[i], (void*)op.w, 2, wdims, wstr, wbox);
    }
    P.total_tiles = tile;
    L->tiles = tile;
    L->block_n = block_n;
    L->mode = mode;
    L->grid = tile < num_sms ? tile : num_sms;
    return L.release();
}

}  // namespace opb
