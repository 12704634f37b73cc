// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// Replaces torch.conv2d (+ ReLU, + MaxPool2d, + torch.cat) as used by the reference networks
// (src/model.py:7-22 make_layers, :106-133 bodypose_model.forward, :197-214 handpose_model.forward).
//
//   GEMM view   D[M=pixels, N=cout] = sum over taps (dy,dx) and channel chunks of
//               A[pixels shifted by (dy,dx), 64 channels] * W[cout, tap, 64 channels]^T
//   A operand   NHWC bf16 activations, fetched by TMA as a 4-D box {64 ch, tw, th, 1 image} whose
//               coordinates are shifted by the tap offset; out-of-image elements are zero-filled by
//               the TMA unit, which is exactly the convolution's zero padding.  tw*th = 128 pixels,
//               so the box lands in shared memory as the canonical K-major SWIZZLE_128B UMMA tile
//               (128 rows x 128 B).
//   B operand   weights repacked to [cout_pad][tap][cin] bf16, TMA box {64, BLOCK_N}.
//   MMA         tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16, fp32 accumulators in TMEM,
//               two accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
//   Epilogue    tcgen05.ld -> +bias -> ReLU -> (2x2 max-pool by warp shuffles) -> bf16/fp32 NHWC stores
//               at a channel offset / stride, i.e. straight into the next stage's concat buffer.
//   Grouping    one launch carries up to 8 problems (scales x branches) that share the kernel size;
//               CTAs are persistent and walk a flat tile list.
//
// Warp roles (192 threads): warp 0 = TMA producer (1 lane), warp 1 = TMEM owner + MMA issuer (1 lane),
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
#include "opb_common.cuh"
#include "tc_ptx.cuh"

namespace opb {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements: 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KiB
constexpr int kThreads = 192;
constexpr int kAccStages = 2;

constexpr int FLAG_RELU = 1, FLAG_F32 = 2, FLAG_POOL = 4;

struct Prob {
    void* out;
    const float* bias;
    int H, W, N;                 // conv output (= input) rows, cols, images
    int tiles_x, tiles_y;
    int tw_log2;                 // tile = (1 << tw_log2) cols x (128 >> tw_log2) rows
    int out_cstride;             // elements per output pixel
    int cout_store;              // channels written
    int n_tiles_n;               // cout_pad / BLOCK_N
    int cin_chunks;              // cin / 64
    int tile_begin;              // first flat tile index of this problem
    int flags;
};

struct alignas(64) ConvParams {
    CUtensorMap tmA[kConvMaxProblems];
    CUtensorMap tmW[kConvMaxProblems];
    Prob prob[kConvMaxProblems];
    int nprob, total_tiles, ks;
};
static_assert(sizeof(ConvParams) <= 4000, "kernel parameter space");

template <int BLOCK_N>
struct Cfg {
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N == 128) ? 6 : 8;
    static constexpr int kTmemCols = kAccStages * BLOCK_N;        // 256 or 128: power of two >= 32
    static constexpr int kBarBytes = (2 * kStages + 2 * kAccStages) * 8 + 16;
    static constexpr int kBiasBytes = kAccStages * BLOCK_N * 4;
    static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kBarBytes + kBiasBytes;
};

using namespace tc;

struct TileCoord {
    int pi, img, x0, y0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t, int block_n) {
    int pi = 0;
    while (pi + 1 < p.nprob && t >= p.prob[pi + 1].tile_begin) ++pi;
    const Prob& q = p.prob[pi];
    int local = t - q.tile_begin;
    int nt = local % q.n_tiles_n;
    int mt = local / q.n_tiles_n;
    int per_img = q.tiles_x * q.tiles_y;
    int img = mt / per_img;
    int r = mt - img * per_img;
    int tyi = r / q.tiles_x;
    int txi = r - tyi * q.tiles_x;
    TileCoord c;
    c.pi = pi;
    c.img = img;
    c.x0 = txi << q.tw_log2;
    c.y0 = tyi * (kBlockM >> q.tw_log2);
    c.n0 = nt * block_n;
    return c;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 needs 1024-B alignment
    uint8_t* tiles = smem;
    uint64_t* full_bar = (uint64_t*)(smem + C::kStages * C::kStageBytes);
    uint64_t* empty_bar = full_bar + C::kStages;
    uint64_t* tfull_bar = empty_bar + C::kStages;
    uint64_t* tempty_bar = tfull_bar + kAccStages;
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + kAccStages);
    float* sbias = (float*)((uint8_t*)full_bar + C::kBarBytes);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform role index
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nprob; ++i) {
            prefetch_tensormap(&p.tmA[i]);
            prefetch_tensormap(&p.tmW[i]);
        }
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // broadcast so the compiler knows the TMEM base is warp-uniform (keeps UTCHMMA operands in uniform registers)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const int taps = p.ks * p.ks;
    const int pad = p.ks >> 1;

    pdl_launch_dependents();
    if (warp == 0) {
        // ================= TMA producer =================
        {   // all 32 lanes walk the loop (warp-uniform); one elected lane issues each async op
            int stage = 0;
            uint32_t phase = 0;
            pdl_wait();                 // programmatic dependent launch: the prologue above overlapped the predecessor
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t, BLOCK_N);
                const int cin_chunks = p.prob[tc.pi].cin_chunks;
                const CUtensorMap* tmA = &p.tmA[tc.pi];
                const CUtensorMap* tmW = &p.tmW[tc.pi];
                for (int tap = 0; tap < taps; ++tap) {
                    const int dy = tap / p.ks - pad;
                    const int dx = tap - (tap / p.ks) * p.ks - pad;
                    for (int cc = 0; cc < cin_chunks; ++cc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1, 0);
                        uint8_t* a_dst = tiles + stage * C::kStageBytes;
                        uint8_t* b_dst = a_dst + kABytes;
                        mbar_arrive_expect_tx_elect(&full_bar[stage], C::kStageBytes);
                        tma_load_4d_elect(a_dst, tmA, &full_bar[stage], cc * kBlockK, tc.x0 + dx, tc.y0 + dy, tc.img);
                        tma_load_2d_elect(b_dst, tmW, &full_bar[stage], (tap * cin_chunks + cc) * kBlockK, tc.n0);
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        {   // all 32 lanes walk the loop (warp-uniform); one elected lane issues each async op
            constexpr uint32_t idesc = make_idesc(BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t, BLOCK_N);
                const int num_kb = taps * p.prob[tc.pi].cin_chunks;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 2);           // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + stage * C::kStageBytes);
                    const uint64_t adesc = make_sw128_desc(a_addr);
                    const uint64_t bdesc = make_sw128_desc(a_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // +32 bytes per K=16 step inside the 128-byte swizzle row: +2 in the >>4 address field
                        umma_bf16_elect(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit_elect(&empty_bar[stage]);                  // smem slot reusable once these MMAs retire
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_elect(&tfull_bar[acc]);                        // accumulator complete -> epilogue
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (warps 2..5) =================
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;              // accumulator row = pixel within the tile
        const int ep_tid = threadIdx.x - 64;           // 0..127
        int acc = 0;
        uint32_t acc_phase = 0;
        int bias_loaded[kAccStages];
#pragma unroll
        for (int a = 0; a < kAccStages; ++a) bias_loaded[a] = -1;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(p, t, BLOCK_N);
            const Prob& q = p.prob[tc.pi];
            float* bias_s = sbias + acc * BLOCK_N;
            // the bias slice changes only with the problem / n tile: reloading it for every tile put a global load and a
            // barrier at the head of each epilogue, which is what paced the short-K (1x1, 3x3) layers
            const int bias_key = (tc.pi << 16) | tc.n0;
            if (bias_loaded[acc] != bias_key) {
                asm volatile("bar.sync 1, 128;" ::: "memory");       // nobody still reads this slot (two tiles back)
                if (ep_tid < BLOCK_N) bias_s[ep_tid] = __ldg(q.bias + tc.n0 + ep_tid);
                asm volatile("bar.sync 1, 128;" ::: "memory");       // bias visible to the 4 epilogue warps
                bias_loaded[acc] = bias_key;
            }           // bias visible to the 4 epilogue warps

            const int tw_mask = (1 << q.tw_log2) - 1;
            const int tx = row & tw_mask, ty = row >> q.tw_log2;
            const int x = tc.x0 + tx, y = tc.y0 + ty;
            const bool inside = (x < q.W) && (y < q.H);
            const bool relu = q.flags & FLAG_RELU;
            const bool pool = q.flags & FLAG_POOL;
            const bool f32 = q.flags & FLAG_F32;
            const int n_valid = q.cout_store - tc.n0;                // channels of this N tile that are stored
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

            mbar_wait(&tfull_bar[acc], acc_phase, 3);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                if (c0 >= n_valid) break;                            // warp-uniform
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float a = __uint_as_float(v[j]) + bias_s[c0 + j];
                    f[j] = relu ? fmaxf(a, 0.f) : a;
                }
                if (pool) {
                    const int up = 1 << q.tw_log2;                   // lane distance of the row below (tw <= 16)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float m = fmaxf(f[j], __shfl_xor_sync(0xffffffffu, f[j], 1));
                        f[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, up));
                    }
                }
                if (writer) {
                    if (f32) {
                        float* o = (float*)q.out + out_off + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (c0 + j < n_valid) *(float4*)(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    } else {
                        store_bf16x32((__nv_bfloat16*)q.out + out_off + c0, f, n_valid - c0);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);                           // 128 arrivals free the accumulator
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

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        OPB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (!ptr || qres != cudaDriverEntryPointSuccess)
            throw Error(OPB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

}  // namespace

void tensor_map_encode_bf16(CUtensorMap* tm, void* addr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                            const cuuint32_t* box) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    OPB_REQUIRE(((uintptr_t)addr & 15) == 0, "TMA base address must be 16-byte aligned");
    CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, addr, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(OPB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

namespace {

// tile = tw x th output pixels with tw*th = 128; pick the shape that wastes the fewest pixels
int choose_tw_log2(int H, int W, bool pool) {
    int best = -1;
    long best_tiles = 0;
    for (int l = pool ? 1 : 0; l <= (pool ? 4 : 7); ++l) {
        int tw = 1 << l, th = 128 >> l;
        long tiles = (long)cdiv(W, tw) * cdiv(H, th);
        // prefer wider tiles on ties (longer contiguous runs per TMA row)
        if (best < 0 || tiles < best_tiles || (tiles == best_tiles && l > best && l <= 5)) {
            best = l;
            best_tiles = tiles;
        }
    }
    return best;
}

}  // namespace

struct TapLaunch : ConvLaunch {
    ConvParams params;
    int grid = 0;
    int block_n = 128;
    void run(cudaStream_t stream) const override;
};

static void conv_tc_prepare(const std::vector<ConvOp>& ops, int block_n, int num_sms, TapLaunch& L) {
    OPB_REQUIRE(!ops.empty() && (int)ops.size() <= kConvMaxProblems, "conv_tc: 1..8 problems per launch");
    OPB_REQUIRE(block_n == 64 || block_n == 128, "conv_tc: block_n must be 64 or 128");
    ConvParams& P = L.params;
    memset(&P, 0, sizeof(P));
    P.nprob = (int)ops.size();
    P.ks = ops[0].ks;
    int tile = 0;
    for (int i = 0; i < P.nprob; ++i) {
        const ConvOp& op = ops[i];
        OPB_REQUIRE(op.ks == P.ks, "conv_tc: grouped problems must share the kernel size");
        OPB_REQUIRE(op.ks == 1 || op.ks == 3 || op.ks == 7, "conv_tc: kernel size 1, 3 or 7");
        OPB_REQUIRE(op.in.elem == 2 && op.in.c % kBlockK == 0, "conv_tc: input must be bf16 with C % 64 == 0");
        OPB_REQUIRE(op.in.cstride % 8 == 0 && op.in.coff % 8 == 0, "conv_tc: input slice must be 16-byte aligned");
        OPB_REQUIRE(op.cout_pad % block_n == 0 && op.cout_store % 8 == 0 && op.cout_store <= op.cout_pad,
                    "conv_tc: bad output channel padding");
        OPB_REQUIRE(op.out.elem == 2 || op.out.elem == 4, "conv_tc: output must be bf16 or fp32");
        OPB_REQUIRE((op.out.coff * op.out.elem) % 16 == 0 && (op.out.cstride * op.out.elem) % 16 == 0,
                    "conv_tc: output slice must be 16-byte aligned");
        int H = op.in.h, W = op.in.w, N = op.in.n;
        // a 1x1 convolution has no spatial structure: run it over the flattened pixel list (tiles of 128 consecutive
        // pixels) instead of 2-D tiles, which waste 20-40 % of their rows on maps as narrow as 41 or 82 columns
        const bool flat = op.ks == 1 && !op.pool && (long long)H * W * N < (1ll << 31);
        if (op.pool) {
            OPB_REQUIRE(H % 2 == 0 && W % 2 == 0, "conv_tc: fused pool needs even dims");
            OPB_REQUIRE(op.out.h == H / 2 && op.out.w == W / 2 && op.out.n == N, "conv_tc: pooled output dims");
            OPB_REQUIRE(op.relu, "conv_tc: fused pool is defined after ReLU");
        } else {
            OPB_REQUIRE(op.out.h == H && op.out.w == W && op.out.n == N, "conv_tc: output dims");
        }
        if (flat) {
            W = H * W * N;
            H = 1;
            N = 1;
        }
        Prob& q = P.prob[i];
        q.out = op.out.ptr();
        q.bias = op.bias;
        q.H = H; q.W = W; q.N = N;
        q.tw_log2 = flat ? 7 : choose_tw_log2(H, W, op.pool);
        const int tw = 1 << q.tw_log2, th = 128 >> q.tw_log2;
        q.tiles_x = cdiv(W, tw);
        q.tiles_y = cdiv(H, th);
        q.out_cstride = op.out.cstride;
        q.cout_store = op.cout_store;
        q.n_tiles_n = op.cout_pad / block_n;
        q.cin_chunks = op.in.c / kBlockK;
        q.tile_begin = tile;
        q.flags = (op.relu ? FLAG_RELU : 0) | (op.out.elem == 4 ? FLAG_F32 : 0) | (op.pool ? FLAG_POOL : 0);
        tile += q.tiles_x * q.tiles_y * N * q.n_tiles_n;

        // A: NHWC bf16 view {C, W, H, N}
        cuuint64_t adims[4] = {(cuuint64_t)op.in.c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t astr[3] = {(cuuint64_t)op.in.cstride * 2, (cuuint64_t)op.in.cstride * 2 * W,
                              (cuuint64_t)op.in.cstride * 2 * W * H};
        cuuint32_t abox[4] = {(cuuint32_t)kBlockK, (cuuint32_t)tw, (cuuint32_t)th, 1};
        tensor_map_encode_bf16(&P.tmA[i], op.in.ptr(), 4, adims, astr, abox);
        // W: [cout_pad][K] bf16
        const cuuint64_t K = (cuuint64_t)op.ks * op.ks * op.in.c;
        cuuint64_t wdims[2] = {K, (cuuint64_t)op.cout_pad};
        cuuint64_t wstr[1] = {K * 2};
        cuuint32_t wbox[2] = {(cuuint32_t)kBlockK, (cuuint32_t)block_n};
        tensor_map_encode_bf16(&P.tmW[i], (void*)op.w, 2, wdims, wstr, wbox);
    }
    P.total_tiles = tile;
    L.tiles = tile;
    L.block_n = block_n;
    L.grid = tile < num_sms ? tile : num_sms;
}

static void conv_tc_run(const TapLaunch& L, cudaStream_t stream) {
    static bool attr_set[64] = {};
    if (first_use_on_device(attr_set)) {
        OPB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg<128>::kSmemBytes));
        OPB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg<64>::kSmemBytes));
    }
    if (L.block_n == 128)
        launch_pdl(conv_tc_kernel<128>, L.grid, kThreads, Cfg<128>::kSmemBytes, stream, L.params);
    else
        launch_pdl(conv_tc_kernel<64>, L.grid, kThreads, Cfg<64>::kSmemBytes, stream, L.params);
    OPB_CUDA(cudaGetLastError());
}

void TapLaunch::run(cudaStream_t stream) const { conv_tc_run(*this, stream); }

ConvLaunch* conv_tc_plan(const std::vector<ConvOp>& ops, int block_n, int num_sms) {
    auto L = std::make_unique<TapLaunch>();
    conv_tc_prepare(ops, block_n, num_sms, *L);
    return L.release();
}

void conv_tc_launch(const std::vector<ConvOp>& ops, int block_n, cudaStream_t stream, int num_sms) {
    TapLaunch L;
    conv_tc_prepare(ops, block_n, num_sms, L);
    conv_tc_run(L, stream);
}

}  // namespace opb
