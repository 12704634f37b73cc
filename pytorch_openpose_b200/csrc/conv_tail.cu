// Fused tail of a refinement stage: Mconv6 (1x1, 128 -> 128, ReLU) followed by Mconv7 (1x1, 128 -> 38 | 19 | 22)
// (src/model.py:76-77, 86-87, 187-188) as ONE kernel -- two back-to-back GEMMs per 128-pixel tile.
//
// Why: run as two launches the 1x1 layers are HBM-bound at ~40 % of the copy bandwidth (ncu: tensor pipe 11 %): each
// reads a 128-channel activation of the whole batch from HBM and the first writes one back.  Here the 128-channel
// intermediate never leaves the SM: GEMM1 accumulates in TMEM, the epilogue warps add the bias, apply the ReLU, round
// to bf16 (exactly what the separate kernel stored) and write the tile into shared memory in the K-major SWIZZLE_128B
// layout a UMMA A operand needs; GEMM2 consumes it from there.  Both weight matrices (32 KB + 16 KB) stay resident in
// shared memory for the whole kernel: a CTA only ever works for one weight set (branch), chosen by blockIdx.x % groups.
//
// Pixels are taken as a flat list (a 1x1 convolution has no spatial structure): tile = 128 consecutive pixels.
//
//   warp 0      : TMA producer (weights once, then the A tiles through a 4-stage ring, 2 stages per tile)
//   warp 1      : MMA issuer, software-pipelined: GEMM1(i) is issued before GEMM2(i-1)
//   warps 2..5  : epilogues: epi1(i) (TMEM -> bias/ReLU/bf16 -> swizzled smem), then epi2(i-1) (TMEM -> bias -> global)
#include "opb_common.cuh"
#include "tc_ptx.cuh"

namespace opb {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kTilePx = 128;
constexpr int kChunkBytes = kTilePx * 128;          // 128 rows x 64 bf16
constexpr int kAStages = 4;
constexpr int kW1Bytes = 2 * 128 * 128;             // 2 K-chunks x 128 rows x 128 B
constexpr int kW2Bytes = 2 * 64 * 128;              // 2 K-chunks x  64 rows x 128 B
constexpr int kMidBytes = 2 * kChunkBytes;          // one tile of the intermediate (2 K-chunks)
constexpr int kNumBars = 2 * kAStages + 1 + 6 * 2;
constexpr int kSmemBytes = 1024 + kAStages * kChunkBytes + kW1Bytes + kW2Bytes + 2 * kMidBytes + kNumBars * 8 + 16 +
                           (128 + 64) * 4;
constexpr int kTmemCols = 512;                      // acc1: 2 x 128 columns, acc2: 2 x 64 columns (384 -> power of two)
constexpr int FLAG_RELU2 = 1, FLAG_F32 = 2;
constexpr int kMaxProb = 8, kMaxGroups = 2;

struct TailProb {
    void* out;
    int npix, tiles, tile_begin;     // tile_begin: first tile index inside the problem's group
    int out_cstride, cout_store, flags, group;
};
struct alignas(64) TailParams {
    CUtensorMap tmA[kMaxProb];
    CUtensorMap tmW1[kMaxGroups];
    CUtensorMap tmW2[kMaxGroups];
    TailProb prob[kMaxProb];
    const float* bias1[kMaxGroups];
    const float* bias2[kMaxGroups];
    int group_tiles[kMaxGroups];
    int nprob, ngroups;
};
static_assert(sizeof(TailParams) <= 4000, "kernel parameter space");

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tile `t` (index inside group g) -> problem and first pixel
__device__ __forceinline__ void decode(const TailParams& p, int g, int t, int& pi, int& pix0) {
    pi = -1;
    for (int i = 0; i < p.nprob; ++i)
        if (p.prob[i].group == g && t >= p.prob[i].tile_begin && t < p.prob[i].tile_begin + p.prob[i].tiles) pi = i;
    pix0 = (t - p.prob[pi].tile_begin) * kTilePx;
}

__global__ void __launch_bounds__(kThreads, 1) conv_tail_kernel(const __grid_constant__ TailParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;
    uint8_t* w1s = a_ring + kAStages * kChunkBytes;
    uint8_t* w2s = w1s + kW1Bytes;
    uint8_t* mids = w2s + kW2Bytes;
    uint64_t* a_full = (uint64_t*)(mids + 2 * kMidBytes);
    uint64_t* a_empty = a_full + kAStages;
    uint64_t* w_full = a_empty + kAStages;
    uint64_t* t1_full = w_full + 1;
    uint64_t* t1_empty = t1_full + 2;
    uint64_t* mid_full = t1_empty + 2;
    uint64_t* mid_empty = mid_full + 2;
    uint64_t* t2_full = mid_empty + 2;
    uint64_t* t2_empty = t2_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(t2_empty + 2);
    float* sbias1 = (float*)((uint8_t*)a_full + kNumBars * 8 + 16);
    float* sbias2 = sbias1 + 128;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x % p.ngroups;
    const int cta = blockIdx.x / p.ngroups, ctas = gridDim.x / p.ngroups;
    const int gt = p.group_tiles[g];
    const int n_my = cta < gt ? (gt - cta + ctas - 1) / ctas : 0;          // tiles cta, cta + ctas, ...

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nprob; ++i) prefetch_tensormap(&p.tmA[i]);
        prefetch_tensormap(&p.tmW1[g]);
        prefetch_tensormap(&p.tmW2[g]);
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&t1_full[s], 1);
            mbar_init(&t1_empty[s], 128);
            mbar_init(&mid_full[s], 128);
            mbar_init(&mid_empty[s], 1);
            mbar_init(&t2_full[s], 1);
            mbar_init(&t2_empty[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    if (threadIdx.x >= 64) {
        const int e = threadIdx.x - 64;
        sbias1[e] = __ldg(p.bias1[g] + e);
        if (e < 64) sbias2[e] = __ldg(p.bias2[g] + e);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer =================
        mbar_arrive_expect_tx_elect(w_full, kW1Bytes + kW2Bytes);
        for (int cc = 0; cc < 2; ++cc) {
            tma_load_2d_elect(w1s + cc * (128 * 128), &p.tmW1[g], w_full, cc * 64, 0);
            tma_load_2d_elect(w2s + cc * (64 * 128), &p.tmW2[g], w_full, cc * 64, 0);
        }
        int stage = 0;
        uint32_t phase = 0;
        pdl_wait();                     // weights are on their way; the activations are the previous kernel's output
        for (int i = 0; i < n_my; ++i) {
            int pi, pix0;
            decode(p, g, cta + i * ctas, pi, pix0);
            for (int cc = 0; cc < 2; ++cc) {
                mbar_wait(&a_empty[stage], phase ^ 1, 20);
                mbar_arrive_expect_tx_elect(&a_full[stage], kChunkBytes);
                tma_load_2d_elect(a_ring + stage * kChunkBytes, &p.tmA[pi], &a_full[stage], cc * 64, pix0);
                if (++stage == kAStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc1 = make_idesc(128), idesc2 = make_idesc(64);
        mbar_wait(w_full, 0, 21);
        tc_fence_after();
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i <= n_my; ++i) {
            if (i < n_my) {                                            // GEMM1 of tile i -> acc1[i & 1]
                const int s = i & 1;
                mbar_wait(&t1_empty[s], ((i >> 1) & 1) ^ 1, 22);
                tc_fence_after();
                const uint32_t d1 = tmem_base + s * 128;
                for (int cc = 0; cc < 2; ++cc) {
                    mbar_wait(&a_full[stage], phase, 23);
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_desc(smem_u32(a_ring + stage * kChunkBytes));
                    const uint64_t bdesc = make_sw128_desc(smem_u32(w1s + cc * (128 * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_elect(d1, adesc + 2 * k, bdesc + 2 * k, idesc1, (cc | k) != 0);
                    umma_commit_elect(&a_empty[stage]);
                    if (++stage == kAStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_elect(&t1_full[s]);
            }
            if (i > 0) {                                               // GEMM2 of tile i-1 -> acc2[(i-1) & 1]
                const int j = i - 1, s = j & 1;
                mbar_wait(&mid_full[s], (j >> 1) & 1, 24);
                mbar_wait(&t2_empty[s], ((j >> 1) & 1) ^ 1, 25);
                tc_fence_after();
                const uint32_t d2 = tmem_base + 256 + s * 64;
                for (int cc = 0; cc < 2; ++cc) {
                    const uint64_t adesc = make_sw128_desc(smem_u32(mids + s * kMidBytes + cc * kChunkBytes));
                    const uint64_t bdesc = make_sw128_desc(smem_u32(w2s + cc * (64 * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_elect(d2, adesc + 2 * k, bdesc + 2 * k, idesc2, (cc | k) != 0);
                }
                umma_commit_elect(&mid_empty[s]);
                umma_commit_elect(&t2_full[s]);
            }
        }
    } else {
        // ================= epilogues (warps 2..5) =================
        const int quad = warp & 3;                                     // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;                              // pixel within the tile
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int i = 0; i <= n_my; ++i) {
            if (i < n_my) {                                            // epi1: acc1 -> bias, ReLU, bf16 -> mid[i & 1]
                const int s = i & 1;
                mbar_wait(&t1_full[s], (i >> 1) & 1, 26);
                mbar_wait(&mid_empty[s], ((i >> 1) & 1) ^ 1, 27);      // GEMM2 of tile i-2 has finished reading this slot
                tc_fence_after();
                uint8_t* mid = mids + s * kMidBytes;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + s * 128 + c0, v);
                    float f[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) f[q] = fmaxf(__uint_as_float(v[q]) + sbias1[c0 + q], 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ch = c0 + 8 * u;                     // first of 8 channels = one 16-byte unit
                        const int half = ch >> 6, unit = (ch & 63) >> 3;
                        *(uint4*)(mid + half * kChunkBytes + row * 128 + ((unit ^ (row & 7)) << 4)) = pack_bf16x8(&f[8 * u]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&t1_empty[s]);                             // acc1 slot may be overwritten
                fence_proxy_async_smem();                              // generic-proxy stores -> visible to the UMMA
                mbar_arrive(&mid_full[s]);
            }
            if (i > 0) {                                               // epi2: acc2 -> bias (ReLU) -> global
                const int j = i - 1, s = j & 1;
                int pi, pix0;
                decode(p, g, cta + j * ctas, pi, pix0);
                const TailProb& q = p.prob[pi];
                mbar_wait(&t2_full[s], (j >> 1) & 1, 28);
                tc_fence_after();
                const int pix = pix0 + row;
                const bool relu2 = q.flags & FLAG_RELU2;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    if (c0 >= q.cout_store) break;
                    uint32_t v[32];
                    tmem_ld32(lane_base + 256 + s * 64 + c0, v);
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const float a = __uint_as_float(v[k]) + sbias2[c0 + k];
                        f[k] = relu2 ? fmaxf(a, 0.f) : a;
                    }
                    if (pix < q.npix) {
                        const size_t off = (size_t)pix * q.out_cstride + c0;
                        if (q.flags & FLAG_F32) {
                            float* o = (float*)q.out + off;
#pragma unroll
                            for (int k = 0; k < 32; k += 4)
                                if (c0 + k < q.cout_store) *(float4*)(o + k) = make_float4(f[k], f[k + 1], f[k + 2], f[k + 3]);
                        } else {
                            store_bf16x32((__nv_bfloat16*)q.out + off, f, q.cout_store - c0);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&t2_empty[s]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ================================================================================================================
// Wide tail: conv5_4 (1x1, 128 -> 512, ReLU) + conv5_5 (1x1, 512 -> 38 | 19) of stage 1 (src/model.py:55-62) and
// conv6_1 / conv6_2 of the hand net (:163-164) as one kernel.  The 512-channel intermediate (232 MB per batch of 8
// frames when it goes through HBM) is produced and consumed in four 128-channel slices:
//     for q in 0..3:  acc1 = A x W1[q]       (K = 128, N = 128)    -> bias, ReLU, bf16 -> shared memory ("mid")
//                     acc2 += mid x W2[:, q]  (K = 128, N = 64)
// (tile, q) pairs form one linear sequence of steps; GEMM1(g) is issued before GEMM2(g - 1), the epilogue of step g runs
// while the tensor core works on step g + 1.  The weights (128 KB + 64 KB per branch) do not fit beside the tiles: the
// slices stream through two-slot rings (L2 resident), the A tile of 128 pixels stays for its four steps.
//   warp 0: TMA producer (A tile; W1 / W2 slice rings)   warp 1: MMA issuer   warps 2..5: epilogues
// ================================================================================================================
constexpr int kWSlices = 4;
constexpr int kW1SliceBytes = 2 * 128 * 128;        // 2 K-chunks x 128 rows x 128 B
constexpr int kW2SliceBytes = 2 * 64 * 128;         // 2 K-chunks x  64 rows x 128 B
constexpr int kWideBars = 2 + 2 * 2 + 2 * 2 + 6 * 2;
constexpr int kWideSmem = 1024 + kMidBytes /*A tile*/ + 2 * kW1SliceBytes + 2 * kW2SliceBytes + 2 * kMidBytes + kWideBars * 8 + 16 +
                          (512 + 64) * 4;
static_assert(kWideSmem <= 232448, "shared memory budget");

struct alignas(64) WideParams {
    CUtensorMap tmA[kMaxProb];
    CUtensorMap tmW1[kMaxGroups];
    CUtensorMap tmW2[kMaxGroups];
    TailProb prob[kMaxProb];
    const float* bias1[kMaxGroups];
    const float* bias2[kMaxGroups];
    int group_tiles[kMaxGroups];
    int nprob, ngroups;
};
static_assert(sizeof(WideParams) <= 4000, "kernel parameter space");

__device__ __forceinline__ void decode_wide(const WideParams& p, int g, int t, int& pi, int& pix0) {
    pi = -1;
    for (int i = 0; i < p.nprob; ++i)
        if (p.prob[i].group == g && t >= p.prob[i].tile_begin && t < p.prob[i].tile_begin + p.prob[i].tiles) pi = i;
    pix0 = (t - p.prob[pi].tile_begin) * kTilePx;
}

__global__ void __launch_bounds__(kThreads, 1) conv_tail_wide_kernel(const __grid_constant__ WideParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_tile = smem;
    uint8_t* w1s = a_tile + kMidBytes;
    uint8_t* w2s = w1s + 2 * kW1SliceBytes;
    uint8_t* mids = w2s + 2 * kW2SliceBytes;
    uint64_t* a_full = (uint64_t*)(mids + 2 * kMidBytes);
    uint64_t* a_empty = a_full + 1;
    uint64_t* w1_full = a_empty + 1;
    uint64_t* w1_empty = w1_full + 2;
    uint64_t* w2_full = w1_empty + 2;
    uint64_t* w2_empty = w2_full + 2;
    uint64_t* t1_full = w2_empty + 2;
    uint64_t* t1_empty = t1_full + 2;
    uint64_t* mid_full = t1_empty + 2;
    uint64_t* mid_empty = mid_full + 2;
    uint64_t* t2_full = mid_empty + 2;
    uint64_t* t2_empty = t2_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(t2_empty + 2);
    float* sbias1 = (float*)((uint8_t*)a_full + kWideBars * 8 + 16);
    float* sbias2 = sbias1 + 512;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x % p.ngroups;
    const int cta = blockIdx.x / p.ngroups, ctas = gridDim.x / p.ngroups;
    const int gt = p.group_tiles[g];
    const int n_my = cta < gt ? (gt - cta + ctas - 1) / ctas : 0;          // tiles cta, cta + ctas, ...
    const int n_steps = n_my * kWSlices;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nprob; ++i) prefetch_tensormap(&p.tmA[i]);
        prefetch_tensormap(&p.tmW1[g]);
        prefetch_tensormap(&p.tmW2[g]);
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&w1_full[s], 1);
            mbar_init(&w1_empty[s], 1);
            mbar_init(&w2_full[s], 1);
            mbar_init(&w2_empty[s], 1);
            mbar_init(&t1_full[s], 1);
            mbar_init(&t1_empty[s], 128);
            mbar_init(&mid_full[s], 128);
            mbar_init(&mid_empty[s], 1);
            mbar_init(&t2_full[s], 1);
            mbar_init(&t2_empty[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    for (int e = threadIdx.x; e < 512 + 64; e += kThreads) {
        if (e < 512) sbias1[e] = __ldg(p.bias1[g] + e);
        else sbias2[e - 512] = __ldg(p.bias2[g] + (e - 512));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer =================
        pdl_wait();                     // the activations are the previous kernel's output
        for (int i = 0; i < n_my; ++i) {
            int pi, pix0;
            decode_wide(p, g, cta + i * ctas, pi, pix0);
            mbar_wait(a_empty, (i & 1) ^ 1, 40);
            mbar_arrive_expect_tx_elect(a_full, kMidBytes);
            for (int cc = 0; cc < 2; ++cc) tma_load_2d_elect(a_tile + cc * kChunkBytes, &p.tmA[pi], a_full, cc * 64, pix0);
            for (int q = 0; q < kWSlices; ++q) {
                const int st = i * kWSlices + q, slot = st & 1;
                const uint32_t ph = ((st >> 1) & 1) ^ 1;
                mbar_wait(&w1_empty[slot], ph, 41);
                mbar_arrive_expect_tx_elect(&w1_full[slot], kW1SliceBytes);
                for (int cc = 0; cc < 2; ++cc)
                    tma_load_2d_elect(w1s + slot * kW1SliceBytes + cc * (128 * 128), &p.tmW1[g], &w1_full[slot], cc * 64, q * 128);
                mbar_wait(&w2_empty[slot], ph, 42);
                mbar_arrive_expect_tx_elect(&w2_full[slot], kW2SliceBytes);
                for (int cc = 0; cc < 2; ++cc)
                    tma_load_2d_elect(w2s + slot * kW2SliceBytes + cc * (64 * 128), &p.tmW2[g], &w2_full[slot], q * 128 + cc * 64, 0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc1 = make_idesc(128), idesc2 = make_idesc(64);
        for (int st = 0; st <= n_steps; ++st) {
            if (st < n_steps) {                                        // GEMM1 of step st -> acc1[st & 1]
                const int tile = st / kWSlices, q = st - tile * kWSlices, slot = st & 1;
                const uint32_t ph = (st >> 1) & 1;
                if (q == 0) mbar_wait(a_full, tile & 1, 43);
                mbar_wait(&w1_full[slot], ph, 44);
                mbar_wait(&t1_empty[slot], ph ^ 1, 45);
                tc_fence_after();
                const uint32_t d1 = tmem_base + slot * 128;
                for (int cc = 0; cc < 2; ++cc) {
                    const uint64_t adesc = make_sw128_desc(smem_u32(a_tile + cc * kChunkBytes));
                    const uint64_t bdesc = make_sw128_desc(smem_u32(w1s + slot * kW1SliceBytes + cc * (128 * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_elect(d1, adesc + 2 * k, bdesc + 2 * k, idesc1, (cc | k) != 0);
                }
                umma_commit_elect(&w1_empty[slot]);
                umma_commit_elect(&t1_full[slot]);
                if (q == kWSlices - 1) umma_commit_elect(a_empty);
            }
            if (st > 0) {                                              // GEMM2 of step st-1 -> acc2[tile & 1]
                const int j = st - 1, tile = j / kWSlices, q = j - tile * kWSlices, slot = j & 1, ts = tile & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&mid_full[slot], ph, 46);
                mbar_wait(&w2_full[slot], ph, 47);
                if (q == 0) mbar_wait(&t2_empty[ts], ((tile >> 1) & 1) ^ 1, 48);
                tc_fence_after();
                const uint32_t d2 = tmem_base + 256 + ts * 64;
                for (int cc = 0; cc < 2; ++cc) {
                    const uint64_t adesc = make_sw128_desc(smem_u32(mids + slot * kMidBytes + cc * kChunkBytes));
                    const uint64_t bdesc = make_sw128_desc(smem_u32(w2s + slot * kW2SliceBytes + cc * (64 * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_elect(d2, adesc + 2 * k, bdesc + 2 * k, idesc2, (q | cc | k) != 0);
                }
                umma_commit_elect(&mid_empty[slot]);
                umma_commit_elect(&w2_empty[slot]);
                if (q == kWSlices - 1) umma_commit_elect(&t2_full[ts]);
            }
        }
    } else {
        // ================= epilogues (warps 2..5) =================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int st = 0; st <= n_steps; ++st) {
            if (st < n_steps) {                                        // epi1: acc1 -> bias, ReLU, bf16 -> mid[st & 1]
                const int q = st % kWSlices, slot = st & 1;
                const uint32_t ph = (st >> 1) & 1;
                mbar_wait(&t1_full[slot], ph, 49);
                mbar_wait(&mid_empty[slot], ph ^ 1, 50);
                tc_fence_after();
                uint8_t* mid = mids + slot * kMidBytes;
                const float* b1 = sbias1 + q * 128;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + slot * 128 + c0, v);
                    float f[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) f[e] = fmaxf(__uint_as_float(v[e]) + b1[c0 + e], 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ch = c0 + 8 * u;
                        const int half = ch >> 6, unit = (ch & 63) >> 3;
                        *(uint4*)(mid + half * kChunkBytes + row * 128 + ((unit ^ (row & 7)) << 4)) = pack_bf16x8(&f[8 * u]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&t1_empty[slot]);
                fence_proxy_async_smem();
                mbar_arrive(&mid_full[slot]);
            }
            if (st > 0 && (st - 1) % kWSlices == kWSlices - 1) {       // epi2 of the tile whose last slice was step st-1
                const int tile = (st - 1) / kWSlices, ts = tile & 1;
                int pi, pix0;
                decode_wide(p, g, cta + tile * ctas, pi, pix0);
                const TailProb& q = p.prob[pi];
                mbar_wait(&t2_full[ts], (tile >> 1) & 1, 51);
                tc_fence_after();
                const int pix = pix0 + row;
                const bool relu2 = q.flags & FLAG_RELU2;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    if (c0 >= q.cout_store) break;
                    uint32_t v[32];
                    tmem_ld32(lane_base + 256 + ts * 64 + c0, v);
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const float a = __uint_as_float(v[k]) + sbias2[c0 + k];
                        f[k] = relu2 ? fmaxf(a, 0.f) : a;
                    }
                    if (pix < q.npix) {
                        const size_t off = (size_t)pix * q.out_cstride + c0;
                        if (q.flags & FLAG_F32) {
                            float* o = (float*)q.out + off;
#pragma unroll
                            for (int k = 0; k < 32; k += 4)
                                if (c0 + k < q.cout_store) *(float4*)(o + k) = make_float4(f[k], f[k + 1], f[k + 2], f[k + 3]);
                        } else {
                            store_bf16x32((__nv_bfloat16*)q.out + off, f, q.cout_store - c0);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&t2_empty[ts]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

struct WideLaunch : ConvLaunch {
    WideParams params;
    int grid = 0;
    void run(cudaStream_t stream) const override {
        static bool attr[64] = {};
        if (first_use_on_device(attr))
            OPB_CUDA(cudaFuncSetAttribute(conv_tail_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideSmem));
        launch_pdl(conv_tail_wide_kernel, grid, kThreads, kWideSmem, stream, params);
        OPB_CUDA(cudaGetLastError());
    }
};

struct TailLaunch : ConvLaunch {
    TailParams params;
    int grid = 0;
    void run(cudaStream_t stream) const override {
        static bool attr[64] = {};
        if (first_use_on_device(attr))
            OPB_CUDA(cudaFuncSetAttribute(conv_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        launch_pdl(conv_tail_kernel, grid, kThreads, kSmemBytes, stream, params);
        OPB_CUDA(cudaGetLastError());
    }
};

}  // namespace

bool conv_tail_supported(const std::vector<TailOp>& ops) {
    if (ops.empty() || (int)ops.size() > kMaxProb) return false;
    int groups = 0;
    const void* seen[kMaxGroups] = {nullptr, nullptr};
    for (const TailOp& op : ops) {
        if (op.in.elem != 2 || op.in.c != 128 || op.in.cstride % 8 || op.in.coff % 8) return false;
        if (op.cout_pad2 != 64 || op.cout_store % 8 || op.cout_store > 64) return false;
        if (op.out.elem != 2 && op.out.elem != 4) return false;
        if ((op.out.coff * op.out.elem) % 16 || (op.out.cstride * op.out.elem) % 16) return false;
        if (op.out.n != op.in.n || op.out.h != op.in.h || op.out.w != op.in.w) return false;
        if ((long long)op.in.n * op.in.h * op.in.w >= (1ll << 31)) return false;
        bool known = false;
        for (int k = 0; k < groups; ++k) known = known || seen[k] == (const void*)op.w1;
        if (!known) {
            if (groups == kMaxGroups) return false;
            seen[groups++] = (const void*)op.w1;
        }
    }
    return true;
}

// wide tail: in 128 ch -> 512 ch (ReLU) -> cout_store (<= 64) channels; w1 [512][128], w2 [64][512], K-major
bool conv_tail_wide_supported(const std::vector<TailOp>& ops) {
    return conv_tail_supported(ops);            // same constraints on the views; the weights differ in shape only
}

ConvLaunch* conv_tail_wide_plan(const std::vector<TailOp>& ops, int num_sms) {
    OPB_REQUIRE(conv_tail_wide_supported(ops), "conv_tail_wide: unsupported problem set");
    auto L = std::make_unique<WideLaunch>();
    WideParams& P = L->params;
    memset(&P, 0, sizeof(P));
    P.nprob = (int)ops.size();
    const void* gw[kMaxGroups] = {nullptr, nullptr};
    int total = 0;
    for (int i = 0; i < P.nprob; ++i) {
        const TailOp& op = ops[i];
        int g = -1;
        for (int k = 0; k < P.ngroups; ++k)
            if (gw[k] == (const void*)op.w1) g = k;
        if (g < 0) {
            g = P.ngroups++;
            gw[g] = (const void*)op.w1;
            P.bias1[g] = op.b1;
            P.bias2[g] = op.b2;
            cuuint64_t d1[2] = {128, 512}, s1[1] = {128 * 2};
            cuuint32_t b1[2] = {64, 128};
            tensor_map_encode_bf16(&P.tmW1[g], (void*)op.w1, 2, d1, s1, b1);
            cuuint64_t d2[2] = {512, 64}, s2[1] = {512 * 2};
            cuuint32_t b2[2] = {64, 64};
            tensor_map_encode_bf16(&P.tmW2[g], (void*)op.w2, 2, d2, s2, b2);
        }
        TailProb& q = P.prob[i];
        q.out = op.out.ptr();
        q.npix = op.in.n * op.in.h * op.in.w;
        q.tiles = cdiv(q.npix, kTilePx);
        q.tile_begin = P.group_tiles[g];
        P.group_tiles[g] += q.tiles;
        q.out_cstride = op.out.cstride;
        q.cout_store = op.cout_store;
        q.flags = (op.relu2 ? FLAG_RELU2 : 0) | (op.out.elem == 4 ? FLAG_F32 : 0);
        q.group = g;
        total += q.tiles;
        cuuint64_t ad[2] = {128, (cuuint64_t)q.npix}, as[1] = {(cuuint64_t)op.in.cstride * 2};
        cuuint32_t ab[2] = {64, (cuuint32_t)kTilePx};
        tensor_map_encode_bf16(&P.tmA[i], op.in.ptr(), 2, ad, as, ab);
    }
    int max_gt = 0;
    for (int g = 0; g < P.ngroups; ++g) max_gt = std::max(max_gt, P.group_tiles[g]);
    const int per_group = std::max(1, std::min(num_sms / P.ngroups, max_gt));
    L->grid = per_group * P.ngroups;
    L->tiles = total;
    return L.release();
}

ConvLaunch* conv_tail_plan(const std::vector<TailOp>& ops, int num_sms) {
    OPB_REQUIRE(conv_tail_supported(ops), "conv_tail: unsupported problem set");
    auto L = std::make_unique<TailLaunch>();
    TailParams& P = L->params;
    memset(&P, 0, sizeof(P));
    P.nprob = (int)ops.size();
    const void* gw[kMaxGroups] = {nullptr, nullptr};
    int total = 0;
    for (int i = 0; i < P.nprob; ++i) {
        const TailOp& op = ops[i];
        int g = -1;
        for (int k = 0; k < P.ngroups; ++k)
            if (gw[k] == (const void*)op.w1) g = k;
        if (g < 0) {
            g = P.ngroups++;
            gw[g] = (const void*)op.w1;
            P.bias1[g] = op.b1;
            P.bias2[g] = op.b2;
            // W1: [128 rows][K = 128], W2: [64 rows][K = 128], K-major
            cuuint64_t d1[2] = {128, 128}, s1[1] = {128 * 2};
            cuuint32_t b1[2] = {64, 128};
            tensor_map_encode_bf16(&P.tmW1[g], (void*)op.w1, 2, d1, s1, b1);
            cuuint64_t d2[2] = {128, 64}, s2[1] = {128 * 2};
            cuuint32_t b2[2] = {64, 64};
            tensor_map_encode_bf16(&P.tmW2[g], (void*)op.w2, 2, d2, s2, b2);
        }
        TailProb& q = P.prob[i];
        q.out = op.out.ptr();
        q.npix = op.in.n * op.in.h * op.in.w;
        q.tiles = cdiv(q.npix, kTilePx);
        q.tile_begin = P.group_tiles[g];
        P.group_tiles[g] += q.tiles;
        q.out_cstride = op.out.cstride;
        q.cout_store = op.cout_store;
        q.flags = (op.relu2 ? FLAG_RELU2 : 0) | (op.out.elem == 4 ? FLAG_F32 : 0);
        q.group = g;
        total += q.tiles;
        // A: flat pixel list {C = 128, pixels}
        cuuint64_t ad[2] = {128, (cuuint64_t)q.npix}, as[1] = {(cuuint64_t)op.in.cstride * 2};
        cuuint32_t ab[2] = {64, (cuuint32_t)kTilePx};
        tensor_map_encode_bf16(&P.tmA[i], op.in.ptr(), 2, ad, as, ab);
    }
    int max_gt = 0;
    for (int g = 0; g < P.ngroups; ++g) max_gt = std::max(max_gt, P.group_tiles[g]);
    const int per_group = std::max(1, std::min(num_sms / P.ngroups, max_gt));
    L->grid = per_group * P.ngroups;
    L->tiles = total;
    return L.release();
}

}  // namespace opb
