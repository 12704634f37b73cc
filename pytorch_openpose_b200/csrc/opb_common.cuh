// Shared declarations for the sm_100a OpenPose hot path (internal; the public C ABI is
// include/openpose_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include <memory>
#include <cstring>

#include "../../include/openpose_b200.h"

namespace opb {

// ---- error plumbing: every C-ABI entry converts exceptions into a code + thread-local message
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);

#define OPB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            throw opb::Error(OPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define OPB_REQUIRE(cond, msg)                                                \
    do {                                                                      \
        if (!(cond)) throw opb::Error(OPB_ERR_INVALID, std::string(msg));     \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- activation tensor view: NHWC, `c` channels starting at element offset `coff` inside rows of
// `cstride` elements (so a conv can read or write a channel slice of a concat buffer in place)
struct TensorView {
    void* base = nullptr;     // start of the underlying buffer (pixel 0, channel 0)
    int n = 1, h = 0, w = 0;  // images, rows, cols
    int c = 0;                // logical channels of this view
    int cstride = 0;          // elements per pixel in the underlying buffer
    int coff = 0;             // first channel of the view
    int elem = 2;             // bytes per element (2 = bf16, 4 = fp32, 1 = u8)
    size_t pixels() const { return (size_t)n * h * w; }
    void* ptr() const { return (char*)base + (size_t)coff * elem; }
};

// ---- tcgen05 implicit-GEMM convolution (conv_tc.cu)
constexpr int kConvMaxProblems = 8;

struct ConvOp {                 // one problem of a grouped launch
    TensorView in;              // bf16 NHWC, c % 64 == 0
    TensorView out;             // bf16 or fp32 NHWC (pooled dims when pool is set)
    const __nv_bfloat16* w;     // [cout_pad][ks*ks*cin] K-major, cin fastest within a tap
    const float* bias;          // [cout_pad]
    int cout_pad = 0;           // multiple of block_n
    int cout_store = 0;         // channels written (multiple of 8, <= cout_pad)
    int ks = 3;                 // 1, 3 or 7 (stride 1, same padding)
    bool relu = true;
    bool pool = false;          // fused 2x2/2 max-pool of the ReLU output
};

void conv_tc_launch(const std::vector<ConvOp>& ops, int block_n, cudaStream_t stream, int num_sms);
// pre-encoded launch (tensor maps + tile list built once per plan, replayed per frame)
struct ConvLaunch {
    int tiles = 0;
    virtual void run(cudaStream_t stream) const = 0;
    virtual ~ConvLaunch() {}
};
ConvLaunch* conv_tc_plan(const std::vector<ConvOp>& ops, int block_n, int num_sms);               // per-tap tiles (any ks)
ConvLaunch* conv_patch_plan(const std::vector<ConvOp>& ops, int block_n, int num_sms, int mode);   // patch-resident (ks 3/7)
ConvLaunch* conv_pair_plan(const std::vector<ConvOp>& ops, int block_n, int num_sms);              // CTA-pair cta_group::2 (ks 3/7)
// fused 1x1 -> ReLU -> 1x1 tail of a refinement stage (conv_tail.cu): in 128 ch -> 128 ch -> cout_store (<= 64) channels
struct TailOp {
    TensorView in, out;
    const __nv_bfloat16 *w1, *w2;   // [128][128] and [64][128], K-major
    const float *b1, *b2;
    int cout_pad2 = 64, cout_store = 0;
    bool relu2 = false;             // ReLU after the second layer (the stage-6 L2 quirk of src/model.py:30-33)
};
bool conv_tail_supported(const std::vector<TailOp>& ops);
ConvLaunch* conv_tail_plan(const std::vector<TailOp>& ops, int num_sms);
inline void conv_tc_plan_run(const ConvLaunch* L, cudaStream_t stream) { L->run(stream); }
inline void conv_tc_plan_free(ConvLaunch* L) { delete L; }
void tensor_map_encode_bf16(CUtensorMap* tm, void* addr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                            const cuuint32_t* box);

// ---- SIMT kernels (conv_simt.cu)
void conv_direct_launch(const ConvOp& op, cudaStream_t stream);                   // debug / cross-check
void conv_first_launch(const TensorView& in_u8, const TensorView& out, const float* w27x64,
                       const float* bias, cudaStream_t stream);                    // conv1_1 (Cin=3) + ReLU
void maxpool2_launch(const TensorView& in, const TensorView& out, cudaStream_t stream);

// Function attributes (dynamic shared memory size) are per device: true the first time `flags` is consulted on the
// current device (one flag array per kernel; contexts on several GPUs may live in one process).
inline bool first_use_on_device(bool (&flags)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (flags[dev]) return false;
    flags[dev] = true;
    return true;
}

// ---- pre/post processing (prepost.cu)
void preprocess_launch(const uint8_t* img, int H, int W, uint8_t* out, int h, int w, int hp, int wp,
                       const int* x_first, const short* x_coef, const int* y_first, const short* y_coef,
                       cudaStream_t stream);
void preprocess_launch_batched(const uint8_t* img, int n, int H, int W, uint8_t* out, int h, int w, int hp, int wp,
                               const int* x_first, const short* x_coef, const int* y_first, const short* y_coef,
                               cudaStream_t stream);

// batched estimators (srcmx/Batch_model.py): float frames -> bf16 HWC3 net input; 5x5 blur of planar maps
void preprocess_f32_launch(const void* frames, bool frames_u8_hwc, int n, int H, int W, void* out_bf16, int h, int w,
                           int hp, int wp, const int* x_first, const float* x_w, const int* y_first, const float* y_w,
                           cudaStream_t stream);
void blur5_launch(const float* maps_planar, float* out_planar, int n_maps, int H, int W, cudaStream_t stream);

constexpr int kUpStrip = 8;          // output rows per register-blocked strip of the y upsample pass (tables + kernel)
struct UpsampleScale {
    const float* src;         // fp32 NHWC net output
    int ho, wo, cstride;      // source dims and per-pixel stride (elements)
    const int* x_first;       // [W]    first source column of the composite footprint
    const float* x_w;         // [W][6] composite weights
    const int* y_first;       // [H]
    const float* y_w;         // [H][6]
    // optional 16-row strip tables for the register-blocked y pass (null -> generic kernel)
    const int* yb_first = nullptr;     // [ceil(H/16)]
    const int* yb_rows = nullptr;      // [ceil(H/16)]
    const float* yb_w = nullptr;       // [ceil(H/16)][yb_rs][16], 1/n_scales folded in
    int yb_rs = 0;
};
constexpr int kUpTaps = 6;
constexpr int kMaxScales = 8;
void upsample_avg_launch2(const UpsampleScale* scales, int n_scales, int n_img, int C, int H, int W, float* scratch,
                          float* out_planar, cudaStream_t stream);

// ---- peaks (peaks.cu)
struct PeakBuffers {
    unsigned long long* keys;   // [capacity] unordered (part<<40 | y<<20 | x)
    float* scores;              // [capacity] raw map value, parallel to keys
    int* count;                 // [1] number appended (may exceed capacity -> overflow)
    double* candidates;         // [capacity][4] sorted (x, y, score, id)
    int* part_begin;            // [19] prefix offsets per part (18 parts + total)
    int capacity;
};
void smooth_nms_launch(const float* heat_planar, int H, int W, int parts, double thre, PeakBuffers pb,
                       double* smoothed_out /* optional [parts][H][W] or null */, cudaStream_t stream);
void sort_peaks_launch2(PeakBuffers pb, int parts, int* part_count_scratch, cudaStream_t stream);
void nms_f32_launch(const float* blurred_planar, int H, int W, int parts, float thre, PeakBuffers pb, cudaStream_t stream);

// ---- PAF grouping (paf.cu)
struct LimbBuffers {
    double* cand_score;       // [19][pair_capacity]
    int* cand_ij;             // [19][pair_capacity][2] (i, j)
    int* cand_count;          // [19]
    double* conn;             // [19][conn_capacity][5]  (idA, idB, score, i, j)
    int* conn_count;          // [19]
    double* subset;           // [subset_capacity][20]
    int* subset_count;        // [1] rows after pruning
    int* status;              // [4]: bit flags (overflow / IndexError edge), rows before pruning, ...
    int pair_capacity, conn_capacity, subset_capacity;
};
void paf_group_launch2(const float* paf_planar, int H, int W, const double* candidates, const int* part_begin,
                       LimbBuffers lb, double thre2, int* scratch_order, unsigned char* scratch_used, int max_part,
                       cudaStream_t stream);
constexpr int kStPairOverflow = 1, kStConnOverflow = 2, kStSubsetOverflow = 4, kStIndexError = 8;

// ---- hand peaks (hand.cu)
struct HandBuffers {
    int* labels;              // [crops*21][h][w]
    double* sums;             // [crops*21][h][w] per-root sums of raw values
    double* peaks;            // [crops*21][3] output
};
void hand_peaks_launch2(const float* heat_planar, int n_crops, int chan_stride_maps, int h, int w, double thre,
                        HandBuffers hb, double* smoothed_out, cudaStream_t stream);
void hand_peaks_blurred_launch(const float* blurred_planar, int n_crops, int chan_stride_maps, int h, int w, float thre,
                               HandBuffers hb, cudaStream_t stream);

// ---- gaussian taps shared by peaks.cu / hand.cu: scipy.ndimage.gaussian_filter(sigma=3) uses
// radius int(4*3+0.5) = 12 and weights exp(-x^2/18) / sum (src/body.py:75, src/hand.py:62).  The
// constants are the exact float64 bit patterns numpy produces (w[d] = weight at distance d), so the
// device filter can be bit-identical to scipy's.
constexpr int kGaussRadius = 12;
struct GaussTaps {
    double w[kGaussRadius + 1];
};
inline GaussTaps gauss_taps_sigma3() {
    return GaussTaps{{0x1.105a329f98197p-3, 0x1.01a25f86eb137p-3, 0x1.b42a57d56c0bep-4, 0x1.4a614d1afd337p-4,
                      0x1.bfde9c12bec92p-5, 0x1.0fa58939b528fp-5, 0x1.26defcaeb0202p-6, 0x1.1e6bccad344bap-7,
                      0x1.f1e9915139406p-9, 0x1.8345966f69518p-10, 0x1.0d8a5ad43c165p-11, 0x1.4fbe39149e277p-13,
                      0x1.763a210dfb306p-15}};
}

}  // namespace opb
