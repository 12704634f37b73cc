// Per-frame caller on the device (srcmx/MotionEstimation.py:126-216): from the body results of
// every frame of a batch -- still in device memory -- pick the person, fill the body rows of PoseMat and derive the two
// hand boxes (util.handDetect, src/util.py:133-201); after the hand network, move the hand key points into frame
// coordinates.  PoseMat (60, 3): rows 0-17 body, 18-38 left hand, 39-59 right hand; zeros mean "missing".
// float64 throughout, operation order of the Python code (this file is compiled with --fmad=false).
#include "opb_common.cuh"

namespace opb {
namespace {

// one thread per frame: the work is a few dozen scalar operations
__global__ void pose_select_kernel(const FramePost* __restrict__ frames, int n_frames, int H, int W,
                                   const int* __restrict__ fixed_boxes /* [n][2][3] (x, y, w) left, right or null */,
                                   double* __restrict__ pose, HandBox* __restrict__ boxes, int* __restrict__ dims) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const FramePost& fr = frames[f];
    const double* cand = fr.pb.candidates;
    const int n_cand = fr.pb.part_begin[18];
    const double* subset = fr.lb.subset;
    const int n_sub = *fr.lb.subset_count;
    double* pm = pose + (size_t)f * 180;
    for (int i = 0; i < 180; ++i) pm[i] = 0.0;

    // the person with the largest left-shoulder x (MotionEstimation.py:144-150); a missing shoulder (-1) indexes the
    // LAST candidate, like candidate[-1]; np.argmax keeps the first maximum
    int chosen = -1;
    double best = 0.0;
    for (int p = 0; p < n_sub; ++p) {
        int idx = (int)subset[(size_t)p * 20 + 5];
        if (idx < 0) idx += n_cand;
        const double x = cand[(size_t)idx * 4];
        if (chosen < 0 || x > best) {
            chosen = p;
            best = x;
        }
    }
    HandBox hb[2];
    for (int k = 0; k < 2; ++k) {
        hb[k].frame = f;
        hb[k].x = hb[k].y = hb[k].w = 0;
        hb[k].left = k == 0;
        hb[k].valid = 0;
    }
    if (chosen >= 0) {
        const double* row = subset + (size_t)chosen * 20;
        for (int k = 0; k < 18; ++k) {
            const int idx = (int)row[k];
            if (idx != -1) {
                pm[k * 3 + 0] = cand[(size_t)idx * 4 + 0];
                pm[k * 3 + 1] = cand[(size_t)idx * 4 + 1];
                pm[k * 3 + 2] = cand[(size_t)idx * 4 + 2];
            }
        }
        // util.handDetect on the chosen person (the caller blanks all other rows first, MotionEstimation.py:160-162):
        // left hand from shoulder / elbow / wrist 5, 6, 7, right hand from 2, 3, 4
        for (int k = 0; k < 2; ++k) {
            const int j0 = k == 0 ? 5 : 2;
            const int is = (int)row[j0], ie = (int)row[j0 + 1], iw = (int)row[j0 + 2];
            if (is == -1 || ie == -1 || iw == -1) continue;
            const double x1 = cand[(size_t)is * 4], y1 = cand[(size_t)is * 4 + 1];
            const double x2 = cand[(size_t)ie * 4], y2 = cand[(size_t)ie * 4 + 1];
            const double x3 = cand[(size_t)iw * 4], y3 = cand[(size_t)iw * 4 + 1];
            double x = __dadd_rn(x3, __dmul_rn(0.33, x3 - x2));
            double y = __dadd_rn(y3, __dmul_rn(0.33, y3 - y2));
            const double dwe = sqrt(__dadd_rn(__dmul_rn(x3 - x2, x3 - x2), __dmul_rn(y3 - y2, y3 - y2)));
            const double des = sqrt(__dadd_rn(__dmul_rn(x2 - x1, x2 - x1), __dmul_rn(y2 - y1, y2 - y1)));
            const double m = __dmul_rn(0.9, des);
            double width = __dmul_rn(1.5, dwe > m ? dwe : m);          // max(a, b): a if a > b ... Python's max keeps the first on ties
            x = __dadd_rn(x, -(width / 2));
            y = __dadd_rn(y, -(width / 2));
            if (x < 0) x = 0;
            if (y < 0) y = 0;
            double w1 = width, w2 = width;
            if (__dadd_rn(x, width) > (double)W) w1 = (double)W - x;
            if (__dadd_rn(y, width) > (double)H) w2 = (double)H - y;
            width = w2 < w1 ? w2 : w1;                                 // min(width1, width2)
            hb[k].x = (int)x;
            hb[k].y = (int)y;
            hb[k].w = (int)width;
            hb[k].valid = hb[k].w > 0;        // the reference's Hand divides by the crop height: an empty box raises there
        }
    }
    if (fixed_boxes) {                        // benchmarks / tests: random weights find nobody
        for (int k = 0; k < 2; ++k) {
            const int* b = fixed_boxes + ((size_t)f * 2 + k) * 3;
            hb[k].x = b[0];
            hb[k].y = b[1];
            hb[k].w = b[2];
            // a box that is empty or leaves the frame is no box (util.handDetect never produces one)
            hb[k].valid = b[2] > 0 && b[0] >= 0 && b[1] >= 0 && b[0] + b[2] <= W && b[1] + b[2] <= H;
        }
    }
    for (int k = 0; k < 2; ++k) {
        boxes[f * 2 + k] = hb[k];
        dims[f * 2 + k] = hb[k].valid ? hb[k].w : 0;
    }
}

// hand key points of slot (frame, k) -> PoseMat rows (MotionEstimation.py:185-194): coordinate 0 means "missing" and
// is not offset; left hands were mirrored: x -> w - x - 1 + x0
__global__ void pose_finish_kernel(const HandBox* __restrict__ boxes, const double* __restrict__ peaks, int n_frames,
                                   double* __restrict__ pose) {
    const int slot = blockIdx.x;
    const int j = threadIdx.x;
    if (j >= 21) return;
    const HandBox b = boxes[slot];
    const int f = slot >> 1;
    double* out = pose + (size_t)f * 180 + (size_t)((b.left ? 18 : 39) + j) * 3;
    if (!b.valid) return;                     // rows stay zero
    const double* p = peaks + ((size_t)slot * 21 + j) * 3;
    double x = p[0], y = p[1];
    if (b.left) x = x == 0 ? x : __dadd_rn(__dadd_rn((double)b.w - x, -1.0), (double)b.x);
    else x = x == 0 ? x : __dadd_rn(x, (double)b.x);
    y = y == 0 ? y : __dadd_rn(y, (double)b.y);
    out[0] = x;
    out[1] = y;
    out[2] = p[2];
}

}  // namespace

void pose_select_launch(const FramePost* frames_dev, int n_frames, int H, int W, const int* fixed_boxes_dev, double* pose_dev,
                        HandBox* boxes_dev, int* dims_dev, cudaStream_t stream) {
    pose_select_kernel<<<cdiv(n_frames, 64), 64, 0, stream>>>(frames_dev, n_frames, H, W, fixed_boxes_dev, pose_dev, boxes_dev,
                                                             dims_dev);
    OPB_CUDA(cudaGetLastError());
}

void pose_finish_launch(const HandBox* boxes_dev, const double* hand_peaks_dev, int n_frames, double* pose_dev,
                        cudaStream_t stream) {
    pose_finish_kernel<<<n_frames * 2, 32, 0, stream>>>(boxes_dev, hand_peaks_dev, n_frames, pose_dev);
    OPB_CUDA(cudaGetLastError());
}

}  // namespace opb
