// PAF line-integral scoring, greedy per-limb matching and person (subset) assembly -- src/body.py:96-212,
// entirely on the device, in float64 with the reference's operation order so that every threshold decision,
// sort position and accumulated score is bit-identical to the float64 Python code run on the same maps.
// Compiled with --fmad=false; multiplies and adds that numpy performs as separate roundings use _rn intrinsics.
//
//   paf_score_kernel  one warp per candidate pair (i in part A, j in part B) of a limb: lanes 0..9 gather the
//                     mid_num=10 nearest-pixel samples of the limb's two PAF channels (np.linspace + round
//                     half-even), lane 0 folds them in index order (Python's builtin sum), applies the distance
//                     prior min(0.5*H/norm-1, 0) and both criteria, and appends survivors to the limb's list.
//   limb_match_kernel one CTA per limb: rank-sorts survivors by (score desc, i asc, j asc) == Python's stable
//                     sorted(..., reverse=True) over the (i, j) loop order, then walks them greedily.
//   assemble_kernel   one warp: the reference's sequential row merge (found==1 / found==2 / new row, k < 17),
//                     rows kept in shared memory (in a global work buffer beyond 1024 rows), row search parallel
//                     over lanes, then pruning.
// The PAF values come either from materialised planes or straight from the low-resolution net outputs
// (composite.cuh: the same fmaf chains as the materialising kernels, so the samples are bit-identical to the planes
// opb_body_maps returns) -- a frame needs 10 samples per candidate pair, not 38 full-resolution planes.
// Every kernel handles all frames of a batch: the frame is a grid dimension.
#include "composite.cuh"
#include <algorithm>

namespace opb {
namespace {

constexpr int kLimbs = 19;
constexpr int kMid = 10;
__constant__ int c_limb_a[kLimbs] = {1, 1, 2, 3, 5, 6, 1, 8, 9, 1, 11, 12, 1, 0, 14, 0, 15, 2, 5};     // limbSeq-1
__constant__ int c_limb_b[kLimbs] = {2, 5, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 0, 14, 16, 15, 17, 16, 17};
__constant__ int c_paf_x[kLimbs] = {12, 20, 14, 16, 22, 24, 0, 2, 4, 6, 8, 10, 28, 30, 34, 32, 36, 18, 26};  // mapIdx-19

constexpr int ST_PAIR_OVERFLOW = 1, ST_CONN_OVERFLOW = 2, ST_SUBSET_OVERFLOW = 4, ST_INDEX_ERROR = 8;

template <bool PLANAR>
__global__ void __launch_bounds__(256) paf_score_kernel(const __grid_constant__ MapSource src, int H, int W,
                                                        const FramePost* __restrict__ frames, double thre2) {
    const int k = blockIdx.y;
    const int frame = blockIdx.z;
    const FramePost& fr = frames[frame];
    const double* __restrict__ cand = fr.pb.candidates;
    const int* __restrict__ part_begin = fr.pb.part_begin;
    const LimbBuffers& lb = fr.lb;
    const int pa = c_limb_a[k], pb = c_limb_b[k];
    const int a0 = part_begin[pa], nA = part_begin[pa + 1] - a0;
    const int b0 = part_begin[pb], nB = part_begin[pb + 1] - b0;
    const long long pairs = (long long)nA * nB;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float* px_map = nullptr;
    const float* py_map = nullptr;
    if (PLANAR) {
        px_map = src.planar + ((size_t)(frame + src.frame_base) * src.planes_per_frame + c_paf_x[k]) * H * W;
        py_map = px_map + (size_t)H * W;
    }

    for (long long pr = warp0; pr < pairs; pr += nwarps) {
        const int i = (int)(pr / nB), j = (int)(pr - (long long)i * nB);
        const double ax = cand[(size_t)(a0 + i) * 4], ay = cand[(size_t)(a0 + i) * 4 + 1];
        const double bx = cand[(size_t)(b0 + j) * 4], by = cand[(size_t)(b0 + j) * 4 + 1];
        const double vx = bx - ax, vy = by - ay;                               // exact (integers)
        const double norm = __dadd_rn(sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy))), 1e-10);
        const double ux = vx / norm, uy = vy / norm;
        double dot = 0.0;
        if (lane < kMid) {
            // np.linspace(a, b, 10): arange(10) * ((b-a)/9) + a, last sample forced to b
            double sx, sy;
            if (lane == kMid - 1) {
                sx = bx;
                sy = by;
            } else {
                sx = __dadd_rn(__dmul_rn((double)lane, vx / 9.0), ax);
                sy = __dadd_rn(__dmul_rn((double)lane, vy / 9.0), ay);
            }
            const int xi = (int)rint(sx), yi = (int)rint(sy);                  // int(round()): half to even
            double fx, fy;
            if (PLANAR) {
                fx = (double)px_map[(size_t)yi * W + xi];
                fy = (double)py_map[(size_t)yi * W + xi];
            } else {
                const float2 f = composite_at2(src.comp, frame + src.frame_base, c_paf_x[k], yi, xi);
                fx = (double)f.x;
                fy = (double)f.y;
            }
            dot = __dadd_rn(__dmul_rn(fx, ux), __dmul_rn(fy, uy));
        }
        const unsigned above = __ballot_sync(0xffffffffu, lane < kMid && dot > thre2);
        double total = 0.0;                                                    // sum(): 0 + d0 + d1 + ...
#pragma unroll
        for (int s = 0; s < kMid; ++s) total = __dadd_rn(total, __shfl_sync(0xffffffffu, dot, s));
        if (lane == 0) {
            double prior = __dadd_rn(__dmul_rn(0.5, (double)H) / norm, -1.0);
            if (!(prior < 0.0)) prior = 0.0;                                   // min(x, 0)
            const double score = __dadd_rn(total / (double)kMid, prior);
            if (__popc(above) > 8 && score > 0.0) {                            // > 0.8 * 10
                const int slot = atomicAdd(&lb.cand_count[k], 1);
                if (slot < lb.pair_capacity) {
                    lb.cand_score[(size_t)k * lb.pair_capacity + slot] = score;
                    lb.cand_ij[((size_t)k * lb.pair_capacity + slot) * 2] = i;
                    lb.cand_ij[((size_t)k * lb.pair_capacity + slot) * 2 + 1] = j;
                } else {
                    atomicOr(lb.status, ST_PAIR_OVERFLOW);
                }
            }
        }
    }
}

// order: score descending, then original (i, j) loop order
__device__ __forceinline__ bool cand_before(double s1, long long o1, double s2, long long o2) {
    return s1 > s2 || (s1 == s2 && o1 < o2);
}

__global__ void __launch_bounds__(256) limb_match_kernel(const FramePost* __restrict__ frames) {
    __shared__ double s_score[256];
    __shared__ long long s_ord[256];
    const int k = blockIdx.x;
    const FramePost& fr = frames[blockIdx.y];
    const int* __restrict__ part_begin = fr.pb.part_begin;
    const LimbBuffers& lb = fr.lb;
    const int max_part = lb.max_part;
    const int pa = c_limb_a[k], pb = c_limb_b[k];
    const int a0 = part_begin[pa], nA = part_begin[pa + 1] - a0;
    const int b0 = part_begin[pb], nB = part_begin[pb + 1] - b0;
    const int n = min(lb.cand_count[k], lb.pair_capacity);
    const double* sc = lb.cand_score + (size_t)k * lb.pair_capacity;
    const int* ij = lb.cand_ij + (size_t)k * lb.pair_capacity * 2;
    int* ord = lb.order + (size_t)k * lb.pair_capacity;
    unsigned char* usedA = lb.used + (size_t)k * 2 * max_part;
    unsigned char* usedB = usedA + max_part;

    if (threadIdx.x == 0) lb.conn_count[k] = (nA == 0 || nB == 0) ? -1 : 0;   // -1: limb in special_k
    if (nA == 0 || nB == 0 || n == 0) return;

    for (int t = threadIdx.x; t < nA; t += blockDim.x) usedA[t] = 0;
    for (int t = threadIdx.x; t < nB; t += blockDim.x) usedB[t] = 0;

    // rank sort (candidates are unique in (i, j), so ranks are a permutation)
    for (int base_i = 0; base_i < n; base_i += blockDim.x) {
        const int c = base_i + threadIdx.x;
        const double my_s = c < n ? sc[c] : 0.0;
        const long long my_o = c < n ? (long long)ij[2 * c] * nB + ij[2 * c + 1] : 0;
        int rank = 0;
        for (int base = 0; base < n; base += 256) {
            const int t = base + threadIdx.x;
            if (t < n) {
                s_score[threadIdx.x] = sc[t];
                s_ord[threadIdx.x] = (long long)ij[2 * t] * nB + ij[2 * t + 1];
            }
            __syncthreads();
            const int m = min(256, n - base);
            if (c < n)
                for (int q = 0; q < m; ++q) rank += cand_before(s_score[q], s_ord[q], my_s, my_o);
            __syncthreads();
        }
        if (c < n) ord[rank] = c;
    }
    __syncthreads();

    // greedy walk (src/body.py:143-150) -- inherently sequential; one thread, inputs are L1/L2 resident
    if (threadIdx.x == 0) {
        const int limit = min(nA, nB);
        int count = 0;
        double* conn = lb.conn + (size_t)k * lb.conn_capacity * 5;
        for (int r = 0; r < n && count < limit; ++r) {
            const int c = ord[r];
            const int i = ij[2 * c], j = ij[2 * c + 1];
            if (usedA[i] || usedB[j]) continue;
            usedA[i] = 1;
            usedB[j] = 1;
            if (count < lb.conn_capacity) {
                double* row = conn + (size_t)count * 5;
                row[0] = (double)(a0 + i);      // candidate id of A (ids are global sorted positions)
                row[1] = (double)(b0 + j);
                row[2] = sc[c];
                row[3] = (double)i;
                row[4] = (double)j;
            } else {
                atomicOr(lb.status, ST_CONN_OVERFLOW);
            }
            ++count;
        }
        lb.conn_count[k] = min(count, lb.conn_capacity);
    }
}

// ---- subset assembly: one warp, rows in dynamic shared memory -----------------------------------------
__global__ void __launch_bounds__(32) assemble_kernel(const FramePost* __restrict__ frames) {
    extern __shared__ double rows_shared[];          // [min(subset_capacity, kSubsetRowsShared)][20]
    const FramePost& fr = frames[blockIdx.x];
    const double* __restrict__ cand = fr.pb.candidates;
    const LimbBuffers& lb = fr.lb;
    // work rows: shared memory, or the frame's global work buffer once a frame has needed more rows than fit
    double* rows = lb.rows_global ? lb.rows_global : rows_shared;
    const int lane = threadIdx.x;
    int nrows = 0;
    bool fail = false;

    // connections are staged through shared memory in chunks (coalesced loads, candidate scores gathered in
    // parallel): the sequential merge below then never waits on global memory (it used to pay two dependent
    // global round trips per connection: 730 us for 950 connections)
    constexpr int CH = 256;
    __shared__ double sconn[CH][5];                  // idA, idB, limb score, score(candA), score(candB)
    for (int k = 0; k < kLimbs && !fail; ++k) {
        const int ncon = lb.conn_count[k];
        if (ncon < 0) continue;                      // special_k
        const int ia = c_limb_a[k], ib = c_limb_b[k];
        const double* conn = lb.conn + (size_t)k * lb.conn_capacity * 5;
        for (int base_c = 0; base_c < ncon && !fail; base_c += CH) {
            const int nch = min(CH, ncon - base_c);
            __syncwarp();
            for (int t = lane; t < nch; t += 32) {
                const double* row = conn + (size_t)(base_c + t) * 5;
                const double a = row[0], b = row[1];
                sconn[t][0] = a;
                sconn[t][1] = b;
                sconn[t][2] = row[2];
                sconn[t][3] = cand[(size_t)(int)a * 4 + 2];
                sconn[t][4] = cand[(size_t)(int)b * 4 + 2];
            }
            __syncwarp();
        for (int c = 0; c < nch && !fail; ++c) {
            const double idA = sconn[c][0], idB = sconn[c][1], limb_score = sconn[c][2];
            const double scoreA = sconn[c][3], scoreB = sconn[c][4];
            // rows j with subset[j][indexA] == partAs[i] or subset[j][indexB] == partBs[i]
            int found = 0, j1 = -1, j2 = -1;
            for (int base = 0; base < nrows; base += 32) {
                const int j = base + lane;
                const bool m = j < nrows && (rows[j * 20 + ia] == idA || rows[j * 20 + ib] == idB);
                unsigned mask = __ballot_sync(0xffffffffu, m);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (found == 0) j1 = base + b;
                    else if (found == 1) j2 = base + b;
                    ++found;
                }
            }
            if (found > 2) {                         // reference: IndexError at src/body.py:173
                fail = true;
                if (lane == 0) atomicOr(lb.status, ST_INDEX_ERROR);
                break;
            }
            if (found == 2) {
                // disjoint?  (membership == 2 nowhere over the 18 part slots)
                const bool both = lane < 18 && rows[j1 * 20 + lane] >= 0.0 && rows[j2 * 20 + lane] >= 0.0;
                const bool overlap = __ballot_sync(0xffffffffu, both) != 0;
                if (!overlap) {
                    if (lane < 18) rows[j1 * 20 + lane] = rows[j1 * 20 + lane] + (rows[j2 * 20 + lane] + 1.0);
                    if (lane == 18) rows[j1 * 20 + 18] = (rows[j1 * 20 + 18] + rows[j2 * 20 + 18]) + limb_score;
                    if (lane == 19) rows[j1 * 20 + 19] = rows[j1 * 20 + 19] + rows[j2 * 20 + 19];
                    __syncwarp();
                    // np.delete(subset, j2, 0): shift the tail up by one row
                    const int first = j2 * 20, last = (nrows - 1) * 20;
                    for (int t = first; t < last; t += 32) {
                        const int e = t + lane;
                        double v = 0.0;
                        if (e < last) v = rows[e + 20];
                        __syncwarp();
                        if (e < last) rows[e] = v;
                        __syncwarp();
                    }
                    --nrows;
                } else if (lane == 0) {
                    rows[j1 * 20 + ib] = idB;
                    rows[j1 * 20 + 19] += 1.0;
                    rows[j1 * 20 + 18] += scoreB + limb_score;
                }
            } else if (found == 1) {
                if (lane == 0 && rows[j1 * 20 + ib] != idB) {
                    rows[j1 * 20 + ib] = idB;
                    rows[j1 * 20 + 19] += 1.0;
                    rows[j1 * 20 + 18] += scoreB + limb_score;
                }
            } else if (k < 17) {
                if (nrows >= lb.subset_capacity) {
                    fail = true;
                    if (lane == 0) atomicOr(lb.status, ST_SUBSET_OVERFLOW);
                    break;
                }
                if (lane < 18) rows[nrows * 20 + lane] = lane == ia ? idA : (lane == ib ? idB : -1.0);
                if (lane == 18) rows[nrows * 20 + 18] = ((0.0 + scoreA) + scoreB) + limb_score;
                if (lane == 19) rows[nrows * 20 + 19] = 2.0;
                ++nrows;
            }
            __syncwarp();
        }
        }
    }
    __syncwarp();
    // prune (src/body.py:204-208) and write out in order
    int out = 0;
    for (int base = 0; base < nrows; base += 32) {
        const int j = base + lane;
        bool keep = false;
        if (j < nrows) {
            const double parts = rows[j * 20 + 19], score = rows[j * 20 + 18];
            keep = !(parts < 4.0 || score / parts < 0.4);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int dst = out + __popc(mask & ((1u << lane) - 1));
            for (int q = 0; q < 20; ++q) lb.subset[(size_t)dst * 20 + q] = rows[j * 20 + q];
        }
        out += __popc(mask);
    }
    if (lane == 0) {
        *lb.subset_count = out;
        lb.status[1] = nrows;
    }
}

// counts + the first rows of the ordered candidates and of the pruned subsets -> the frame's result block, which one
// device-to-host copy per batch then moves into pinned memory
__global__ void __launch_bounds__(256) pack_results_kernel(const FramePost* __restrict__ frames) {
    const FramePost& fr = frames[blockIdx.x];
    FrameResults* out = fr.result;
    if (out == nullptr) return;
    const int tid = threadIdx.x;
    if (tid == 0) out->counts[0] = *fr.pb.count;
    if (tid < 19) out->counts[1 + tid] = fr.pb.part_begin[tid];
    if (tid == 32) out->counts[20] = *fr.lb.subset_count;
    if (tid >= 64 && tid < 68) out->counts[21 + tid - 64] = fr.lb.status[tid - 64];
    const int nc = min(min(fr.pb.part_begin[18], fr.pb.capacity), kEagerCand) * 4;
    for (int i = tid; i < nc; i += blockDim.x) out->cand[i] = fr.pb.candidates[i];
    const int ns = min(min(*fr.lb.subset_count, fr.lb.subset_capacity), kEagerSubset) * 20;
    for (int i = tid; i < ns; i += blockDim.x) out->subset[i] = fr.lb.subset[i];
}

}  // namespace

// The per-frame counters (cand_count, status) must be zero on entry: the caller clears the plan's counter slab once
// per batch.  subset_capacity: the largest FramePost::lb.subset_capacity of the batch.
void paf_group_launch(const MapSource& paf, int n_frames, int H, int W, const FramePost* frames_dev, double thre2,
                      int subset_capacity, cudaStream_t stream) {
    dim3 grid(64, kLimbs, n_frames);
    if (paf.planar != nullptr) paf_score_kernel<true><<<grid, 256, 0, stream>>>(paf, H, W, frames_dev, thre2);
    else paf_score_kernel<false><<<grid, 256, 0, stream>>>(paf, H, W, frames_dev, thre2);
    OPB_CUDA(cudaGetLastError());
    limb_match_kernel<<<dim3(kLimbs, n_frames), 256, 0, stream>>>(frames_dev);
    OPB_CUDA(cudaGetLastError());
    static bool attr[64] = {};
    if (first_use_on_device(attr)) {
        OPB_CUDA(cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSubsetRowsShared * 20 * 8));
    }
    const size_t rows_smem = (size_t)std::min(subset_capacity, kSubsetRowsShared) * 20 * 8;
    assemble_kernel<<<n_frames, 32, rows_smem, stream>>>(frames_dev);
    OPB_CUDA(cudaGetLastError());
}

void pack_results_launch(const FramePost* frames_dev, int n_frames, cudaStream_t stream) {
    pack_results_kernel<<<n_frames, 256, 0, stream>>>(frames_dev);
    OPB_CUDA(cudaGetLastError());
}

}  // namespace opb
