// PAF line-integral scoring, greedy per-limb matching and person (subset) assembly -- src/body.py:96-212,
// entirely on the device, in float64 with the reference's operation order so that every threshold decision,
// sort position and accumulated score is bit-identical to the float64 Python code run on the same maps.
// Compiled with --fmad=false; multiplies and adds that numpy performs as separate roundings use _rn intrinsics.
//
//   paf_score_kernel  one warp per candidate pair (i in part A, j in part B) of a limb: lanes 0..9 gather the
//                     mid_num=10 nearest-pixel samples of the limb's two PAF channels (np.linspace + round
//                     half-even), lane 0 folds them in index order (Python's builtin sum), applies the distance
//                     prior min(0.5*H/norm-1, 0) and both criteria, and appends survivors to the limb's list.
//   limb_sort_kernel  segmented sort: every (frame, limb) segment of survivors is ordered by (score desc, i asc, j asc)
//                     == Python's stable sorted(..., reverse=True) over the (i, j) loop order, in one launch.
//   limb_match_kernel one CTA per (frame, limb): the greedy walk over the sorted survivors.
//   assemble_kernel   one warp: the reference's sequential row merge (found==1 / found==2 / new row, k < 17),
//                     rows kept in shared memory (in a global work buffer beyond 1024 rows), matching rows found
//                     through per-candidate owner lists, then pruning.
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

// Segmented sort of the survivors: one launch orders every (frame, limb) segment.  Each CTA ranks 256 survivors of its
// segment against the whole segment (keys are unique in (i, j), so the ranks are a permutation); the segment's CTAs
// work in parallel, so a crowded limb is spread over the GPU instead of one SM.
__global__ void __launch_bounds__(256) limb_sort_kernel(const FramePost* __restrict__ frames) {
    __shared__ double s_score[256];
    __shared__ long long s_ord[256];
    const int k = blockIdx.y;
    const FramePost& fr = frames[blockIdx.z];
    const LimbBuffers& lb = fr.lb;
    const int n = min(lb.cand_count[k], lb.pair_capacity);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    const int* __restrict__ part_begin = fr.pb.part_begin;
    const int pb = c_limb_b[k];
    const int nB = part_begin[pb + 1] - part_begin[pb];
    const double* sc = lb.cand_score + (size_t)k * lb.pair_capacity;
    const int* ij = lb.cand_ij + (size_t)k * lb.pair_capacity * 2;
    int* ord = lb.order + (size_t)k * lb.pair_capacity;
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

// greedy walk over the sorted survivors (src/body.py:143-150) -- sequential by definition: one thread walks, the CTA
// stages the sorted entries through shared memory in chunks (coalesced gathers instead of dependent global loads) and
// the "already used" sets are bitmaps in shared memory
__global__ void __launch_bounds__(256) limb_match_kernel(const FramePost* __restrict__ frames) {
    extern __shared__ unsigned used_bits[];          // [2][ceil(max_part / 32)]
    constexpr int CH = 1024;
    __shared__ int s_i[CH], s_j[CH], s_c[CH];
    __shared__ int s_count, s_done;
    const int k = blockIdx.x;
    const FramePost& fr = frames[blockIdx.y];
    const int* __restrict__ part_begin = fr.pb.part_begin;
    const LimbBuffers& lb = fr.lb;
    const int pa = c_limb_a[k], pb = c_limb_b[k];
    const int a0 = part_begin[pa], nA = part_begin[pa + 1] - a0;
    const int b0 = part_begin[pb], nB = part_begin[pb + 1] - b0;
    const int n = min(lb.cand_count[k], lb.pair_capacity);
    const double* sc = lb.cand_score + (size_t)k * lb.pair_capacity;
    const int* ij = lb.cand_ij + (size_t)k * lb.pair_capacity * 2;
    const int* ord = lb.order + (size_t)k * lb.pair_capacity;
    const int words = (lb.max_part + 31) / 32;
    unsigned* usedA = used_bits;
    unsigned* usedB = used_bits + words;

    if (threadIdx.x == 0) lb.conn_count[k] = (nA == 0 || nB == 0) ? -1 : 0;   // -1: limb in special_k
    if (nA == 0 || nB == 0 || n == 0) return;
    for (int t = threadIdx.x; t < 2 * words; t += blockDim.x) used_bits[t] = 0u;
    if (threadIdx.x == 0) {
        s_count = 0;
        s_done = 0;
    }
    const int limit = min(nA, nB);
    double* conn = lb.conn + (size_t)k * lb.conn_capacity * 5;
    for (int base = 0; base < n; base += CH) {
        __syncthreads();
        if (s_done) break;
        const int m = min(CH, n - base);
        for (int t = threadIdx.x; t < m; t += blockDim.x) {
            const int c = ord[base + t];
            s_c[t] = c;
            s_i[t] = ij[2 * c];
            s_j[t] = ij[2 * c + 1];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int count = s_count;
            for (int r = 0; r < m && count < limit; ++r) {
                const int i = s_i[r], j = s_j[r];
                if (((usedA[i >> 5] >> (i & 31)) & 1u) || ((usedB[j >> 5] >> (j & 31)) & 1u)) continue;
                usedA[i >> 5] |= 1u << (i & 31);
                usedB[j >> 5] |= 1u << (j & 31);
                if (count < lb.conn_capacity) {
                    double* row = conn + (size_t)count * 5;
                    row[0] = (double)(a0 + i);      // candidate id of A (ids are global sorted positions)
                    row[1] = (double)(b0 + j);
                    row[2] = sc[s_c[r]];
                    row[3] = (double)i;
                    row[4] = (double)j;
                } else {
                    atomicOr(lb.status, ST_CONN_OVERFLOW);
                }
                ++count;
            }
            s_count = count;
            if (count >= limit) s_done = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) lb.conn_count[k] = min(s_count, lb.conn_capacity);
}

// ---- subset assembly (src/body.py:157-212): one warp ----------------------------------------------------------------
// The reference walks the limbs and their connections in order; for each it looks for the rows that hold candidate A
// in column a or candidate B in column b and extends / merges / creates rows.  The walk is sequential by definition,
// so its cost is the latency of one connection.  Two things keep that short:
//   * owner lists: for every candidate the (at most 3) rows that currently hold it, so "which rows match" is two
//     16-byte loads instead of a scan over all rows (a candidate held by more rows falls back to the scan);
//   * np.delete(subset, j2, 0) after a merge is a tombstone (slot 20 of the row): deleting only shifts later rows up,
//     it never reorders them, so skipping dead rows when the result is written gives the same row order.
// Rows live in shared memory (21 doubles apart: conflict-free column reads in the fallback scan) or, beyond 1024 rows,
// in the frame's global work buffer; the owner lists likewise (shared up to kOwnerShared candidates).
constexpr int kOwnerShared = 2048;

__device__ __forceinline__ void owner_add(int4* own, int id, int row, bool& overflow) {
    int4 o = own[id];
    if (o.x < 0) o.x = row;
    else if (o.y < 0) o.y = row;
    else if (o.z < 0) o.z = row;
    else {
        o.w = 1;                      // more than three rows hold this candidate: scan for it from now on
        overflow = true;
    }
    own[id] = o;
}
__device__ __forceinline__ void owner_remove(int4* own, int id, int row) {
    int4 o = own[id];
    if (o.x == row) o.x = -1;
    else if (o.y == row) o.y = -1;
    else if (o.z == row) o.z = -1;
    own[id] = o;
}
__device__ __forceinline__ void owner_replace(int4* own, int id, int from, int to) {
    int4 o = own[id];
    if (o.x == from) o.x = to;
    else if (o.y == from) o.y = to;
    else if (o.z == from) o.z = to;
    own[id] = o;
}
// smallest, second smallest and number of distinct non-negative values among six entries
__device__ __forceinline__ void two_smallest(const int e[6], int& found, int& j1, int& j2) {
    j1 = 0x7fffffff;
    j2 = 0x7fffffff;
    found = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int v = e[i];
        if (v < 0 || v == j1 || v == j2) continue;
        bool dup = false;
#pragma unroll
        for (int q = 0; q < i; ++q) dup |= e[q] == v;
        if (dup) continue;
        ++found;
        if (v < j1) {
            j2 = j1;
            j1 = v;
        } else if (v < j2) {
            j2 = v;
        }
    }
    if (found < 2) j2 = -1;
    if (found < 1) j1 = -1;
}

__global__ void __launch_bounds__(32) assemble_kernel(const FramePost* __restrict__ frames) {
    constexpr int RS = kSubsetRowStride;
    extern __shared__ double rows_shared[];          // [min(subset_capacity, kSubsetRowsShared)][RS], owner lists, row claims
    const FramePost& fr = frames[blockIdx.x];
    const double* __restrict__ cand = fr.pb.candidates;
    const LimbBuffers& lb = fr.lb;
    double* rows = lb.rows_global ? lb.rows_global : rows_shared;
    const int n_cand = min(fr.pb.part_begin[18], fr.pb.capacity);
    const int rows_in_smem = lb.rows_global ? 0 : min(lb.subset_capacity, kSubsetRowsShared);
    const size_t rows_bytes = ((size_t)rows_in_smem * RS * 8 + 15) & ~(size_t)15;
    int4* own_shared = (int4*)((uint8_t*)rows_shared + rows_bytes);
    int4* own = n_cand <= kOwnerShared ? own_shared : lb.owner_global;
    // claim[row]: lowest lane of the current round that touches the row (INT_MAX when idle)
    int* claim = lb.rows_global ? lb.claim_global : (int*)(own_shared + kOwnerShared);
    const int lane = threadIdx.x;
    int nrows = 0;
    bool fail = false, own_overflow = false;
    for (int t = lane; t < n_cand; t += 32) own[t] = make_int4(-1, -1, -1, 0);
    for (int t = lane; t < lb.subset_capacity; t += 32) claim[t] = 0x7fffffff;

    // connections are staged through shared memory in chunks (coalesced loads, candidate scores gathered in
    // parallel): the walk below then never waits on global memory
    constexpr int CH = 256;
    __shared__ double sconn[CH][5];                  // idA, idB, limb score, score(candA), score(candB)
    for (int k = 0; k < kLimbs && !fail; ++k) {
        const int ncon = lb.conn_count[k];
        if (ncon < 0) continue;                      // special_k
        const int ia = c_limb_a[k], ib = c_limb_b[k];
        const double* conn = lb.conn + (size_t)k * lb.conn_capacity * 5;

        // one connection, by the whole warp, exactly as the reference's loop body (src/body.py:166-202)
        auto serial_one = [&](int c) {
            const double idA = sconn[c][0], idB = sconn[c][1], limb_score = sconn[c][2];
            const double scoreA = sconn[c][3], scoreB = sconn[c][4];
            const int a_id = (int)idA, b_id = (int)idB;
            int found, j1, j2;
            const int4 oa = own[a_id], ob = own[b_id];
            if ((oa.w | ob.w) == 0) {
                const int e[6] = {oa.x, oa.y, oa.z, ob.x, ob.y, ob.z};
                two_smallest(e, found, j1, j2);
            } else {                                 // a candidate held by more than three rows: scan
                found = 0;
                j1 = j2 = -1;
                for (int base = 0; base < nrows; base += 32) {
                    const int j = base + lane;
                    const bool m = j < nrows && rows[j * RS + 20] == 0.0 &&
                                   (rows[j * RS + ia] == idA || rows[j * RS + ib] == idB);
                    unsigned mask = __ballot_sync(0xffffffffu, m);
                    while (mask) {
                        const int bpos = __ffs(mask) - 1;
                        mask &= mask - 1;
                        if (found == 0) j1 = base + bpos;
                        else if (found == 1) j2 = base + bpos;
                        ++found;
                    }
                }
            }
            if (found > 2) {                         // reference: IndexError at src/body.py:173
                fail = true;
                if (lane == 0) atomicOr(lb.status, ST_INDEX_ERROR);
                return;
            }
            bool extend = found == 1;                // "row j1 gets candidate B" (also the overlapping found == 2 case)
            if (found == 2) {
                // disjoint?  (membership == 2 nowhere over the 18 part slots)
                const bool both = lane < 18 && rows[j1 * RS + lane] >= 0.0 && rows[j2 * RS + lane] >= 0.0;
                const bool overlap = __ballot_sync(0xffffffffu, both) != 0;
                if (!overlap) {
                    // subset[j1][:-2] += subset[j2][:-2] + 1 ; tails summed ; + limb score ; row j2 deleted
                    double moved = -1.0;
                    if (lane < 18) {
                        moved = rows[j2 * RS + lane];
                        rows[j1 * RS + lane] = rows[j1 * RS + lane] + (moved + 1.0);
                    }
                    if (lane == 18) rows[j1 * RS + 18] = (rows[j1 * RS + 18] + rows[j2 * RS + 18]) + limb_score;
                    if (lane == 19) rows[j1 * RS + 19] = rows[j1 * RS + 19] + rows[j2 * RS + 19];
                    if (lane == 20) rows[j2 * RS + 20] = 1.0;                     // tombstone
                    __syncwarp();
                    if (lane < 18 && moved >= 0.0) owner_replace(own, (int)moved, j2, j1);
                    __syncwarp();
                } else {
                    extend = true;
                }
            } else if (found == 0 && k < 17) {
                if (nrows >= lb.subset_capacity) {
                    fail = true;
                    if (lane == 0) atomicOr(lb.status, ST_SUBSET_OVERFLOW);
                    return;
                }
                if (lane < 18) rows[nrows * RS + lane] = lane == ia ? idA : (lane == ib ? idB : -1.0);
                if (lane == 18) rows[nrows * RS + 18] = ((0.0 + scoreA) + scoreB) + limb_score;
                if (lane == 19) rows[nrows * RS + 19] = 2.0;
                if (lane == 20) rows[nrows * RS + 20] = 0.0;
                if (lane == 0) {
                    owner_add(own, a_id, nrows, own_overflow);
                    owner_add(own, b_id, nrows, own_overflow);
                }
                ++nrows;
            }
            if (extend && lane == 0) {
                // found == 1 (src/body.py:176-181) and the overlapping found == 2 case (:189-193)
                const double old = rows[j1 * RS + ib];
                if (found == 2 || old != idB) {
                    if (old != idB) {
                        if (old >= 0.0) owner_remove(own, (int)old, j1);          // overwritten, as in the reference
                        owner_add(own, b_id, j1, own_overflow);
                    }
                    rows[j1 * RS + ib] = idB;
                    rows[j1 * RS + 19] += 1.0;
                    rows[j1 * RS + 18] += scoreB + limb_score;
                }
            }
            __syncwarp();
        };

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
            // 32 connections at a time, one per lane.  Connections of one limb have distinct A and distinct B candidates,
            // so two of them interact only through a row both touch.  Every pending lane claims its rows (atomicMin of
            // the lane index); a lane that holds all its claims has no earlier pending connection on its rows and runs
            // now -- the simple cases (new row, plain extension) in parallel, merges / overwrites / owner-list overflows
            // one at a time by the whole warp -- the others wait for the next round.  New rows never conflict and get
            // their indices from a prefix count in connection order, so the result is the sequential walk's, bit for bit.
            for (int g0 = 0; g0 < nch && !fail; g0 += 32) {
                const int c = g0 + lane;
                unsigned pending = __ballot_sync(0xffffffffu, c < nch);
                while (pending && !fail) {
                    const bool mine = (pending >> lane) & 1u;
                    int found = 0, j1 = -1, j2 = -1, a_id = 0, b_id = 0;
                    bool complex_case = false, noop = false, scan = false;
                    double idB = 0.0;
                    if (mine) {
                        a_id = (int)sconn[c][0];
                        idB = sconn[c][1];
                        b_id = (int)idB;
                        const int4 oa = own[a_id], ob = own[b_id];
                        if ((oa.w | ob.w) != 0) {
                            complex_case = scan = true;      // owner lists overflowed: the rows are only found by a scan
                        } else {
                            const int e[6] = {oa.x, oa.y, oa.z, ob.x, ob.y, ob.z};
                            two_smallest(e, found, j1, j2);
                            if (found >= 2) complex_case = true;
                            if (found == 1) {
                                const double old = rows[j1 * RS + ib];
                                noop = old == idB;
                                if (!noop && old >= 0.0) complex_case = true;        // overwrite: owner_remove on a shared entry
                            }
                        }
                    }
                    // a lane on the scan path cannot name its rows: it may only run when it is the lowest pending lane
                    const unsigned scan_lanes = __ballot_sync(0xffffffffu, mine && scan);
                    if (mine && j1 >= 0 && found <= 2) {
                        atomicMin(&claim[j1], lane);
                        if (found == 2) atomicMin(&claim[j2], lane);
                    }
                    __syncwarp();
                    const int lowest = __ffs(pending) - 1;
                    bool first = mine;
                    if (mine && j1 >= 0 && found <= 2) first = claim[j1] == lane && (found < 2 || claim[j2] == lane);
                    if (mine && found > 2) first = lane == lowest;                   // IndexError: reported in order
                    if (mine && ((scan_lanes >> lane) & 1u)) first = lane == lowest;
                    // nobody may run ahead of a pending scan-path lane below it (its rows are unknown)
                    const unsigned below_scan = scan_lanes & ((1u << lane) - 1);
                    if (below_scan) first = false;
                    __syncwarp();
                    if (mine && j1 >= 0 && found <= 2) {                             // release the claims for the next round
                        claim[j1] = 0x7fffffff;
                        if (found == 2) claim[j2] = 0x7fffffff;
                    }
                    const unsigned run = __ballot_sync(0xffffffffu, first);
                    // ---- simple cases in parallel
                    const bool simple = first && !complex_case;
                    const bool make_row = simple && found == 0 && k < 17;
                    const unsigned newmask = __ballot_sync(0xffffffffu, make_row);
                    if (nrows + __popc(newmask) > lb.subset_capacity) {
                        fail = true;
                        if (lane == 0) atomicOr(lb.status, ST_SUBSET_OVERFLOW);
                        break;
                    }
                    if (make_row) {
                        const int r = nrows + __popc(newmask & ((1u << lane) - 1));
                        const double idA = sconn[c][0];
                        for (int q = 0; q < 18; ++q) rows[r * RS + q] = q == ia ? idA : (q == ib ? idB : -1.0);
                        rows[r * RS + 18] = ((0.0 + sconn[c][3]) + sconn[c][4]) + sconn[c][2];
                        rows[r * RS + 19] = 2.0;
                        rows[r * RS + 20] = 0.0;
                        owner_add(own, a_id, r, own_overflow);
                        owner_add(own, b_id, r, own_overflow);
                    }
                    nrows += __popc(newmask);
                    if (simple && found == 1 && !noop) {
                        rows[j1 * RS + ib] = idB;
                        rows[j1 * RS + 19] += 1.0;
                        rows[j1 * RS + 18] += sconn[c][4] + sconn[c][2];
                        owner_add(own, b_id, j1, own_overflow);
                    }
                    __syncwarp();
                    // ---- the rest one connection at a time, in lane order, by the whole warp
                    unsigned serial = __ballot_sync(0xffffffffu, first && complex_case);
                    while (serial && !fail) {
                        const int l = __ffs(serial) - 1;
                        serial &= serial - 1;
                        serial_one(g0 + l);
                    }
                    pending &= ~run;
                }
            }
        }
    }
    __syncwarp();
    // prune (src/body.py:204-208) and write out in order, skipping deleted rows
    int out = 0;
    for (int base = 0; base < nrows; base += 32) {
        const int j = base + lane;
        bool keep = false;
        if (j < nrows && rows[j * RS + 20] == 0.0) {
            const double parts = rows[j * RS + 19], score = rows[j * RS + 18];
            keep = !(parts < 4.0 || score / parts < 0.4);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int dst = out + __popc(mask & ((1u << lane) - 1));
            for (int q = 0; q < 20; ++q) lb.subset[(size_t)dst * 20 + q] = rows[j * RS + q];
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
                      int subset_capacity, int pair_capacity, int max_part, cudaStream_t stream) {
    dim3 grid(64, kLimbs, n_frames);
    if (paf.planar != nullptr) paf_score_kernel<true><<<grid, 256, 0, stream>>>(paf, H, W, frames_dev, thre2);
    else paf_score_kernel<false><<<grid, 256, 0, stream>>>(paf, H, W, frames_dev, thre2);
    OPB_CUDA(cudaGetLastError());
    limb_sort_kernel<<<dim3(cdiv(pair_capacity, 256), kLimbs, n_frames), 256, 0, stream>>>(frames_dev);
    OPB_CUDA(cudaGetLastError());
    const size_t used_smem = (size_t)2 * ((max_part + 31) / 32) * sizeof(unsigned);
    OPB_REQUIRE(used_smem <= 160 * 1024, "limb matching: too many peaks per part for the shared-memory bitmaps");
    static bool mattr[64] = {};
    if (first_use_on_device(mattr))
        OPB_CUDA(cudaFuncSetAttribute(limb_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    limb_match_kernel<<<dim3(kLimbs, n_frames), 256, used_smem, stream>>>(frames_dev);
    OPB_CUDA(cudaGetLastError());
    static bool attr[64] = {};
    if (first_use_on_device(attr)) {
        OPB_CUDA(cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kSubsetRowsShared * kSubsetRowStride * 8 + 16 + kOwnerShared * 16 + kSubsetRowsShared * 4));
    }
    // work rows (shared up to kSubsetRowsShared rows, else the frames' global buffers) + owner lists of the candidates
    const size_t rows_smem = (size_t)std::min(subset_capacity, kSubsetRowsShared) * kSubsetRowStride * 8 + 16;
    assemble_kernel<<<n_frames, 32, rows_smem + kOwnerShared * 16 + kSubsetRowsShared * 4, stream>>>(frames_dev);
    OPB_CUDA(cudaGetLastError());
}

void pack_results_launch(const FramePost* frames_dev, int n_frames, cudaStream_t stream) {
    pack_results_kernel<<<n_frames, 256, 0, stream>>>(frames_dev);
    OPB_CUDA(cudaGetLastError());
}

}  // namespace opb
