// K1 -- neighbour search emitting sorted CSR edge lists, plus the int32 scan it needs.
// Replaces torch_cluster.radius / radius_graph (see include/cb200.h for the call sites).
//
// One warp per query point; the 32 lanes sweep the candidate segment of the query's graph in
// ascending index order, so a ballot + popc gives every hit its rank and the output comes out
// sorted by (query, candidate) with no atomics and no sort.  Candidate segments are a few hundred
// to a few thousand points (one receptor), i.e. L1/L2 resident: the kernel is latency-, not
// bandwidth-bound, and the grid (one warp per query, B*N_lig warps) covers the 148 SMs.
// Bit-exactness: the predicate uses explicitly rounded mul/add/div intrinsics (no FMA contraction),
// the same operation order as the reference's  ((y/c) - (x/c))^2 summed x,y,z  <  r*r.
#include "common.cuh"
#include "../../include/cb200.h"

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ bool in_range(float qx, float qy, float qz, const float* __restrict__ p, float c,
                                         bool scaled, float r2) {
    float px = p[0], py = p[1], pz = p[2];
    if (scaled) {
        px = __fdiv_rn(px, c);
        py = __fdiv_rn(py, c);
        pz = __fdiv_rn(pz, c);
    }
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return d2 < r2;
}

// forward: CSR over queries y, candidates x ascending.
template <bool FILL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
radius_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ x_ptr, const float* __restrict__ y,
                  const int32_t* __restrict__ y_batch, const float* __restrict__ cutoff, float r2, int n_y,
                  int max_nb, int exclude_self, const int32_t* __restrict__ rowptr, int32_t* __restrict__ row,
                  int32_t* __restrict__ col, int32_t* __restrict__ count) {
    const int q = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= n_y) return;
    const int b = y_batch[q];
    const int x0 = x_ptr[b], x1 = x_ptr[b + 1];
    const bool scaled = cutoff != nullptr;
    const float c = scaled ? cutoff[b] : 1.0f;
    float qx = y[3 * q], qy = y[3 * q + 1], qz = y[3 * q + 2];
    if (scaled) {
        qx = __fdiv_rn(qx, c);
        qy = __fdiv_rn(qy, c);
        qz = __fdiv_rn(qz, c);
    }
    const unsigned lt = (1u << lane) - 1u;
    int found = 0, emitted = 0;
    const int base_out = FILL ? rowptr[q] : 0;
    for (int base = x0; base < x1 && found < max_nb; base += 32) {
        const int cand = base + lane;
        const bool hit = cand < x1 && in_range(qx, qy, qz, x + 3 * (size_t)cand, c, scaled, r2);
        const unsigned hb = __ballot_sync(0xffffffffu, hit);
        const int rank = found + __popc(hb & lt);
        const bool keep = hit && rank < max_nb && !(exclude_self && cand == q);
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) {
            const int pos = base_out + emitted + __popc(kb & lt);
            row[pos] = q;
            col[pos] = cand;
        }
        emitted += __popc(kb);
        found += __popc(hb);
    }
    if (!FILL && lane == 0) count[q] = emitted;
}

// is candidate `xc` among the kept neighbours of query `q` (ascending list)?
__device__ __forceinline__ bool kept_contains(const int32_t* __restrict__ kr, const int32_t* __restrict__ kc, int q,
                                              int xc) {
    int lo = kr[q], hi = kr[q + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int v = kc[mid];
        if (v == xc) return true;
        if (v < xc) lo = mid + 1; else hi = mid;
    }
    return false;
}

// transposed: CSR over candidates x, queries y ascending (same edge set as the forward search).
template <bool FILL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
radius_t_kernel(const float* __restrict__ x, const int32_t* __restrict__ x_batch, const float* __restrict__ y,
                const int32_t* __restrict__ y_ptr, const float* __restrict__ cutoff, float r2, int n_x,
                int exclude_self, const int32_t* __restrict__ kept_rowptr, const int32_t* __restrict__ kept_col,
                const int32_t* __restrict__ rowptr_t, int32_t* __restrict__ row_t, int32_t* __restrict__ col_t,
                int32_t* __restrict__ count) {
    const int xc = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (xc >= n_x) return;
    const int b = x_batch[xc];
    const int y0 = y_ptr[b], y1 = y_ptr[b + 1];
    const bool scaled = cutoff != nullptr;
    const float c = scaled ? cutoff[b] : 1.0f;
    float px = x[3 * xc], py = x[3 * xc + 1], pz = x[3 * xc + 2];
    if (scaled) {
        px = __fdiv_rn(px, c);
        py = __fdiv_rn(py, c);
        pz = __fdiv_rn(pz, c);
    }
    const unsigned lt = (1u << lane) - 1u;
    int emitted = 0;
    const int base_out = FILL ? rowptr_t[xc] : 0;
    for (int base = y0; base < y1; base += 32) {
        const int q = base + lane;
        bool hit = false;
        if (q < y1 && !(exclude_self && q == xc)) {
            float qx = y[3 * (size_t)q], qy = y[3 * (size_t)q + 1], qz = y[3 * (size_t)q + 2];
            if (scaled) {
                qx = __fdiv_rn(qx, c);
                qy = __fdiv_rn(qy, c);
                qz = __fdiv_rn(qz, c);
            }
            // same operand order as the forward kernel: (query - candidate)
            const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            hit = d2 < r2;
            if (hit && kept_rowptr != nullptr) hit = kept_contains(kept_rowptr, kept_col, q, xc);
        }
        const unsigned hb = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) {
            const int pos = base_out + emitted + __popc(hb & lt);
            row_t[pos] = xc;
            col_t[pos] = q;
        }
        emitted += __popc(hb);
    }
    if (!FILL && lane == 0) count[xc] = emitted;
}

// ---------------------------------------------------------------- exclusive scan (3 small kernels)
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*[32]*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = smem[lane];
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        smem[lane] = si - s;  // exclusive warp offsets
        if (lane == 31) *total = si;
    }
    __syncthreads();
    return inc - v + smem[w];
}

__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n, int32_t* __restrict__ tile_sums) {
    __shared__ int sm[32];
    __shared__ int total;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems], s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int off = block_exclusive_scan(s, &total, sm);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i + 1] = off + v[i];  // inclusive value at i goes to slot i+1
        off += v[i];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(int32_t* __restrict__ tile_sums, int n_tiles) {
    __shared__ int sm[32];
    __shared__ int total;
    const int base = threadIdx.x * kScanItems;
    int v[kScanItems], s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n_tiles) ? tile_sums[base + i] : 0;
        s += v[i];
    }
    int off = block_exclusive_scan(s, &total, sm);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n_tiles) tile_sums[base + i] = off;
        off += v[i];
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(int32_t* __restrict__ out, int n, const int32_t* __restrict__ tile_sums) {
    const int add = tile_sums[blockIdx.x];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) out[base + i + 1] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

}  // namespace

extern "C" int cb_radius_count(const float* x, const int32_t* x_ptr, const float* y, const int32_t* y_batch,
                               const float* cutoff, float r, int32_t n_y, int32_t max_neighbors,
                               int32_t exclude_self, int32_t* count, void* stream) {
    CB_CHECK_ARG(n_y >= 0 && max_neighbors > 0, "cb_radius_count: bad sizes n_y=%d max=%d", n_y, max_neighbors);
    if (n_y == 0) return CB_OK;
    CB_CHECK_ARG(x && x_ptr && y && y_batch && count, "cb_radius_count: null pointer");
    const float r2 = r * r;
    radius_fwd_kernel<false><<<cb_div_up(n_y, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        x, x_ptr, y, y_batch, cutoff, r2, n_y, max_neighbors, exclude_self, nullptr, nullptr, nullptr, count);
    CB_CHECK_LAUNCH("cb_radius_count");
    return CB_OK;
}

extern "C" int cb_radius_fill(const float* x, const int32_t* x_ptr, const float* y, const int32_t* y_batch,
                              const float* cutoff, float r, int32_t n_y, int32_t max_neighbors,
                              int32_t exclude_self, const int32_t* rowptr, int32_t* row, int32_t* col,
                              void* stream) {
    CB_CHECK_ARG(n_y >= 0 && max_neighbors > 0, "cb_radius_fill: bad sizes n_y=%d max=%d", n_y, max_neighbors);
    if (n_y == 0) return CB_OK;
    CB_CHECK_ARG(x && x_ptr && y && y_batch && rowptr && row && col, "cb_radius_fill: null pointer");
    const float r2 = r * r;
    radius_fwd_kernel<true><<<cb_div_up(n_y, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        x, x_ptr, y, y_batch, cutoff, r2, n_y, max_neighbors, exclude_self, rowptr, row, col, nullptr);
    CB_CHECK_LAUNCH("cb_radius_fill");
    return CB_OK;
}

extern "C" int cb_radius_count_t(const float* x, const int32_t* x_batch, const float* y, const int32_t* y_ptr,
                                 const float* cutoff, float r, int32_t n_x, int32_t exclude_self,
                                 const int32_t* kept_rowptr, const int32_t* kept_col, int32_t* count,
                                 void* stream) {
    CB_CHECK_ARG(n_x >= 0, "cb_radius_count_t: bad size n_x=%d", n_x);
    if (n_x == 0) return CB_OK;
    CB_CHECK_ARG(x && x_batch && y && y_ptr && count, "cb_radius_count_t: null pointer");
    CB_CHECK_ARG((kept_rowptr == nullptr) == (kept_col == nullptr), "cb_radius_count_t: kept lists must come in pairs");
    const float r2 = r * r;
    radius_t_kernel<false><<<cb_div_up(n_x, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        x, x_batch, y, y_ptr, cutoff, r2, n_x, exclude_self, kept_rowptr, kept_col, nullptr, nullptr, nullptr, count);
    CB_CHECK_LAUNCH("cb_radius_count_t");
    return CB_OK;
}

extern "C" int cb_radius_fill_t(const float* x, const int32_t* x_batch, const float* y, const int32_t* y_ptr,
                                const float* cutoff, float r, int32_t n_x, int32_t exclude_self,
                                const int32_t* kept_rowptr, const int32_t* kept_col, const int32_t* rowptr_t,
                                int32_t* row_t, int32_t* col_t, void* stream) {
    CB_CHECK_ARG(n_x >= 0, "cb_radius_fill_t: bad size n_x=%d", n_x);
    if (n_x == 0) return CB_OK;
    CB_CHECK_ARG(x && x_batch && y && y_ptr && rowptr_t && row_t && col_t, "cb_radius_fill_t: null pointer");
    CB_CHECK_ARG((kept_rowptr == nullptr) == (kept_col == nullptr), "cb_radius_fill_t: kept lists must come in pairs");
    const float r2 = r * r;
    radius_t_kernel<true><<<cb_div_up(n_x, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        x, x_batch, y, y_ptr, cutoff, r2, n_x, exclude_self, kept_rowptr, kept_col, rowptr_t, row_t, col_t, nullptr);
    CB_CHECK_LAUNCH("cb_radius_fill_t");
    return CB_OK;
}

extern "C" int cb_exclusive_scan_i32(const int32_t* in, int32_t* out, int32_t n, int32_t* scratch, void* stream) {
    CB_CHECK_ARG(n >= 0 && n <= (1 << 24), "cb_exclusive_scan_i32: n=%d out of range", n);
    CB_CHECK_ARG(out && scratch && (in || n == 0), "cb_exclusive_scan_i32: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(int32_t), st);
        return CB_OK;
    }
    const int tiles = cb_div_up(n, kScanTile);
    scan_tiles_kernel<<<tiles, kScanThreads, 0, st>>>(in, out, n, scratch);
    scan_sums_kernel<<<1, kScanThreads, 0, st>>>(scratch, tiles);
    scan_add_kernel<<<tiles, kScanThreads, 0, st>>>(out, n, scratch);
    CB_CHECK_LAUNCH("cb_exclusive_scan_i32");
    return CB_OK;
}
