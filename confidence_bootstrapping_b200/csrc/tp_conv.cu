// K3 -- fused tensor-product convolution layer (the kernel that matters).
// Replaces TensorProductConvLayer.forward (models/tensor_layers.py:195-217): gather, radial MLP,
// FasterTensorProduct / e3nn FullyConnectedTensorProduct, scatter-mean, BatchNorm(eval), residual.
//
// Formulation (exact up to fp re-association).  For aggregation node i and slot q (= edge group)
//     sum_{e in q, e->i} tp_e  =  T_q( A ),   A[r][j] = sum_e f_e[r] * h~_e[j]
//   f_e[r]  : the CG "intermediates" of the tensor product, sum of coef * x[col[e]][.] * sh_e[.]
//             (R rows; FasterTensorProduct's out_dict entries tensor_layers.py:72-85, or one row
//             per (e3nn instruction, u, k))
//   h~_e    : [relu(W1 a_e + b1) ; 1]  (H+1 columns), the hidden layer of the radial MLP (layers.py:8-15)
//   T_q(A)[o(r,m)] += sum_j W2[w(r)+m][j] A[r][j] + b2[w(r)+m] A[r][H]     (W2a = [W2 | b2 | 0 0 0] row-wise)
// so the [E, weight_numel] per-edge weight tensor of the reference (6.6 KB per edge) is never formed.
//
// Two kernels:
//  (a) tp_accumulate_kernel  one CTA per (node, slot): R x H rank-1 updates per edge held in registers
//      (NRT x NCT threads, each a TR x TC tile), the finished tile written once to the workspace
//      (R x (H+4) floats per (node, slot); column H carries sum_e f_e for the bias term).
//  (b) tp_transform_kernel   one CTA per 32 nodes: for every slot and every run of rows it streams the
//      W2 rows once per 32 nodes through shared memory (the v1 single-kernel design re-read all of W2
//      per node and serialised the contraction on the 3 warps owning the 0e rows: profiles/r1),
//      then mean over all incoming edges, BatchNorm(eval) affine, residual.
// Everything is deterministic: fixed-order register / shared-memory reductions, no atomics.
#include "common.cuh"
#include "../../include/cb200.h"

namespace {

constexpr int CH = 32;   // edges staged per chunk in the accumulate kernel
constexpr int PADC = 4;  // extra workspace columns per row: [H] = sum_e f_e, [H+1..H+3] = 0
// Workspace layout.  The aggregation nodes are cut into tiles of WS_TILE consecutive nodes (counted from
// node_begin; one transform CTA per tile).  For every (slot, tile) the accumulators of the tile's active nodes
// (>= 1 edge in the slot) are stored row-major over (row r, rank of the node among the tile's active nodes):
//     A[(tile_off[q] + tile) * R * WS_TILE * HA  +  (r * n_act + rank) * HA  +  j]
// so the transform kernel fetches a whole row group of a tile with ONE contiguous bulk copy.
constexpr int WS_TILE = 32;

struct SlotTable {
    int n_slots;
    int first_seg[CB_MAX_SEGS], n_segs[CB_MAX_SEGS];
    int lo[CB_MAX_SEGS], hi[CB_MAX_SEGS];   // node range of the slot clipped to [node_begin, node_end)
    int item_off[CB_MAX_SEGS + 1];
    int tile0[CB_MAX_SEGS];                 // first node tile of the slot (tiles of WS_TILE nodes counted from node_begin)
    int tile_off[CB_MAX_SEGS + 1];          // workspace tiles before the slot
};

__host__ __device__ inline void build_slots(const cb_tp_conv_args& a, SlotTable& t) {
    t.n_slots = 0;
    t.item_off[0] = 0;
    t.tile_off[0] = 0;
    for (int s = 0; s < a.n_segs; ++s) {
        if (s > 0 && a.segs[s].slot == a.segs[s - 1].slot) {
            t.n_segs[t.n_slots - 1]++;
            continue;
        }
        const int q = t.n_slots++;
        t.first_seg[q] = s;
        t.n_segs[q] = 1;
        const int lo = a.segs[s].n0 > a.node_begin ? a.segs[s].n0 : a.node_begin;
        const int hi = a.segs[s].n1 < a.node_end ? a.segs[s].n1 : a.node_end;
        t.lo[q] = lo;
        t.hi[q] = hi > lo ? hi : lo;
        t.item_off[q + 1] = t.item_off[q] + (t.hi[q] - t.lo[q]);
        t.tile0[q] = (t.lo[q] - a.node_begin) / WS_TILE;
        t.tile_off[q + 1] = t.tile_off[q] + (t.hi[q] > t.lo[q] ? (t.hi[q] - a.node_begin + WS_TILE - 1) / WS_TILE - t.tile0[q] : 0);
    }
}

// edge range of `node` in segment sg; empty when the segment's gate says the node's output is never read
__device__ __forceinline__ void seg_edges(const cb_tp_segment& sg, int node, int& e0, int& e1) {
    const int i = node - sg.n0;
    e0 = __ldg(sg.rowptr + i);
    e1 = __ldg(sg.rowptr + i + 1);
    if (sg.gate_rowptr != nullptr && __ldg(sg.gate_rowptr + i + 1) <= __ldg(sg.gate_rowptr + i)) e1 = e0;
    if (sg.gate_mask != nullptr && __ldg(sg.gate_mask + i) == 0) e1 = e0;
}
__device__ __forceinline__ int seg_degree(const cb_tp_segment& sg, int node) {
    int e0, e1;
    seg_edges(sg, node, e0, e1);
    return e1 - e0;
}

// Where the accumulator of (slot q, node) lives: called by all 32 lanes of a warp (lane = node of the tile).
// Returns the float offset of element (row 0, column 0) and the row stride in floats.
__device__ __forceinline__ size_t ws_place(const cb_tp_conv_args& a, const SlotTable& st, int q, int node, int n_rows, int HA, int lane,
                                           int& row_stride) {
    const int tile = (node - a.node_begin) / WS_TILE;
    const int other = a.node_begin + tile * WS_TILE + lane;
    int deg = 0;
    if (other >= st.lo[q] && other < st.hi[q]) {
#pragma unroll 1
        for (int s = st.first_seg[q]; s < st.first_seg[q] + st.n_segs[q]; ++s) {
            const cb_tp_segment& sg = a.segs[s];
            deg += seg_degree(sg, other);
        }
    }
    const unsigned act = __ballot_sync(0xffffffffu, deg > 0);
    const int rank = __popc(act & ((1u << (node - a.node_begin - tile * WS_TILE)) - 1u));
    row_stride = __popc(act) * HA;
    return ((size_t)(st.tile_off[q] + tile - st.tile0[q]) * n_rows * WS_TILE + rank) * HA;
}

// ------------------------------------------------------------------------------------------ (a)
template <int NRT_, int NCT_, int TR_, int TC_>
struct Cfg {
    static constexpr int NRT = NRT_, NCT = NCT_, TR = TR_, TC = TC_;
    static constexpr int THREADS = NRT * NCT, RP = NRT * TR, HP = NCT * TC;
};

struct SmemLayout {
    int rows, terms, w1e, hbase, cols, xs, shs, es, F, Hs, total;  // offsets in 4-byte words
    int nep, dxp;
};

__host__ __device__ inline SmemLayout make_layout(int n_rows, int n_terms, int ne, int d_in, int S, int RP, int HP) {
    SmemLayout L;
    auto al4 = [](int v) { return (v + 3) & ~3; };
    int o = 0;
    L.nep = ne + 4;
    L.dxp = d_in | 1;
    L.rows = o;   o += al4(n_rows * 8);
    L.terms = o;  o += al4(n_terms * 2);
    L.w1e = o;    o += al4(HP * L.nep);
    L.hbase = o;  o += al4(HP);
    L.cols = o;   o += al4(CH);
    L.xs = o;     o += al4(CH * L.dxp);
    L.shs = o;    o += al4(CH * S);
    L.es = o;     o += al4(CH * ne);
    L.F = o;      o += al4(CH * RP);
    L.Hs = o;     o += al4(CH * HP);
    L.total = o;
    return L;
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
tp_accumulate_kernel(const __grid_constant__ cb_tp_conv_args a, int n_items) {
    constexpr int NCT = C::NCT, TR = C::TR, TC = C::TC, RP = C::RP, HP = C::HP, THREADS = C::THREADS;
    static_assert(TR == 4 && TC % 4 == 0, "tile shape");
    extern __shared__ __align__(16) float sm[];
    const SmemLayout L = make_layout(a.n_rows, a.n_terms, a.ne, a.d_in, a.S, RP, HP);
    cb_tp_row* rows_s = reinterpret_cast<cb_tp_row*>(sm + L.rows);
    cb_tp_term* terms_s = reinterpret_cast<cb_tp_term*>(sm + L.terms);
    float* W1e_s = sm + L.w1e;
    float* hbase = sm + L.hbase;
    int* cols_s = reinterpret_cast<int*>(sm + L.cols);
    float* xs = sm + L.xs;
    float* shs = sm + L.shs;
    float* es = sm + L.es;
    float* F = sm + L.F;
    float* Hs = sm + L.Hs;
    __shared__ SlotTable st;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr = tid / NCT, tc = tid % NCT;
    const int H = a.H, ne = a.ne, S = a.S, d_in = a.d_in, n_rows = a.n_rows, nep = L.nep, dxp = L.dxp;
    const int HA = H + PADC;

    // once per (persistent) CTA: slot table, TP program, zeroed padding of the staging tiles
    if (tid == 0) build_slots(a, st);
    {
        const int* src = reinterpret_cast<const int*>(a.rows);
        int* dst = reinterpret_cast<int*>(rows_s);
        for (int i = tid; i < n_rows * 8; i += THREADS) dst[i] = src[i];
        const int2* tsrc = reinterpret_cast<const int2*>(a.terms);
        int2* tdst = reinterpret_cast<int2*>(terms_s);
        for (int i = tid; i < a.n_terms; i += THREADS) tdst[i] = tsrc[i];
        for (int i = tid; i < CH * RP; i += THREADS) F[i] = 0.0f;
        for (int i = tid; i < CH * HP; i += THREADS) Hs[i] = 0.0f;
    }
    __syncthreads();
    // f-rows: thread -> fixed row, strided over the chunk's edges
    const int my_row = tid % RP, my_par = tid / RP;
    constexpr int EP = THREADS / RP > 0 ? THREADS / RP : 1;
    constexpr int MAXT = 4;  // terms of the thread's row kept in registers (longer rows finish from smem)
    int tb = 0, te = 0, t_xi[MAXT], t_si[MAXT];
    float t_cf[MAXT];
    const bool f_thread = my_row < n_rows && my_par < EP;
    if (f_thread) {
        tb = rows_s[my_row].term_begin;
        te = rows_s[my_row].term_end;
    }
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const bool on = f_thread && tb + t < te;
        const cb_tp_term tm = on ? terms_s[tb + t] : cb_tp_term{0, 0, 0.0f};
        t_xi[t] = tm.x_idx;
        t_si[t] = tm.sh_idx;
        t_cf[t] = on ? tm.coef : 0.0f;   // inactive slots multiply x[0]*sh[0] by zero
    }
    int staged_slot = -1;
    int q = 0;  // items are visited in increasing order: the slot index only moves forward

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        // item -> (slot, node)
        while (q + 1 < st.n_slots && item >= st.item_off[q + 1]) ++q;
        const int node = st.lo[q] + (item - st.item_off[q]);
        const int seg0 = st.first_seg[q], nseg = st.n_segs[q];
        int deg = 0;
        for (int s = seg0; s < seg0 + nseg; ++s) {
            const cb_tp_segment& sg = a.segs[s];
            deg += seg_degree(sg, node);
        }
        if (deg == 0) continue;  // block-uniform; the transform kernel skips (node, slot) pairs without edges

        const cb_tp_segment& s0 = a.segs[seg0];
        __syncthreads();  // previous item's h-phase readers of W1e_s / hbase are done
        if (staged_slot != q) {
            for (int i = tid; i < HP * (ne / 4); i += THREADS) {
                const int qq = i / (ne / 4), c4 = i - qq * (ne / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qq < H) v = __ldg(reinterpret_cast<const float4*>(s0.W1e + (size_t)qq * s0.ldw1) + c4);
                *reinterpret_cast<float4*>(W1e_s + qq * nep + 4 * c4) = v;
            }
            staged_slot = q;
            __syncthreads();
        }
        const int graph = a.agg_graph ? a.agg_graph[node] : 0;
        for (int qq = tid; qq < HP; qq += THREADS) {
            float v = 0.0f;
            if (qq < H) {
                v = s0.b1[qq];
                if (s0.P_agg) v += s0.P_agg[(size_t)node * s0.ldp_agg + qq];
                if (s0.e_post) {
                    const float* ep = s0.e_post + (size_t)graph * ne;
                    for (int c = 0; c < ne; ++c) v = fmaf(W1e_s[qq * nep + c], ep[c], v);
                }
            }
            hbase[qq] = v;
        }

        float acc[TR][TC];
        float fsum[TR];
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            fsum[i] = 0.0f;
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = 0.0f;
        }

        for (int s = seg0; s < seg0 + nseg; ++s) {
            const cb_tp_segment& sg = a.segs[s];
            int e0, e1;
            seg_edges(sg, node, e0, e1);
            for (int base = e0; base < e1; base += CH) {
                const int n = min(CH, e1 - base);
                // ---- gather raw operands of the chunk
                if (tid < n) cols_s[tid] = sg.col[base + tid] + sg.col_off;
                __syncthreads();  // also orders hbase / the previous accumulate phase
                for (int e = warp; e < n; e += THREADS / 32) {
                    const float* xr = a.x + (size_t)cols_s[e] * d_in;
                    for (int k = lane; k < d_in; k += 32) xs[e * dxp + k] = __ldg(xr + k);
                    if (sg.P_nbr) {  // neighbour-side projection of the first Linear, staged where h will be written
                        const float* pr = sg.P_nbr + (size_t)cols_s[e] * sg.ldp_nbr;
                        for (int k = lane; k < H; k += 32) Hs[e * HP + k] = __ldg(pr + k);
                    }
                }
                for (int i = tid; i < n * S; i += THREADS) shs[i] = __ldg(sg.sh + (size_t)base * S + i);
                for (int i = tid; i < n * (ne / 4); i += THREADS)
                    reinterpret_cast<float4*>(es)[i] = __ldg(reinterpret_cast<const float4*>(sg.e_attr + (size_t)base * ne) + i);
                __syncthreads();
                // ---- f rows: CG products of the gathered node features with the edge harmonics
                if (f_thread) {
                    for (int e = my_par; e < n; e += EP) {
                        const float* xe = xs + e * dxp;
                        const float* se = shs + e * S;
                        float v = 0.0f;
#pragma unroll
                        for (int t = 0; t < MAXT; ++t) v = fmaf(t_cf[t] * xe[t_xi[t]], se[t_si[t]], v);
                        for (int t = tb + MAXT; t < te; ++t) {
                            const cb_tp_term tm = terms_s[t];
                            v = fmaf(tm.coef * xe[tm.x_idx], se[tm.sh_idx], v);
                        }
                        F[e * RP + my_row] = v;
                    }
                }
                // ---- hidden layer of the radial MLP
                {
                    const int EG = THREADS / H;  // edge groups processed concurrently
                    const int qq = tid % H, eg = tid / H;
                    if (eg < EG) {
                        const float4* w4 = reinterpret_cast<const float4*>(W1e_s + qq * nep);
                        for (int e = eg; e < n; e += EG) {
                            float v = hbase[qq];
                            if (sg.P_nbr) v += Hs[e * HP + qq];
                            const float4* e4 = reinterpret_cast<const float4*>(es + e * ne);
                            for (int c = 0; c < ne / 4; ++c) {
                                const float4 w = w4[c], x4 = e4[c];
                                v = fmaf(w.x, x4.x, v);
                                v = fmaf(w.y, x4.y, v);
                                v = fmaf(w.z, x4.z, v);
                                v = fmaf(w.w, x4.w, v);
                            }
                            Hs[e * HP + qq] = fmaxf(v, 0.0f);
                        }
                    }
                }
                __syncthreads();
                // ---- rank-1 updates of the register tile
#pragma unroll 4
                for (int e = 0; e < n; ++e) {
                    const float4 f4 = *reinterpret_cast<const float4*>(F + e * RP + tr * TR);
                    const float f[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
                    for (int j4 = 0; j4 < TC / 4; ++j4) {
                        const float4 h4 = *reinterpret_cast<const float4*>(Hs + e * HP + tc * TC + 4 * j4);
#pragma unroll
                        for (int i = 0; i < TR; ++i) {
                            acc[i][4 * j4 + 0] = fmaf(f[i], h4.x, acc[i][4 * j4 + 0]);
                            acc[i][4 * j4 + 1] = fmaf(f[i], h4.y, acc[i][4 * j4 + 1]);
                            acc[i][4 * j4 + 2] = fmaf(f[i], h4.z, acc[i][4 * j4 + 2]);
                            acc[i][4 * j4 + 3] = fmaf(f[i], h4.w, acc[i][4 * j4 + 3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < TR; ++i) fsum[i] += f[i];
                }
                // the barrier after the next chunk's cols load protects F / Hs / xs
            }
        }

        // ---- write the finished tile: A[item][r][0..H) and column H = sum_e f_e[r]
        int row_stride;
        float* Aout = a.workspace + ws_place(a, st, q, node, n_rows, HA, threadIdx.x & 31, row_stride);
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            const int r = tr * TR + i;
            if (r < n_rows) {
                float* dst = Aout + (size_t)r * row_stride + tc * TC;
#pragma unroll
                for (int j4 = 0; j4 < TC / 4; ++j4)
                    if (tc * TC + 4 * j4 < H)
                        *reinterpret_cast<float4*>(dst + 4 * j4) =
                            make_float4(acc[i][4 * j4], acc[i][4 * j4 + 1], acc[i][4 * j4 + 2], acc[i][4 * j4 + 3]);
                if (tc == 0) *reinterpret_cast<float4*>(Aout + (size_t)r * row_stride + H) = make_float4(fsum[i], 0.f, 0.f, 0.f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ (a')
// Tensor-core variant of the accumulate kernel: A[R x (H+1)] += F^T[R x K] . H~[K x (H+1)] per chunk of
// K = 16 edges as tcgen05.mma.kind::tf32 with the accumulator in TMEM.  fp32 accuracy is kept with the
// 3xTF32 split (x = hi + lo, hi = top 19 bits; D += Fhi.Hhi + Fhi.Hlo + Flo.Hhi, error ~2^-21).  The
// constant-1 column of H~ makes column H of the accumulator the sum of f (bias term of the second Linear).
// CUDA cores only produce the two operand tiles (CG products, hidden layer of the radial MLP) into shared
// memory in the no-swizzle K-major core-matrix layout; one thread issues the MMAs; the tile is read back
// once per (node, slot) with tcgen05.ld and written to the workspace.  256 threads, 2 CTAs per SM
// (256 TMEM columns each).
namespace tc {

constexpr int KC = 16;          // edges per chunk = 2 MMA k-steps of 8
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 256;
constexpr int LBO = 128;        // bytes between the two 16-byte K-chunks of a k-step (adjacent core matrices)
constexpr int SBO = (KC / 4) * 128;  // bytes between 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version(1)<<46
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}

// 3xTF32 operand split: hi = x rounded to the 10-bit TF32 mantissa, lo = (x - hi) rounded likewise (the tensor
// core ignores the low 13 bits, so rounding here keeps the error unbiased)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}

// cheaper split for the F tiles (values of both signs, so truncation does not bias the sums): hi = x truncated to
// TF32, lo = x - hi exactly (the tensor core truncates lo's low bits); error ~2^-21 relative per product
__device__ __forceinline__ void split_tf32_trunc(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t saddr, int sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity);
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_addr(smem_u32(bar), parity); }
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {   // addr = shared-space address of the barrier
#pragma unroll 1
    for (int spin = 0; spin < (1 << 26); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (ok) return;
    }
    __trap();  // never hang the GPU: a lost MMA completion aborts the kernel instead
}

struct Layout {
    int rows, terms, hbase, cols, xs, shs, es, ps, fhi, flo, hhi, hlo, w1hi, w1lo, ehi, elo, total;  // byte offsets
    int dxp, npad, mtiles, sbow;
};

__host__ __device__ inline Layout make_layout(int n_rows, int n_terms, int ne, int d_in, int S, int H, bool transposed = false) {
    Layout L;
    auto al = [](int v, int a) { return (v + a - 1) / a * a; };
    L.sbow = (ne / 4) * LBO;     // stride between 8-row groups of the [.., ne] operand tiles of the hidden-layer MMA
    L.dxp = d_in + (d_in & 1);   // even: rows are filled with 8-byte cp.async
    L.npad = transposed ? 128 : al(H + 1, 16);   // transposed: H~ is the M = 128 operand (rows >= H stay zero)
    L.mtiles = (n_rows + 127) / 128;
    int o = 0;
    const int f_groups = transposed ? (n_rows + 15) / 16 * 2 : L.mtiles * 16;   // 8-row groups of the F^T tile (transposed: N = rows padded to 16)
    L.fhi = o;   o += f_groups * SBO;
    L.flo = o;   o += f_groups * SBO;
    L.hhi = o;   o += (L.npad / 8) * SBO;
    L.hlo = o;   o += (L.npad / 8) * SBO;
    L.ehi = o;   o += (KC / 8) * L.sbow;
    L.elo = o;   o += (KC / 8) * L.sbow;
    L.w1hi = o;  o += ((H + 7) / 8) * L.sbow;     // the MMA reads 128 rows: rows >= H overlap what follows and only
    L.w1lo = o;  o += ((H + 7) / 8) * L.sbow;     // feed accumulator lanes that are never read back
    L.rows = o;  o += al(n_rows * 32, 16);
    L.terms = o; o += al(n_terms * 8, 16);
    L.hbase = o; o += al(H * 4, 16);
    L.cols = o;  o += al(2 * KC * 4, 16);          // raw-operand staging is double buffered
    L.xs = o;    o += al(2 * KC * L.dxp * 4, 16);
    L.shs = o;   o += al(2 * KC * S * 4, 16);
    L.es = o;    o += al(2 * KC * ne * 4, 16);
    L.ps = o;    o += al(2 * KC * H * 4, 16);
    L.total = o;
    return L;
}

// One chunk of <= KC edges of one (node, slot) item; every thread tracks the same iterator.
struct Chunk {
    int item, q, node, seg, base, n;
    bool valid, first, last;
};

__device__ __forceinline__ void cp_async_bytes8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_bytes4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_bytes16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}

constexpr int MAXT = 4;   // terms of a thread's f-row kept in registers
constexpr int ITEM_BLOCK = 1;   // consecutive (node, slot) items dealt to a persistent CTA at a time

struct FRowCtx {
    const float* xs; const float* shs; const cb_tp_term* terms_s; unsigned char* Fhi; unsigned char* Flo;
    int dxp, S, n, nq, tb0, te0, rbase;
    int e4_stride = LBO, e4_xor = 0;   // 16-byte chunk j of a row lands at rbase + (j ^ e4_xor) * e4_stride (swizzled tiles: stride 16, xor = row bits)
};

// F^T tile row of one thread: F[r][e] = sum_t coef_t * x[col e][xi_t] * sh_e[si_t] for the chunk's edges, hi/lo split,
// four consecutive edges per 16-byte core-matrix row.  NT = register-resident terms evaluated (warp-uniform).
template <int NT>
__device__ __forceinline__ float f_row(const FRowCtx& c, const int (&t_xi)[MAXT], const int (&t_si)[MAXT], const float (&t_cf)[MAXT]) {
    const int dx4 = 4 * c.dxp, s4 = 4 * c.S;
    float total = 0.0f;          // sum_e f_e[r] over the chunk (bias term of the second Linear in the transposed kernel)
    const float* xe = c.xs;      // first edge of the current 4-edge group
    const float* se = c.shs;
#pragma unroll 1
    for (int e4 = 0; e4 < c.nq; ++e4, xe += dx4, se += s4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // edges >= n read stale (possibly non-finite) staging rows: the select below discards them
            float acc = 0.0f;
#pragma unroll
            for (int t = 0; t < NT; ++t) acc = fmaf(t_cf[t] * xe[j * c.dxp + t_xi[t]], se[j * c.S + t_si[t]], acc);
            v[j] = (4 * e4 + j < c.n) ? acc : 0.0f;
        }
        if (c.te0 - c.tb0 > MAXT) {   // rare long rows: finish from the shared-memory term table
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = 4 * e4 + j;
                if (e < c.n)
#pragma unroll 1
                    for (int t = c.tb0 + MAXT; t < c.te0; ++t) {
                        const cb_tp_term tm = c.terms_s[t];
                        v[j] = fmaf(tm.coef * c.xs[e * c.dxp + tm.x_idx], c.shs[e * c.S + tm.sh_idx], v[j]);
                    }
            }
        }
        float4 hi, lo;
        split_tf32_trunc(v[0], hi.x, lo.x);
        split_tf32_trunc(v[1], hi.y, lo.y);
        split_tf32_trunc(v[2], hi.z, lo.z);
        split_tf32_trunc(v[3], hi.w, lo.w);
        const int coff = c.rbase + (e4 ^ c.e4_xor) * c.e4_stride;
        *reinterpret_cast<float4*>(c.Fhi + coff) = hi;   // 4 consecutive edges = one 16-byte chunk of the operand row
        *reinterpret_cast<float4*>(c.Flo + coff) = lo;
        total += (v[0] + v[1]) + (v[2] + v[3]);
    }
    return total;
}

#ifdef CB_PHASE_TIMING
__device__ unsigned long long cb_dbg_phase[8][12];
#define PH_MARK(k) do { if (ph_on) { const long long t_ = clock64(); ph[k] += t_ - ph_t; ph_t = t_; } } while (0)
#else
#define PH_MARK(k) do { } while (0)
#endif

// TR = false: D[f-row][h~ column] in MT x NP TMEM columns, read back through a shared-memory transpose (any layer shape).
// TR = true : the accumulator is kept TRANSPOSED, D[h~ unit (128 lanes)][f-row (<= 240 columns)] = H~ . F^T -- one MMA per
//             k-step and product instead of one per 128-row tile, 23 % fewer operand bytes read from shared memory, and a
//             tcgen05.ld now returns 32 f-rows of ONE hidden unit per lane, so 32 lanes hold 128 contiguous workspace bytes
//             of a row: the finished tile goes TMEM -> registers -> global with no shared-memory transpose (which was 19 %
//             of the kernel's shared-memory traffic and 18 % of its stall samples, profiles/r1/k3_ncu_summary.txt).  The
//             constant-1 row of H~ is dropped; column H of the workspace (sum_e f_e) is summed by the f-row threads.
//             Needs n_rows <= 240 (TMEM: 240 + 16 columns) -- every layer of the score model.
template <bool TR>
__global__ void __launch_bounds__(THREADS, 2)
tp_accumulate_tc_kernel(const __grid_constant__ cb_tp_conv_args a, int n_items) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int H = a.H, ne = a.ne, S = a.S, d_in = a.d_in, n_rows = a.n_rows;
    const Layout L = make_layout(n_rows, a.n_terms, ne, d_in, S, H, TR);
    const int dxp = L.dxp, NP = L.npad, MT = L.mtiles, HA = H + PADC, SBOW = L.sbow;
    unsigned char* Fhi = smraw + L.fhi;
    unsigned char* Flo = smraw + L.flo;
    unsigned char* Hhi = smraw + L.hhi;
    unsigned char* Hlo = smraw + L.hlo;
    cb_tp_row* rows_s = reinterpret_cast<cb_tp_row*>(smraw + L.rows);
    cb_tp_term* terms_s = reinterpret_cast<cb_tp_term*>(smraw + L.terms);
    unsigned char* W1hi = smraw + L.w1hi;
    unsigned char* W1lo = smraw + L.w1lo;
    unsigned char* Ehi = smraw + L.ehi;
    unsigned char* Elo = smraw + L.elo;
    float* hbase = reinterpret_cast<float*>(smraw + L.hbase);
    __shared__ SlotTable st;
    __shared__ __align__(8) uint64_t mma_bar, h_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // double-buffered staging of the raw per-edge operands (filled by cp.async one chunk ahead)
    auto cols_of = [&](int b) { return reinterpret_cast<int*>(smraw + L.cols) + b * KC; };
    auto xs_of = [&](int b) { return reinterpret_cast<float*>(smraw + L.xs) + b * KC * dxp; };
    auto shs_of = [&](int b) { return reinterpret_cast<float*>(smraw + L.shs) + b * KC * S; };
    auto es_of = [&](int b) { return reinterpret_cast<float*>(smraw + L.es) + b * KC * ne; };
    auto ps_of = [&](int b) { return reinterpret_cast<float*>(smraw + L.ps) + b * KC * H; };

    // ---- once per persistent CTA
    if (tid == 0) {
        build_slots(a, st);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&h_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    {
        const int* src = reinterpret_cast<const int*>(a.rows);
        int* dst = reinterpret_cast<int*>(rows_s);
#pragma unroll 1
        for (int i = tid; i < n_rows * 8; i += THREADS) dst[i] = src[i];
        const int2* tsrc = reinterpret_cast<const int2*>(a.terms);
        int2* tdst = reinterpret_cast<int2*>(terms_s);
#pragma unroll 1
        for (int i = tid; i < a.n_terms; i += THREADS) tdst[i] = tsrc[i];
        // operand tiles start as zeros: padding rows (r >= n_rows, q > H) are never written again
#pragma unroll 1
        for (int i = tid; i < (L.rows - L.fhi) / 16; i += THREADS) reinterpret_cast<float4*>(smraw + L.fhi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int NRP = (n_rows + 15) & ~15;       // transposed: accumulator columns (f-rows padded to the MMA's N granularity)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((TR ? NRP : NP) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_h = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t tmem_h = tmem_base + (uint32_t)(TR ? NRP : MT * NP);   // [128 hidden lanes] x [KC edge columns] pre-activations
    float fsum = 0.0f;                         // transposed: sum_e f_e[tid] of the open item
    uint32_t commits = 0, waited = 0;   // block-uniform bookkeeping of the MMA barrier phases
    uint32_t h_phase = 0;
    int staged_slot = -1;
    // terms of the thread's (first) f-row live in registers; rows beyond THREADS use the generic loop
    int tb0 = 0, te0 = 0, t_xi[MAXT], t_si[MAXT];
    float t_cf[MAXT];
    if (tid < n_rows) {
        tb0 = rows_s[tid].term_begin;
        te0 = rows_s[tid].term_end;
    }
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const bool on = tid < n_rows && tb0 + t < te0;
        const cb_tp_term tm = on ? terms_s[tb0 + t] : cb_tp_term{0, 0, 0.0f};
        t_xi[t] = tm.x_idx;
        t_si[t] = tm.sh_idx;
        t_cf[t] = on ? tm.coef : 0.0f;
    }
    // number of register-resident terms the warp's rows need (warp-uniform)
    const int nt_warp = __reduce_max_sync(0xffffffffu, min(te0 - tb0, MAXT));
    const uint32_t fhi_a = smem_u32(Fhi), flo_a = smem_u32(Flo), hhi_a = smem_u32(Hhi), hlo_a = smem_u32(Hlo);
    const uint64_t d_fhi = make_desc(fhi_a), d_flo = make_desc(flo_a), d_hhi = make_desc(hhi_a), d_hlo = make_desc(hlo_a);
    const uint64_t d_w1hi = make_desc_sbo(smem_u32(W1hi), SBOW), d_w1lo = make_desc_sbo(smem_u32(W1lo), SBOW);
    const uint64_t d_ehi = make_desc_sbo(smem_u32(Ehi), SBOW), d_elo = make_desc_sbo(smem_u32(Elo), SBOW);
    static_assert(KC == 16, "the MMA issue code is written for two k-steps per chunk");

    // ---- chunk iterator (block-uniform): items of this CTA in increasing order, their segments, KC edges at a time
    auto seg_range = [&](int seg, int node, int& e0, int& e1) {
        const cb_tp_segment& sg = a.segs[seg];
        seg_edges(sg, node, e0, e1);
    };
    // items are dealt to the persistent CTAs in blocks of ITEM_BLOCK consecutive (node, slot) pairs: neighbouring nodes
    // share their graph (cached per-graph bias), their rowptr / projection cache lines and their workspace tile
    auto next_item = [&](int item) { return (item % ITEM_BLOCK) < ITEM_BLOCK - 1 ? item + 1 : item + 1 + (int)(gridDim.x - 1) * ITEM_BLOCK; };
    auto open_item = [&](Chunk& c, int item, int q) {
        // first non-empty segment of the first non-empty item at or after `item`
        c.valid = false;
#pragma unroll 1
        for (; item < n_items; item = next_item(item)) {
            while (q + 1 < st.n_slots && item >= st.item_off[q + 1]) ++q;
            const int node = st.lo[q] + (item - st.item_off[q]);
#pragma unroll 1
            for (int seg = st.first_seg[q]; seg < st.first_seg[q] + st.n_segs[q]; ++seg) {
                int e0, e1;
                seg_range(seg, node, e0, e1);
                if (e1 > e0) {
                    c.item = item; c.q = q; c.node = node; c.seg = seg; c.base = e0; c.n = min(KC, e1 - e0);
                    c.valid = true; c.first = true;
                    return e1;
                }
            }
        }
        return 0;
    };
    auto has_more = [&](const Chunk& c, int e1) {   // more edges of the same item after this chunk?
        if (c.base + KC < e1) return true;
#pragma unroll 1
        for (int seg = c.seg + 1; seg < st.first_seg[c.q] + st.n_segs[c.q]; ++seg) {
            int e0, e1b;
            seg_range(seg, c.node, e0, e1b);
            if (e1b > e0) return true;
        }
        return false;
    };
    auto advance = [&](const Chunk& c, int& e1) {   // returns the chunk after c; e1 = end of its segment
        Chunk n = c;
        n.first = false;
        if (c.base + KC < e1) {
            n.base = c.base + KC;
            n.n = min(KC, e1 - n.base);
            return n;
        }
#pragma unroll 1
        for (int seg = c.seg + 1; seg < st.first_seg[c.q] + st.n_segs[c.q]; ++seg) {
            int e0, e1b;
            seg_range(seg, c.node, e0, e1b);
            if (e1b > e0) {
                n.seg = seg; n.base = e0; n.n = min(KC, e1b - e0);
                e1 = e1b;
                return n;
            }
        }
        e1 = open_item(n, next_item(c.item), c.q);
        return n;
    };
    auto load_cols = [&](const Chunk& c, int b) {
        const cb_tp_segment& sg = a.segs[c.seg];
        if (tid < c.n) cols_of(b)[tid] = __ldg(sg.col + c.base + tid) + sg.col_off;
    };
    auto issue_gather = [&](const Chunk& c, int b) {   // cp.async: lands while the previous chunk is being multiplied
        const cb_tp_segment& sg = a.segs[c.seg];
        const int* cols = cols_of(b);
        float* xs = xs_of(b);
        float* Ps = ps_of(b);
#pragma unroll 1
        for (int e = warp; e < c.n; e += THREADS / 32) {
            const float* xr = a.x + (size_t)cols[e] * d_in;
#pragma unroll 1
            for (int k = lane; k < d_in / 2; k += 32) cp_async_bytes8(xs + e * dxp + 2 * k, xr + 2 * k);
            if (sg.P_nbr) {
                const float* pr = sg.P_nbr + (size_t)cols[e] * sg.ldp_nbr;
#pragma unroll 1
                for (int k = lane; k < H / 4; k += 32) cp_async_bytes16(Ps + e * H + 4 * k, pr + 4 * k);
            }
        }
#pragma unroll 1
        for (int i = tid; i < c.n * S; i += THREADS) cp_async_bytes4(shs_of(b) + i, sg.sh + (size_t)c.base * S + i);
#pragma unroll 1
        for (int i = tid; i < c.n * (ne / 4); i += THREADS) cp_async_bytes16(es_of(b) + 4 * i, sg.e_attr + (size_t)c.base * ne + 4 * i);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    Chunk cur;
    int cur_e1 = open_item(cur, blockIdx.x * ITEM_BLOCK, 0);
    int buf = 0;
    // per-item constants of the hidden layer, prefetched one chunk ahead / cached in registers (threads tid < H):
    //   hbase[q] = b1[q] + W1e[q] . e_post[graph]   (changes with the slot or the graph)   +  P_agg[node][q]
    auto load_pagg = [&](const Chunk& c) {
        const cb_tp_segment& s0 = a.segs[st.first_seg[c.q]];
        return (tid < H && s0.P_agg) ? __ldg(s0.P_agg + (size_t)c.node * s0.ldp_agg + tid) : 0.0f;
    };
    auto load_graph = [&](const Chunk& c) { return a.agg_graph ? __ldg(a.agg_graph + c.node) : 0; };
    float pagg_cur = 0.0f, pagg_nxt = 0.0f, hb_const = 0.0f;
    int graph_cur = 0, graph_nxt = 0, hb_q = -1, hb_graph = -1;
    if (cur.valid) {
        pagg_cur = load_pagg(cur);
        graph_cur = load_graph(cur);
    }
    size_t ws_off = 0;
    int ws_stride = 0;
    if (cur.valid) {
        load_cols(cur, 0);
        __syncthreads();
        issue_gather(cur, 0);
    }
#ifdef CB_PHASE_TIMING
    const bool ph_on = (tid == 0 || tid == 224);
    long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ph_t = clock64();
#endif
#pragma unroll 1
    while (cur.valid) {
        PH_MARK(9);
        const cb_tp_segment& sg = a.segs[cur.seg];
        const int n = cur.n;
        cur.last = !has_more(cur, cur_e1);
        int nxt_e1 = cur_e1;
        Chunk nxt = advance(cur, nxt_e1);
        if (nxt.valid) load_cols(nxt, buf ^ 1);
        if (nxt.valid && nxt.first) {       // the next item's per-node constants arrive while this chunk is processed
            pagg_nxt = load_pagg(nxt);
            graph_nxt = load_graph(nxt);
        }
        if (cur.first) {
            // ---- per-item setup: edge-embedding slice of the first Linear (on slot change), constant part of h
            const cb_tp_segment& s0 = a.segs[st.first_seg[cur.q]];
            if (staged_slot != cur.q) {
#pragma unroll 1
                for (int i = tid; i < H * (ne / 4); i += THREADS) {   // hi/lo operand tiles of W1e: row = hidden unit, K = edge channel
                    const int qq = i / (ne / 4), c4 = i - qq * (ne / 4);
                    const float4 w = __ldg(reinterpret_cast<const float4*>(s0.W1e + (size_t)qq * s0.ldw1) + c4);
                    float4 hi, lo;
                    split_tf32(w.x, hi.x, lo.x);
                    split_tf32(w.y, hi.y, lo.y);
                    split_tf32(w.z, hi.z, lo.z);
                    split_tf32(w.w, hi.w, lo.w);
                    const int off = (qq >> 3) * SBOW + c4 * LBO + (qq & 7) * 16;
                    *reinterpret_cast<float4*>(W1hi + off) = hi;
                    *reinterpret_cast<float4*>(W1lo + off) = lo;
                }
                staged_slot = cur.q;
            }
            if (tid < H) {
                if (hb_q != cur.q || (s0.e_post && hb_graph != graph_cur)) {
                    float v = __ldg(s0.b1 + tid);
                    if (s0.e_post) {
                        const float* ep = s0.e_post + (size_t)graph_cur * ne;
#pragma unroll 4
                        for (int c = 0; c < ne; ++c) v = fmaf(__ldg(s0.W1e + (size_t)tid * s0.ldw1 + c), __ldg(ep + c), v);
                    }
                    hb_const = v;
                    hb_q = cur.q;
                    hb_graph = graph_cur;
                }
                hbase[tid] = hb_const + pagg_cur;
            }
            // where the finished tile goes (rowptr lookups of the node's workspace tile): needed only by the epilogue
            ws_off = ws_place(a, st, cur.q, cur.node, n_rows, HA, lane, ws_stride);
        }
        PH_MARK(0);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        PH_MARK(10);
        const float* xs = xs_of(buf);
        const float* shs = shs_of(buf);
        const float* es = es_of(buf);
        const float* Ps = ps_of(buf);
        const int ksteps = (n + 7) >> 3;
        // ---- hidden layer on the tensor core: pre[q][e] = sum_c W1e[q][c] * e_attr[e][c]  (M = 128 hidden lanes,
        // N = KC edges, K = ne, 3xTF32); the edge-embedding rows of the chunk become the hi/lo B-operand tiles.
        // Every thread splits exactly the 16-byte pieces it copied itself (same index map as issue_gather), so this
        // needs no barrier after the cp.async wait; the barrier below publishes the tiles to the MMA.
#pragma unroll 1
        for (int i = tid; i < KC * (ne / 4); i += THREADS) {
            const int e = i / (ne / 4), c4 = i - e * (ne / 4);
            const float4 v = *reinterpret_cast<const float4*>(es + e * ne + 4 * c4);   // rows >= n hold stale data: zeroed after the MMA
            float4 hi, lo;
            split_tf32(v.x, hi.x, lo.x);
            split_tf32(v.y, hi.y, lo.y);
            split_tf32(v.z, hi.z, lo.z);
            split_tf32(v.w, hi.w, lo.w);
            const int off = (e >> 3) * SBOW + c4 * LBO + (e & 7) * 16;
            *reinterpret_cast<float4*>(Ehi + off) = hi;
            *reinterpret_cast<float4*>(Elo + off) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                   // gather(cur), cols(nxt), hbase, E / W1e tiles visible to everyone
        PH_MARK(1);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors differ only in the start-address field: advance it by one k-step (2 core matrices) per iteration
            uint64_t dwh = d_w1hi, dwl = d_w1lo, deh = d_ehi, del = d_elo;
#pragma unroll 2
            for (int ks = 0; ks < ne / 8; ++ks) {
                mma_tf32(tmem_h, dwh, deh, idesc_h, ks > 0 ? 1u : 0u);
                mma_tf32(tmem_h, dwh, del, idesc_h, 1u);
                mma_tf32(tmem_h, dwl, deh, idesc_h, 1u);
                dwh += (2 * LBO) >> 4; dwl += (2 * LBO) >> 4; deh += (2 * LBO) >> 4; del += (2 * LBO) >> 4;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&h_bar)) : "memory");
        }
        PH_MARK(2);
        if (nxt.valid) issue_gather(nxt, buf ^ 1);         // overlaps everything below
        PH_MARK(3);
        // the F / H~ operand tiles are free once the previous chunk's MMAs completed; their latency has been overlapped with
        // the iterator, the gather issue and the hidden-layer MMA above (which only touch the E / W1e tiles and D_h)
        if (waited < commits) {
            mbar_wait(&mma_bar, waited & 1);
            ++waited;
        }
        // ---- F^T tile: row r = f-row, column = edge; hi/lo TF32 split.  Loops stay rolled on purpose:
        // the kernel must fit the 32 KB instruction cache (an unrolled build stalled on instruction fetch).
        const int nq = 2 * ((n + 7) >> 3);   // 4-edge groups covered by the MMA k-steps of this chunk
        if (tid < n_rows) {
            const FRowCtx fc{xs, shs, terms_s, Fhi, Flo, dxp, S, n, nq, tb0, te0, (tid >> 3) * SBO + (tid & 7) * 16};
            float part;
            switch (nt_warp) {      // warp-uniform: most f-rows have a single term
                case 1: part = f_row<1>(fc, t_xi, t_si, t_cf); break;
                case 2: part = f_row<2>(fc, t_xi, t_si, t_cf); break;
                case 3: part = f_row<3>(fc, t_xi, t_si, t_cf); break;
                default: part = f_row<MAXT>(fc, t_xi, t_si, t_cf); break;
            }
            fsum = cur.first ? part : fsum + part;
        }
#pragma unroll 1
        for (int r = tid + THREADS; r < n_rows; r += THREADS) {   // rows beyond the first 256 (lmax-2 layers)
            const int tb = rows_s[r].term_begin, te = rows_s[r].term_end;
            const int rbase = (r >> 3) * SBO + (r & 7) * 16;
#pragma unroll 1
            for (int e4 = 0; e4 < nq; ++e4) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = 4 * e4 + j;
                    float acc = 0.0f;
                    if (e < n)
#pragma unroll 1
                        for (int t = tb; t < te; ++t) {
                            const cb_tp_term tm = terms_s[t];
                            acc = fmaf(tm.coef * xs[e * dxp + tm.x_idx], shs[e * S + tm.sh_idx], acc);
                        }
                    v[j] = acc;
                }
                float4 hi, lo;
                split_tf32(v[0], hi.x, lo.x);
                split_tf32(v[1], hi.y, lo.y);
                split_tf32(v[2], hi.z, lo.z);
                split_tf32(v[3], hi.w, lo.w);
                *reinterpret_cast<float4*>(Fhi + rbase + e4 * LBO) = hi;
                *reinterpret_cast<float4*>(Flo + rbase + e4 * LBO) = lo;
            }
        }
        // ---- H~ tile: row q = hidden unit (row H = constant 1), column = edge.  The pre-activations come back from
        // TMEM one hidden unit per lane (warps 0-3: edges 0-7, warps 4-7: edges 8-15) -- exactly one row of the
        // K-major operand tile -- and get the node / neighbour projections, the ReLU and the hi/lo split.
        PH_MARK(4);
        mbar_wait(&h_bar, h_phase);
        h_phase ^= 1u;
        PH_MARK(5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const int qq = (warp & 3) * 32 + lane, eh = warp >> 2;
            if (eh < ksteps) {     // warp-uniform
                uint32_t v[8];
                const uint32_t taddr = tmem_h + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(8 * eh);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (TR ? qq < H : qq < NP) {      // transposed: rows >= H keep the zeros written at kernel start
                    float h[8];
                    if (qq < H) {
                        const float hb = hbase[qq];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float pre = __uint_as_float(v[j]) + hb;
                            if (sg.P_nbr) pre += Ps[(8 * eh + j) * H + qq];
                            h[j] = 8 * eh + j < n ? fmaxf(pre, 0.0f) : 0.0f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) h[j] = qq == H ? 1.0f : 0.0f;
                    }
                    const int rbase = (qq >> 3) * SBO + (qq & 7) * 16 + 2 * eh * LBO;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        float4 hi, lo;
                        split_tf32(h[4 * g + 0], hi.x, lo.x);
                        split_tf32(h[4 * g + 1], hi.y, lo.y);
                        split_tf32(h[4 * g + 2], hi.z, lo.z);
                        split_tf32(h[4 * g + 3], hi.w, lo.w);
                        *reinterpret_cast<float4*>(Hhi + rbase + g * LBO) = hi;
                        *reinterpret_cast<float4*>(Hlo + rbase + g * LBO) = lo;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
        PH_MARK(6);
        __syncthreads();
        PH_MARK(11);
        // ---- one thread issues the MMAs of this chunk and commits them to the barrier
        if (TR && tid == 0) {
            // D[h~ unit][f-row] += H~ . F^T : A = H~ tile (M = 128 rows), B = F^T tile (N = NRP rows), one MMA per product
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc0 = cur.first ? 0u : 1u;
            mma_tf32(tmem_base, d_hhi, d_fhi, idesc, acc0);
            mma_tf32(tmem_base, d_hhi, d_flo, idesc, 1u);
            mma_tf32(tmem_base, d_hlo, d_fhi, idesc, 1u);
            if (ksteps > 1) {
                const uint64_t ks = (uint64_t)((2 * LBO) >> 4);
                mma_tf32(tmem_base, d_hhi + ks, d_fhi + ks, idesc, 1u);
                mma_tf32(tmem_base, d_hhi + ks, d_flo + ks, idesc, 1u);
                mma_tf32(tmem_base, d_hlo + ks, d_fhi + ks, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar))
                         : "memory");
        }
        if (!TR && tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t d = tmem_base + (uint32_t)(mt * NP);
                const uint32_t acc0 = cur.first ? 0u : 1u;
                uint64_t dfh = d_fhi + (uint64_t)((mt * 16 * SBO) >> 4), dfl = d_flo + (uint64_t)((mt * 16 * SBO) >> 4);
                mma_tf32(d, dfh, d_hhi, idesc, acc0);
                mma_tf32(d, dfh, d_hlo, idesc, 1u);
                mma_tf32(d, dfl, d_hhi, idesc, 1u);
                if (ksteps > 1) {
                    dfh += (2 * LBO) >> 4; dfl += (2 * LBO) >> 4;
                    mma_tf32(d, dfh, d_hhi + ((2 * LBO) >> 4), idesc, 1u);
                    mma_tf32(d, dfh, d_hlo + ((2 * LBO) >> 4), idesc, 1u);
                    mma_tf32(d, dfl, d_hhi + ((2 * LBO) >> 4), idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar))
                         : "memory");
        }
        ++commits;
        PH_MARK(7);
        if (cur.last) {
            // ---- epilogue: wait for the last MMAs, read the tile back from TMEM, write it to the workspace
            while (waited < commits) {
                mbar_wait(&mma_bar, waited & 1);
                ++waited;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row_stride = ws_stride;
            float* Aout = a.workspace + ws_off;
            const int lg = warp & 3;                 // TMEM lane group this warp may access
            if constexpr (TR) {
                // lane = hidden unit j = 32 lg + lane, columns = f-rows: a warp-wide store of one column is 128 contiguous
                // bytes of workspace row r.  The two warps of a lane group split the columns.
                const int j = lg * 32 + lane;
                if (lg * 32 < H) {                   // warp-uniform: lane groups beyond the hidden width hold nothing
                    const int c_half = ((NRP / 2) + 15) & ~15;
                    const int c_lo = (warp >> 2) ? c_half : 0, c_hi = (warp >> 2) ? NRP : c_half;
                    float* dst = Aout + j;
#pragma unroll 1
                    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                        uint32_t v[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                            : "r"(taddr));
                        const bool second = c0 + 16 < c_hi;
                        if (second)
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                                : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                : "r"(taddr + 16u));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (j < H) {
                            const int nr = min(second ? 32 : 16, n_rows - c0);      // f-rows of this batch that exist
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < nr) dst[(size_t)(c0 + i) * row_stride] = __uint_as_float(v[i]);
                        }
                    }
                }
                // column H: sum_e f_e[r] (bias of the second Linear), then the three zero pad columns the transform multiplies by 0
                if (tid < n_rows) *reinterpret_cast<float4*>(Aout + (size_t)tid * row_stride + H) = make_float4(fsum, 0.f, 0.f, 0.f);
            } else {
            constexpr int STG_LD = 36;               // floats per staged row: 144-byte stride keeps float4 accesses conflict-free
            float* stg = reinterpret_cast<float*>(smraw + L.fhi) + warp * 32 * STG_LD;
#pragma unroll 1
            for (int mt = warp >> 2; mt < MT; mt += THREADS / 128) {
#pragma unroll 1
                for (int c0 = 0; c0 < NP; c0 += 32) {   // NP is a multiple of 16; a trailing half chunk reads 16 spare columns
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mt * NP + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr));
                    if (c0 + 16 < NP)
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                            : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                            : "r"(taddr + 16u));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // transpose the warp's 32 rows x 32 columns through shared memory (the operand tiles are idle
                    // here) so that 8 lanes write 128 contiguous bytes of one row instead of 32 lanes hitting 32 rows
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(stg + lane * STG_LD + j) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    __syncwarp();
                    const int sub = lane >> 3, ch = (lane & 7) * 4;
                    if (c0 + ch < HA) {
                        const int r0 = mt * 128 + lg * 32 + sub;       // this lane's rows: r0, r0 + 4, ...
                        float* dst = Aout + (size_t)r0 * row_stride + c0 + ch;
                        const float* src = stg + sub * STG_LD + ch;
                        const int nr = min(8, (n_rows - r0 + 3) >> 2);  // rows of the lane inside the tile
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (i < nr) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src + i * 4 * STG_LD);
                            dst += 4 * (size_t)row_stride;
                        }
                    }
                    __syncwarp();
                }
            }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();   // TMEM tile fully read before the next item's first MMA overwrites it
        }
        PH_MARK(8);
        if (nxt.valid && nxt.first) {
            pagg_cur = pagg_nxt;
            graph_cur = graph_nxt;
        }
        cur = nxt;
        cur_e1 = nxt_e1;
        buf ^= 1;
    }
#ifdef CB_PHASE_TIMING
    if (ph_on)
        for (int k = 0; k < 12; ++k) atomicAdd(&cb_dbg_phase[tid == 0 ? 0 : 1][k], (unsigned long long)ph[k]);
#endif
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

#include "tp_accumulate_ws.cuh"

}  // namespace tc

// ------------------------------------------------------------------------------------------ (b)
// Transform kernel.  One CTA per 32 aggregation nodes.  For every slot with edges into the tile and every run
// (rows with the same output block and multiplicity `mul`) it computes
//     out[n][out_base + m*out_step] += sum_r sum_j A[n][r][j] * W2a[w(r)+m][j]      (column H of W2a = b2)
// Warp specialised: one producer warp streams the A rows (one 1-D bulk copy per node and row) and the
// contiguous W2a rows of a row group into a 4-stage shared-memory ring with cp.async.bulk + mbarrier
// complete_tx; 14 consumer warps wait on the stage's "full" barrier, compute, and release it through its
// "empty" barrier -- no CTA-wide barrier per row group.
// Register tiling of the consumers: a thread owns 4 nodes x MT outputs x one 4-column chunk of K (8 node lanes
// x 28 K-chunks x 2 combos; a combo is the second row of a 2-row group for mul <= 8, or the upper half of the
// outputs for mul > 8) and keeps its partial sums in registers over all rows of the run: per row 4 LDS.128
// of A (node lanes interleaved => conflict-free), MT broadcast LDS.128 of W2a and 16*MT FMAs.  At the end of a
// run the partials are reduced in a fixed order (shuffles over the 4 K-chunks of a warp, then shared memory
// over warps).  Epilogue: mean over all incoming edges, BatchNorm(eval) affine, residual.
namespace tf {

constexpr int NB = WS_TILE;     // nodes per CTA = one workspace tile
static_assert(NB == 32, "the producer warp and the rank ballots map one lane to one node of the tile");
constexpr int NT = 4;           // nodes per thread: node lane nl owns nodes nl, nl+8, nl+16, nl+24
constexpr int WPC = 7;          // warps per combo (each warp: 8 node lanes x 4 K-chunks)
constexpr int KSTRIDE = 4 * WPC;
constexpr int CWARPS = 2 * WPC; // consumer warps
constexpr int CT = 32 * CWARPS; // 448 consumer threads
constexpr int TT = CT + 32;     // + the producer warp
constexpr int STAGES = 4;
constexpr int MAX_RS = 2;
constexpr int MAX_WROWS = 32;
constexpr int MTMAX = 16;
constexpr int MAX_RUNS = 96;

using tc::mbar_wait;
using tc::smem_u32;

struct Layout {
    int a_stage, w_stage, stage, zero, part, outacc, items, owner, total;   // floats
};
__host__ __device__ inline Layout make_layout(int H, int d_out, int n_slots) {
    Layout L;
    const int HA = H + PADC;
    L.a_stage = MAX_RS * NB * HA;
    L.w_stage = MAX_WROWS * HA;
    L.stage = L.a_stage + L.w_stage;
    int o = STAGES * L.stage;
    L.zero = o;   o += HA;
    L.part = o;   o += CWARPS * 8 * NT * MTMAX;
    L.outacc = o; o += (NB * d_out + 3) & ~3;
    L.items = o;  o += n_slots * NB;
    L.owner = o;  o += d_out;
    L.total = o;
    return L;
}

struct RunS { int row_begin, row_end, mul, w_base0, out_base, out_step, mt, rs; };

__device__ __forceinline__ void bulk_g2s(float* smem_dst, const float* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// one row of one combo: acc[j][m] += A[node nl+8j][4kc..4kc+3] . W[m][4kc..4kc+3]
template <int MT>
__device__ __forceinline__ void row_fma(const float* const (&ap)[NT], const float* __restrict__ Wb, int HA, int a4, int kc, int nvalid,
                                        float (&acc)[NT][MTMAX]) {
    // a4 <= KSTRIDE for every shipped layer (H <= 108): the loop body runs at most once per thread
#pragma unroll 1
    for (int k = kc; k < a4; k += KSTRIDE) {
        float4 av[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) av[j] = *reinterpret_cast<const float4*>(ap[j] + 4 * k);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            // outputs beyond the run's multiplicity read row 0 and are discarded by the reduction
            const float4 w = *reinterpret_cast<const float4*>(Wb + (m < nvalid ? m : 0) * HA + 4 * k);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                acc[j][m] = fmaf(av[j].x, w.x, acc[j][m]);
                acc[j][m] = fmaf(av[j].y, w.y, acc[j][m]);
                acc[j][m] = fmaf(av[j].z, w.z, acc[j][m]);
                acc[j][m] = fmaf(av[j].w, w.w, acc[j][m]);
            }
        }
    }
}

// Packed-FMA variant for thread tiles of at most 8 outputs (mul <= 8 runs: 5 of every 6 rows).  Blackwell's FFMA2
// (fma.rn.f32x2) does two independent fp32 FMAs per issue slot; the even and the odd K columns of a chunk accumulate in the
// two halves of one register pair, acc[j][2m] / acc[j][2m+1], which the run-end reduction adds.  Same IEEE fp32 FMAs, half
// the FMA instructions.
__device__ __forceinline__ void fma2(float& a0, float& a1, float x0, float x1, float y0, float y1) {
    unsigned long long a, x, y;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(y0), "f"(y1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(x), "l"(y));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}

template <int MT>
__device__ __forceinline__ void row_fma2(const float* const (&ap)[NT], const float* __restrict__ Wb, int HA, int a4, int kc, int nvalid,
                                         float (&acc)[NT][MTMAX]) {
    static_assert(2 * MT <= MTMAX, "pair accumulators need two slots per output");
#pragma unroll 1
    for (int k = kc; k < a4; k += KSTRIDE) {
        float4 av[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) av[j] = *reinterpret_cast<const float4*>(ap[j] + 4 * k);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const float4 w = *reinterpret_cast<const float4*>(Wb + (m < nvalid ? m : 0) * HA + 4 * k);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                fma2(acc[j][2 * m], acc[j][2 * m + 1], av[j].x, av[j].y, w.x, w.y);
                fma2(acc[j][2 * m], acc[j][2 * m + 1], av[j].z, av[j].w, w.z, w.w);
            }
        }
    }
}

// Wide runs (mul > 8: the 0e output block, two thirds of the transform's FMAs): a thread owns 2 nodes x MT outputs and
// visits its K chunks kc, kc + KSTRIDE_W, ...; even / odd K columns accumulate in the two halves of a register pair
// (fma.rn.f32x2), so the 2 x MT x 2 partial sums fill the same 64 registers the 4 x MT scalar tile used while the inner
// loop issues 82 instead of 276 instructions per 256 MACs.
constexpr int NTW = 2;
constexpr int KSTRIDE_W = 2 * WPC;
template <int MT>
__device__ __forceinline__ void row_fma2_wide(const float* const (&ap)[NTW], const float* __restrict__ Wb, int HA, int a4, int kc, int nvalid,
                                              float (&acc)[NT][MTMAX]) {
    static_assert(NTW * MT * 2 <= NT * MTMAX, "pair accumulators must fit the register tile");
#pragma unroll 1
    for (int k = kc; k < a4; k += KSTRIDE_W) {
        float4 av[NTW];
#pragma unroll
        for (int j = 0; j < NTW; ++j) av[j] = *reinterpret_cast<const float4*>(ap[j] + 4 * k);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const float4 w = *reinterpret_cast<const float4*>(Wb + (m < nvalid ? m : 0) * HA + 4 * k);
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                constexpr int dummy = 0; (void)dummy;
                float& lo = acc[(j * 2 * MT + 2 * m) / MTMAX][(j * 2 * MT + 2 * m) % MTMAX];
                float& hi = acc[(j * 2 * MT + 2 * m + 1) / MTMAX][(j * 2 * MT + 2 * m + 1) % MTMAX];
                fma2(lo, hi, av[j].x, av[j].y, w.x, w.y);
                fma2(lo, hi, av[j].z, av[j].w, w.z, w.w);
            }
        }
    }
}

struct ConsumerCtx {
    float* sm; const int* items; const int* n_items; uint32_t full_bar, empty_bar;   // barriers: shared-space addresses of [STAGES]
    int stage, a_stage, zero, n_active, HA, a4, nl, kc, lane;
    int rs; bool mine; int w_off, nvalid, row_begin, row_end, run_rs;
};

// all row groups of one run (over every active slot) for one consumer thread
template <int MT>
__device__ __forceinline__ void run_groups(const ConsumerCtx& c, int& G, float (&acc)[NT][MTMAX]) {
    constexpr bool WIDE = 2 * MT > MTMAX;       // mul > 8 runs: 2 nodes per thread, packed FMAs (row_fma2_wide)
    constexpr int NN = WIDE ? NTW : NT;
    const int nl = WIDE ? (c.lane & 15) : c.nl, nstep = WIDE ? 16 : 8;
    const int kc = WIDE ? (c.kc >> 2) * 2 + (c.lane >> 4) : c.kc;     // c.kc = 4 * warp-in-combo + (lane >> 3)
#pragma unroll 1
    for (int k = 0; k < c.n_active; ++k) {
        const int n_act = c.n_items[k];
        int a_off[NN];      // offset of the thread's nodes inside the A stage, or -1: nodes without edges read zeros
#pragma unroll
        for (int j = 0; j < NN; ++j) {
            const int rank = c.items[k * NB + nl + nstep * j];
            a_off[j] = rank >= 0 ? (c.rs * n_act + rank) * c.HA : -1;
        }
#pragma unroll 1
        for (int rg = c.row_begin; rg < c.row_end; rg += c.run_rs, ++G) {
            const int s = G % STAGES;
            const float* Ab = c.sm + s * c.stage;
            tc::mbar_wait_addr(c.full_bar + 8u * s, (G / STAGES) & 1);
            if (c.mine && rg + c.rs < c.row_end) {
                const float* ap[NN];
#pragma unroll
                for (int j = 0; j < NN; ++j) ap[j] = a_off[j] >= 0 ? Ab + a_off[j] : c.sm + c.zero;
                if constexpr (WIDE) row_fma2_wide<MT>(ap, Ab + c.a_stage + c.w_off, c.HA, c.a4, kc, c.nvalid, acc);
                else row_fma2<MT>(ap, Ab + c.a_stage + c.w_off, c.HA, c.a4, c.kc, c.nvalid, acc);
            }
            __syncwarp();
            if (c.lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c.empty_bar + 8u * s) : "memory");
        }
    }
}

__global__ void __launch_bounds__(TT, 1)
tp_transform_kernel(const __grid_constant__ cb_tp_conv_args a) {
    extern __shared__ __align__(128) float sm[];
    const int H = a.H, HA = H + PADC, d_out = a.d_out, n_rows = a.n_rows;
    const int a4 = HA / 4;
    __shared__ SlotTable st;
    __shared__ int active[CB_MAX_SEGS], n_items_s[CB_MAX_SEGS], deg_tot[NB];
    __shared__ RunS run_s[MAX_RUNS];
    __shared__ signed char run_split[MAX_RUNS];    // which CTA of the tile (blockIdx.y) processes the run
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = a.node_begin + blockIdx.x * NB;
    if (tid == 0) {
        build_slots(a, st);
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full_bar[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar[s])), "r"(CWARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const Layout L = make_layout(H, d_out, st.n_slots);
    float* part = sm + L.part;        // [CWARPS][8 node lanes][NT][MTMAX]
    float* outacc = sm + L.outacc;    // [NB][d_out]
    int* items = reinterpret_cast<int*>(sm + L.items);   // [active slot][NB] rank of the node among the tile's active nodes, or -1

    // ---- per-CTA tables
#pragma unroll 1
    for (int ri = tid; ri < a.n_runs; ri += TT) {
        const cb_tp_run rn = a.runs[ri];
        RunS s;
        s.row_begin = rn.row_begin; s.row_end = rn.row_end; s.mul = rn.mul; s.w_base0 = rn.w_base0;
        s.out_base = rn.out_base; s.out_step = rn.out_step;
        // mul <= 8: all outputs in one thread tile, two rows per group; else the outputs are split over the two combos
        s.mt = rn.mul <= 8 ? ((rn.mul + 1) & ~1) : (rn.mul <= 16 ? 8 : (rn.mul <= 24 ? 12 : 16));
        s.rs = rn.mul <= 8 ? 2 : 1;
        run_s[ri] = s;
    }
#pragma unroll 1
    for (int i = tid; i < NB * d_out; i += TT) outacc[i] = 0.0f;
    if (tid < HA) sm[L.zero + tid] = 0.0f;
    if (tid < NB) deg_tot[tid] = 0;
    int n_active = 0;
#pragma unroll 1
    for (int q = 0; q < st.n_slots; ++q) {
        int it = -1;
        if (tid < NB) {
            const int node = t0 + tid;
            int deg = 0;
            if (node >= st.lo[q] && node < st.hi[q]) {
#pragma unroll 1
                for (int s = st.first_seg[q]; s < st.first_seg[q] + st.n_segs[q]; ++s) {
                    const cb_tp_segment& sg = a.segs[s];
                    deg += seg_degree(sg, node);
                }
            }
            if (deg > 0) it = 0;
            deg_tot[tid] += deg;
        }
        if (tid < 32) {                            // NB == 32: warp 0 ranks the active nodes
            const unsigned act = __ballot_sync(0xffffffffu, it >= 0);
            items[n_active * NB + tid] = it >= 0 ? __popc(act & ((1u << tid) - 1u)) : -1;
        }
        const int cnt = __syncthreads_count(it >= 0);
        if (cnt > 0) {                             // slots without any edge into this tile are skipped
            if (tid == 0) { active[n_active] = q; n_items_s[n_active] = cnt; }
            ++n_active;
        }
    }
    if (a.n_runs == 0) n_active = 0;
    __syncthreads();
    // ---- small launches (few node tiles) split a tile's runs over gridDim.y CTAs: runs that feed the same output
    // block stay together, every output channel is written by exactly one CTA (channels no run writes: CTA 0)
    int* chan_owner = reinterpret_cast<int*>(sm + L.owner);
    const int n_split = (int)gridDim.y, split = (int)blockIdx.y;
    if (n_split == 1) {
        for (int i = tid; i < d_out; i += TT) chan_owner[i] = 0;
        for (int i = tid; i < a.n_runs; i += TT) run_split[i] = 0;
    } else if (tid == 0) {
        int* blk_of = reinterpret_cast<int*>(part);      // scratch: the reduction buffer is idle during set-up
        int* blk_cost = blk_of + MAX_RUNS;
        int* blk_split = blk_cost + MAX_RUNS;
        int load[8] = {0, 0, 0, 0, 0, 0, 0, 0}, n_blk = 0;
        for (int ri = 0; ri < a.n_runs; ++ri) {
            int b = -1;
            for (int rj = 0; rj < ri; ++rj)
                if (run_s[rj].out_base == run_s[ri].out_base && run_s[rj].out_step == run_s[ri].out_step) b = blk_of[rj];
            if (b < 0) { b = n_blk++; blk_cost[b] = 0; }
            blk_of[ri] = b;
            blk_cost[b] += (run_s[ri].row_end - run_s[ri].row_begin + run_s[ri].rs - 1) / run_s[ri].rs;   // row groups = chain length
        }
        for (int it = 0; it < n_blk; ++it) {        // heaviest block first onto the least loaded CTA
            int best = -1;
            for (int b = 0; b < n_blk; ++b)
                if (blk_cost[b] >= 0 && (best < 0 || blk_cost[b] > blk_cost[best])) best = b;
            int tgt = 0;
            for (int k = 1; k < n_split; ++k)
                if (load[k] < load[tgt]) tgt = k;
            blk_split[best] = tgt;
            load[tgt] += blk_cost[best];
            blk_cost[best] = -1;
        }
        for (int o = 0; o < d_out; ++o) chan_owner[o] = 0;
        for (int ri = 0; ri < a.n_runs; ++ri) {
            run_split[ri] = (signed char)blk_split[blk_of[ri]];
            for (int m = 0; m < run_s[ri].mul; ++m) chan_owner[run_s[ri].out_base + m * run_s[ri].out_step] = run_split[ri];
        }
    }
    __syncthreads();

    // ---- group sequence: for every run, for every active slot, the run's rows in groups of run.rs rows.  The
    // partial sums of a run stay in registers across the slots (they add into the same output channels).
    if (warp == CWARPS) {
        // ================= producer warp
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)HA * 4u;
            int G = 0;
#pragma unroll 1
            for (int ri = 0; ri < a.n_runs; ++ri) {
                if (run_split[ri] != split) continue;
                const RunS run = run_s[ri];
#pragma unroll 1
                for (int k = 0; k < n_active; ++k) {
                    const int q = active[k], n_act = n_items_s[k];
                    const float* w2a = a.segs[st.first_seg[q]].W2a + (size_t)run.w_base0 * HA;
                    // the tile's active rows are contiguous in the workspace: [row][rank][HA]
                    const float* ws = a.workspace + (size_t)(st.tile_off[q] + (int)blockIdx.x - st.tile0[q]) * n_rows * WS_TILE * HA;
#pragma unroll 1
                    for (int rg = run.row_begin; rg < run.row_end; rg += run.rs, ++G) {
                        const int s = G % STAGES;
                        mbar_wait(&empty_bar[s], ((G / STAGES) & 1) ^ 1);
                        const int nrs = min(run.rs, run.row_end - rg);
                        float* Ab = sm + s * L.stage;
                        const uint32_t a_bytes = (uint32_t)(nrs * n_act) * row_bytes, w_bytes = (uint32_t)(nrs * run.mul) * row_bytes;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(a_bytes + w_bytes)
                                     : "memory");
                        bulk_g2s(Ab, ws + (size_t)rg * n_act * HA, a_bytes, &full_bar[s]);
                        bulk_g2s(Ab + L.a_stage, w2a + (size_t)(rg - run.row_begin) * run.mul * HA, w_bytes, &full_bar[s]);
                    }
                }
            }
        }
    } else {
        // ================= consumer warps
        const int nl = lane & 7, combo = warp / WPC, kc = (warp % WPC) * 4 + (lane >> 3);
        float acc[NT][MTMAX];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int m = 0; m < MTMAX; ++m) acc[j][m] = 0.0f;
        int G = 0;
#pragma unroll 1
        for (int ri = 0; ri < a.n_runs; ++ri) {
            if (run_split[ri] != split) continue;
            const RunS run = run_s[ri];
            // combo -> (row of the group, first output of the thread tile)
            const int rs = run.rs == 2 ? combo : 0;
            const int m0 = run.rs == 2 ? 0 : combo * run.mt;
            const ConsumerCtx cx{sm, items, n_items_s, smem_u32(full_bar), smem_u32(empty_bar), L.stage, L.a_stage, L.zero, n_active, HA, a4, nl, kc, lane,
                                 rs, m0 < run.mul, (rs * run.mul + m0) * HA, run.mul - m0, run.row_begin, run.row_end, run.rs};
            switch (run.mt) {
                case 2: run_groups<2>(cx, G, acc); break;
                case 4: run_groups<4>(cx, G, acc); break;
                case 6: run_groups<6>(cx, G, acc); break;
                case 8: run_groups<8>(cx, G, acc); break;
                case 12: run_groups<12>(cx, G, acc); break;
                default: run_groups<16>(cx, G, acc); break;
            }
            // ---- end of the run: reduce the partial sums in a fixed order and add into the output channels
            asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");   // the previous run's reducers are done with `part`
            if (2 * run.mt <= MTMAX) {      // uniform.  Thread tiles of <= 8 outputs: 4 nodes, even / odd K columns in two slots per output (row_fma2)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int m = 0; m < MTMAX / 2; ++m)
                        if (m < run.mt) {
                            float v = acc[j][2 * m] + acc[j][2 * m + 1];
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            if (lane < 8) part[((warp * 8 + nl) * NT + j) * MTMAX + m] = v;
                        }
            } else {                        // wide tiles: 2 nodes (lane & 15, + 16) x run.mt outputs, pairs at 2 * (j * mt + m)
#pragma unroll
                for (int j = 0; j < NTW; ++j)
#pragma unroll
                    for (int m = 0; m < MTMAX; ++m)
                        if (m < run.mt) {
                            // flat pair index 2 * (j * mt + m): the run's mt is one of 12 / 16 (dispatch below)
                            const int f12 = 2 * (j * 12 + m), f16 = 2 * (j * 16 + m);
                            const float v12 = m < 12 ? acc[(f12 / MTMAX) % NT][f12 % MTMAX] + acc[((f12 + 1) / MTMAX) % NT][(f12 + 1) % MTMAX] : 0.0f;
                            const float v16 = acc[f16 / MTMAX][f16 % MTMAX] + acc[(f16 + 1) / MTMAX][(f16 + 1) % MTMAX];
                            float v = run.mt == 12 ? v12 : v16;
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            const int n = (lane & 15) + 16 * j;
                            if (lane < 16) part[((warp * 8 + (n & 7)) * NT + (n >> 3)) * MTMAX + m] = v;
                        }
            }
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int m = 0; m < MTMAX; ++m) acc[j][m] = 0.0f;
            asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");
#pragma unroll 1
            for (int i = tid; i < NB * run.mul; i += CT) {
                const int m = i % run.mul, n = i / run.mul;
                const int c_lo = run.rs == 2 ? 0 : m / run.mt, c_hi = run.rs == 2 ? 2 : c_lo + 1;
                const int mm = run.rs == 2 ? m : m - c_lo * run.mt;
                float v = 0.0f;
#pragma unroll 1
                for (int w = c_lo * WPC; w < c_hi * WPC; ++w) v += part[((w * 8 + (n & 7)) * NT + (n >> 3)) * MTMAX + mm];
                outacc[n * d_out + run.out_base + m * run.out_step] += v;
            }
        }
    }
    __syncthreads();
    // ---- epilogue: mean over all incoming edges, BatchNorm (eval) affine, residual
#pragma unroll 1
    for (int i = tid; i < NB * d_out; i += TT) {
        const int n = i / d_out, o = i - n * d_out;
        const int node = t0 + n;
        if (node < a.node_end && chan_owner[o] == split) {
            float v = outacc[i];
            if (!(a.flags & CB_TP_RAW_SUM)) {
                int deg = deg_tot[n];
                if (a.pre_sum != nullptr && node >= a.pre_n0 && node < a.pre_n1) {   // sample-invariant part computed once (cb200.h)
                    const int k = (node - a.pre_n0) % a.pre_period;
                    v += __ldg(a.pre_sum + (size_t)k * d_out + o);
                    deg += __ldg(a.pre_deg + k);
                }
                v = v / (float)max(deg, 1);
                if (a.bn_scale) v = fmaf(v, a.bn_scale[o], a.bn_shift[o]);
                if (a.residual && o < a.d_res) v += a.residual[(size_t)node * a.ld_res + o];
            }
            a.out[(size_t)node * d_out + o] = v;
        }
    }
}

#include "tp_transform_tc.cuh"

}  // namespace tf

template <class C>
int launch_accumulate(const cb_tp_conv_args* a, int items, cudaStream_t st) {
    const SmemLayout L = make_layout(a->n_rows, a->n_terms, a->ne, a->d_in, a->S, C::RP, C::HP);
    const size_t smem = (size_t)L.total * 4;
    CB_CHECK_ARG(smem <= 220 * 1024, "cb_tp_conv_forward: accumulate kernel needs %zu B of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(tp_accumulate_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cb_set_error("cb_tp_conv_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return CB_ERR_CUDA;
    }
    const int grid = items < CB_NUM_SMS * 4 ? items : CB_NUM_SMS * 4;  // persistent CTAs, items strided across them
    tp_accumulate_kernel<C><<<grid, C::THREADS, smem, st>>>(*a, items);
    CB_CHECK_LAUNCH("cb_tp_conv_forward(accumulate)");
    return CB_OK;
}

}  // namespace

#ifdef CB_PHASE_TIMING
extern "C" int cb_debug_phases(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tc::cb_dbg_phase, sizeof(unsigned long long) * 96);
    if (reset) {
        unsigned long long z[96] = {0};
        cudaMemcpyToSymbol(tc::cb_dbg_phase, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" int64_t cb_tp_conv_items(const cb_tp_conv_args* a) {
    if (a == nullptr || a->n_segs < 0 || a->n_segs > CB_MAX_SEGS) return -1;
    SlotTable t;
    build_slots(*a, t);
    return (int64_t)t.tile_off[t.n_slots] * WS_TILE;
}

extern "C" int cb_tp_conv_forward(const cb_tp_conv_args* a, void* stream) {
    CB_CHECK_ARG(a != nullptr, "cb_tp_conv_forward: null args");
    if (a->n_out <= 0 || a->node_end <= a->node_begin) return CB_OK;
    CB_CHECK_ARG(a->x && a->rows && a->terms && a->runs && a->out, "cb_tp_conv_forward: null pointer");
    CB_CHECK_ARG(a->n_segs >= 0 && a->n_segs <= CB_MAX_SEGS, "cb_tp_conv_forward: n_segs=%d", a->n_segs);
    CB_CHECK_ARG(a->H > 0 && a->H % 4 == 0 && a->ne > 0 && a->ne % 4 == 0, "cb_tp_conv_forward: H=%d ne=%d must be multiples of 4",
                 a->H, a->ne);
    CB_CHECK_ARG(a->n_rows > 0 && a->n_runs > 0 && a->d_out > 0 && a->d_in > 0 && a->S > 0, "cb_tp_conv_forward: bad sizes");
    CB_CHECK_ARG((a->bn_scale == nullptr) == (a->bn_shift == nullptr), "cb_tp_conv_forward: bn_scale/bn_shift must come together");
    CB_CHECK_ARG((a->pre_sum == nullptr) == (a->pre_deg == nullptr), "cb_tp_conv_forward: pre_sum/pre_deg must come together");
    CB_CHECK_ARG(a->pre_sum == nullptr || (a->pre_period > 0 && 0 <= a->pre_n0 && a->pre_n0 <= a->pre_n1 && a->pre_n1 <= a->n_out),
                 "cb_tp_conv_forward: pre_sum range [%d,%d) period %d outside [0,%d)", a->pre_n0, a->pre_n1, a->pre_period, a->n_out);
    CB_CHECK_ARG(0 <= a->node_begin && a->node_end <= a->n_out, "cb_tp_conv_forward: node range [%d,%d) outside [0,%d)",
                 a->node_begin, a->node_end, a->n_out);
    for (int s = 0; s < a->n_segs; ++s) {
        const cb_tp_segment& g = a->segs[s];
        CB_CHECK_ARG(g.rowptr && g.col && g.e_attr && g.sh && g.W1e && g.b1 && g.W2a, "cb_tp_conv_forward: segment %d has a null pointer", s);
        CB_CHECK_ARG(0 <= g.n0 && g.n0 <= g.n1 && g.n1 <= a->n_out, "cb_tp_conv_forward: segment %d node range [%d,%d) outside [0,%d)", s, g.n0, g.n1, a->n_out);
        if (s > 0) {
            const cb_tp_segment& p = a->segs[s - 1];
            CB_CHECK_ARG(g.slot == p.slot || g.slot == p.slot + 1, "cb_tp_conv_forward: slots must be consecutive in segment order");
            if (g.slot == p.slot)
                CB_CHECK_ARG(g.n0 == p.n0 && g.n1 == p.n1 && g.W2a == p.W2a && g.W1e == p.W1e && g.P_agg == p.P_agg && g.e_post == p.e_post,
                             "cb_tp_conv_forward: segments of slot %d must share node range and radial MLP", g.slot);
        } else {
            CB_CHECK_ARG(g.slot == 0, "cb_tp_conv_forward: first segment must be slot 0");
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    SlotTable t;
    build_slots(*a, t);
    const int64_t ws_items = (int64_t)t.tile_off[t.n_slots] * WS_TILE;   // accumulator slots of the workspace (cb_tp_conv_items)
    const int64_t items = t.item_off[t.n_slots];                          // (node, slot) pairs the accumulate kernel visits
    const int H = a->H, R = a->n_rows;
    if (items > 0) {
        CB_CHECK_ARG(a->workspace != nullptr && a->workspace_floats >= ws_items * (int64_t)R * (H + PADC),
                     "cb_tp_conv_forward: workspace too small (%lld floats for %lld accumulators)",
                     (long long)a->workspace_floats, (long long)ws_items);
        int rc;
        bool ws_ok = false;
        if (a->accum_mode == 4) {
            // warp-specialised kernel (tp_accumulate_ws.cuh): transposed accumulator, 1 CTA per SM, 512 TMEM columns
            const tc::ws::LayoutWS LW = tc::ws::make_layout_ws(R, a->n_terms, a->ne, a->d_in, a->S, H);
            ws_ok = R <= tc::ws::ACC_COLS && H % 32 == 0 && H <= 128 && a->ne == 32 && a->d_in % 2 == 0 && LW.total <= 225 * 1024;   // ne = 32: 128-byte swizzled rows
            if (ws_ok) {
                const size_t smem = (size_t)LW.total;
                cudaError_t e = cudaFuncSetAttribute(tc::ws::tp_accumulate_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) {
                    cb_set_error("cb_tp_conv_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                    return CB_ERR_CUDA;
                }
                const int grid = items < CB_NUM_SMS ? (int)items : CB_NUM_SMS;
                tc::ws::tp_accumulate_ws_kernel<<<grid, tc::ws::THREADS_WS, smem, st>>>(*a, (int)items);
                CB_CHECK_LAUNCH("cb_tp_conv_forward(accumulate, warp-specialised)");
                rc = CB_OK;
            }
        }
        if (ws_ok) {
        } else if (a->accum_mode >= 2) {
            CB_CHECK_ARG(a->d_in % 2 == 0, "cb_tp_conv_forward: tcgen05 accumulate needs an even node-feature width (d_in=%d)", a->d_in);
            CB_CHECK_ARG(R <= 384 && ((R + 127) / 128) * ((H + 1 + 15) / 16 * 16) + tc::KC <= tc::TMEM_COLS,
                         "cb_tp_conv_forward: tcgen05 accumulate supports rows<=384 and tiles within 256 TMEM columns (rows=%d H=%d)", R, H);
            CB_CHECK_ARG(a->ne % 8 == 0 && H <= 120, "cb_tp_conv_forward: tcgen05 accumulate needs ne %% 8 == 0 and H <= 120 (ne=%d H=%d)", a->ne, H);
            // transposed accumulator (no shared-memory epilogue) whenever the f-rows fit 240 TMEM columns and the hidden width
            // is a multiple of 32 (a lane group is either full or empty)
            const bool transposed = a->accum_mode >= 3 && R <= 240 && H % 32 == 0 && H <= 128;
            const tc::Layout L = tc::make_layout(R, a->n_terms, a->ne, a->d_in, a->S, H, transposed);
            // at least 80 KB so that never more than 2 CTAs (2 x 256 TMEM columns) share an SM
            const size_t smem = (size_t)(L.total > 80 * 1024 ? L.total : 80 * 1024);
            // <= 112 KB: two CTAs per SM; wider edge embeddings (generic layer call) run one CTA per SM
            CB_CHECK_ARG(smem <= 220 * 1024, "cb_tp_conv_forward: tcgen05 accumulate needs %zu B of shared memory", smem);
            auto kern = transposed ? tc::tp_accumulate_tc_kernel<true> : tc::tp_accumulate_tc_kernel<false>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                cb_set_error("cb_tp_conv_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return CB_ERR_CUDA;
            }
            const int per_sm = smem <= 112 * 1024 ? 2 : 1;
            const int blocks = (int)((items + tc::ITEM_BLOCK - 1) / tc::ITEM_BLOCK);
            const int grid = blocks < CB_NUM_SMS * per_sm ? blocks : CB_NUM_SMS * per_sm;
            kern<<<grid, tc::THREADS, smem, st>>>(*a, (int)items);
            CB_CHECK_LAUNCH("cb_tp_conv_forward(accumulate, tcgen05)");
            rc = CB_OK;
        } else
        if (H <= 64 && R <= 256) rc = launch_accumulate<Cfg<64, 8, 4, 8>>(a, (int)items, st);
        else if (H <= 96 && R <= 256) rc = launch_accumulate<Cfg<64, 8, 4, 12>>(a, (int)items, st);
        else if (H <= 80 && R <= 320) rc = launch_accumulate<Cfg<80, 4, 4, 20>>(a, (int)items, st);
        else if (H <= 96 && R <= 320) rc = launch_accumulate<Cfg<80, 8, 4, 12>>(a, (int)items, st);
        else {
            CB_CHECK_ARG(false, "cb_tp_conv_forward: no tile configuration for H=%d rows=%d", H, R);
            return CB_ERR_ARG;
        }
        if (rc != CB_OK) return rc;
    }
    // transform + epilogue
    CB_CHECK_ARG(a->n_runs <= tf::MAX_RUNS, "cb_tp_conv_forward: %d runs (max %d)", a->n_runs, tf::MAX_RUNS);
    const tf::Layout L = tf::make_layout(H, a->d_out, t.n_slots);
    const size_t smem = sizeof(float) * (size_t)L.total;
    CB_CHECK_ARG(smem <= 220 * 1024, "cb_tp_conv_forward: transform kernel needs %zu B of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(tf::tp_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cb_set_error("cb_tp_conv_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return CB_ERR_CUDA;
    }
    const int tiles = cb_div_up(a->node_end - a->node_begin, tf::NB);
    // tensor-core transform: big launches of layers with a plan (cb200.h); small ones keep the FFMA kernel and its run split
    bool use_tc = a->chains != nullptr && a->blocks != nullptr && a->n_chains > 0 && tiles * 2 > CB_NUM_SMS && (H + PADC) % 4 == 0;
    for (int s = 0; s < a->n_segs && use_tc; ++s) use_tc = a->segs[s].W2t != nullptr;
    if (use_tc) {
        const tf::tt::LayoutTT LT = tf::tt::make_layout_tt(H + PADC, a->kp, a->d_out, a->n_chains, a->n_blocks, t.n_slots, a->max_chain_bytes);
        use_tc = LT.total <= 225 * 1024;
        if (use_tc) {
            cudaError_t e2 = cudaFuncSetAttribute(tf::tt::tp_transform_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT.total);
            if (e2 != cudaSuccess) {
                cb_set_error("cb_tp_conv_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e2));
                return CB_ERR_CUDA;
            }
            tf::tt::tp_transform_tc_kernel<<<tiles, tf::tt::THREADS_TT, LT.total, st>>>(*a, a->max_chain_bytes);
            CB_CHECK_LAUNCH("cb_tp_conv_forward(transform, tcgen05)");
            return CB_OK;
        }
    }
    // launches with few node tiles are chains of row groups on a handful of SMs: split each tile's runs over up to 8 CTAs
    int n_split = 1;
    while (n_split < 8 && n_split < a->n_runs && tiles * n_split * 2 <= CB_NUM_SMS * 3) n_split *= 2;
    tf::tp_transform_kernel<<<dim3(tiles, n_split), tf::TT, smem, st>>>(*a);
    CB_CHECK_LAUNCH("cb_tp_conv_forward(transform)");
    return CB_OK;
}
