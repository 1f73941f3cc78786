// K3 -- fused tensor-product convolution layer (the kernel that matters).
// Replaces TensorProductConvLayer.forward (models/tensor_layers.py:195-217): gather, radial MLP,
// FasterTensorProduct / e3nn FullyConnectedTensorProduct, scatter-mean, BatchNorm(eval), residual.
//
// Formulation (exact up to fp re-association).  For aggregation node i and edge segment s
//     sum_{e in s, e->i} tp_e  =  T_s( A ),   A[r][j] = sum_e f_e[r] * h~_e[j]
//   f_e[r]  : the CG "intermediates" of the tensor product, sum of coef * x[col[e]][.] * sh_e[.]
//             (R rows; FasterTensorProduct's out_dict entries tensor_layers.py:72-85, or one row
//             per (e3nn instruction, u, k))
//   h~_e    : [relu(W1 a_e + b1) ; 1]  (H+1 columns), the hidden layer of the radial MLP (layers.py:8-15)
//   T_s(A)[o(r,m)] += sum_j W2[w(r)+m][j] A[r][j] + b2[w(r)+m] A[r][H]
// so the [E, weight_numel] per-edge weight tensor of the reference (6.6 KB per edge) is never
// formed: per edge the kernel does an R x H rank-1 update held in registers (one CTA per node,
// NRT x NCT threads each owning a TR x TC tile of A), and the weight_numel x H contraction with W2
// happens once per (node, segment) instead of once per edge.
//
// One CTA per aggregation node; heavy nodes (ligand atoms with hundreds of cross edges) come first in
// the node ordering so the tail of the grid is made of light receptor nodes.  Everything is
// deterministic: segmented in-register / shuffle / fixed-order shared-memory reductions, no atomics.
#include "common.cuh"
#include "../../include/cb200.h"

namespace {

constexpr int CH = 16;  // edges staged per chunk

template <int NRT_, int NCT_, int TR_, int TC_>
struct Cfg {
    static constexpr int NRT = NRT_, NCT = NCT_, TR = TR_, TC = TC_;
    static constexpr int THREADS = NRT * NCT, RP = NRT * TR, HP = NCT * TC;
};

struct SmemLayout {
    int rows, terms, w1e, hbase, cols, xs, shs, es, F, Hs, P, outacc, total;  // offsets in 4-byte words
    int nep, dxp;
};

__host__ __device__ inline SmemLayout make_layout(int n_rows, int n_terms, int ne, int d_in, int S, int n_slots,
                                                  int d_out, int RP, int HP) {
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
    L.P = o;      o += al4(n_slots);
    L.outacc = o; o += al4(d_out);
    L.total = o;
    return L;
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
tp_conv_kernel(const __grid_constant__ cb_tp_conv_args a) {
    constexpr int NCT = C::NCT, TR = C::TR, TC = C::TC, RP = C::RP, HP = C::HP, THREADS = C::THREADS;
    static_assert(TR == 4 && TC % 4 == 0, "tile shape");
    extern __shared__ __align__(16) float sm[];
    const SmemLayout L = make_layout(a.n_rows, a.n_terms, a.ne, a.d_in, a.S, a.n_slots, a.d_out, RP, HP);
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
    float* P = sm + L.P;
    float* outacc = sm + L.outacc;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr = tid / NCT, tc = tid % NCT;
    const int node = blockIdx.x;
    const int H = a.H, ne = a.ne, S = a.S, d_in = a.d_in, n_rows = a.n_rows, nep = L.nep, dxp = L.dxp;

    // one-time staging: TP program, zeroed tiles
    {
        const int* src = reinterpret_cast<const int*>(a.rows);
        int* dst = reinterpret_cast<int*>(rows_s);
        for (int i = tid; i < n_rows * 8; i += THREADS) dst[i] = src[i];
        const int2* tsrc = reinterpret_cast<const int2*>(a.terms);
        int2* tdst = reinterpret_cast<int2*>(terms_s);
        for (int i = tid; i < a.n_terms; i += THREADS) tdst[i] = tsrc[i];
        for (int i = tid; i < CH * RP; i += THREADS) F[i] = 0.0f;
        for (int i = tid; i < CH * HP; i += THREADS) Hs[i] = 0.0f;
        for (int i = tid; i < a.d_out; i += THREADS) outacc[i] = 0.0f;
    }
    const int graph = a.agg_graph ? a.agg_graph[node] : 0;
    int deg_total = 0;
    __syncthreads();

    for (int s = 0; s < a.n_segs; ++s) {
        const cb_tp_segment& sg = a.segs[s];
        if (node < sg.n0 || node >= sg.n1) continue;  // block-uniform
        const int e0 = sg.rowptr[node - sg.n0], e1 = sg.rowptr[node - sg.n0 + 1];
        if (e1 <= e0) continue;
        deg_total += e1 - e0;

        // ---- per-segment setup: edge-embedding slice of the first Linear, constant part of h
        for (int i = tid; i < HP * ne; i += THREADS) {
            const int q = i / ne, c = i - q * ne;
            W1e_s[q * nep + c] = q < H ? sg.W1e[(size_t)q * sg.ldw1 + c] : 0.0f;
        }
        __syncthreads();
        for (int q = tid; q < HP; q += THREADS) {
            float v = 0.0f;
            if (q < H) {
                v = sg.b1[q];
                if (sg.P_agg) v += sg.P_agg[(size_t)node * sg.ldp_agg + q];
                if (sg.e_post) {
                    const float* ep = sg.e_post + (size_t)graph * ne;
                    for (int c = 0; c < ne; ++c) v = fmaf(W1e_s[q * nep + c], ep[c], v);
                }
            }
            hbase[q] = v;
        }

        float acc[TR][TC];
        float fsum[TR];
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            fsum[i] = 0.0f;
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = 0.0f;
        }

        for (int base = e0; base < e1; base += CH) {
            const int n = min(CH, e1 - base);
            // ---- gather raw operands of the chunk
            if (tid < n) cols_s[tid] = sg.col[base + tid] + sg.col_off;
            __syncthreads();  // also orders hbase / previous accumulate phase
            for (int i = tid; i < n * d_in; i += THREADS) {
                const int e = i / d_in, k = i - e * d_in;
                xs[e * dxp + k] = a.x[(size_t)cols_s[e] * d_in + k];  // cols_s already includes col_off
            }
            for (int i = tid; i < n * S; i += THREADS) shs[i] = sg.sh[(size_t)base * S + i];
            for (int i = tid; i < n * ne; i += THREADS) es[i] = sg.e_attr[(size_t)base * ne + i];
            __syncthreads();
            // ---- f rows: CG products of the gathered node features with the edge harmonics
            for (int i = tid; i < n * n_rows; i += THREADS) {
                const int e = i / n_rows, r = i - e * n_rows;
                const int tb = rows_s[r].term_begin, te = rows_s[r].term_end;
                float v = 0.0f;
                for (int t = tb; t < te; ++t) {
                    const cb_tp_term tm = terms_s[t];
                    v = fmaf(tm.coef * xs[e * dxp + tm.x_idx], shs[e * S + tm.sh_idx], v);
                }
                F[e * RP + r] = v;
            }
            // ---- hidden layer of the radial MLP
            {
                const int EG = THREADS / H;  // edge groups processed concurrently
                const int q = tid % H, eg = tid / H;
                if (eg < EG) {
                    for (int e = eg; e < n; e += EG) {
                        float v = hbase[q];
                        if (sg.P_nbr) v += sg.P_nbr[(size_t)cols_s[e] * sg.ldp_nbr + q];
                        const float4* w4 = reinterpret_cast<const float4*>(W1e_s + q * nep);
                        const float4* e4 = reinterpret_cast<const float4*>(es + e * ne);
                        for (int c = 0; c < ne / 4; ++c) {
                            const float4 w = w4[c], x4 = e4[c];
                            v = fmaf(w.x, x4.x, v);
                            v = fmaf(w.y, x4.y, v);
                            v = fmaf(w.z, x4.z, v);
                            v = fmaf(w.w, x4.w, v);
                        }
                        Hs[e * HP + q] = fmaxf(v, 0.0f);
                    }
                }
            }
            __syncthreads();
            // ---- rank-1 updates of the register tile
#pragma unroll 2
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
            // the barrier at the top of the next chunk (after cols) protects F / Hs / xs
        }

        // ---- transform: contract the tile with the second Linear (weights streamed from L2)
        {
            const unsigned gmask = (NCT == 32 ? 0xffffffffu : ((1u << NCT) - 1u)) << (lane & ~(NCT - 1));
#pragma unroll
            for (int i = 0; i < TR; ++i) {
                const int r = tr * TR + i;
                if (r < n_rows) {  // uniform across the NCT lanes sharing tr
                    const cb_tp_row rw = rows_s[r];
                    for (int m = 0; m < rw.mul; ++m) {
                        const float* wp = sg.W2 + (size_t)(rw.w_base + m) * H + tc * TC;
                        float sacc = 0.0f;
#pragma unroll
                        for (int j4 = 0; j4 < TC / 4; ++j4) {
                            if (tc * TC + 4 * j4 < H) {
                                const float4 w = __ldg(reinterpret_cast<const float4*>(wp) + j4);
                                sacc = fmaf(w.x, acc[i][4 * j4 + 0], sacc);
                                sacc = fmaf(w.y, acc[i][4 * j4 + 1], sacc);
                                sacc = fmaf(w.z, acc[i][4 * j4 + 2], sacc);
                                sacc = fmaf(w.w, acc[i][4 * j4 + 3], sacc);
                            }
                        }
                        if (tc == 0) sacc = fmaf(__ldg(sg.b2 + rw.w_base + m), fsum[i], sacc);
#pragma unroll
                        for (int o = NCT >> 1; o > 0; o >>= 1) sacc += __shfl_xor_sync(gmask, sacc, o);
                        if (tc == 0) P[rw.p_off + m] = sacc;
                    }
                }
            }
        }
        __syncthreads();
        // ---- fixed-order reduction of the partial sums into the output channels
        for (int o = warp; o < a.d_out; o += THREADS / 32) {
            const int p0 = a.out_ptr[o], p1 = a.out_ptr[o + 1];
            float v = 0.0f;
            for (int p = p0 + lane; p < p1; p += 32) v += P[a.out_idx[p]];
            v = cb_warp_sum(v);
            if (lane == 0) outacc[o] += v;
        }
        __syncthreads();
    }

    // ---- epilogue: mean over all incoming edges, BatchNorm (eval) affine, residual
    const float inv_deg = 1.0f / (float)max(deg_total, 1);
    for (int o = tid; o < a.d_out; o += THREADS) {
        float v = outacc[o] * inv_deg;
        if (a.bn_scale) v = fmaf(v, a.bn_scale[o], a.bn_shift[o]);
        if (a.residual && o < a.d_res) v += a.residual[(size_t)node * a.ld_res + o];
        a.out[(size_t)node * a.d_out + o] = v;
    }
}

template <class C>
int launch(const cb_tp_conv_args* a, cudaStream_t st) {
    const SmemLayout L = make_layout(a->n_rows, a->n_terms, a->ne, a->d_in, a->S, a->n_slots, a->d_out, C::RP, C::HP);
    const size_t smem = (size_t)L.total * 4;
    CB_CHECK_ARG(smem <= 227 * 1024, "cb_tp_conv_forward: needs %zu B of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(tp_conv_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cb_set_error("cb_tp_conv_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return CB_ERR_CUDA;
    }
    tp_conv_kernel<C><<<a->n_out, C::THREADS, smem, st>>>(*a);
    CB_CHECK_LAUNCH("cb_tp_conv_forward");
    return CB_OK;
}

}  // namespace

extern "C" int cb_tp_conv_forward(const cb_tp_conv_args* a, void* stream) {
    CB_CHECK_ARG(a != nullptr, "cb_tp_conv_forward: null args");
    if (a->n_out <= 0) return CB_OK;
    CB_CHECK_ARG(a->x && a->rows && a->terms && a->out_ptr && a->out_idx && a->out, "cb_tp_conv_forward: null pointer");
    CB_CHECK_ARG(a->n_segs >= 0 && a->n_segs <= CB_MAX_SEGS, "cb_tp_conv_forward: n_segs=%d", a->n_segs);
    CB_CHECK_ARG(a->H > 0 && a->H % 4 == 0 && a->ne > 0 && a->ne % 4 == 0, "cb_tp_conv_forward: H=%d ne=%d must be multiples of 4",
                 a->H, a->ne);
    CB_CHECK_ARG(a->n_rows > 0 && a->n_slots > 0 && a->d_out > 0 && a->d_in > 0 && a->S > 0, "cb_tp_conv_forward: bad sizes");
    CB_CHECK_ARG((a->bn_scale == nullptr) == (a->bn_shift == nullptr), "cb_tp_conv_forward: bn_scale/bn_shift must come together");
    for (int s = 0; s < a->n_segs; ++s) {
        const cb_tp_segment& g = a->segs[s];
        CB_CHECK_ARG(g.rowptr && g.col && g.e_attr && g.sh && g.W1e && g.b1 && g.W2 && g.b2, "cb_tp_conv_forward: segment %d has a null pointer", s);
        CB_CHECK_ARG(0 <= g.n0 && g.n0 <= g.n1 && g.n1 <= a->n_out, "cb_tp_conv_forward: segment %d node range [%d,%d) outside [0,%d)", s, g.n0, g.n1, a->n_out);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int H = a->H, R = a->n_rows;
    if (H <= 64 && R <= 256) return launch<Cfg<64, 8, 4, 8>>(a, st);
    if (H <= 96 && R <= 256) return launch<Cfg<64, 8, 4, 12>>(a, st);
    if (H <= 80 && R <= 320) return launch<Cfg<80, 4, 4, 20>>(a, st);
    if (H <= 96 && R <= 320) return launch<Cfg<80, 8, 4, 12>>(a, st);
    CB_CHECK_ARG(false, "cb_tp_conv_forward: no tile configuration for H=%d rows=%d", H, R);
    return CB_ERR_ARG;
}
