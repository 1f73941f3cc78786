// K2 -- fused edge featurisation: length -> Gaussian smearing -> 2-layer edge-embedding MLP, plus the
// real spherical harmonics of the edge direction, one thread per edge, weights staged in shared
// memory.  Replaces GaussianSmearing + o3.spherical_harmonics + *_edge_embedding
// (models/score_model.py:492-664; see include/cb200.h).
//
// HBM-bound by design: per edge it reads 8 B of indices + two positions (L2-resident gathers) and
// writes 4*(ns + S) B; the 36..68-wide raw feature vector never leaves the SM.
#include "common.cuh"
#include "../../include/cb200.h"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxNs = 64;

// smem layout (floats):
//   W1t [(n_extra + n_gauss)][ns]   transposed first Linear restricted to the in-kernel feature blocks
//   W2t [ns][ns]                    transposed second Linear
//   b2  [ns]
//   feat[(n_extra + n_gauss)][kThreads]
//   hid [ns][kThreads]
__global__ void __launch_bounds__(kThreads)
edge_feat_kernel(cb_edge_feat_args a) {
    extern __shared__ __align__(16) float sm[];
    const int ns = a.ns, nf = a.n_extra + a.n_gauss;
    float* W1t = sm;
    float* W2t = W1t + nf * ns;
    float* b2s = W2t + ns * ns;
    float* feat = b2s + ns;
    float* hid = feat + nf * kThreads;
    float* offs = hid + ns * kThreads;
    const int tid = threadIdx.x;
    for (int i = tid; i < a.n_gauss; i += kThreads) offs[i] = a.smear_offset[i];

    for (int i = tid; i < nf * ns; i += kThreads) {
        const int k = i / ns, o = i % ns;
        const int colw = k < a.n_extra ? a.extra_off + k : a.smear_off + (k - a.n_extra);
        W1t[i] = a.W1[o * a.ldw1 + colw];
    }
    for (int i = tid; i < ns * ns; i += kThreads) {
        const int k = i / ns, o = i % ns;
        W2t[i] = a.W2[o * ns + k];
    }
    for (int i = tid; i < ns; i += kThreads) b2s[i] = a.b2[i];
    __syncthreads();

    const int E = min(*a.n_edges_dev, a.e_cap);
    const int S = (a.lmax + 1) * (a.lmax + 1);
    const float coeff = a.smear_coeff;
    const int NS4 = ns / 4;

    // every thread of the block runs the same number of iterations (block-wide barriers inside)
    const int n_iter = (E + gridDim.x * kThreads - 1) / (gridDim.x * kThreads);
    for (int it = 0; it < n_iter; ++it) {
        const int e = (it * gridDim.x + blockIdx.x) * kThreads + tid;
        const bool live = e < E;
        int g = 0;
        if (live) {
            const int ia = a.row[e], ib = a.col[e];
            g = a.agg_graph ? a.agg_graph[ia] : 0;
            const float vx = a.sh_sign * (a.pos_nbr[3 * ib] - a.pos_agg[3 * ia]);
            const float vy = a.sh_sign * (a.pos_nbr[3 * ib + 1] - a.pos_agg[3 * ia + 1]);
            const float vz = a.sh_sign * (a.pos_nbr[3 * ib + 2] - a.pos_agg[3 * ia + 2]);
            const float len = sqrtf(vx * vx + vy * vy + vz * vz);
            // spherical harmonics of the unit vector, 'component' normalisation
            const float inv = 1.0f / fmaxf(len, 1e-12f);
            const float ux = vx * inv, uy = vy * inv, uz = vz * inv;
            float* sh = a.out_sh + (size_t)e * S;
            const float s3 = 1.7320508075688772f;
            sh[0] = 1.0f;
            sh[1] = s3 * ux;
            sh[2] = s3 * uy;
            sh[3] = s3 * uz;
            if (a.lmax >= 2) {
                const float s5 = 2.23606797749979f, s15 = 3.872983346207417f;
                sh[4] = s15 * ux * uz;
                sh[5] = s15 * ux * uy;
                sh[6] = s5 * (uy * uy - 0.5f * (ux * ux + uz * uz));
                sh[7] = s15 * uy * uz;
                sh[8] = 0.5f * s15 * (uz * uz - ux * ux);
            }
            for (int k = 0; k < a.n_extra; ++k)
                feat[k * kThreads + tid] = a.extra ? a.extra[(size_t)e * a.n_extra + k] : 0.0f;
            for (int k = 0; k < a.n_gauss; ++k) {
                const float d = len - offs[k];
                feat[(a.n_extra + k) * kThreads + tid] = expf(coeff * d * d);
            }
        }
        // layer 1: 4 hidden units at a time, weights broadcast from smem as float4
        const float* b1 = a.b1_graph + (size_t)g * a.b1_graph_stride;
#pragma unroll 1
        for (int ob = 0; ob < NS4; ++ob) {
            float4 acc = live ? *reinterpret_cast<const float4*>(b1 + 4 * ob) : make_float4(0, 0, 0, 0);
            for (int k = 0; k < nf; ++k) {
                const float f = feat[k * kThreads + tid];
                const float4 w = *reinterpret_cast<const float4*>(W1t + k * ns + 4 * ob);
                acc.x = fmaf(f, w.x, acc.x);
                acc.y = fmaf(f, w.y, acc.y);
                acc.z = fmaf(f, w.z, acc.z);
                acc.w = fmaf(f, w.w, acc.w);
            }
            hid[(4 * ob + 0) * kThreads + tid] = fmaxf(acc.x, 0.0f);
            hid[(4 * ob + 1) * kThreads + tid] = fmaxf(acc.y, 0.0f);
            hid[(4 * ob + 2) * kThreads + tid] = fmaxf(acc.z, 0.0f);
            hid[(4 * ob + 3) * kThreads + tid] = fmaxf(acc.w, 0.0f);
        }
        // layer 2 (each thread only touches its own feat/hid columns: no barrier needed)
#pragma unroll 1
        for (int ob = 0; ob < NS4; ++ob) {
            float4 acc = *reinterpret_cast<const float4*>(b2s + 4 * ob);
            for (int k = 0; k < ns; ++k) {
                const float h = hid[k * kThreads + tid];
                const float4 w = *reinterpret_cast<const float4*>(W2t + k * ns + 4 * ob);
                acc.x = fmaf(h, w.x, acc.x);
                acc.y = fmaf(h, w.y, acc.y);
                acc.z = fmaf(h, w.z, acc.z);
                acc.w = fmaf(h, w.w, acc.w);
            }
            if (live) *reinterpret_cast<float4*>(a.out_attr + (size_t)e * ns + 4 * ob) = acc;
        }
    }
}

}  // namespace

extern "C" int cb_edge_featurize(const cb_edge_feat_args* a, void* stream) {
    CB_CHECK_ARG(a != nullptr, "cb_edge_featurize: null args");
    CB_CHECK_ARG(a->ns > 0 && a->ns <= kMaxNs && a->ns % 4 == 0, "cb_edge_featurize: ns=%d must be a multiple of 4 <= %d",
                 a->ns, kMaxNs);
    CB_CHECK_ARG(a->lmax == 1 || a->lmax == 2, "cb_edge_featurize: lmax=%d unsupported", a->lmax);
    CB_CHECK_ARG(a->n_gauss >= 2 && a->n_extra >= 0, "cb_edge_featurize: bad feature sizes");
    CB_CHECK_ARG(a->row && a->col && a->n_edges_dev && a->pos_agg && a->pos_nbr && a->b1_graph && a->W1 && a->W2 && a->smear_offset &&
                     a->b2 && a->out_attr && a->out_sh,
                 "cb_edge_featurize: null pointer");
    if (a->e_cap <= 0) return CB_OK;
    const int ns = a->ns, nf = a->n_extra + a->n_gauss;
    const size_t smem = sizeof(float) * ((size_t)nf * ns + (size_t)ns * ns + ns + (size_t)nf * kThreads + (size_t)ns * kThreads + a->n_gauss);
    int blocks = cb_div_up(a->e_cap, kThreads);
    const int max_blocks = CB_NUM_SMS * 4;
    if (blocks > max_blocks) blocks = max_blocks;
    cudaStream_t st = (cudaStream_t)stream;
    cudaFuncSetAttribute(edge_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    edge_feat_kernel<<<blocks, kThreads, smem, st>>>(*a);
    CB_CHECK_LAUNCH("cb_edge_featurize");
    return CB_OK;
}
