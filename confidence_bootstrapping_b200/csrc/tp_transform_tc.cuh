// K3 (b'): transform + epilogue on the tensor cores -- included inside namespace tf of tp_conv.cu.
//
// Same contraction as tp_transform_kernel,   out[n][o(r, m)] += sum_j A[n][r][j] * W2a[w(r) + m][j],   but as tcgen05 MMAs with
// the NODES of a 32-node tile along M.  The FFMA kernel is latency-bound (one third of the FMA pipe, 45 % issue slots, no
// resource saturated: profiles/r2/k3_ncu_summary_mid.txt) because an accumulator row only meets `mul` = 6 weight rows for 5
// of every 6 rows -- too little arithmetic per shared-memory load.  Here the rows that share their weights (the components of
// one output irrep: cb_tp_chain) are stacked along M, D[(component, node rank)][m] += A . W^T, so one chain of K/8 x 3 MMAs
// (3xTF32: A = hi + lo with hi taken by the tensor core's own truncation, W = hi + lo rounded on the host) does the work of
// 3 x 32 x mul x K FMAs, and the CUDA cores only re-lay the streamed accumulator rows into the MMA's operand layout.
//
//   warp 0       P  producer   : per chain one bulk copy per component (the tile's accumulator rows of workspace row
//                                row_c + u: n_act x HA floats, contiguous) + one of the chain's pre-laid W tiles (hi | lo)
//   warps 1-8    C  converters : raw rows -> K-major core-matrix tiles, 32 K columns per ring stage: hi = the raw fp32 word,
//                                lo = x - trunc_tf32(x)
//   warp 9       M  MMA issuer : per ring stage 4 k-steps x 3 MMAs (M = 128, N = 16 | 32) into the chain's TMEM partial sum
//   warps 10-13  E  epilogue   : per slot, component c = TMEM lane group c: adds the partial sums and scatters them to the
//                                tile's output accumulator; at the end mean / BatchNorm / residual like the FFMA kernel
//
// Accumulation order is fixed (one issuer, chains in table order), so results stay bit-reproducible.

namespace tt {

constexpr int C_WARPS = 8, E_WARPS = 4;
constexpr int W_P = 0, W_C0 = 1, W_M = W_C0 + C_WARPS, W_E0 = W_M + 1;
constexpr int THREADS_TT = 32 * (W_E0 + E_WARPS);
constexpr int NT_STAGES = 2;                 // A-tile ring: one stage = 32 K columns of one chain, hi + lo
constexpr int TILE_ROWS = 96;                // 3 components x 32 node ranks
constexpr int TILE_BYTES = TILE_ROWS * 128;  // one hi (or lo) tile of a stage: [12 groups of 8 rows][8 x 16-byte K chunks][8 rows]
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int NCB = 3;                       // chain staging buffers (raw accumulator rows + W tile): copies run NCB - 1 chains ahead
constexpr int SET_COLS = 256;                // TMEM columns of one accumulator set (two sets: slots alternate)

struct Chain { int row[3]; int n_comp, w_off, npad, acc_col, first; };
struct Block { int n_comp, mul, npad, acc_col0, n_partials, out_step, out_base[3], pad[3]; };

struct LayoutTT {
    int ring, chain_stage[NCB], outacc, chains, blocks, items, node_of, total;   // byte offsets
    int chain_bytes;
};
__host__ __device__ inline LayoutTT make_layout_tt(int HA, int kp, int d_out, int n_chains, int n_blocks, int n_slots, int max_chain_bytes) {
    LayoutTT L;
    auto al = [](int v, int a) { return (v + a - 1) / a * a; };
    int o = 0;
    L.ring = o; o += NT_STAGES * STAGE_BYTES + 4096;      // + the rows an M = 128 MMA reads beyond the last 96-row tile
    L.chain_bytes = al(max_chain_bytes, 128);
    for (int b = 0; b < NCB; ++b) { L.chain_stage[b] = o; o += L.chain_bytes; }
    L.outacc = o; o += al(32 * d_out * 4, 16);
    L.chains = o; o += al(n_chains * (int)sizeof(Chain), 16);
    L.blocks = o; o += al(n_blocks * (int)sizeof(Block), 16);
    L.items = o;  o += n_slots * 32 * 4;
    L.node_of = o; o += n_slots * 32 * 4;
    L.total = o;
    (void)HA; (void)kp;
    return L;
}

using tc::ws::mbar_arrive;
using tc::ws::mbar_init;
using tc::ws::mbar_wait_ws;
using tc::ws::umma_commit;

__global__ void __launch_bounds__(THREADS_TT, 1)
tp_transform_tc_kernel(const __grid_constant__ cb_tp_conv_args a, int max_chain_bytes) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int H = a.H, HA = H + PADC, d_out = a.d_out, n_rows = a.n_rows, KP = a.kp;
    __shared__ SlotTable st;
    __shared__ int active[CB_MAX_SEGS], n_items_s[CB_MAX_SEGS], deg_tot[NB];
    __shared__ __align__(8) uint64_t chain_full[NCB], chain_free[NCB], tile_full[NT_STAGES], tile_free[NT_STAGES], acc_full[2], acc_free[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ int n_active_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = a.node_begin + blockIdx.x * NB;
    if (tid == 0) {
        build_slots(a, st);
        for (int b = 0; b < NCB; ++b) { mbar_init(&chain_full[b], 1); mbar_init(&chain_free[b], C_WARPS + 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_free[b], E_WARPS); }
        for (int s = 0; s < NT_STAGES; ++s) { mbar_init(&tile_full[s], C_WARPS); mbar_init(&tile_free[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * SET_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    __syncthreads();
    const LayoutTT L = make_layout_tt(HA, KP, d_out, a.n_chains, a.n_blocks, st.n_slots, max_chain_bytes);
    float* outacc = reinterpret_cast<float*>(smraw + L.outacc);     // [NB][d_out]
    Chain* chains = reinterpret_cast<Chain*>(smraw + L.chains);
    Block* blocks = reinterpret_cast<Block*>(smraw + L.blocks);
    int* items = reinterpret_cast<int*>(smraw + L.items);           // [active slot][NB] rank of the node among the tile's active nodes, or -1
    int* node_of = reinterpret_cast<int*>(smraw + L.node_of);       // [active slot][NB] node (0..31) of a rank

    // ---- per-CTA tables
    {
        const int* src = reinterpret_cast<const int*>(a.chains);
        int* dst = reinterpret_cast<int*>(chains);
        for (int i = tid; i < a.n_chains * (int)(sizeof(Chain) / 4); i += THREADS_TT) dst[i] = src[i];
        const int* bsrc = reinterpret_cast<const int*>(a.blocks);
        int* bdst = reinterpret_cast<int*>(blocks);
        for (int i = tid; i < a.n_blocks * (int)(sizeof(Block) / 4); i += THREADS_TT) bdst[i] = bsrc[i];
        for (int i = tid; i < NB * d_out; i += THREADS_TT) outacc[i] = 0.0f;
        // the rows an M = 128 MMA reads beyond a 96-row tile must at least be finite on first use: zero the ring once
        for (int i = tid; i < (NT_STAGES * STAGE_BYTES + 4096) / 16; i += THREADS_TT) reinterpret_cast<float4*>(smraw + L.ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < NB) deg_tot[tid] = 0;
    }
    int n_active = 0;
#pragma unroll 1
    for (int q = 0; q < st.n_slots; ++q) {
        int it = -1;
        if (tid < NB) {
            const int node = t0 + tid;
            int deg = 0;
            if (node >= st.lo[q] && node < st.hi[q]) {
#pragma unroll 1
                for (int s = st.first_seg[q]; s < st.first_seg[q] + st.n_segs[q]; ++s) deg += seg_degree(a.segs[s], node);
            }
            if (deg > 0) it = 0;
            deg_tot[tid] += deg;
        }
        if (tid < 32) {
            const unsigned act = __ballot_sync(0xffffffffu, it >= 0);
            const int rank = it >= 0 ? __popc(act & ((1u << tid) - 1u)) : -1;
            items[n_active * NB + tid] = rank;
            if (rank >= 0) node_of[n_active * NB + rank] = tid;
        }
        const int cnt = __syncthreads_count(it >= 0);
        if (cnt > 0) {
            if (tid == 0) { active[n_active] = q; n_items_s[n_active] = cnt; }
            ++n_active;
        }
    }
    if (tid == 0) n_active_s = n_active;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int n_chains = a.n_chains;
    const int n_kc = (KP + 31) / 32;

    if (warp == W_P) {
        // =============================================================== P: producer
        if (lane == 0) {
            int G = 0;
#pragma unroll 1
            for (int k = 0; k < n_active; ++k) {
                const int q = active[k], n_act = n_items_s[k];
                const float* w2t = a.segs[st.first_seg[q]].W2t;
                const float* ws = a.workspace + (size_t)(st.tile_off[q] + (int)blockIdx.x - st.tile0[q]) * n_rows * WS_TILE * HA;
                const uint32_t a_bytes = (uint32_t)(n_act * HA * 4);
#pragma unroll 1
                for (int g = 0; g < n_chains; ++g, ++G) {
                    const Chain ch = chains[g];
                    const int cb = G % NCB;
                    mbar_wait_ws(&chain_free[cb], ((G / NCB) & 1) ^ 1);
                    unsigned char* stg = smraw + L.chain_stage[cb];
                    const uint32_t w_bytes = (uint32_t)(2 * ch.npad * KP * 4);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&chain_full[cb])),
                                 "r"(a_bytes * (uint32_t)ch.n_comp + w_bytes)
                                 : "memory");
#pragma unroll 1
                    for (int c = 0; c < ch.n_comp; ++c)
                        bulk_g2s(reinterpret_cast<float*>(stg + c * NB * HA * 4), ws + (size_t)ch.row[c] * n_act * HA, a_bytes, &chain_full[cb]);
                    bulk_g2s(reinterpret_cast<float*>(stg + ch.n_comp * NB * HA * 4), w2t + ch.w_off, w_bytes, &chain_full[cb]);
                }
            }
        }
    } else if (warp >= W_C0 && warp < W_M) {
        // =============================================================== C: converters (raw fp32 rows -> K-major hi / lo operand tiles)
        int G = 0, T = 0;
#pragma unroll 1
        for (int k = 0; k < n_active; ++k) {
            const int n_act = n_items_s[k];
#pragma unroll 1
            for (int g = 0; g < n_chains; ++g, ++G) {
                const int n_comp = chains[g].n_comp;
                const int cb = G % NCB;
                mbar_wait_ws(&chain_full[cb], (G / NCB) & 1);
                const float* raw = reinterpret_cast<const float*>(smraw + L.chain_stage[cb]);
#pragma unroll 1
                for (int kc = 0; kc < n_kc; ++kc, ++T) {
                    const int ts = T % NT_STAGES;
                    mbar_wait_ws(&tile_free[ts], ((T / NT_STAGES) & 1) ^ 1);
                    unsigned char* hi_t = smraw + L.ring + ts * STAGE_BYTES;
                    unsigned char* lo_t = hi_t + TILE_BYTES;
                    const int nk4 = min(8, (KP - 32 * kc) / 4);
                    // lane = node rank; a warp takes (component, 16-byte K chunk) pairs: conflict-free 400-byte-stride reads of the raw
                    // rows and 128-byte contiguous writes of a core-matrix column, no per-element index arithmetic
                    const int cw = warp - W_C0;
#pragma unroll 1
                    for (int p2 = cw; p2 < n_comp * nk4; p2 += C_WARPS) {
                        const int c = nk4 == 8 ? (p2 >> 3) : p2 / nk4, k4l = nk4 == 8 ? (p2 & 7) : p2 % nk4;
                        if (lane < n_act) {
                            const int kk = 32 * kc + 4 * k4l;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (kk < HA) v = *reinterpret_cast<const float4*>(raw + (c * NB + lane) * HA + kk);
                            float4 lo;
                            lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                            lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                            lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                            lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                            const int m = c * 32 + lane;
                            const int off = (m >> 3) * 1024 + k4l * 128 + (m & 7) * 16;
                            *reinterpret_cast<float4*>(hi_t + off) = v;       // the tensor core ignores the low 13 mantissa bits
                            *reinterpret_cast<float4*>(lo_t + off) = lo;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tile_full[ts]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&chain_free[cb]);       // raw rows consumed (the W tile is released by the issuer's commit)
            }
        }
    } else if (warp == W_M) {
        // =============================================================== M: MMA issuer
        // One thread issues ~33 000 small MMAs per CTA: its instruction count per MMA is what paces the kernel (the first
        // version rebuilt four 64-bit descriptors per k-step and ran at 6 300 cycles per chain).  Descriptors are now
        // advanced by adding the k-step's constant to their low word, and the three products of a k-step go out in one asm block.
        if (lane == 0) {
            int G = 0, T = 0;
            const uint32_t sbo_w = (uint32_t)(KP / 4) * 128u;
            uint64_t d_ahi[NT_STAGES], d_alo[NT_STAGES];
#pragma unroll
            for (int s = 0; s < NT_STAGES; ++s) {
                d_ahi[s] = tc::make_desc_sbo(smem_u32(smraw + L.ring + s * STAGE_BYTES), 1024);
                d_alo[s] = tc::make_desc_sbo(smem_u32(smraw + L.ring + s * STAGE_BYTES + TILE_BYTES), 1024);
            }
            auto kstep3 = [](uint32_t dacc, uint64_t ah, uint64_t al, uint64_t wh, uint64_t wl, uint32_t idesc, uint32_t accf) {
                asm volatile(
                    "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 q, 0, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %3, %5, p;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %4, %5, q;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %3, %5, q;\n\t}\n" ::"r"(dacc),
                    "l"(ah), "l"(al), "l"(wh), "l"(wl), "r"(idesc), "r"(accf)
                    : "memory");
            };
#pragma unroll 1
            for (int k = 0; k < n_active; ++k) {
                const int set = k & 1;
                if (k >= 2) mbar_wait_ws(&acc_free[set], ((k >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator set
#pragma unroll 1
                for (int g = 0; g < n_chains; ++g, ++G) {
                    const Chain ch = chains[g];
                    const int cb = G % NCB;
                    mbar_wait_ws(&chain_full[cb], (G / NCB) & 1);
                    const uint32_t w_hi = smem_u32(smraw + L.chain_stage[cb] + ch.n_comp * NB * HA * 4);
                    uint64_t dwh = tc::make_desc_sbo(w_hi, (int)sbo_w), dwl = tc::make_desc_sbo(w_hi + (uint32_t)(ch.npad * KP * 4), (int)sbo_w);
                    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ch.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                    const uint32_t dacc = tmem_base + (uint32_t)(set * SET_COLS + ch.acc_col);
                    uint32_t accf = ch.first ? 0u : 1u;
#pragma unroll 1
                    for (int kc = 0; kc < n_kc; ++kc, ++T) {
                        const int ts = T % NT_STAGES;
                        mbar_wait_ws(&tile_full[ts], (T / NT_STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t ah = d_ahi[ts], al = d_alo[ts];
                        if (KP - 32 * kc >= 32) {      // a full stage: 4 k-steps (a k-step = two 16-byte K chunks = +16 in descriptor units)
                            kstep3(dacc, ah, al, dwh, dwl, idesc, accf);
                            kstep3(dacc, ah + 16, al + 16, dwh + 16, dwl + 16, idesc, 1u);
                            kstep3(dacc, ah + 32, al + 32, dwh + 32, dwl + 32, idesc, 1u);
                            kstep3(dacc, ah + 48, al + 48, dwh + 48, dwl + 48, idesc, 1u);
                            dwh += 64; dwl += 64;
                        } else {
                            const int ksteps = (KP - 32 * kc) / 8;
#pragma unroll 1
                            for (int ks = 0; ks < ksteps; ++ks) {
                                kstep3(dacc, ah + 16 * ks, al + 16 * ks, dwh, dwl, idesc, ks == 0 ? accf : 1u);
                                dwh += 16; dwl += 16;
                            }
                        }
                        accf = 1u;
                        umma_commit(&tile_free[ts]);
                    }
                    umma_commit(&chain_free[cb]);
                }
                umma_commit(&acc_full[set]);
            }
        }
    } else {
        // =============================================================== E: epilogue warps (component c = TMEM lane group c)
        const int c = warp - W_E0;               // W_E0 is a multiple of 4?  no: the lane group is warp % 4 (see below)
        const int lg = warp & 3;
        (void)c;
#pragma unroll 1
        for (int k = 0; k < n_active; ++k) {
            const int set = k & 1;
            mbar_wait_ws(&acc_full[set], (k >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int n_act = n_items_s[k];
            const int node = lane < n_act ? node_of[k * NB + lane] : -1;     // lane = node rank
#pragma unroll 1
            for (int b = 0; b < a.n_blocks; ++b) {
                const Block bl = blocks[b];
                if (lg >= bl.n_comp) continue;       // warp-uniform
                float sum[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[i] = 0.0f;
#pragma unroll 1
                for (int p = 0; p < bl.n_partials; ++p) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(set * SET_COLS + bl.acc_col0 + p * bl.npad);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr));
                    if (bl.npad > 16)
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                            : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                            : "r"(taddr + 16u));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < bl.npad) sum[i] += __uint_as_float(v[i]);
                }
                if (node >= 0) {
                    float* dst = outacc + node * d_out + bl.out_base[lg];
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < bl.mul) dst[i * bl.out_step] += sum[i];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[set]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * SET_COLS));
    // ---- epilogue: mean over all incoming edges, BatchNorm (eval) affine, residual (same arithmetic as tp_transform_kernel)
#pragma unroll 1
    for (int i = tid; i < NB * d_out; i += THREADS_TT) {
        const int n = i / d_out, o = i - n * d_out;
        const int node = t0 + n;
        if (node < a.node_end) {
            float v = outacc[i];
            if (!(a.flags & CB_TP_RAW_SUM)) {
                int deg = deg_tot[n];
                if (a.pre_sum != nullptr && node >= a.pre_n0 && node < a.pre_n1) {
                    const int kk = (node - a.pre_n0) % a.pre_period;
                    v += __ldg(a.pre_sum + (size_t)kk * d_out + o);
                    deg += __ldg(a.pre_deg + kk);
                }
                v = v / (float)max(deg, 1);
                if (a.bn_scale) v = fmaf(v, a.bn_scale[o], a.bn_shift[o]);
                if (a.residual && o < a.d_res) v += a.residual[(size_t)node * a.ld_res + o];
            }
            a.out[(size_t)node * d_out + o] = v;
        }
    }
}

}  // namespace tt
