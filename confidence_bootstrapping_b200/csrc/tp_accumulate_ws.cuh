// K3 (a''): warp-specialised accumulate kernel -- included inside namespace tc of tp_conv.cu.
//
// Same arithmetic as tp_accumulate_tc_kernel<true> (transposed accumulator D[h~ unit][f-row] = H~ . F^T in TMEM, 3xTF32,
// hidden layer of the radial MLP on the tensor core, column H of the workspace summed by the f-row threads), but the chunk
// pipeline is cut into ROLES that run concurrently on different warps of ONE CTA per SM and hand work over through
// mbarriers only -- no CTA-wide barrier per 16 edges (those were 25 % of the warp-stall samples of the lock-step kernel,
// profiles/r2/k3_ncu_summary_mid.txt), the chunk iterator and the operand gather run ahead on their own warp, and two
// TMEM accumulators let the epilogue of item i overlap the chunks of item i+1:
//
//   warps 0-7    F  f-rows             : CG products of the gathered features -> F^T operand tile (hi/lo), buffer c & 1
//   warps 8-15   H  hidden layer       : e_attr split -> E tile; after the hidden-layer MMA: pre-activations back from TMEM
//                                        (two warps per lane group, 8 edges each) + node / neighbour projections, ReLU,
//                                        hi/lo split -> H~ operand tile, buffer c & 1
//   warps 16-19  E  epilogue           : finished accumulator TMEM -> registers -> coalesced global stores (workspace)
//   warp 20      S  scheduler          : walks the CTA's (node, slot) items; per chunk of 16 edges a descriptor + the 16
//                                        neighbour indices into stage c % NS (a single warp issuing the whole gather was the
//                                        bottleneck of the first version: 1300 instructions per chunk at one warp's issue rate)
//   warps 21-24  G  gather             : cp.async the raw operands (x[col], sh, e_attr, P_nbr[col], P_agg[node]) of the chunk,
//                                        four edges per warp, published two chunks later when the copies have landed
//   warp 25      I  MMA issuer         : ONE thread issues every tcgen05.mma of the CTA.  The H warps publish the E tile of
//                                        chunk c+1 BEFORE they await the pre-activations of chunk c, so the tensor pipe runs
//                                        hidden(c+1) ahead of main(c) and the H~ tile of chunk c+1 is produced while main(c)
//                                        executes (v2 served hidden(c+1) after main(c): the H warps then waited 36 % of their
//                                        time for pre-activations queued behind a 720-cycle rank-16 update).  (In v1 hidden unit 0 issued the MMAs: its timeline -- tile production AND
//                                        the issue back-pressure of the tensor pipe -- was 82 % busy and paced the whole CTA,
//                                        profiles/r2/phase_ws_v1.log)
//
//   barrier           producer -> consumer            count
//   desc_full[s]      S (descriptor)   -> G           1
//   raw_full[s]       G (data landed)  -> F, H        4
//   raw_empty[s]      F, H (stage read) -> S          16 (one lane per warp)
//   f_full[b]         F (tile written) -> I           8
//   f_free[b]         tcgen05.commit   -> F           1
//   e_full[b]         H (E tile)       -> I           8   (ictl_e[b] = 0 marks the end of the work)
//   hid_done[b]       tcgen05.commit   -> H           1
//   h_full[b]         H (H~ tile)      -> I           8
//   h_free[b]         tcgen05.commit   -> H           1
//   acc_full[a]       tcgen05.commit   -> E           1   (or a plain arrive carrying the end-of-work sentinel)
//   acc_empty[a]      E (TMEM read)    -> I           4
//
// TMEM (512 columns, one CTA per SM): accumulator a at column 240 a, hidden pre-activations of chunk k at 480 + 16 (k & 1).
// Every wait is a bounded spin that traps instead of hanging the GPU.

namespace ws {

#ifdef CB_PHASE_TIMING
#define WS_T0() long long wph[12] = {0,0,0,0,0,0,0,0,0,0,0,0}, wt = clock64(); long long wcount = 0
#define WS_MARK(k) do { const long long t_ = clock64(); wph[k] += t_ - wt; wt = t_; } while (0)
#define WS_FLUSH(row) do { for (int k_ = 0; k_ < 12; ++k_) atomicAdd(&cb_dbg_phase[row][k_], (unsigned long long)wph[k_]); } while (0)
#else
#define WS_T0() do { } while (0)
#define WS_MARK(k) do { } while (0)
#define WS_FLUSH(row) do { } while (0)
#endif

constexpr int NS = 6;                 // raw-operand stages the scheduler may run ahead
constexpr int LAG = 2;                // a chunk's gather is published LAG chunks after it was issued (its copies have landed)
constexpr int NI = 8;                 // item-descriptor ring (>= NS + 2 items can be open between S and E)
constexpr int F_WARPS = 8, H_WARPS = 8, E_WARPS = 4, G_WARPS = 4;
// warp order: F | H | E | S | G... | I
constexpr int W_S = F_WARPS + H_WARPS + E_WARPS, W_G0 = W_S + 1, W_I = W_G0 + G_WARPS;
constexpr int THREADS_WS = 32 * (W_I + 1);
constexpr int ACC_COLS = 240;
constexpr int TMEM_COLS_WS = 512;

struct ChunkDesc { int valid, seg, n, flags, item_seq, node, q, pad; };     // flags: 1 = first chunk of its item, 2 = last
struct ItemDesc { unsigned long long ws_off; int row_stride, valid; };

struct LayoutWS {
    int fhi[2], flo[2], hhi[2], hlo[2], ehi[2], elo[2], w1hi, w1lo, raw, rows, terms, cdesc, idesc, total;   // byte offsets
    int raw_stage, o_xs, o_shs, o_es, o_ps, o_pa;     // bytes per raw stage and offsets inside it
    int dxp, sbow, f_groups;
};

__host__ __device__ inline LayoutWS make_layout_ws(int n_rows, int n_terms, int ne, int d_in, int S, int H) {
    LayoutWS L;
    auto al = [](int v, int a) { return (v + a - 1) / a * a; };
    L.sbow = (ne / 4) * LBO;
    L.dxp = d_in + (d_in & 1);
    L.f_groups = (n_rows + 15) / 16 * 2;
    int o = 0;
    for (int b = 0; b < 2; ++b) { L.fhi[b] = o; o += L.f_groups * SBO; L.flo[b] = o; o += L.f_groups * SBO; }
    for (int b = 0; b < 2; ++b) { L.hhi[b] = o; o += 16 * SBO; L.hlo[b] = o; o += 16 * SBO; }
    for (int b = 0; b < 2; ++b) { L.ehi[b] = o; o += (KC / 8) * L.sbow; L.elo[b] = o; o += (KC / 8) * L.sbow; }
    L.w1hi = o; o += 16 * L.sbow;          // full 128 rows: rows >= H stay zero (nothing may follow that a stale read could hit)
    L.w1lo = o; o += 16 * L.sbow;
    L.o_xs = 64;                           // the stage starts with the chunk's 16 neighbour indices
    L.o_shs = L.o_xs + al(KC * L.dxp * 4, 16);
    L.o_es = L.o_shs + al(KC * S * 4, 16);
    L.o_ps = L.o_es + al(KC * ne * 4, 16);
    L.o_pa = L.o_ps + al(KC * H * 4, 16);          // P_agg[node] row of the item (first chunk of an item only)
    L.raw_stage = L.o_pa + al(H * 4, 16);
    L.raw = o;  o += NS * L.raw_stage;
    L.rows = o; o += al(n_rows * 32, 16);
    L.terms = o; o += al(n_terms * 8, 16);
    L.cdesc = o; o += NS * (int)sizeof(ChunkDesc);
    L.idesc = o; o += NI * (int)sizeof(ItemDesc);
    L.total = o;
    return L;
}

constexpr uint32_t WS_SUSPEND_NS = 100000u;      // per try_wait; a lost arrival traps after WS_SUSPEND_TRIES failed waits
constexpr uint32_t WS_SUSPEND_TRIES = 4000000u;   // >= 1.5 s even if the hardware caps the suspension at a few hundred ns

// bounded wait: a lost arrival traps after ~1.5 s of spinning instead of hanging the GPU (the kernel itself runs ~1 ms)
// (An out-of-line copy and a sleep between polls were both measured slower.)
__device__ __forceinline__ void mbar_wait_ws(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#ifndef CB_WS_SUSPEND_WAIT
    const long long t0 = clock64();
#pragma unroll 1
    for (unsigned spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (ok) return;
        if ((spin & 255u) == 255u && clock64() - t0 > 3000000000ll) __trap();
    }
#else
    // experiment (-DCB_WS_SUSPEND_WAIT): try_wait with a suspend-time hint, so that the hardware parks the warp instead of
    // re-issuing the poll (the spin loops are 70 % of the executed instructions of this kernel).  Measured: no change
    // (conv layer 1: 1.047 ms vs 1.055 ms) -- the polls do not take issue slots anybody else wants.
#pragma unroll 1
    for (unsigned spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity), "r"(WS_SUSPEND_NS)
                     : "memory");
        if (ok) return;
        if (spin > WS_SUSPEND_TRIES) __trap();
    }
#endif
}

// UMMA shared-memory descriptor of a K-major tile in a SWIZZLED layout (rows of 64 or 128 bytes, 8-row atoms, 16-byte chunks
// XOR-ed with the row index): start >> 4 | LBO (unused for K-major swizzled: 1) << 16 | SBO >> 4 << 32 | version 1 << 46 | layout << 61
// (layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B).  The tensor core reads a no-swizzle tile one 128-byte core matrix at a time --
// measured ~100 cycles per M = 128 MMA whatever N -- while a swizzled row is fetched at full shared-memory width.
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, int sbo_bytes, int layout) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__global__ void __launch_bounds__(THREADS_WS, 1)
tp_accumulate_ws_kernel(const __grid_constant__ cb_tp_conv_args a, int n_items) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int H = a.H, ne = a.ne, S = a.S, d_in = a.d_in, n_rows = a.n_rows;
    const LayoutWS L = make_layout_ws(n_rows, a.n_terms, ne, d_in, S, H);
    const int dxp = L.dxp, SBOW = L.sbow, HA = H + PADC;
    const int NRP = (n_rows + 15) & ~15;
    cb_tp_row* rows_s = reinterpret_cast<cb_tp_row*>(smraw + L.rows);
    cb_tp_term* terms_s = reinterpret_cast<cb_tp_term*>(smraw + L.terms);
    ChunkDesc* cdesc = reinterpret_cast<ChunkDesc*>(smraw + L.cdesc);
    ItemDesc* idesc_ring = reinterpret_cast<ItemDesc*>(smraw + L.idesc);
    __shared__ SlotTable st;
    __shared__ __align__(8) uint64_t desc_full[NS], raw_full[NS], raw_empty[NS], f_full[2], f_free[2], h_full[2], h_free[2], e_full[2], hid_done[2], acc_full[2], acc_empty[2];
    __shared__ int ictl_e[2];                    // H -> I: does chunk k (buffer k & 1) exist?
    __shared__ int4 ictl_m[2];                   // H -> I: (flags, item_seq, n) of the chunk whose H~ tile sits in buffer k & 1
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- once per CTA
    if (tid == 0) {
        build_slots(a, st);
        for (int s = 0; s < NS; ++s) { mbar_init(&desc_full[s], 1); mbar_init(&raw_full[s], G_WARPS); mbar_init(&raw_empty[s], F_WARPS + H_WARPS); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&f_full[b], F_WARPS); mbar_init(&f_free[b], 1); mbar_init(&h_full[b], H_WARPS); mbar_init(&h_free[b], 1); mbar_init(&e_full[b], H_WARPS); mbar_init(&hid_done[b], 1);
            mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], E_WARPS);
        }

        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS_WS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    {
        const int* src = reinterpret_cast<const int*>(a.rows);
        int* dst = reinterpret_cast<int*>(rows_s);
#pragma unroll 1
        for (int i = tid; i < n_rows * 8; i += THREADS_WS) dst[i] = src[i];
        const int2* tsrc = reinterpret_cast<const int2*>(a.terms);
        int2* tdst = reinterpret_cast<int2*>(terms_s);
#pragma unroll 1
        for (int i = tid; i < a.n_terms; i += THREADS_WS) tdst[i] = tsrc[i];
        // operand tiles start as zeros: padding rows (f-rows >= n_rows, hidden units >= H) are never written again
#pragma unroll 1
        for (int i = tid; i < L.raw / 16; i += THREADS_WS) reinterpret_cast<float4*>(smraw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tmem_hid = tmem_base + 2u * ACC_COLS;

    auto stage_ptr = [&](int s) { return smraw + L.raw + s * L.raw_stage; };

    if (warp == W_S) {
        // =============================================================== S: scheduler (one warp, warp-uniform control)
        auto seg_range = [&](int seg, int node, int& e0, int& e1) { seg_edges(a.segs[seg], node, e0, e1); };
        int item = blockIdx.x, q = 0, item_seq = -1;
        int c = 0;                   // chunk counter of this CTA
        bool done = false;
        WS_T0();
#pragma unroll 1
        while (!done) {
            // ---- next item with at least one edge
            int node = 0, seg = 0, e0 = 0, e1 = 0;
            bool found = false;
#pragma unroll 1
            for (; item < n_items && !found; ) {
                while (q + 1 < st.n_slots && item >= st.item_off[q + 1]) ++q;
                node = st.lo[q] + (item - st.item_off[q]);
#pragma unroll 1
                for (int sg = st.first_seg[q]; sg < st.first_seg[q] + st.n_segs[q]; ++sg) {
                    seg_range(sg, node, e0, e1);
                    if (e1 > e0) { seg = sg; found = true; break; }
                }
                if (!found) item += (int)gridDim.x;
            }
            if (!found) {
                done = true;
            } else {
                ++item_seq;
                // where the finished tile goes: written to the item ring before the item's first chunk is published
                int ws_stride;
                const size_t ws_off = ws_place(a, st, q, node, n_rows, HA, lane, ws_stride);
                if (lane == 0) {
                    ItemDesc d; d.ws_off = (unsigned long long)ws_off; d.row_stride = ws_stride; d.valid = 1;
                    idesc_ring[item_seq % NI] = d;
                }
            }
            // ---- the item's chunks (or the end-of-work sentinel)
            bool first = true;
#pragma unroll 1
            while (true) {
                const int s = c % NS;
                WS_MARK(1);
                mbar_wait_ws(&raw_empty[s], ((c / NS) & 1) ^ 1);
                WS_MARK(0);
                ChunkDesc d;
                d.valid = done ? 0 : 1; d.seg = seg; d.item_seq = item_seq; d.node = node; d.q = q; d.pad = e0;   // pad = first edge
                bool last = true;
                d.n = 0; d.flags = 0;
                if (!done) {
                    const int n = min(KC, e1 - e0);
                    // more edges of this item after this chunk?
                    int nseg = seg, ne0 = e0 + KC, ne1 = e1;
                    bool more = ne0 < ne1;
                    if (!more) {
#pragma unroll 1
                        for (int sg = seg + 1; sg < st.first_seg[q] + st.n_segs[q]; ++sg) {
                            int b0, b1;
                            seg_range(sg, node, b0, b1);
                            if (b1 > b0) { nseg = sg; ne0 = b0; ne1 = b1; more = true; break; }
                        }
                    }
                    last = !more;
                    d.n = n; d.flags = (first ? 1 : 0) | (last ? 2 : 0);
                    const cb_tp_segment& sg = a.segs[seg];
                    if (lane < KC) reinterpret_cast<int*>(stage_ptr(s))[lane] = lane < n ? __ldg(sg.col + e0 + lane) + sg.col_off : 0;
                    seg = nseg; e0 = ne0; e1 = ne1;
                }
                if (lane == 0) cdesc[s] = d;
                __syncwarp();
                if (lane == 0) mbar_arrive(&desc_full[s]);
                ++c;
                first = false;
                if (done || last) break;
            }
            if (!done) item += (int)gridDim.x;
        }
        if (lane == 0) WS_FLUSH(3);
    } else if (warp >= W_G0 && warp < W_I) {
        // =============================================================== G: gather warps (edges g, g+4, g+8, g+12 of every chunk)
        const int g = warp - W_G0;
        int pending = 0;
        int c = 0;
        WS_T0();
#pragma unroll 1
        for (;; ++c) {
            const int s = c % NS;
            WS_MARK(2);
            mbar_wait_ws(&desc_full[s], (c / NS) & 1);
            WS_MARK(0);
            const ChunkDesc d = cdesc[s];
            if (d.valid) {
                const cb_tp_segment& sg = a.segs[d.seg];
                unsigned char* stg = stage_ptr(s);
                const int* cols = reinterpret_cast<const int*>(stg);
                float* xs = reinterpret_cast<float*>(stg + L.o_xs);
                float* shs = reinterpret_cast<float*>(stg + L.o_shs);
                float* es = reinterpret_cast<float*>(stg + L.o_es);
                float* ps = reinterpret_cast<float*>(stg + L.o_ps);
                const int n = d.n, base = d.pad;
#pragma unroll 1
                for (int e = g; e < n; e += G_WARPS) {
                    const int col = cols[e];
                    const float* xr = a.x + (size_t)col * d_in;
                    float* xd = xs + e * dxp;
#pragma unroll 1
                    for (int k = lane; k < d_in / 2; k += 32) cp_async_bytes8(xd + 2 * k, xr + 2 * k);
                    if (sg.P_nbr) {
                        const float* pr = sg.P_nbr + (size_t)col * sg.ldp_nbr;
#pragma unroll 1
                        for (int k = lane; k < H / 4; k += 32) cp_async_bytes16(ps + e * H + 4 * k, pr + 4 * k);
                    }
                }
                if (g == G_WARPS - 1 && (d.flags & 1)) {      // per-item constant of the hidden layer: the aggregation node's projection row
                    const cb_tp_segment& s0 = a.segs[st.first_seg[d.q]];
                    if (s0.P_agg) {
                        float* pa = reinterpret_cast<float*>(stg + L.o_pa);
                        const float* pr = s0.P_agg + (size_t)d.node * s0.ldp_agg;
#pragma unroll 1
                        for (int k = lane; k < H / 4; k += 32) cp_async_bytes16(pa + 4 * k, pr + 4 * k);
                    }
                }
#pragma unroll 1
                for (int i = g * 32 + lane; i < n * S; i += 32 * G_WARPS) cp_async_bytes4(shs + i, sg.sh + (size_t)base * S + i);
#pragma unroll 1
                for (int i = g * 32 + lane; i < n * (ne / 4); i += 32 * G_WARPS) cp_async_bytes16(es + 4 * i, sg.e_attr + (size_t)base * ne + 4 * i);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            WS_MARK(1);
            ++pending;
            // publish the chunk issued LAG iterations ago: this warp's copies for it have landed
            if (pending > LAG) {
                asm volatile("cp.async.wait_group %0;\n" ::"n"(LAG) : "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_full[(c - LAG) % NS]);
                --pending;
            }
            if (!d.valid) break;
        }
        // drain: publish what is still pending (the sentinel included)
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
#pragma unroll 1
        for (; pending > 0; --pending)
            if (lane == 0) mbar_arrive(&raw_full[(c + 1 - pending) % NS]);
        if (g == 0 && lane == 0) WS_FLUSH(4);
    } else if (warp < F_WARPS) {
        // =============================================================== F: f-rows -> F^T operand tiles
        int tb0 = 0, te0 = 0, t_xi[MAXT], t_si[MAXT];
        float t_cf[MAXT];
        if (tid < n_rows) { tb0 = rows_s[tid].term_begin; te0 = rows_s[tid].term_end; }
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            const bool on = tid < n_rows && tb0 + t < te0;
            const cb_tp_term tm = on ? terms_s[tb0 + t] : cb_tp_term{0, 0, 0.0f};
            t_xi[t] = tm.x_idx; t_si[t] = tm.sh_idx; t_cf[t] = on ? tm.coef : 0.0f;
        }
        const int nt_warp = __reduce_max_sync(0xffffffffu, min(te0 - tb0, MAXT));
        float fsum = 0.0f;
        WS_T0();
#pragma unroll 1
        for (int c = 0;; ++c) {
            const int s = c % NS, fb = c & 1;
            WS_MARK(3);
            mbar_wait_ws(&raw_full[s], (c / NS) & 1);
            WS_MARK(0);
            const ChunkDesc d = cdesc[s];
            if (!d.valid) break;
            mbar_wait_ws(&f_free[fb], ((c >> 1) & 1) ^ 1);
            WS_MARK(1);         // the MMAs that read this tile buffer two chunks ago are done
            const unsigned char* stg = stage_ptr(s);
            const float* xs = reinterpret_cast<const float*>(stg + L.o_xs);
            const float* shs = reinterpret_cast<const float*>(stg + L.o_shs);
            const int n = d.n, nq = 2 * ((n + 7) >> 3);
            if (tid < n_rows) {
                const FRowCtx fc{xs, shs, terms_s, smraw + L.fhi[fb], smraw + L.flo[fb], dxp, S, n, nq, tb0, te0, (tid >> 3) * 512 + (tid & 7) * 64,
                                 16, (tid & 7) >> 1};       // SWIZZLE_64B operand tile
                float part;
                switch (nt_warp) {
                    case 1: part = f_row<1>(fc, t_xi, t_si, t_cf); break;
                    case 2: part = f_row<2>(fc, t_xi, t_si, t_cf); break;
                    case 3: part = f_row<3>(fc, t_xi, t_si, t_cf); break;
                    default: part = f_row<MAXT>(fc, t_xi, t_si, t_cf); break;
                }
                fsum = (d.flags & 1) ? part : fsum + part;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy tile writes -> async proxy (UMMA)
            __syncwarp();
            if (lane == 0) { mbar_arrive(&f_full[fb]); mbar_arrive(&raw_empty[s]); }
            WS_MARK(2);
            if ((d.flags & 2) && tid < n_rows) {       // column H of the workspace: sum_e f_e[r], then the zero pad the transform multiplies by 0
                const ItemDesc it = idesc_ring[d.item_seq % NI];
                *reinterpret_cast<float4*>(a.workspace + it.ws_off + (size_t)tid * it.row_stride + H) = make_float4(fsum, 0.f, 0.f, 0.f);
            }
        }
        if (tid == 0) WS_FLUSH(2);
        if (tid == 224) WS_FLUSH(7);
    } else if (warp < F_WARPS + H_WARPS) {
        // =============================================================== H: E tile (one chunk ahead), then pre-activations -> H~ tile
        const int hw = warp - F_WARPS;              // 0..7
        const int lg = warp & 3;                    // TMEM lane group (warp 8 is lane group 0)
        const int eh = hw >> 2;                     // which 8 of the chunk's 16 edges
        const int q_unit = lg * 32 + lane;          // hidden unit = TMEM lane
        const int ht = tid - 32 * F_WARPS;          // 0..255 within the role
        unsigned char* W1hi = smraw + L.w1hi; unsigned char* W1lo = smraw + L.w1lo;
        int staged_slot = -1, hb_q = -1, hb_graph = -1;
        float hb_const = 0.0f;
        WS_T0();
        // E-phase of chunk k (descriptor d, stage s): per-item constants when the chunk opens an item, then the chunk's
        // edge-embedding rows as hi/lo B tiles of the hidden-layer MMA (buffer k & 1).  Returns the item's hidden-layer bias.
        auto e_phase = [&](int k, const ChunkDesc& d, int s, float hb_prev) -> float {
            const unsigned char* stg = stage_ptr(s);
            const float* es = reinterpret_cast<const float*>(stg + L.o_es);
            float hb = hb_prev;
            if (d.flags & 1) {
                const cb_tp_segment& s0 = a.segs[st.first_seg[d.q]];
                if (staged_slot != d.q) {       // only reached when no hidden-layer MMA is in flight (see the loop below)
#pragma unroll 1
                    for (int i = ht; i < H * (ne / 4); i += 32 * H_WARPS) {
                        const int qq = i / (ne / 4), c4 = i - qq * (ne / 4);
                        const float4 w = __ldg(reinterpret_cast<const float4*>(s0.W1e + (size_t)qq * s0.ldw1) + c4);
                        float4 hi, lo;
                        split_tf32(w.x, hi.x, lo.x); split_tf32(w.y, hi.y, lo.y); split_tf32(w.z, hi.z, lo.z); split_tf32(w.w, hi.w, lo.w);
                        const int off = (qq >> 3) * 1024 + (qq & 7) * 128 + ((c4 ^ (qq & 7)) << 4);      // SWIZZLE_128B (ne = 32: 128-byte rows)
                        *reinterpret_cast<float4*>(W1hi + off) = hi;
                        *reinterpret_cast<float4*>(W1lo + off) = lo;
                    }
                    staged_slot = d.q;
                }
                if (q_unit < H) {
                    const int graph = (s0.e_post && a.agg_graph) ? __ldg(a.agg_graph + d.node) : 0;
                    if (hb_q != d.q || (s0.e_post && hb_graph != graph)) {
                        float v = __ldg(s0.b1 + q_unit);
                        if (s0.e_post) {
                            const float* ep = s0.e_post + (size_t)graph * ne;
#pragma unroll 4
                            for (int kk = 0; kk < ne; ++kk) v = fmaf(__ldg(s0.W1e + (size_t)q_unit * s0.ldw1 + kk), __ldg(ep + kk), v);
                        }
                        hb_const = v; hb_q = d.q; hb_graph = graph;
                    }
                    hb = hb_const + (s0.P_agg ? reinterpret_cast<const float*>(stg + L.o_pa)[q_unit] : 0.0f);   // row gathered with the chunk
                }
            }
            unsigned char* Ehi = smraw + L.ehi[k & 1]; unsigned char* Elo = smraw + L.elo[k & 1];
#pragma unroll 1
            for (int i = ht; i < KC * (ne / 4); i += 32 * H_WARPS) {
                const int e = i / (ne / 4), c4 = i - e * (ne / 4);
                const float4 v = *reinterpret_cast<const float4*>(es + e * ne + 4 * c4);     // rows >= n hold stale data: discarded later
                float4 hi, lo;
                split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
                const int off = (e >> 3) * 1024 + (e & 7) * 128 + ((c4 ^ (e & 7)) << 4);             // SWIZZLE_128B
                *reinterpret_cast<float4*>(Ehi + off) = hi;
                *reinterpret_cast<float4*>(Elo + off) = lo;
            }
            if (ht == 0) ictl_e[k & 1] = 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&e_full[k & 1]);
            return hb;
        };
        auto end_of_work = [&](int k) {          // tell the issuer that chunk k does not exist
            if (ht == 0) ictl_e[k & 1] = 0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&e_full[k & 1]);
        };
        mbar_wait_ws(&raw_full[0], 0);
        ChunkDesc d = cdesc[0];
        float hb = 0.0f;
        if (!d.valid) {
            end_of_work(0);
        } else {
            hb = e_phase(0, d, 0, 0.0f);
#pragma unroll 1
            for (int c = 0;; ++c) {
                const int s = c % NS, b = c & 1, s1 = (c + 1) % NS;
                // ---- look ahead: the E tile of chunk c+1 goes out before this chunk's pre-activations are awaited, so the tensor
                // pipe runs hidden(c+1) BEFORE main(c) and the H~ tile of chunk c+1 is produced while main(c) executes.  Not when
                // chunk c+1 opens a new slot: re-staging the W1 tiles must wait for hidden(c) (done below, after the wait).
                mbar_wait_ws(&raw_full[s1], ((c + 1) / NS) & 1);
                WS_MARK(0);
                const ChunkDesc d1 = cdesc[s1];
                const bool early = d1.valid && !((d1.flags & 1) && staged_slot != d1.q);
                float hb1 = hb;
                if (early) hb1 = e_phase(c + 1, d1, s1, hb);
                WS_MARK(1);
                // ---- H-phase of chunk c
                const cb_tp_segment& sg = a.segs[d.seg];
                const float* Ps = reinterpret_cast<const float*>(stage_ptr(s) + L.o_ps);
                const int n = d.n, ksteps = (n + 7) >> 3;
                mbar_wait_ws(&hid_done[b], (c >> 1) & 1);
                WS_MARK(4);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v[8];
                if (eh < ksteps) {       // warp-uniform
                    const uint32_t taddr = tmem_hid + (uint32_t)(16 * b) + ((uint32_t)(lg * 32) << 16) + (uint32_t)(8 * eh);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                                 : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                WS_MARK(5);
                mbar_wait_ws(&h_free[b], ((c >> 1) & 1) ^ 1);        // the MMAs that read this H~ buffer two chunks ago are done
                WS_MARK(6);
                if (eh < ksteps && q_unit < H) {
                    unsigned char* Hhi = smraw + L.hhi[b]; unsigned char* Hlo = smraw + L.hlo[b];
                    const int rbase = (q_unit >> 3) * 512 + (q_unit & 7) * 64;          // SWIZZLE_64B: 16-byte chunk j of the row at (j ^ row bits) * 16
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        float h[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int e = 8 * eh + 4 * g + j;
                            float pre = __uint_as_float(v[4 * g + j]) + hb;
                            if (sg.P_nbr) pre += Ps[e * H + q_unit];
                            h[j] = e < n ? fmaxf(pre, 0.0f) : 0.0f;
                        }
                        float4 hi, lo;
                        split_tf32(h[0], hi.x, lo.x); split_tf32(h[1], hi.y, lo.y); split_tf32(h[2], hi.z, lo.z); split_tf32(h[3], hi.w, lo.w);
                        const int coff = rbase + (((2 * eh + g) ^ ((q_unit & 7) >> 1)) << 4);
                        *reinterpret_cast<float4*>(Hhi + coff) = hi;
                        *reinterpret_cast<float4*>(Hlo + coff) = lo;
                    }
                }
                if (ht == 0) ictl_m[b] = make_int4(d.flags, d.item_seq, d.n, 0);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) { mbar_arrive(&h_full[b]); mbar_arrive(&raw_empty[s]); }
                WS_MARK(7);
                if (!early) {
                    if (!d1.valid) { end_of_work(c + 1); break; }
                    hb1 = e_phase(c + 1, d1, s1, hb);      // new slot: hidden(c) has completed, the W1 tiles may be re-staged
                    WS_MARK(1);
                }
                hb = hb1;
                d = d1;
            }
        }
        if (ht == 0) WS_FLUSH(0);
        if (ht == 32) WS_FLUSH(1);
    } else if (warp == W_I) {
        // =============================================================== I: the CTA's MMA issuer (one thread)
        if (lane == 0) {
            const uint64_t d_w1hi = make_desc_sw(smem_u32(smraw + L.w1hi), 1024, 2), d_w1lo = make_desc_sw(smem_u32(smraw + L.w1lo), 1024, 2);
            const uint64_t d_e0h = make_desc_sw(smem_u32(smraw + L.ehi[0]), 1024, 2), d_e0l = make_desc_sw(smem_u32(smraw + L.elo[0]), 1024, 2);
            const uint64_t d_e1h = make_desc_sw(smem_u32(smraw + L.ehi[1]), 1024, 2), d_e1l = make_desc_sw(smem_u32(smraw + L.elo[1]), 1024, 2);
            const uint64_t d_f0h = make_desc_sw(smem_u32(smraw + L.fhi[0]), 512, 4), d_f0l = make_desc_sw(smem_u32(smraw + L.flo[0]), 512, 4);
            const uint64_t d_f1h = make_desc_sw(smem_u32(smraw + L.fhi[1]), 512, 4), d_f1l = make_desc_sw(smem_u32(smraw + L.flo[1]), 512, 4);
            const uint64_t d_h0h = make_desc_sw(smem_u32(smraw + L.hhi[0]), 512, 4), d_h0l = make_desc_sw(smem_u32(smraw + L.hlo[0]), 512, 4);
            const uint64_t d_h1h = make_desc_sw(smem_u32(smraw + L.hhi[1]), 512, 4), d_h1l = make_desc_sw(smem_u32(smraw + L.hlo[1]), 512, 4);
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NRP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_h = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // Two in-order queues -- hidden(k) waits for the E tile of chunk k, main(k) for the H~ and F^T tiles of chunk k -- are
            // polled and served in whichever order they become ready.  The interleaving only decides overlap: each queue writes
            // its own TMEM columns in a fixed order, so the sums stay bit-reproducible.
            int kh = 0, km = 0, end = 0x7fffffff, last_seq = -1;
            long long t0 = clock64();
            WS_T0();
#pragma unroll 1
            while (km < end) {
                bool progressed = false;
                WS_MARK(0);
                if (kh < end && mbar_test(&e_full[kh & 1], (kh >> 1) & 1)) {
                    asm volatile("fence.acq_rel.cta;" ::: "memory");
                    if (ictl_e[kh & 1] == 0) {
                        end = kh;
                    } else {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const int eb = kh & 1;
                        uint64_t dwh = d_w1hi, dwl = d_w1lo, deh = eb ? d_e1h : d_e0h, del = eb ? d_e1l : d_e0l;
                        const uint32_t dh = tmem_hid + (uint32_t)(16 * eb);
#pragma unroll 2
                        for (int ks = 0; ks < ne / 8; ++ks) {
                            mma_tf32(dh, dwh, deh, idesc_h, ks > 0 ? 1u : 0u);
                            mma_tf32(dh, dwh, del, idesc_h, 1u);
                            mma_tf32(dh, dwl, deh, idesc_h, 1u);
                            dwh += 2; dwl += 2; deh += 2; del += 2;      // next k-step: 32 bytes along the swizzled row
                        }
                        umma_commit(&hid_done[eb]);
                        ++kh;
                    }
                    progressed = true;
                    WS_MARK(1);
                }
                if (km < kh && km < end && mbar_test(&h_full[km & 1], (km >> 1) & 1) && mbar_test(&f_full[km & 1], (km >> 1) & 1)) {
                    asm volatile("fence.acq_rel.cta;" ::: "memory");
                    const int b = km & 1;
                    const int4 ctl = ictl_m[b];          // (flags, item_seq, n)
                    const int acc = ctl.y & 1;
                    WS_MARK(0);
                    if (ctl.x & 1) mbar_wait_ws(&acc_empty[acc], ((ctl.y >> 1) & 1) ^ 1);   // the epilogue of item_seq - 2 has drained it
                    WS_MARK(2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t dacc = tmem_base + (uint32_t)(acc * ACC_COLS);
                    const uint32_t acc0 = (ctl.x & 1) ? 0u : 1u;
                    const uint64_t fh = b ? d_f1h : d_f0h, fl = b ? d_f1l : d_f0l, hh = b ? d_h1h : d_h0h, hl = b ? d_h1l : d_h0l;
                    mma_tf32(dacc, hh, fh, idesc, acc0);
                    mma_tf32(dacc, hh, fl, idesc, 1u);
                    mma_tf32(dacc, hl, fh, idesc, 1u);
                    if (((ctl.z + 7) >> 3) > 1) {
                        const uint64_t ks = 2;      // second k-step (edges 8-15): 32 bytes along the swizzled row
                        mma_tf32(dacc, hh + ks, fh + ks, idesc, 1u);
                        mma_tf32(dacc, hh + ks, fl + ks, idesc, 1u);
                        mma_tf32(dacc, hl + ks, fh + ks, idesc, 1u);
                    }
                    umma_commit(&f_free[b]);
                    umma_commit(&h_free[b]);
                    if (ctl.x & 2) umma_commit(&acc_full[acc]);
                    last_seq = ctl.y;
                    ++km;
                    progressed = true;
                    WS_MARK(3);
                }
                if (progressed) t0 = clock64();
                else if (clock64() - t0 > 6000000000ll) __trap();      // bounded: ~3 s without progress
            }
            // end of work: hand the sentinel to the epilogue warps through the next accumulator's barrier, with the flow control of
            // a real item (the epilogue must have consumed the previous phase of that barrier)
            const int seq = last_seq + 1;
            mbar_wait_ws(&acc_empty[seq & 1], ((seq >> 1) & 1) ^ 1);
            ItemDesc it; it.ws_off = 0; it.row_stride = 0; it.valid = 0;
            idesc_ring[seq % NI] = it;
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            mbar_arrive(&acc_full[seq & 1]);
#ifdef CB_PHASE_TIMING
            wph[4] = km;
#endif
            WS_FLUSH(5);
        }
    } else {
        // =============================================================== E: epilogue (TMEM -> registers -> workspace)
        const int lg = warp & 3;                    // warps 16..19 are lane groups 0..3
        const int j = lg * 32 + lane;
        const int c_lo = 0, c_hi = NRP;        // one warp per lane group: all f-row columns
        WS_T0();
#pragma unroll 1
        for (int item_seq = 0;; ++item_seq) {
            const int acc = item_seq & 1;
            WS_MARK(1);
            mbar_wait_ws(&acc_full[acc], (item_seq >> 1) & 1);
            WS_MARK(0);
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            const ItemDesc it = idesc_ring[item_seq % NI];
            if (!it.valid) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lg * 32 < H) {                      // warp-uniform: lane groups beyond the hidden width hold nothing
                const size_t row_stride = (size_t)it.row_stride;
                float* dst = a.workspace + it.ws_off + j + (size_t)c_lo * row_stride;
#pragma unroll 1
                for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * ACC_COLS + c0);
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
                        const int nr = min(second ? 32 : 16, n_rows - c0);
                        float* p = dst;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
#ifndef CB_WS_NO_STORE
                            if (i < nr) *p = __uint_as_float(v[i]);
#else
                            if (i < nr && v[i] == 0x7fc12345u) *p = 1.0f;     // experiment: keep the TMEM reads, drop the stores
#endif
                            p += row_stride;
                        }
                    }
                    dst += 32 * row_stride;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
#ifdef CB_PHASE_TIMING
            wph[2] += 1;
#endif
        }
        if (warp == F_WARPS + H_WARPS && lane == 0) WS_FLUSH(6);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS_WS));
    }
}

}  // namespace ws
