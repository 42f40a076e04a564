// apx_ipattn_fused.cu — one adapter site in ONE tcgen05 kernel:
//
//   Y = ( softmax(Q K_t^T / sqrt(d)) V_t + ap_scale * softmax(Q K_a^T / sqrt(d)) V_a ) Wo^T + bo,   Q = X Wq^T
//
// i.e. the to_q GEMM, the two SDPAs, the ap_scale add and the to_out GEMM of
// APadapter/ap_adapter/attention_processor.py:387-459 with no intermediate ever leaving the SM.
//
// One CTA owns a 128-query tile of one batch element (and, when the fp32 output accumulator does not fit in
// tensor memory, one of `nsplit` column ranges of Y).  The X tile is TMA-staged once (SWIZZLE_128B) and stays
// resident; per head the CTA streams, through mbarrier rings fed by 1-D bulk TMA copies,
//   * the head's rows of Wq      (pre-packed K-major core-matrix chunks, 64 input channels each),
//   * the head's K and V^T tiles of BOTH contexts (the packed context cache of apx_ctx_kv),
//   * the head's columns of Wo   (pre-packed [d_pad/8][CN][8]),
// and runs, all with fp32 accumulators in tensor memory:
//   Q_h   = X · Wq_h^T                  tcgen05.mma 128 x d_pad x 16   (K = C)
//   S     = Q_h · [K_t ; K_a]^T         tcgen05.mma 128 x L_pad x 16   (K = d_pad), one tile for both contexts
//   P     = two row softmaxes           one query row per thread: tcgen05.ld -> exp2 -> 16-bit P in shared memory
//   O_t  += P_t · V_t,  O_a += P_a · V_a   tcgen05.mma 128 x d_pad x 16
//   O_h   = O_t / l_t + ap_scale · O_a / l_a     in registers (the ap_scale add costs nothing)
//   Y    += O_h · Wo[:, head]^T         tcgen05.mma 128 x CN x 16      (K = d_pad), accumulated over the heads
// Warp roles (192 threads): warps 0..3 = softmax / conversion / epilogue warpgroup (TMEM lane quarter = warp), warp 4 =
// TMEM allocator + single-thread MMA issuer, warp 5 = TMA producer.  The two single-thread roles carry the HIGHER warp ids on
// purpose: each shares an SM sub-partition with one ALU-heavy warpgroup warp, and the issue arbiter prefers the higher warp
// id — with the roles the other way round the MMA issuer starved (measured: 215 cycles per tcgen05.mma instead of 65).
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace apx {

FusedGeom fused_geom(int C, int H) {
    FusedGeom g;
    memset(&g, 0, sizeof(g));
    if (C <= 0 || H <= 0 || C % H || C % 64) return g;
    g.C = C; g.H = H;
    g.d = C / H;
    if (g.d % 8) return g;
    g.d_pad = (g.d + 15) / 16 * 16;
    g.nsplit = C <= 256 ? 1 : 2;
    g.CN = C / g.nsplit;
    if (g.CN % 16 || g.CN > 256) { g.nsplit = 0; return g; }
    g.wq_chunk_bytes = static_cast<size_t>(g.d_pad) * 128;               // [8 col-groups][d_pad][8] x 2 B
    g.wq_head_bytes = g.wq_chunk_bytes * (C / 64);
    g.wo_tile_bytes = static_cast<size_t>(g.CN) * g.d_pad * 2;           // [d_pad/8][CN][8] x 2 B
    g.wq_off = 0;
    g.wo_off = g.wq_off + g.wq_head_bytes * H;
    g.total = g.wo_off + g.wo_tile_bytes * H * g.nsplit;
    return g;
}

namespace {

constexpr int X_BOX_BYTES = 128 * 64 * 2;     // one SWIZZLE_128B box of the X tile: 128 rows x 64 channels
constexpr int MAX_WQ_STAGES = 8;

// ---- weight packing (once per weight version) ---------------------------------------------------------
// WQ section: [H][C/64][8][d_pad][8]: element (head h, row n < d_pad, input channel k) — rows n >= d are zero.
// WO section: [nsplit][H][d_pad/8][CN][8]: element (part p, head h, output column n < CN, head dim k < d_pad) —
//             head dims k >= d are zero.
__global__ void ipattn_pack_kernel(const uint16_t* __restrict__ Wq, const uint16_t* __restrict__ Wo,
                                   uint16_t* __restrict__ wq_out, uint16_t* __restrict__ wo_out, FusedGeom g) {
    const size_t n_wq = g.wq_head_bytes / 2 * g.H;
    const size_t n_wo = g.wo_tile_bytes / 2 * g.H * g.nsplit;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n_wq + n_wo;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        if (idx < n_wq) {
            size_t r = idx;
            const int e = r % 8; r /= 8;
            const int n = r % g.d_pad; r /= g.d_pad;
            const int grp = r % 8; r /= 8;
            const int chunk = r % (g.C / 64);
            const int h = r / (g.C / 64);
            const int k = chunk * 64 + grp * 8 + e;
            wq_out[idx] = n < g.d ? Wq[static_cast<size_t>(h * g.d + n) * g.C + k] : uint16_t(0);
        } else {
            size_t r = idx - n_wq;
            const int e = r % 8; r /= 8;
            const int n = r % g.CN; r /= g.CN;
            const int kg = r % (g.d_pad / 8); r /= (g.d_pad / 8);
            const int h = r % g.H;
            const int p = r / g.H;
            const int k = kg * 8 + e;
            wo_out[idx - n_wq] = k < g.d ? Wo[static_cast<size_t>(p * g.CN + n) * g.C + h * g.d + k] : uint16_t(0);
        }
    }
}

// ---- the fused kernel -----------------------------------------------------------------------------------
struct FusedArgs {
    const uint8_t* wq;     // packed WQ section
    const uint8_t* wo;     // packed WO section
    const uint8_t* tiles;  // packed context K / V^T tiles, [Bc][H]{K tile, V^T tile}
    const float* bo;       // fp32[C] or null
    uint16_t* Y;           // [Bc, S, C]
    int Bc, S, C, H, d_pad, CN;
    int Lt, La, Lt_pad, La_pad, L_pad;
    float ap_scale, scale_log2;
    int fmt;
    uint32_t kv_bytes, k_tile_bytes, wq_chunk_bytes, wo_tile_bytes;
    int nwq, nkv, nwo;                                   // ring depths
    uint32_t off_wq, off_kv, off_wo, off_p, off_o, off_bar;  // shared-memory offsets from the 1024-aligned base
    long long* trace;  // diagnostic (APX_FUSED_TRACE=1): clock64 stamps of CTA (0,0,0), [role][head][8]
};

#define APX_TRACE(role, h, slot)                                                                        \
    do {                                                                                                \
        if (TRACE && a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)                  \
            a.trace[((role) * 64 + (h)) * 8 + (slot)] = clock64();                                      \
    } while (0)

// Adding a byte offset to a UMMA shared-memory descriptor: the start-address field (bits 0-13, units of 16 B) cannot carry
// out of its field because every operand lives below 256 KB.
__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

// n dependent tcgen05.mma steps along K: A advances by a_step bytes, B by b_step bytes per 16-element k-step.
__device__ __forceinline__ void mma_chain(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t a_step, uint32_t b_step, int n,
                                          uint32_t idesc, uint32_t acc_first) {
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        umma_ss(tmem_d, da, db, idesc, acc_first | static_cast<uint32_t>(k));
        da = desc_add(da, a_step);
        db = desc_add(db, b_step);
    }
}

// The kernel is latency-bound, one warp per scheduler: the code is kept compact on purpose (loops over 16-column groups
// are NOT unrolled across groups) so that the three roles together stay inside the SM's instruction cache.
template <int FMT, bool TRACE>
__global__ void __launch_bounds__(192, 1)
ipattn_fused_kernel(const __grid_constant__ CUtensorMap tmX, FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sX = smem;
    uint8_t* sWQ = smem + a.off_wq;
    uint8_t* sKV = smem + a.off_kv;
    uint8_t* sWO = smem + a.off_wo;
    uint8_t* sP = smem + a.off_p;   // also holds the 16-bit Q_h tile between the Q conversion and S = Q K^T
    uint8_t* sO = smem + a.off_o;
    float* sBias = reinterpret_cast<float*>(smem + a.off_bar + 320);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint64_t* bar_x = bars;
    uint64_t* wq_full = bars + 1;
    uint64_t* wq_empty = wq_full + MAX_WQ_STAGES;
    uint64_t* kv_full = wq_empty + MAX_WQ_STAGES;
    uint64_t* kv_empty = kv_full + 2;
    uint64_t* wo_full = kv_empty + 2;
    uint64_t* wo_empty = wo_full + 2;
    uint64_t* bar_q = wo_empty + 2;   // Q_h accumulator complete            (MMA  -> warpgroup)
    uint64_t* bar_sq = bar_q + 1;     // 16-bit Q_h tile in shared memory     (warpgroup -> MMA)
    uint64_t* bar_s = bar_q + 2;      // S complete                           (MMA  -> warpgroup)
    uint64_t* bar_p = bar_q + 3;      // P tile in shared memory              (warpgroup -> MMA)
    uint64_t* bar_pv = bar_q + 4;     // O_t, O_a complete                    (MMA  -> warpgroup)
    uint64_t* bar_so = bar_q + 5;     // 16-bit O_h tile in shared memory     (warpgroup -> MMA)
    uint64_t* bar_yacc = bar_q + 6;   // Y += O_h Wo_h^T complete             (MMA  -> warpgroup)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 7);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int part = blockIdx.y;
    const int b = blockIdx.z;
    const int n_kb = a.C / 64;

    if (threadIdx.x == 160) {
        tma_prefetch_desc(&tmX);
        mbar_init(bar_x, 1);
#pragma unroll 1
        for (int i = 0; i < 2 * MAX_WQ_STAGES + 8; ++i) mbar_init(&wq_full[i], 1);   // all rings: one producer / one commit
        mbar_init(bar_q, 1);   mbar_init(bar_s, 1);  mbar_init(bar_pv, 1);  mbar_init(bar_yacc, 1);
        mbar_init(bar_sq, 128); mbar_init(bar_p, 128); mbar_init(bar_so, 128);  // every warpgroup thread arrives
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_Y = tmem_base;
    const uint32_t tmem_Q = tmem_base + a.CN;
    const uint32_t tmem_O0 = tmem_Q + a.d_pad;
    const uint32_t tmem_O1 = tmem_O0 + a.d_pad;
    const uint32_t tmem_S = tmem_O1 + a.d_pad;

    if (warp == 5) {
        // ================================================================ TMA producer
        if (lane == 0) {
            mbar_expect_tx(bar_x, static_cast<uint32_t>(n_kb) * X_BOX_BYTES);
#pragma unroll 1
            for (int kb = 0; kb < n_kb; ++kb) tma_load_2d(sX + kb * X_BOX_BYTES, &tmX, kb * 64, b * a.S + m0, bar_x);
            const uint8_t* wq_src = a.wq;
            const uint8_t* kv_src = a.tiles + static_cast<size_t>(b) * a.H * a.kv_bytes;
            const uint8_t* wo_src = a.wo + static_cast<size_t>(part) * a.H * a.wo_tile_bytes;
            int wq_s = 0, kv_s = 0, wo_s = 0;
            uint32_t wq_ph = 1, kv_ph = 1, wo_ph = 1;   // "empty" parity of the current ring pass (first pass: free)
#pragma unroll 1
            for (int h = 0; h < a.H; ++h) {
                APX_TRACE(0, h, 0);
#pragma unroll 1
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&wq_empty[wq_s], wq_ph);
                    mbar_expect_tx(&wq_full[wq_s], a.wq_chunk_bytes);
                    tma_load_1d(sWQ + wq_s * a.wq_chunk_bytes, wq_src, a.wq_chunk_bytes, &wq_full[wq_s]);
                    wq_src += a.wq_chunk_bytes;
                    if (++wq_s == a.nwq) { wq_s = 0; wq_ph ^= 1; }
                }
                APX_TRACE(0, h, 1);
                mbar_wait(&kv_empty[kv_s], kv_ph);
                mbar_expect_tx(&kv_full[kv_s], a.kv_bytes);
                tma_load_1d(sKV + kv_s * a.kv_bytes, kv_src, a.kv_bytes, &kv_full[kv_s]);
                kv_src += a.kv_bytes;
                if (++kv_s == a.nkv) { kv_s = 0; kv_ph ^= 1; }
                APX_TRACE(0, h, 2);
                mbar_wait(&wo_empty[wo_s], wo_ph);
                mbar_expect_tx(&wo_full[wo_s], a.wo_tile_bytes);
                tma_load_1d(sWO + wo_s * a.wo_tile_bytes, wo_src, a.wo_tile_bytes, &wo_full[wo_s]);
                wo_src += a.wo_tile_bytes;
                if (++wo_s == a.nwo) { wo_s = 0; wo_ph ^= 1; }
                APX_TRACE(0, h, 3);
            }
        }
    } else if (warp == 4) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc_d = umma_idesc_f16(FMT, 128, a.d_pad);   // Q_h and O accumulators
            const uint32_t idesc_s = umma_idesc_f16(FMT, 128, a.L_pad);
            const uint32_t idesc_y = umma_idesc_f16(FMT, 128, a.CN);
            const uint64_t dX = umma_desc(smem_u32(sX), 16, 1024, kLayoutSW128);
            const uint64_t dWQ = umma_desc(smem_u32(sWQ), a.d_pad * 16, 128, kLayoutNone);
            const uint64_t dQP = umma_desc(smem_u32(sP), 2048, 128, kLayoutNone);     // Q_h / P as the A operand
            const uint64_t dK = umma_desc(smem_u32(sKV), a.L_pad * 16, 128, kLayoutNone);
            const uint64_t dV = umma_desc(smem_u32(sKV) + a.k_tile_bytes, a.d_pad * 16, 128, kLayoutNone);
            const uint64_t dO = umma_desc(smem_u32(sO), 2048, 128, kLayoutNone);
            const uint64_t dWO = umma_desc(smem_u32(sWO), a.CN * 16, 128, kLayoutNone);
            const uint32_t wq_kstep = 2 * a.d_pad * 16, k_kstep = 2 * a.L_pad * 16, v_kstep = 2 * a.d_pad * 16,
                           wo_kstep = 2 * a.CN * 16;
            int wq_s = 0, kv_s = 0, wo_s = 0;
            uint32_t wq_ph = 0, kv_ph = 0, wo_ph = 0;   // "full" parity of the current ring pass
            // Q_h = X · Wq_h^T: A = resident SW128 X boxes, B = ring of packed Wq chunks (no swizzle, K-major)
            auto issue_q = [&]() {
#pragma unroll 1
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&wq_full[wq_s], wq_ph);
                    tc_fence_after();
                    mma_chain(tmem_Q, desc_add(dX, kb * X_BOX_BYTES), desc_add(dWQ, wq_s * a.wq_chunk_bytes), 32, wq_kstep, 4,
                              idesc_d, static_cast<uint32_t>(kb));
                    umma_commit(&wq_empty[wq_s]);
                    if (++wq_s == a.nwq) { wq_s = 0; wq_ph ^= 1; }
                }
                umma_commit(bar_q);
            };
            APX_TRACE(1, 63, 0);
            mbar_wait(bar_x, 0);
            APX_TRACE(1, 63, 1);
            issue_q();
            APX_TRACE(1, 63, 2);
#pragma unroll 1
            for (int h = 0; h < a.H; ++h) {
                const uint32_t ph = h & 1;
                const uint32_t kv_off = kv_s * a.kv_bytes;
                APX_TRACE(1, h, 0);
                mbar_wait(bar_sq, ph);  // 16-bit Q_h is in shared memory; the Q accumulator is free again
                APX_TRACE(1, h, 1);
                mbar_wait(&kv_full[kv_s], kv_ph);
                tc_fence_after();
                APX_TRACE(1, h, 2);
                mma_chain(tmem_S, dQP, desc_add(dK, kv_off), 2 * 2048, k_kstep, a.d_pad / 16, idesc_s, 0);
                umma_commit(bar_s);
                APX_TRACE(1, h, 3);
                if (h + 1 < a.H) issue_q();  // next head's Q projection runs under this head's softmax
                APX_TRACE(1, h, 4);
                mbar_wait(bar_p, ph);
                tc_fence_after();
                APX_TRACE(1, h, 5);
                // text columns [0, Lt_pad) -> O_t, audio columns [Lt_pad, L_pad) -> O_a
                mma_chain(tmem_O0, dQP, desc_add(dV, kv_off), 2 * 2048, v_kstep, a.Lt_pad / 16, idesc_d, 0);
                mma_chain(tmem_O1, desc_add(dQP, (a.Lt_pad / 8) * 2048), desc_add(dV, kv_off + (a.Lt_pad / 8) * (a.d_pad * 16)),
                          2 * 2048, v_kstep, a.La_pad / 16, idesc_d, 0);
                umma_commit(bar_pv);
                umma_commit(&kv_empty[kv_s]);
                if (++kv_s == a.nkv) { kv_s = 0; kv_ph ^= 1; }
                APX_TRACE(1, h, 6);
                mbar_wait(bar_so, ph);
                mbar_wait(&wo_full[wo_s], wo_ph);
                tc_fence_after();
                APX_TRACE(1, h, 7);
                mma_chain(tmem_Y, dO, desc_add(dWO, wo_s * a.wo_tile_bytes), 2 * 2048, wo_kstep, a.d_pad / 16, idesc_y,
                          static_cast<uint32_t>(h));
                umma_commit(bar_yacc);
                umma_commit(&wo_empty[wo_s]);
                if (++wo_s == a.nwo) { wo_s = 0; wo_ph ^= 1; }
            }
        }
    } else {
        // ================================================================ softmax / conversion / epilogue warpgroup
        // One query row per thread; each warp sits alone on its SM sub-partition.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // tile row == TMEM lane
        const uint32_t lane_taddr = static_cast<uint32_t>(q * 32) << 16;
        const int la_end = a.Lt_pad + a.La;
        const uint32_t sP_row = smem_u32(sP) + row * 16;   // element (row, c) of a K-major core-matrix tile:
        const uint32_t sO_row = smem_u32(sO) + row * 16;   //   (c / 8) * 2048 + row * 16 + (c % 8) * 2
        const float sl = a.scale_log2;
        const int n0 = part * a.CN;
        if (a.bo) {
            for (int i = threadIdx.x; i < a.CN; i += 128) sBias[i] = a.bo[n0 + i];
        }

        auto store16 = [&](uint32_t dst, const float* v) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack2<FMT>(v[2 * i], v[2 * i + 1]);
            st_shared_v4(dst, w[0], w[1], w[2], w[3]);
            st_shared_v4(dst + 2048, w[4], w[5], w[6], w[7]);
        };

#pragma unroll 1
        for (int h = 0; h < a.H; ++h) {
            const uint32_t ph = h & 1;
            // ---- Q_h: fp32 accumulator -> 16-bit K-major core-matrix tile (the A operand of S = Q K^T)
            if (threadIdx.x == 0) APX_TRACE(2, h, 0);
            mbar_wait(bar_q, ph);
            tc_fence_after();
            if (threadIdx.x == 0) APX_TRACE(2, h, 1);
#pragma unroll 1
            for (int c = 0; c < a.d_pad; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_Q + lane_taddr + c, r);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                store16(sP_row + (c >> 3) * 2048, v);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_sq);
            if (threadIdx.x == 0) APX_TRACE(2, h, 2);

            // ---- two softmaxes over one S row: text columns [0, Lt), audio columns [Lt_pad, Lt_pad + La)
            mbar_wait(bar_s, ph);
            tc_fence_after();
            if (threadIdx.x == 0) APX_TRACE(2, h, 3);
            float mx_t = -INFINITY, mx_a = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < a.L_pad; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_S + lane_taddr + c, r);
                tmem_ld_wait();
                const bool text = c < a.Lt_pad;
                const int nv = (text ? a.Lt : la_end) - c;   // valid columns in this group (may be <= 0 or >= 16)
                float m4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) m4[j] = -INFINITY;
                if (nv >= 16) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], i < nv ? __uint_as_float(r[i]) : -INFINITY);
                }
                const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                if (text) mx_t = fmaxf(mx_t, m); else mx_a = fmaxf(mx_a, m);
            }
            if (threadIdx.x == 0) APX_TRACE(2, h, 4);
            const float mxs_t = mx_t * sl, mxs_a = mx_a * sl;
            float l_t = 0.f, l_a = 0.f;
            // P = exp2(S * scale_log2 - max * scale_log2), unnormalised, 16 bit; padded columns are exactly zero
#pragma unroll 1
            for (int c = 0; c < a.L_pad; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_S + lane_taddr + c, r);
                tmem_ld_wait();
                const bool text = c < a.Lt_pad;
                const float mxs = text ? mxs_t : mxs_a;
                const int nv = (text ? a.Lt : la_end) - c;
                float e[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x = ex2_approx(fmaf(__uint_as_float(r[i]), sl, -mxs));
                    e[i] = i < nv ? x : 0.f;
                }
                const float sum = (((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]))) +
                                  (((e[8] + e[9]) + (e[10] + e[11])) + ((e[12] + e[13]) + (e[14] + e[15])));
                if (text) l_t += sum; else l_a += sum;
                store16(sP_row + (c >> 3) * 2048, e);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p);
            if (threadIdx.x == 0) APX_TRACE(2, h, 5);

            // ---- O_h = O_t / l_t + ap_scale * O_a / l_a  -> 16-bit tile, the A operand of the out-projection
            const float w0 = a.Lt > 0 ? 1.0f / l_t : 0.f;
            const float w1 = a.La > 0 ? a.ap_scale / l_a : 0.f;
            mbar_wait(bar_pv, ph);
            if (h > 0) mbar_wait(bar_yacc, ph ^ 1);  // the previous head's out-projection has read sO
            tc_fence_after();
            if (threadIdx.x == 0) APX_TRACE(2, h, 6);
#pragma unroll 1
            for (int c = 0; c < a.d_pad; c += 16) {
                uint32_t r0[16], r1[16];
                if (a.Lt > 0) tmem_ld16(tmem_O0 + lane_taddr + c, r0);
                if (a.La > 0) tmem_ld16(tmem_O1 + lane_taddr + c, r1);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float acc = 0.f;
                    if (a.Lt > 0) acc = __uint_as_float(r0[i]) * w0;
                    if (a.La > 0) acc = fmaf(__uint_as_float(r1[i]), w1, acc);
                    v[i] = acc;
                }
                store16(sO_row + (c >> 3) * 2048, v);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_so);
            if (threadIdx.x == 0) APX_TRACE(2, h, 7);
        }

        // ---- epilogue: Y tile = accumulator + bias -> 16 bit -> global
        if (threadIdx.x == 0) APX_TRACE(2, 63, 0);
        mbar_wait(bar_yacc, (a.H - 1) & 1);
        tc_fence_after();
        if (threadIdx.x == 0) APX_TRACE(2, 63, 1);
        const int srow = m0 + row;
        uint16_t* yrow = a.Y + (static_cast<int64_t>(b) * a.S + srow) * a.C + n0;
        const bool row_ok = srow < a.S;
#pragma unroll 1
        for (int c = 0; c < a.CN; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_Y + lane_taddr + c, r);
            tmem_ld_wait();
            uint32_t p[8];
            if (a.bo) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 bb = *reinterpret_cast<const float2*>(sBias + c + 2 * i);
                    p[i] = pack2<FMT>(__uint_as_float(r[2 * i]) + bb.x, __uint_as_float(r[2 * i + 1]) + bb.y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = pack2<FMT>(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            }
            if (row_ok) {
                *reinterpret_cast<uint4*>(yrow + c) = make_uint4(p[0], p[1], p[2], p[3]);
                *reinterpret_cast<uint4*>(yrow + c + 8) = make_uint4(p[4], p[5], p[6], p[7]);
            }
        }
        if (threadIdx.x == 0) APX_TRACE(2, 63, 2);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem_base);
}

struct FusedPlan {
    bool ok;
    int nwq, nkv, nwo;
    uint32_t off_wq, off_kv, off_wo, off_p, off_o, off_bar, smem_bytes;
};

FusedPlan fused_plan(const FusedGeom& g, int Lt, int La) {
    FusedPlan p;
    memset(&p, 0, sizeof(p));
    if (g.nsplit == 0) return p;
    const KvPackGeom kg = kv_pack_geom(Lt, La, g.C, g.H);
    if (kg.L_pad < 16 || kg.L_pad > 256) return p;
    if (g.CN + 3 * g.d_pad + kg.L_pad > 512) return p;  // tensor-memory columns: Y | Q | O_t | O_a | S
    const uint32_t x_bytes = static_cast<uint32_t>(g.C / 64) * X_BOX_BYTES;
    const uint32_t p_bytes = 128u * std::max(kg.L_pad, g.d_pad) * 2;
    const uint32_t o_bytes = 128u * g.d_pad * 2;
    const int tries[][3] = {{8, 2, 2}, {6, 2, 2}, {4, 2, 2}, {4, 2, 1}, {4, 1, 1}, {2, 1, 1}};
    for (const auto& t : tries) {
        uint32_t off = x_bytes;
        p.off_wq = off; off += static_cast<uint32_t>(align_up(t[0] * g.wq_chunk_bytes, 128));
        p.off_kv = off; off += static_cast<uint32_t>(align_up(t[1] * kg.bytes_per_bh, 128));
        p.off_wo = off; off += static_cast<uint32_t>(align_up(t[2] * g.wo_tile_bytes, 128));
        p.off_p = off;  off += p_bytes;
        p.off_o = off;  off += o_bytes;
        p.off_bar = off; off += 256 + 1024 + 64;   // mbarriers + TMEM slot | fp32 bias slice of this CTA
        const uint32_t total = off + 1024;  // + alignment slack of the dynamic shared-memory base
        if (total <= 227u * 1024u) {
            p.ok = true;
            p.nwq = std::min(t[0], (g.C / 64) * g.H); p.nkv = t[1]; p.nwo = t[2];
            p.smem_bytes = total;
            return p;
        }
    }
    return p;
}

}  // namespace

bool ipattn_fused_supported(int C, int H, int Lt, int La) {
    const FusedGeom g = fused_geom(C, H);
    if (g.nsplit == 0 || Lt < 0 || La < 0 || Lt + La <= 0) return false;
    return fused_plan(g, Lt, La).ok;
}

int ipattn_pack_weights(const void* Wq, const void* Wo, void* pack, int C, int H, int dtype, cudaStream_t st) {
    const FusedGeom g = fused_geom(C, H);
    APX_REQUIRE(g.nsplit != 0, APX_EUNSUPPORTED, "apx_ipattn_pack: the fused kernel does not cover C=%d H=%d", C, H);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "apx_ipattn_pack: dtype %d", dtype);
    const size_t total = (g.total - g.wq_off) / 2;
    const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 8));
    uint8_t* pb = static_cast<uint8_t*>(pack);
    ipattn_pack_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(Wq), static_cast<const uint16_t*>(Wo),
                                               reinterpret_cast<uint16_t*>(pb + g.wq_off),
                                               reinterpret_cast<uint16_t*>(pb + g.wo_off), g);
    APX_LAUNCH_CHECK("ipattn_pack_kernel");
    return APX_OK;
}

int ipattn_fused(const void* X, const void* wpack, const void* cache_tiles, const float* bo, void* Y, int Bc, int S,
                 int C, int H, int Lt, int La, float ap_scale, int dtype, cudaStream_t st) {
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "ipattn_fused: dtype %d", dtype);
    const FusedGeom g = fused_geom(C, H);
    APX_REQUIRE(g.nsplit != 0, APX_EUNSUPPORTED, "ipattn_fused: C=%d H=%d is not covered", C, H);
    const FusedPlan p = fused_plan(g, Lt, La);
    APX_REQUIRE(p.ok, APX_EUNSUPPORTED, "ipattn_fused: context %d+%d does not fit the fused kernel at C=%d", Lt, La, C);
    APX_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wpack) & 15) == 0 && (reinterpret_cast<uintptr_t>(cache_tiles) & 15) == 0,
                APX_EINVAL, "ipattn_fused: X / Y / pack / cache must be 16-byte aligned");
    APX_REQUIRE(!bo || (reinterpret_cast<uintptr_t>(bo) & 15) == 0, APX_EINVAL, "ipattn_fused: bias alignment");
    const KvPackGeom kg = kv_pack_geom(Lt, La, C, H);

    CUtensorMap tmX;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(Bc) * S};
        uint64_t strides[1] = {static_cast<uint64_t>(C) * 2};
        uint32_t box[2] = {64, 128};
        int rc = encode_tmap(&tmX, dtype, 2, X, dims, strides, box, 1);
        if (rc) return rc;
    }
    FusedArgs a;
    const uint8_t* pb = static_cast<const uint8_t*>(wpack);
    a.wq = pb + g.wq_off;
    a.wo = pb + g.wo_off;
    a.tiles = static_cast<const uint8_t*>(cache_tiles);
    a.bo = bo;
    a.Y = static_cast<uint16_t*>(Y);
    a.Bc = Bc; a.S = S; a.C = C; a.H = H; a.d_pad = g.d_pad; a.CN = g.CN;
    a.Lt = Lt; a.La = La; a.Lt_pad = kg.Lt_pad; a.La_pad = kg.La_pad; a.L_pad = kg.L_pad;
    a.ap_scale = ap_scale;
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(g.d));
    a.fmt = dtype == APX_F16 ? 0 : 1;
    a.kv_bytes = static_cast<uint32_t>(kg.bytes_per_bh);
    a.k_tile_bytes = static_cast<uint32_t>(kg.k_tile_bytes);
    a.wq_chunk_bytes = static_cast<uint32_t>(g.wq_chunk_bytes);
    a.wo_tile_bytes = static_cast<uint32_t>(g.wo_tile_bytes);
    a.nwq = p.nwq; a.nkv = p.nkv; a.nwo = p.nwo;
    a.off_wq = p.off_wq; a.off_kv = p.off_kv; a.off_wo = p.off_wo; a.off_p = p.off_p; a.off_o = p.off_o; a.off_bar = p.off_bar;

    // diagnostic timeline (APX_FUSED_TRACE=1): synchronises and prints, never enabled on the product path
    static const bool trace_on = [] { const char* e = getenv("APX_FUSED_TRACE"); return e && e[0] == '1'; }();
    static long long* trace_buf = nullptr;
    a.trace = nullptr;
    if (trace_on) {
        if (!trace_buf) APX_CUDA(cudaMalloc(&trace_buf, 3 * 64 * 8 * sizeof(long long)));
        APX_CUDA(cudaMemsetAsync(trace_buf, 0, 3 * 64 * 8 * sizeof(long long), st));
        a.trace = trace_buf;
    }
    static bool attr_set = false;
    if (!attr_set) {
        APX_CUDA(cudaFuncSetAttribute(ipattn_fused_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        APX_CUDA(cudaFuncSetAttribute(ipattn_fused_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        APX_CUDA(cudaFuncSetAttribute(ipattn_fused_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid(ceil_div(S, 128), g.nsplit, Bc);
    if (trace_on && a.fmt == 0) ipattn_fused_kernel<0, true><<<grid, 192, p.smem_bytes, st>>>(tmX, a);
    else if (a.fmt == 0)        ipattn_fused_kernel<0, false><<<grid, 192, p.smem_bytes, st>>>(tmX, a);
    else                        ipattn_fused_kernel<1, false><<<grid, 192, p.smem_bytes, st>>>(tmX, a);
    APX_LAUNCH_CHECK("ipattn_fused_kernel");
    if (trace_on) {
        static int printed = 0;
        std::vector<long long> t(3 * 64 * 8);
        APX_CUDA(cudaStreamSynchronize(st));
        APX_CUDA(cudaMemcpy(t.data(), trace_buf, t.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        if (printed++ < 2) {
            const long long t0 = t[(1 * 64 + 63) * 8 + 0];
            auto rel = [&](int role, int h, int slot) { const long long v = t[(role * 64 + h) * 8 + slot]; return v ? v - t0 : -1; };
            printf("[apx trace] Bc=%d S=%d C=%d Lt=%d La=%d  (cycles since the MMA thread started; CTA 0)\n", Bc, S, C, Lt, La);
            printf("  mma: x_wait %lld x_ready %lld q0_issued %lld\n", rel(1, 63, 0), rel(1, 63, 1), rel(1, 63, 2));
            for (int h = 0; h < H; ++h) {
                printf("  h%d tma : wq %lld kv %lld wo %lld done %lld\n", h, rel(0, h, 0), rel(0, h, 1), rel(0, h, 2), rel(0, h, 3));
                printf("     mma : top %lld sq_seen %lld kv_ok %lld s_issued %lld qnext_issued %lld p_seen %lld pv_issued %lld so_seen %lld\n",
                       rel(1, h, 0), rel(1, h, 1), rel(1, h, 2), rel(1, h, 3), rel(1, h, 4), rel(1, h, 5), rel(1, h, 6), rel(1, h, 7));
                printf("     wg  : top %lld q_ready %lld sq_arrived %lld s_ready %lld max_done %lld p_arrived %lld pv_ready %lld so_arrived %lld\n",
                       rel(2, h, 0), rel(2, h, 1), rel(2, h, 2), rel(2, h, 3), rel(2, h, 4), rel(2, h, 5), rel(2, h, 6), rel(2, h, 7));
            }
            printf("  wg epilogue: wait %lld y_ready %lld stored %lld\n", rel(2, 63, 0), rel(2, 63, 1), rel(2, 63, 2));
        }
    }
    return APX_OK;
}

}  // namespace apx
