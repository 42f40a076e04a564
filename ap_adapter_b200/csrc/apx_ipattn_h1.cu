// apx_ipattn_h1.cu — one adapter site in ONE tcgen05 kernel, "one head per CTA" form:
//
//   Y = ( softmax(Q K_t^T / sqrt(d)) V_t + ap_scale * softmax(Q K_a^T / sqrt(d)) V_a ) Wo^T + bo,   Q = X Wq^T
//
// (to_q, both SDPAs, the ap_scale add and to_out of APadapter/ap_adapter/attention_processor.py:387-459.)
//
// The all-heads kernel (apx_ipattn_site.cu) keeps Q / O of every head of a 128-query tile in one SM: that caps it at
// H * d_pad <= 256 (C = 256) and leaves a B = 1 launch with 16 busy SMs.  Here a thread-block CLUSTER of H CTAs owns a query
// tile and CTA h owns head h end to end:
//
//   phase A   Q_h[128, d_pad] = X[128, C] · Wq_h^T          X chunks (SWIZZLE_128B tensor TMA) + the head's packed Wq rows (1-D bulk
//                                                           TMA) through an mbarrier ring; N = d_pad MMAs, K = C
//   phase B   S = Q_h K_h^T -> two softmaxes -> P · [V | 1]  exactly the all-heads kernel's per-head step (score row in registers,
//                                                           denominators out of the tensor core), one warpgroup, no ring
//             O_h = O_t / l_t + ap_scale · O_a / l_a         16 bit, into slot h of the [128, H*d_pad] O tile
//   gather    every CTA pushes its slot into the same slot of all H-1 peers over distributed shared memory
//             (cp.async.bulk.shared::cluster, bytes completing on the peer's mbarrier)
//   phase C   Y[128, h*d : (h+1)*d] = O[128, H*d_pad] · Wo[h*d : (h+1)*d, :]^T + bo     K = H*d_pad, N = d_pad, Wo rows streamed
//             through a second ring; the d valid columns leave as ONE 4-D tensor-TMA store (or reduce-add into the residual stream)
//
// so any C = H * d with H <= 8 (the portable cluster size), d % 8 == 0 is covered: C = 256 / 320 / 384 / 640 of AudioLDM2
//  with d = 32 / 40 / 48 / 80.  Shared memory, two plans (h1_plan): when everything fits side by side
// (C <= 384) the O tile region is ONLY ever written by its owner and by the peers' pushes, so a CTA pushes its slot the moment it
// exists; at C = 640 the O tile region (160 KB) is also the landing zone of the phase-A ring and of the P tile, and a second
// cluster barrier ("every CTA is past phase B") gates the pushes.
//
// Warp roles (192 threads): warps 0-3 = softmax / conversion / epilogue warpgroup (TMEM lane quarter = warp), warp 4 = TMEM
// allocator + single-thread MMA issuer, warp 5 = TMA producer.
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include "apx_site_common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace apx {

H1Geom h1_geom(int C, int H) {
    H1Geom g;
    memset(&g, 0, sizeof(g));
    if (C <= 0 || H <= 0 || H > 8 || C % H || C % 64) return g;
    const int d = C / H;
    if (d % 8 || d > 112) return g;
    g.ok = 1; g.C = C; g.H = H; g.d = d;
    g.d_pad = (d + 15) / 16 * 16;
    g.dv_pad = g.d_pad + 16;
    g.Nq = H * g.d_pad;
    g.n_kb = C / 64;
    g.wq_off = 0;                                                  // [H][C/64][8][d_pad][8]
    g.wo_off = static_cast<size_t>(g.Nq) * C * 2;                   // [H][Nq/8][d_pad][8]
    g.total = g.wo_off + static_cast<size_t>(g.Nq) * g.Nq * 2;
    return g;
}

namespace {

constexpr int X_CHUNK_BYTES = 128 * 64 * 2;   // one SWIZZLE_128B box: 128 rows x 64 channels
constexpr int KS = 8;                         // k-steps (16 O columns each) per Wo stage
constexpr int NTHREADS = 192;
constexpr int MAX_A = 6, MAX_W = 8;

// WQ: element (h, kb, kg, n, e) = Wq[h*d + n, kb*64 + kg*8 + e]          (rows n >= d are zero)
// WO: element (r, kg, n, e)     = Wo[r*d + n, h'*d + j],  kg*8 + e = h'*d_pad + j   (zero for n >= d or j >= d)
__global__ void h1_pack_kernel(const uint16_t* __restrict__ Wq, const uint16_t* __restrict__ Wo, uint16_t* __restrict__ out, H1Geom g) {
    const size_t n_q = static_cast<size_t>(g.Nq) * g.C, n_o = static_cast<size_t>(g.Nq) * g.Nq;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n_q + n_o;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        if (idx < n_q) {
            size_t r = idx;
            const int e = r % 8; r /= 8;
            const int n = r % g.d_pad; r /= g.d_pad;
            const int kg = r % 8; r /= 8;
            const int kb = r % g.n_kb;
            const int h = static_cast<int>(r / g.n_kb);
            out[g.wq_off / 2 + idx] = n < g.d ? Wq[static_cast<size_t>(h * g.d + n) * g.C + kb * 64 + kg * 8 + e] : uint16_t(0);
        } else {
            size_t r = idx - n_q;
            const int e = r % 8; r /= 8;
            const int n = r % g.d_pad; r /= g.d_pad;
            const int kg = r % (g.Nq / 8);
            const int rank = static_cast<int>(r / (g.Nq / 8));
            const int kk = kg * 8 + e, hh = kk / g.d_pad, j = kk % g.d_pad;
            out[g.wo_off / 2 + (idx - n_q)] = (n < g.d && j < g.d) ? Wo[static_cast<size_t>(rank * g.d + n) * g.C + hh * g.d + j] : uint16_t(0);
        }
    }
}

struct H1Args {
    const uint8_t* wq;     // packed WQ section
    const uint8_t* wo;     // packed WO section
    const uint8_t* tiles;  // packed context K / V^T tiles, [Bc][H]{K tile, V^T tile, extended V^T tile}
    const float* bo;       // fp32[C] or null
    int S, C, H, d, d_pad, dv_pad, Nq, n_kb;
    int Lt, La, Lt_pad, La_pad, L_pad;
    float ap_scale, scale_log2;
    uint32_t kv_bytes, k_tile_bytes, v_tile_bytes, vx_off, wq_chunk_bytes, wo_stage_bytes, a_stage_bytes, p_bytes, slot_bytes;
    int ns_a, n_w, nw_slots, wo_early, n_ko;
    int p_in_r0, overlay;                      // overlay: phase-A ring / P tile share the O tile region -> pushes gated by cluster barrier #2
    uint32_t off_a, off_p, off_k, off_v, off_bar;   // shared-memory offsets from the 1024-aligned base (the O tile is at 0)
    uint32_t wo_off[MAX_W];
    int accumulate;
    long long* trace;      // diagnostic (APX_SITE_TRACE=1): clock64 stamps of CTA (0,0); null on the product path
};

// trace slots: [0,16) MMA thread, [16,24) producer, [24,48) warpgroup thread 0
#define H1_T(slot)                                                                      \
    do {                                                                                \
        if (a.trace && blockIdx.x == 0 && blockIdx.y == 0) a.trace[(slot)] = clock64(); \
    } while (0)

enum {  // mbarrier slots
    B_A_FULL = 0, B_A_EMPTY = MAX_A, B_W_FULL = 2 * MAX_A, B_W_EMPTY = 2 * MAX_A + MAX_W, B_K_FULL = 2 * MAX_A + 2 * MAX_W, B_V_FULL,
    B_Q, B_SQ, B_S_FULL, B_P_FULL, B_O_FULL, B_PV_ALL, B_SO, B_Y, B_OALL, B_DONE, B_COUNT
};
constexpr uint32_t BAR_REGION_BYTES = 512;   // barriers [0, 8 * B_COUNT) | TMEM slot at 480 | bias slice from 512
static_assert(B_COUNT * 8 <= 480, "barrier region");

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// N16 > 0: the score row is N16 16-column groups (the first LT16 of them text) and lives in registers between the two softmax
//          passes; N16 == 0: any context shape, two passes over tensor memory.
template <int FMT, int N16, int LT16>
__global__ void __launch_bounds__(NTHREADS, 1)
ipattn_site_h1_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, H1Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int h = a.H > 1 ? static_cast<int>(cluster_ctarank()) : 0;   // this CTA's head == its rank in the cluster
    const int m0 = (blockIdx.x / a.H) * 128;
    const int b = blockIdx.y;
    const bool clustered = a.H > 1;

    uint8_t* sO = smem;                                        // [Nq/8][128][8] 16 bit: the gathered O tile; slot h = Q_h, then O_h
    uint8_t* sSlot = sO + h * a.slot_bytes;
    // the P tile lives in the larger half of the O tile next to the own slot when that half can hold it
    uint8_t* sP = smem + (a.p_in_r0 ? (static_cast<uint32_t>(h) * a.slot_bytes >= a.p_bytes ? 0u : static_cast<uint32_t>(h + 1) * a.slot_bytes)
                                    : a.off_p);
    uint8_t* sK = smem + a.off_k;
    uint8_t* sV = smem + a.off_v;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.off_bar + 480);
    float* sBias = reinterpret_cast<float*>(smem + a.off_bar + BAR_REGION_BYTES);

    if (threadIdx.x == 160) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
#pragma unroll 1
        for (int i = 0; i < B_COUNT; ++i) {
            const uint32_t cnt = (i == B_SQ || i == B_P_FULL || i == B_SO) ? 128u : i == B_DONE ? static_cast<uint32_t>(clustered ? a.H - 1 : 1) : 1u;
            mbar_init(&bars[i], cnt);
        }
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    pdl_wait();   // programmatic dependent launch: set-up above, global reads below
    if (threadIdx.x < 128) {
        for (int i = threadIdx.x; i < a.d_pad; i += 128) sBias[i] = (a.bo && i < a.d) ? a.bo[h * a.d + i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    pdl_launch_dependents();
    tc_fence_after();
    if (clustered) cluster_arrive();   // #1: this CTA's barriers exist (waited for before the first remote operation)
    const uint32_t tmem_base = *tmem_slot;
    // tensor memory: Q accumulator, then S, then Y at column 0; {O_t, O_a} (dv_pad columns each, column d_pad = row sum) behind S
    const uint32_t tmem_O = tmem_base + a.L_pad;
    const int first_wave = min(a.ns_a, a.n_kb);
    const uint8_t* wq_src = a.wq + static_cast<size_t>(h) * a.n_kb * a.wq_chunk_bytes;
    // one phase-A chunk (64 input channels): X box + the head's Wq rows -> ring slot kb % ns_a
    auto load_a = [&](int kb) {
        const int slot = kb % a.ns_a;
        uint8_t* dst = smem + a.off_a + slot * a.a_stage_bytes;
        mbar_expect_tx(&bars[B_A_FULL + slot], X_CHUNK_BYTES + a.wq_chunk_bytes);
        tma_load_3d(dst, &tmX, kb * 64, m0, b, &bars[B_A_FULL + slot]);
        tma_load_1d(dst + X_CHUNK_BYTES, wq_src + static_cast<size_t>(kb) * a.wq_chunk_bytes, a.wq_chunk_bytes, &bars[B_A_FULL + slot]);
    };

    if (warp == 5) {
        // ================================================================ TMA producer (one thread)
        const uint8_t* wo_src = a.wo + static_cast<size_t>(h) * a.n_ko * (a.wo_stage_bytes / KS);
        auto load_wo = [&](int i) {
            const int slot = i % a.nw_slots;
            if (i >= a.nw_slots) mbar_wait_nocall(&bars[B_W_EMPTY + slot], ((i / a.nw_slots) - 1) & 1);
            const int ks = min(KS, a.n_ko - i * KS);
            const uint32_t bytes = ks * (a.wo_stage_bytes / KS);
            mbar_expect_tx(&bars[B_W_FULL + slot], bytes);
            tma_load_1d(smem + a.wo_off[slot], wo_src + static_cast<size_t>(i) * a.wo_stage_bytes, bytes, &bars[B_W_FULL + slot]);
        };
        if (elect_one()) {
            H1_T(16);
            // The first min(4, first_wave) phase-A chunks are issued by the four (still idle) warpgroup warps, one each: a TMA issue
            // costs this thread ~180 cycles, and the first MMA can only start when its chunk has landed.
#pragma unroll 1
            for (int kb = 4; kb < first_wave; ++kb) load_a(kb);
            // K / V^T / early Wo are not needed before the softmax: keep them out of the way of the first phase-A chunks (the
            // first MMA waits for chunk 0, and everything in flight shares this SM's ~40 B/clk of TMA ingest)
            mbar_wait_nocall(&bars[B_A_FULL + min(first_wave, 2) - 1], 0);
            const uint8_t* kv_src = a.tiles + (static_cast<size_t>(b) * a.H + h) * a.kv_bytes;
            mbar_expect_tx(&bars[B_K_FULL], a.k_tile_bytes);
            tma_load_1d(sK, kv_src, a.k_tile_bytes, &bars[B_K_FULL]);
            mbar_expect_tx(&bars[B_V_FULL], a.v_tile_bytes);
            tma_load_1d(sV, kv_src + a.vx_off, a.v_tile_bytes, &bars[B_V_FULL]);
#pragma unroll 1
            for (int i = 0; i < a.wo_early && i < a.n_w; ++i) load_wo(i);
            H1_T(17);
#pragma unroll 1
            for (int kb = first_wave; kb < a.n_kb; ++kb) {   // the rest of phase A through the ring
                mbar_wait_nocall(&bars[B_A_EMPTY + kb % a.ns_a], ((kb / a.ns_a) - 1) & 1);
                load_a(kb);
            }
        }
        __syncwarp();
        // #1 wait, #2 arrive BEFORE the late Wo stages: those wait for phase C to free ring slots, and phase C waits for #2
        if (clustered) { cluster_wait(); if (a.overlay) cluster_arrive(); }
        if (elect_one()) {
            if (a.n_w > a.wo_early) {
                mbar_wait_nocall(&bars[B_PV_ALL], 0);   // the later Wo stages land over the dead phase-A / P / K / V regions
                H1_T(18);
#pragma unroll 1
                for (int i = a.wo_early; i < a.n_w; ++i) load_wo(i);
            }
            H1_T(19);
            if (clustered) {
                // exit handshake, off the MMA thread's path: once every peer's slot has landed here tell each peer so; leave only
                // when every peer has said the same (then no bulk copy can still be reading this CTA's shared memory)
                mbar_wait_nocall(&bars[B_OALL], 0);
                const uint32_t done = smem_u32(&bars[B_DONE]);
#pragma unroll 1
                for (int r = 0; r < a.H; ++r)
                    if (r != h) mbar_arrive_remote(map_to_cta(done, r));
                mbar_wait_nocall(&bars[B_DONE], 0);
            }
        }
        __syncwarp();
        if (clustered && a.overlay) cluster_wait();   // #2
    } else if (warp == 4) {
        // ================================================================ MMA issuer (one thread); its waits spin (see mbar_wait_spin)
        if (elect_one()) {
            H1_T(0);
            if (clustered) mbar_expect_tx(&bars[B_OALL], static_cast<uint32_t>(a.H - 1) * a.slot_bytes);
            // ---- phase A: Q_h = X · Wq_h^T
            const uint32_t idesc_q = umma_idesc_f16(FMT, 128, a.d_pad);
#pragma unroll 1
            for (int kb = 0; kb < a.n_kb; ++kb) {
                const int slot = kb % a.ns_a;
                mbar_wait_spin(&bars[B_A_FULL + slot], (kb / a.ns_a) & 1);
                tc_fence_after();
                if (kb == 0) H1_T(1);
                const uint32_t base = smem_u32(smem + a.off_a + slot * a.a_stage_bytes);
                mma_chain(tmem_base, umma_desc(base, 16, 1024, kLayoutSW128), umma_desc(base + X_CHUNK_BYTES, a.d_pad * 16, 128, kLayoutNone),
                          32, 2 * a.d_pad * 16, 4, idesc_q, static_cast<uint32_t>(kb));
                if (kb + a.ns_a < a.n_kb) umma_commit(&bars[B_A_EMPTY + slot]);   // only slots that are refilled need the signal
            }
            umma_commit(&bars[B_Q]);
            H1_T(2);
            // ---- phase B: S = Q_h K^T, then O_t / O_a = P · [V | 1]
            mbar_wait_spin(&bars[B_K_FULL], 0);
            mbar_wait_spin(&bars[B_SQ], 0);       // Q_h is in the own slot; the Q accumulator columns are free
            tc_fence_after();
            H1_T(3);
            const uint64_t dQ = umma_desc(smem_u32(sSlot), 2048, 128, kLayoutNone);
            mma_chain(tmem_base, dQ, umma_desc(smem_u32(sK), a.L_pad * 16, 128, kLayoutNone), 2 * 2048, 2 * a.L_pad * 16, a.d_pad / 16,
                      umma_idesc_f16(FMT, 128, a.L_pad), 0);
            umma_commit(&bars[B_S_FULL]);
            H1_T(4);
            mbar_wait_spin(&bars[B_V_FULL], 0);
            mbar_wait_spin(&bars[B_P_FULL], 0);
            tc_fence_after();
            H1_T(5);
            {
                const uint32_t idesc_v = umma_idesc_f16(FMT, 128, a.dv_pad);
                const uint64_t dP = umma_desc(smem_u32(sP), 2048, 128, kLayoutNone), dV = umma_desc(smem_u32(sV), a.dv_pad * 16, 128, kLayoutNone);
                mma_chain(tmem_O, dP, dV, 2 * 2048, 2 * a.dv_pad * 16, a.Lt_pad / 16, idesc_v, 0);
                mma_chain(tmem_O + a.dv_pad, desc_add(dP, (a.Lt_pad / 8) * 2048), desc_add(dV, (a.Lt_pad / 8) * (a.dv_pad * 16)), 2 * 2048,
                          2 * a.dv_pad * 16, a.La_pad / 16, idesc_v, 0);
            }
            umma_commit(&bars[B_O_FULL]);
            umma_commit(&bars[B_PV_ALL]);
            H1_T(6);
        }
        __syncwarp();
        if (clustered) { cluster_wait(); if (a.overlay) { cluster_arrive(); cluster_wait(); } }   // #1; #2 = every CTA's O tile region is free to receive
        if (elect_one()) {
            // ---- phase C: Y[:, own d columns] = O · Wo_h^T over the gathered O tile
            H1_T(7);
            mbar_wait_spin(&bars[B_SO], 0);       // own O_h is in its slot, the accumulators have been read
            H1_T(8);
            if (clustered) mbar_wait_spin(&bars[B_OALL], 0);   // every peer's slot has landed
            tc_fence_after();
            H1_T(9);
            const uint32_t idesc_y = umma_idesc_f16(FMT, 128, a.d_pad);
            const uint64_t dO = umma_desc(smem_u32(sO), 2048, 128, kLayoutNone);
#pragma unroll 1
            for (int i = 0; i < a.n_w; ++i) {
                const int slot = i % a.nw_slots;
                const int ks = min(KS, a.n_ko - i * KS);
                mbar_wait_spin(&bars[B_W_FULL + slot], (i / a.nw_slots) & 1);
                tc_fence_after();
                if (i == 0) H1_T(10);
                mma_chain(tmem_base, desc_add(dO, i * KS * 2 * 2048), umma_desc(smem_u32(smem + a.wo_off[slot]), a.d_pad * 16, 128, kLayoutNone),
                          2 * 2048, 2 * a.d_pad * 16, ks, idesc_y, static_cast<uint32_t>(i));
                if (i + a.nw_slots < a.n_w) umma_commit(&bars[B_W_EMPTY + slot]);
            }
            umma_commit(&bars[B_Y]);
            H1_T(11);
        }
        __syncwarp();
    } else {
        // ================================================================ softmax / conversion / epilogue warpgroup
        const int row = warp * 32 + lane;          // tile row == TMEM lane
        const uint32_t lane_taddr = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t slot_row = smem_u32(sSlot) + row * 16;   // element (row, c) of a K-major core-matrix tile: (c / 8) * 2048 + row * 16 + (c % 8) * 2
        const uint32_t sP_row = smem_u32(sP) + row * 16;
        const float sl = a.scale_log2;
        const int la_end = a.Lt_pad + a.La;
        const int n16q = a.d_pad / 16;             // 2 .. 7

        // ---- phase A: Q accumulator -> 16-bit operand tile (own slot)
        const bool t0 = threadIdx.x == 0;
        if (warp < first_wave && elect_one()) load_a(warp);   // the first phase-A chunks go out from four warps at once
        __syncwarp();
        if (t0) H1_T(24);
        mbar_wait_nocall(&bars[B_Q], 0);
        tc_fence_after();
        if (t0) H1_T(25);
        {
            uint32_t r[7 * 16];
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (j < n16q) tmem_ld16(tmem_base + lane_taddr + j * 16, r + j * 16);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (j < n16q) { reg_tie16(r + j * 16); cvt_store16<FMT>(r + j * 16, slot_row + j * 2 * 2048); }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[B_SQ]);
        if (t0) H1_T(26);

        // ---- phase B: two softmaxes over one score row: text columns [0, Lt), audio columns [Lt_pad, Lt_pad + La); P unnormalised
        const uint32_t tS = tmem_base + lane_taddr;
        float mx_t = -INFINITY, mx_a = -INFINITY;
        mbar_wait_nocall(&bars[B_S_FULL], 0);
        tc_fence_after();
        if (t0) H1_T(27);
        if (N16 > 0) {
            uint32_t s[(N16 > 0 ? N16 : 1) * 16];
#pragma unroll
            for (int c = 0; c < N16; ++c) tmem_ld16(tS + c * 16, s + c * 16);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < N16; ++c) reg_tie16(s + c * 16);
            if (LT16 > 0 && a.Lt < LT16 * 16) mask16(s + (LT16 > 0 ? LT16 - 1 : 0) * 16, a.Lt - (LT16 - 1) * 16);
            if (N16 > LT16 && la_end < N16 * 16) mask16(s + (N16 - 1) * 16, la_end - (N16 - 1) * 16);
#pragma unroll
            for (int c = 0; c < N16; ++c) {
                if (c < LT16) mx_t = fmaxf(mx_t, max16(s + c * 16)); else mx_a = fmaxf(mx_a, max16(s + c * 16));
            }
            const float mxs_t = mx_t * sl, mxs_a = mx_a * sl;
            if (t0) H1_T(28);
#pragma unroll
            for (int c = 0; c < N16; ++c) exp_store16<FMT>(s + c * 16, sl, c < LT16 ? mxs_t : mxs_a, sP_row + c * 2 * 2048);
        } else {
#pragma unroll 1
            for (int c = 0; c < a.L_pad / 16; ++c) {   // pass 1 over tensor memory: maxima
                uint32_t r[16];
                tmem_ld16(tS + c * 16, r);
                tmem_ld_wait16(r);
                const bool text = c * 16 < a.Lt_pad;
                mask16(r, (text ? a.Lt : la_end) - c * 16);
                const float m = max16(r);
                if (text) mx_t = fmaxf(mx_t, m); else mx_a = fmaxf(mx_a, m);
            }
            const float mxs_t = mx_t * sl, mxs_a = mx_a * sl;
#pragma unroll 1
            for (int c = 0; c < a.L_pad / 16; ++c) {   // pass 2: exponentials
                uint32_t r[16];
                tmem_ld16(tS + c * 16, r);
                tmem_ld_wait16(r);
                const bool text = c * 16 < a.Lt_pad;
                mask16(r, (text ? a.Lt : la_end) - c * 16);
                exp_store16<FMT>(r, sl, text ? mxs_t : mxs_a, sP_row + c * 2 * 2048);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[B_P_FULL]);
        if (t0) H1_T(29);

        // ---- O_h = O_t / l_t + ap_scale * O_a / l_a -> 16 bit, over Q_h in the own slot; l_t, l_a = accumulator column d_pad
        //      (ones row of the extended V^T tile).  A segment that does not exist reads the other accumulator with weight 0.
        {
            const uint32_t tO = tmem_O + lane_taddr;
            const uint32_t tOt = a.Lt > 0 ? tO : tO + a.dv_pad, tOa = a.La > 0 ? tO + a.dv_pad : tO;
            mbar_wait_nocall(&bars[B_O_FULL], 0);
            tc_fence_after();
            if (t0) H1_T(30);
            uint32_t lt, la;
            tmem_ld1(tOt + a.d_pad, lt);
            tmem_ld1(tOa + a.d_pad, la);
            float w_t = 0.f, w_a = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < a.d_pad; c0 += 32) {   // two 16-column groups of each accumulator per round trip
                uint32_t rt[32], ra[32];
                const bool two = c0 + 16 < a.d_pad;
                tmem_ld16(tOt + c0, rt);
                tmem_ld16(tOa + c0, ra);
                if (two) { tmem_ld16(tOt + c0 + 16, rt + 16); tmem_ld16(tOa + c0 + 16, ra + 16); }
                tmem_ld_wait();
                if (c0 == 0) {
                    asm volatile("" : "+r"(lt), "+r"(la) : : "memory");
                    w_t = a.Lt > 0 ? 1.0f / __uint_as_float(lt) : 0.f;
                    w_a = a.La > 0 ? a.ap_scale / __uint_as_float(la) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j == 0 || two) {
                        reg_tie16(rt + j * 16);
                        reg_tie16(ra + j * 16);
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(ra[j * 16 + i]), w_a, __uint_as_float(rt[j * 16 + i]) * w_t);
                        store16<FMT>(slot_row + (c0 + j * 16) / 8 * 2048, v);
                    }
                }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[B_SO]);
        if (t0) H1_T(31);
        if (clustered) {
            // all-gather of the O slots over distributed shared memory: push the own slot into the same slot of every peer.  Overlay
            // plan: only once every CTA of the cluster has finished with the shared memory its O tile region overlays (#2).
            cluster_wait();      // #1: the peers' barriers exist
            if (a.overlay) {
                cluster_arrive();    // #2 (release: orders the st.shared of the slot before the pusher's view of it)
                cluster_wait();
            } else {
                named_bar_sync(1, 128);   // every row of the slot is written (and fenced to the async proxy)
            }
            if (t0) H1_T(32);
            if (elect_one()) {   // warp w pushes to ranks w and w + 4: the copies are issued from four warps at once
                const uint32_t slot = smem_u32(sSlot), bar = smem_u32(&bars[B_OALL]);
#pragma unroll 1
                for (int r = warp; r < a.H; r += 4)
                    if (r != h) dsmem_bulk_copy(map_to_cta(slot, r), sSlot, a.slot_bytes, map_to_cta(bar, r));
            }
            __syncwarp();
            if (t0) H1_T(33);
        }

        // ---- phase C epilogue: Y accumulator (+ bias) -> 16 bit -> core-matrix staging tile -> one 4-D tensor-TMA store.
        //      The staging tile is the slot of the NEXT rank: received, consumed by the MMAs (B_Y), and never the source of a push.
        mbar_wait_nocall(&bars[B_Y], 0);
        tc_fence_after();
        if (t0) H1_T(34);
        {
            uint8_t* sStage = sO + ((h + 1) % a.H) * a.slot_bytes;
            const uint32_t stage_row = smem_u32(sStage) + row * 16;
            uint32_t r[7 * 16];
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (j < n16q) tmem_ld16(tmem_base + lane_taddr + j * 16, r + j * 16);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (j < n16q) {
                    reg_tie16(r + j * 16);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[j * 16 + i]) + sBias[j * 16 + i];
                    store16<FMT>(stage_row + j * 2 * 2048, v);
                }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (threadIdx.x == 0) {
                // accumulate: Y holds the block's residual stream, the tile is added to it (one 16-bit rounding, as h + y)
                if (a.accumulate) tma_reduce_add_4d(&tmY, sStage, 0, m0, h * (a.d / 8), b);
                else tma_store_4d(&tmY, sStage, 0, m0, h * (a.d / 8), b);
                tma_store_commit();
                H1_T(35);
                tma_store_wait_read();
                H1_T(36);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem_base);
}

struct H1Plan {
    bool ok;
    int ns_a, n_w, nw_slots, wo_early, p_in_r0, overlay;
    uint32_t off_a, off_p, off_k, off_v, off_bar, a_stage_bytes, p_bytes, wo_stage_bytes, smem_bytes, wo_off[MAX_W];
};

H1Plan h1_plan(const H1Geom& g, int Lt, int La) {
    H1Plan p;
    memset(&p, 0, sizeof(p));
    if (!g.ok || Lt < 0 || La < 0 || Lt + La <= 0) return p;
    const KvPackGeom kg = kv_pack_geom(Lt, La, g.C, g.H);
    if (kg.L_pad < 16 || kg.L_pad > 256) return p;
    if (kg.L_pad + 2 * kg.dv_pad > 512) return p;           // tensor-memory columns of phase B
    const uint32_t slot = static_cast<uint32_t>(g.d_pad) * 256, r0 = static_cast<uint32_t>(g.Nq) * 256;
    p.p_bytes = 128u * kg.L_pad * 2;
    p.a_stage_bytes = static_cast<uint32_t>(align_up(X_CHUNK_BYTES + static_cast<size_t>(g.d_pad) * 128, 1024));   // SWIZZLE_128B boxes: 1024-aligned
    p.wo_stage_bytes = KS * static_cast<uint32_t>(g.d_pad) * 32;
    p.n_w = (g.Nq / 16 + KS - 1) / KS;
    const uint32_t limit = 227u * 1024u - 1024u;            // minus the alignment slack of the dynamic shared-memory base
    const uint32_t tail = BAR_REGION_BYTES + static_cast<uint32_t>(g.d_pad) * 4;
    const uint32_t kv = static_cast<uint32_t>(align_up(kg.k_tile_bytes, 128) + align_up(kg.vx_tile_bytes, 1024));
    // Wo ring over [ovl_begin, ovl_end) (dead after the last P·V) + up to two stages of their own behind `off`
    auto finish = [&](uint32_t off, uint32_t ovl_begin, uint32_t ovl_end) -> bool {
        if (off + tail > limit) return false;
        int early = static_cast<int>(std::min<uint32_t>((limit - tail - off) / p.wo_stage_bytes, 2));
        early = std::min(early, p.n_w);
        int late = static_cast<int>((ovl_end - ovl_begin) / p.wo_stage_bytes);
        late = std::min(late, MAX_W - early);
        late = std::min(late, p.n_w - early);
        if (early + late < std::min(2, p.n_w)) return false;
        p.wo_early = early; p.nw_slots = early + late;
        for (int i = 0; i < early; ++i) { p.wo_off[i] = off; off += p.wo_stage_bytes; }
        for (int i = 0; i < late; ++i) p.wo_off[early + i] = ovl_begin + i * p.wo_stage_bytes;
        p.off_bar = static_cast<uint32_t>(align_up(off, 128));
        p.smem_bytes = p.off_bar + tail + 1024;
        return p.off_bar + tail <= limit;
    };
    // plan 1 (no overlay): [O tile][phase-A ring | P tile][K][V][early Wo]: nothing but its owner and the peers' pushes ever writes the
    // O tile, so a slot is pushed the moment it exists
    for (int ns = std::min(4, g.n_kb); ns >= std::min(2, g.n_kb); --ns) {
        p.overlay = 0; p.p_in_r0 = 0; p.ns_a = ns;
        p.off_a = r0; p.off_p = r0;
        uint32_t off = r0 + static_cast<uint32_t>(align_up(std::max<uint32_t>(ns * p.a_stage_bytes, p.p_bytes), 1024));
        p.off_k = off; p.off_v = off + static_cast<uint32_t>(align_up(kg.k_tile_bytes, 128));
        off += kv;
        if (finish(off, r0, off)) { p.ok = true; return p; }
    }
    // plan 2 (overlay): the phase-A ring lies over the O tile, and so does the P tile when the larger half next to the own slot holds it
    for (int in_r0 = 1; in_r0 >= 0; --in_r0) {
        if (in_r0 && !(g.H >= 2 && static_cast<uint32_t>(g.H / 2) * slot >= p.p_bytes)) continue;
        p.overlay = 1; p.p_in_r0 = in_r0;
        uint32_t off = r0;
        p.off_a = 0; p.off_p = off;
        if (!in_r0) off += static_cast<uint32_t>(align_up(p.p_bytes, 1024));
        const uint32_t a_limit = off;                        // the phase-A ring may use [0, a_limit)
        p.off_k = off; p.off_v = off + static_cast<uint32_t>(align_up(kg.k_tile_bytes, 128));
        off += kv;
        p.ns_a = static_cast<int>(std::min<uint32_t>(std::min<uint32_t>(MAX_A, g.n_kb), a_limit / p.a_stage_bytes));
        if (p.ns_a < std::min(2, g.n_kb)) continue;
        if (finish(off, in_r0 ? p.off_k : p.off_p, off)) { p.ok = true; return p; }
    }
    p.ok = false;
    return p;
}

}  // namespace

bool ipattn_site_h1_supported(int C, int H, int Lt, int La) {
    const H1Geom g = h1_geom(C, H);
    return g.ok && h1_plan(g, Lt, La).ok;
}

int ipattn_site_h1_pack_weights(const void* Wq, const void* Wo, void* pack, int C, int H, int dtype, cudaStream_t st) {
    const H1Geom g = h1_geom(C, H);
    APX_REQUIRE(g.ok, APX_EUNSUPPORTED, "ipattn_site_h1_pack: C=%d H=%d is not covered", C, H);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "ipattn_site_h1_pack: dtype %d", dtype);
    h1_pack_kernel<<<148 * 8, 256, 0, st>>>(static_cast<const uint16_t*>(Wq), static_cast<const uint16_t*>(Wo), static_cast<uint16_t*>(pack), g);
    APX_LAUNCH_CHECK("h1_pack_kernel");
    return APX_OK;
}

int ipattn_site_h1(const void* X, const void* wpack, const void* cache_tiles, const float* bo, void* Y, int Bc, int S, int C, int H,
                   int Lt, int La, float ap_scale, int dtype, bool accumulate, cudaStream_t st) {
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "ipattn_site_h1: dtype %d", dtype);
    const H1Geom g = h1_geom(C, H);
    APX_REQUIRE(g.ok, APX_EUNSUPPORTED, "ipattn_site_h1: C=%d H=%d is not covered", C, H);
    const H1Plan p = h1_plan(g, Lt, La);
    APX_REQUIRE(p.ok, APX_EUNSUPPORTED, "ipattn_site_h1: context %d+%d does not fit at C=%d H=%d", Lt, La, C, H);
    APX_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wpack) & 15) == 0 && (reinterpret_cast<uintptr_t>(cache_tiles) & 15) == 0,
                APX_EINVAL, "ipattn_site_h1: X / Y / pack / cache must be 16-byte aligned");
    APX_REQUIRE(Bc <= 65535, APX_EUNSUPPORTED, "ipattn_site_h1: Bc=%d", Bc);
    const KvPackGeom kg = kv_pack_geom(Lt, La, C, H);

    CUtensorMap tmX, tmY;
    {
        // X: [Bc][S][C] as a 3-D tensor, 64-channel SWIZZLE_128B boxes (rows of a tile past S are zero-filled)
        uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(S), static_cast<uint64_t>(Bc)};
        uint64_t strides[2] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(S) * C * 2};
        uint32_t box[3] = {64, 128, 1};
        int rc = encode_tmap(&tmX, dtype, 3, X, dims, strides, box, 1);
        if (rc) return rc;
        // Y: [Bc][C/8][S][8] view, box {8, 128, d/8, 1} = a K-major core-matrix tile of the head's d columns (rows past S clipped)
        uint64_t dims4[4] = {8, static_cast<uint64_t>(S), static_cast<uint64_t>(C / 8), static_cast<uint64_t>(Bc)};
        uint64_t strides4[3] = {static_cast<uint64_t>(C) * 2, 16, static_cast<uint64_t>(S) * C * 2};
        uint32_t box4[4] = {8, 128, static_cast<uint32_t>(g.d / 8), 1};
        rc = encode_tmap(&tmY, dtype, 4, Y, dims4, strides4, box4, 0);
        if (rc) return rc;
    }
    H1Args a;
    memset(&a, 0, sizeof(a));
    const uint8_t* pb = static_cast<const uint8_t*>(wpack);
    a.wq = pb + g.wq_off;
    a.wo = pb + g.wo_off;
    a.tiles = static_cast<const uint8_t*>(cache_tiles);
    a.bo = bo;
    a.S = S; a.C = C; a.H = H; a.d = g.d; a.d_pad = g.d_pad; a.dv_pad = kg.dv_pad; a.Nq = g.Nq; a.n_kb = g.n_kb;
    a.Lt = Lt; a.La = La; a.Lt_pad = kg.Lt_pad; a.La_pad = kg.La_pad; a.L_pad = kg.L_pad;
    a.ap_scale = ap_scale;
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(g.d));
    a.kv_bytes = static_cast<uint32_t>(kg.bytes_per_bh);
    a.k_tile_bytes = static_cast<uint32_t>(kg.k_tile_bytes);
    a.v_tile_bytes = static_cast<uint32_t>(kg.vx_tile_bytes);   // the EXTENDED V^T tile (ones row -> softmax denominators)
    a.vx_off = static_cast<uint32_t>(kg.k_tile_bytes + kg.v_tile_bytes);
    a.wq_chunk_bytes = static_cast<uint32_t>(g.d_pad) * 128;
    a.wo_stage_bytes = p.wo_stage_bytes;
    a.a_stage_bytes = p.a_stage_bytes;
    a.p_bytes = p.p_bytes;
    a.slot_bytes = static_cast<uint32_t>(g.d_pad) * 256;
    a.ns_a = p.ns_a; a.n_w = p.n_w; a.nw_slots = p.nw_slots; a.wo_early = p.wo_early; a.n_ko = g.Nq / 16;
    a.p_in_r0 = p.p_in_r0; a.overlay = p.overlay;
    a.off_a = p.off_a; a.off_p = p.off_p; a.off_k = p.off_k; a.off_v = p.off_v; a.off_bar = p.off_bar;
    for (int i = 0; i < MAX_W; ++i) a.wo_off[i] = p.wo_off[i];
    a.accumulate = accumulate ? 1 : 0;
    // diagnostic timeline (APX_SITE_TRACE=1): synchronises and prints, never enabled on the product path
    static const bool trace_on = [] { const char* e = getenv("APX_SITE_TRACE"); return e && e[0] == '1'; }();
    static long long* trace_buf = nullptr;
    if (trace_on) {
        if (!trace_buf) APX_CUDA(cudaMalloc(&trace_buf, 64 * sizeof(long long)));
        APX_CUDA(cudaMemsetAsync(trace_buf, 0, 64 * sizeof(long long), st));
        a.trace = trace_buf;
    }

    // kernel variant: score rows of 32 / 48 / 144 columns (La = 8 / 32 / 128 behind 1..16 text tokens) stay in registers
    const int n16 = kg.L_pad / 16;
    const int variant = kg.Lt_pad != 16 ? 0 : n16 == 2 ? 1 : n16 == 3 ? 2 : n16 == 9 ? 3 : 0;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, H1Args);
    static const KernelFn table[2][4] = {
        {ipattn_site_h1_kernel<0, 0, 0>, ipattn_site_h1_kernel<0, 2, 1>, ipattn_site_h1_kernel<0, 3, 1>, ipattn_site_h1_kernel<0, 9, 1>},
        {ipattn_site_h1_kernel<1, 0, 0>, ipattn_site_h1_kernel<1, 2, 1>, ipattn_site_h1_kernel<1, 3, 1>, ipattn_site_h1_kernel<1, 9, 1>}};
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j)
                APX_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(table[i][j]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const KernelFn fn = table[dtype == APX_F16 ? 0 : 1][variant];
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(H) * ceil_div(S, 128), Bc);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = static_cast<unsigned>(H);
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = H > 1 ? 1 : 0;
    // no programmatic-launch attribute on the site kernel itself: its predecessor in the step is a PyTorch kernel (nothing to overlap), and
    // back-to-back site launches overlapping each other would flatter the kernel's stand-alone timing.  It still TRIGGERS its dependents.
    APX_CUDA(cudaLaunchKernelEx(&cfg, fn, tmX, tmY, a));
    APX_LAUNCH_CHECK("ipattn_site_h1_kernel");
    if (trace_on) {
        static int printed = 0;
        long long t[64];
        APX_CUDA(cudaStreamSynchronize(st));
        APX_CUDA(cudaMemcpy(t, trace_buf, sizeof(t), cudaMemcpyDeviceToHost));
        if (printed++ < 3) {
            long long t0 = t[0];
            for (int i = 0; i < 64; ++i) if (t[i] && t[i] < t0) t0 = t[i];
            auto rel = [&](int slot) { return t[slot] ? t[slot] - t0 : -1; };
            printf("[apx h1 trace] Bc=%d S=%d C=%d H=%d Lt=%d La=%d ns_a=%d nw_slots=%d/%d early=%d p_in_r0=%d overlay=%d smem=%u (cycles since the first stamp; CTA 0)\n", Bc, S, C, H,
                   Lt, La, p.ns_a, p.nw_slots, p.n_w, p.wo_early, p.p_in_r0, p.overlay, p.smem_bytes);
            printf("  mma : start %lld a_full0 %lld q_committed %lld sq+k_seen %lld s_issued %lld p+v_seen %lld pv_issued %lld | cluster2 %lld so_seen %lld oall_seen %lld wo_full0 %lld y_committed %lld\n",
                   rel(0), rel(1), rel(2), rel(3), rel(4), rel(5), rel(6), rel(7), rel(8), rel(9), rel(10), rel(11));
            printf("  tma : start %lld first_loads_issued %lld pv_all_seen %lld done %lld\n", rel(16), rel(17), rel(18), rel(19));
            printf("  wg  : top %lld q_seen %lld sq_arrived %lld s_seen %lld max_done %lld p_arrived %lld o_full_seen %lld so_arrived %lld cluster2 %lld pushed %lld | y_seen %lld store_issued %lld store_read %lld\n",
                   rel(24), rel(25), rel(26), rel(27), rel(28), rel(29), rel(30), rel(31), rel(32), rel(33), rel(34), rel(35), rel(36));
        }
    }
    return APX_OK;
}

}  // namespace apx
