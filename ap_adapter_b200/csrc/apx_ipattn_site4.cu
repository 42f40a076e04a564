// apx_ipattn_site4.cu — the all-heads adapter-site kernel (apx_ipattn_site.cu) with FOUR softmax warpgroups, for full grids.
//
//   Y = ( softmax(Q K_t^T / sqrt(d)) V_t + ap_scale * softmax(Q K_a^T / sqrt(d)) V_a ) Wo^T + bo,   Q = X Wq^T
//   (APadapter/ap_adapter/attention_processor.py:387-459), context = 1..16 text tokens + 113..128 audio tokens (time/freq pooling 2).
//
// Why: at batch 8 (Bc = 16) the S = 1000, C = 256 site is 128 query tiles on 148 SMs — one tile per SM, so the kernel's time IS the
// latency of one tile.  With two softmax warpgroups (one score row of 144 columns per thread) each SM sub-partition holds two
// softmax warps and its MUFU pipe (the exp2 of 147 k scores per tile: 9.2 k cycles at 16/clk/SM) is busy half of phase B: while a
// warp loads scores, takes maxima or converts the previous head's output, nothing feeds the pipe (profiles/r1e_ncu_site_b8_raw.csv:
// XU 25 %, tensor 18 %).  Here a score row is split over TWO threads (columns [0, 80) = text + half of the audio tokens, and
// [80, 144)), so four warpgroups fit in the register file (112 registers each, setmaxnreg) and four warps per sub-partition take turns on the
// MUFU pipe; the same four warpgroups convert Q / O and run the output epilogue four boxes at a time.  Heads still alternate
// between two slots (S / O accumulators and a P tile each); slot s is served by the warpgroup pair (s, s + 2).  The two halves of a
// row exchange their audio maxima through shared memory (one 256-thread named barrier per head).
//
// Warp roles (640 threads): warps 0-15 = softmax warpgroups 0-3 (TMEM lane quarter = warp % 4), warp 16 = TMEM allocator +
// single-thread MMA issuer, warp 17 = TMA producer, warps 18-19 idle (they complete the fifth warpgroup for setmaxnreg).
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include "apx_site_common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace apx {
namespace {

constexpr int X_CHUNK_BYTES = 128 * 64 * 2;   // one SWIZZLE_128B box: 128 rows x 64 channels
constexpr int WO_KSTEPS_PER_STAGE = 4;
constexpr int NTHREADS = 640;
// setmaxnreg moves registers inside the CTA's LAUNCH allocation only (640 threads x 96 registers, the count ptxas reports for this
// kernel): what the fifth warpgroup gives up, 128 x (96 - 32), is exactly what the four softmax warpgroups claim, 512 x (112 - 96).
// (An inc that asks for more than has been released blocks forever.)
constexpr int REGS_LAUNCH = 96;
constexpr int REGS_WG = 112;
constexpr int REGS_AUX = 32;
static_assert(128 * (REGS_LAUNCH - REGS_AUX) >= 512 * (REGS_WG - REGS_LAUNCH), "setmaxnreg budget");
constexpr int NL16 = 5, NR16 = 4;      // 16-column groups of the left half (group 0 = text) and of the right half

struct Site4Args {
    const uint8_t* wq;     // packed WQ section (cluster size 1)
    const uint8_t* wo;     // packed WO section (cluster size 1)
    const uint8_t* tiles;  // packed context K / V^T tiles
    const float* bo;       // fp32[C] or null
    int S, C, H, d_pad, dv_pad, Nq, n_kb;
    int Lt, La;
    float ap_scale, scale_log2;
    uint32_t kv_bytes, k_tile_bytes, v_tile_bytes, vx_off, wq_chunk_bytes, wo_kstep_bytes, a_stage_bytes, p_bytes;
    int nk, nv, n_wo, wo_early;
    uint32_t off_p, off_k, off_v, off_bar;
    uint32_t wo_off[4];
    int accumulate;
    long long* trace;      // diagnostic (APX_SITE_TRACE=1): clock64 stamps of CTA (0,0); null on the product path
};
// trace slots: [0,8) MMA thread phases, [8,24) MMA per head (p_full seen, pv issued), [24,32) producer, [32,96) left warpgroup of slot 0
// (8 per head), [96,104) epilogue
#define S4_T(slot_)                                                                      \
    do {                                                                                 \
        if (a.trace && blockIdx.x == 0 && blockIdx.y == 0) a.trace[(slot_)] = clock64(); \
    } while (0)

enum {  // mbarrier slots (40 x 8 bytes reserved)
    B_A_FULL = 0, B_K_FULL = 4, B_K_EMPTY = 7, B_V_FULL = 10, B_V_EMPTY = 13, B_WO_FULL = 16, B_Q = 20, B_SQ = 21, B_S_FULL = 22,
    B_S_FREE = 24, B_P_FULL = 26, B_O_FULL = 28, B_PV_ALL = 30, B_SO_ALL = 31, B_Y = 32, B_COUNT = 33
};
constexpr uint32_t OFF_TMEM_SLOT = 320, OFF_BIAS = 384, OFF_MAX = 384 + 1024, TAIL_BYTES = 384 + 1024 + 4096;
constexpr int L_PAD = 144, LT_PAD = 16, LA_PAD = 128;

template <int FMT>
__global__ void __launch_bounds__(NTHREADS, 1)
ipattn_site4_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, Site4Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQO = smem;                 // [Nq/8][128][8] 16 bit: Q of every head, overwritten head by head with O; then the A operand of phase C
    uint8_t* sP = smem + a.off_p;        // P tiles of the two slots
    uint8_t* sK = smem + a.off_k;        // K tile ring
    uint8_t* sV = smem + a.off_v;        // V^T tile ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.off_bar + OFF_TMEM_SLOT);
    float* sBias = reinterpret_cast<float*>(smem + a.off_bar + OFF_BIAS);
    float* sMax = reinterpret_cast<float*>(smem + a.off_bar + OFF_MAX);   // [2 parities][2 slots][2 halves][128 rows]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int b = blockIdx.y;

    if (threadIdx.x == 544) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
#pragma unroll 1
        for (int i = 0; i < B_COUNT; ++i) {
            const uint32_t cnt = (i == B_SQ || i == B_SO_ALL) ? 512u : (i >= B_S_FREE && i < B_O_FULL) ? 256u : 1u;
            mbar_init(&bars[i], cnt);   // warpgroup -> MMA barriers count every arriving thread
        }
        fence_mbar_init();
    }
    if (warp == 16) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < a.C; i += NTHREADS) sBias[i] = a.bo ? a.bo[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // phase A: Q at [0, Nq).  phase B: S(slot) at slot * 144, then {O_t, O_a}(slot), dv_pad columns each.  phase C: Y at [0, C).
    const uint32_t tmem_O = tmem_base + 2 * L_PAD;
    const int n_ko = a.Nq / 16;

    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
        if (warp == 17) {
            // ================================================================ TMA producer (one thread)
            if (elect_one()) {
#pragma unroll 1
                for (int kb = 0; kb < a.n_kb; ++kb) {   // every phase-A chunk has a stage of its own: all loads are in flight at once
                    uint8_t* dst = smem + kb * a.a_stage_bytes;
                    mbar_expect_tx(&bars[B_A_FULL + kb], X_CHUNK_BYTES + a.wq_chunk_bytes);
                    tma_load_3d(dst, &tmX, kb * 64, m0, b, &bars[B_A_FULL + kb]);
                    tma_load_1d(dst + X_CHUNK_BYTES, a.wq + static_cast<size_t>(kb) * a.wq_chunk_bytes, a.wq_chunk_bytes, &bars[B_A_FULL + kb]);
                }
                if (static_cast<uint32_t>(a.n_kb) * a.a_stage_bytes > a.off_k) mbar_wait_nocall(&bars[B_Q], 0);
                const uint8_t* kv_src = a.tiles + static_cast<size_t>(b) * a.H * a.kv_bytes;
                auto load_wo = [&](int w) {
                    const int ks = min(WO_KSTEPS_PER_STAGE, n_ko - w * WO_KSTEPS_PER_STAGE);
                    mbar_expect_tx(&bars[B_WO_FULL + w], ks * a.wo_kstep_bytes);
                    tma_load_1d(smem + a.wo_off[w], a.wo + static_cast<size_t>(w) * WO_KSTEPS_PER_STAGE * a.wo_kstep_bytes, ks * a.wo_kstep_bytes,
                                &bars[B_WO_FULL + w]);
                };
                // issue order = the MMA thread's consumption order: K(2p), K(2p+1), then V(2p), V(2p+1)
                int sk = 0, sv = 0;
                uint32_t pk = 1, pv = 1;
#pragma unroll 1
                for (int h0 = 0; h0 < a.H; h0 += 2) {
#pragma unroll 1
                    for (int h = h0; h < h0 + 2 && h < a.H; ++h) {
                        mbar_wait_nocall(&bars[B_K_EMPTY + sk], pk);
                        mbar_expect_tx(&bars[B_K_FULL + sk], a.k_tile_bytes);
                        tma_load_1d(sK + sk * a.k_tile_bytes, kv_src + static_cast<size_t>(h) * a.kv_bytes, a.k_tile_bytes, &bars[B_K_FULL + sk]);
                        if (++sk == a.nk) { sk = 0; pk ^= 1; }
                    }
#pragma unroll 1
                    for (int h = h0; h < h0 + 2 && h < a.H; ++h) {
                        mbar_wait_nocall(&bars[B_V_EMPTY + sv], pv);
                        mbar_expect_tx(&bars[B_V_FULL + sv], a.v_tile_bytes);
                        tma_load_1d(sV + sv * a.v_tile_bytes, kv_src + static_cast<size_t>(h) * a.kv_bytes + a.vx_off, a.v_tile_bytes,
                                    &bars[B_V_FULL + sv]);
                        if (++sv == a.nv) { sv = 0; pv ^= 1; }
                    }
                    if (h0 == 0) {
#pragma unroll 1
                        for (int w = 0; w < a.wo_early; ++w) load_wo(w);
                    }
                }
                mbar_wait_nocall(&bars[B_PV_ALL], 0);   // the P tiles and the K / V rings are dead once every P·V has completed
#pragma unroll 1
                for (int w = a.wo_early; w < a.n_wo; ++w) load_wo(w);
            }
        } else if (warp == 16) {
            // ================================================================ MMA issuer (one thread)
            if (elect_one()) {
                S4_T(0);
                const uint32_t idesc_q = umma_idesc_f16(FMT, 128, a.Nq);
                const uint32_t idesc_s = umma_idesc_f16(FMT, 128, L_PAD);
#pragma unroll 1
                for (int kb = 0; kb < a.n_kb; ++kb) {
                    mbar_wait_spin(&bars[B_A_FULL + kb], 0);
                    tc_fence_after();
                    if (kb == 0) S4_T(1);
                    const uint32_t base = smem_u32(smem + kb * a.a_stage_bytes);
                    mma_chain(tmem_base, umma_desc(base, 16, 1024, kLayoutSW128), umma_desc(base + X_CHUNK_BYTES, a.Nq * 16, 128, kLayoutNone), 32,
                              2 * a.Nq * 16, 4, idesc_q, static_cast<uint32_t>(kb));
                }
                umma_commit(&bars[B_Q]);
                S4_T(2);
                const uint64_t dQO = umma_desc(smem_u32(sQO), 2048, 128, kLayoutNone);
                const uint32_t head_bytes = (a.d_pad / 8) * 2048;
                const uint32_t idesc_v = umma_idesc_f16(FMT, 128, a.dv_pad);
                const uint64_t dP0 = umma_desc(smem_u32(sP), 2048, 128, kLayoutNone);
                const uint64_t dK0 = umma_desc(smem_u32(sK), L_PAD * 16, 128, kLayoutNone);
                const uint64_t dV0 = umma_desc(smem_u32(sV), a.dv_pad * 16, 128, kLayoutNone);
                int sk = 0, sv = 0;
                uint32_t pk = 0, pv = 0;
                mbar_wait_spin(&bars[B_SQ], 0);   // every Q column is in sQO; the Q accumulator columns are free
                tc_fence_after();
                S4_T(3);
                // Event-driven issue (as in apx_ipattn_site.cu): two in-order queues, polled without blocking
                {
                    int hs = 0, hp = 0;
                    uint32_t polls = 0;
                    while (hp < a.H) {
                        if (++polls > (1u << 30)) asm volatile("trap;");   // a protocol bug becomes a CUDA error (after ~30 s), not a hung GPU
                        if (hs < a.H) {
                            const int g = hs & 1;
                            if (mbar_test_wait(&bars[B_K_FULL + sk], pk) && (hs < 2 || mbar_test_wait(&bars[B_S_FREE + g], ((hs >> 1) - 1) & 1))) {
                                tc_fence_after();
                                mma_chain(tmem_base + g * L_PAD, desc_add(dQO, hs * head_bytes), desc_add(dK0, sk * a.k_tile_bytes), 2 * 2048,
                                          2 * L_PAD * 16, a.d_pad / 16, idesc_s, 0);
                                umma_commit(&bars[B_S_FULL + g]);
                                umma_commit(&bars[B_K_EMPTY + sk]);
                                if (++sk == a.nk) { sk = 0; pk ^= 1; }
                                ++hs;
                                polls = 0;
                            }
                        }
                        if (hp < hs) {
                            const int g = hp & 1;
                            if (mbar_test_wait(&bars[B_P_FULL + g], (hp >> 1) & 1) && mbar_test_wait(&bars[B_V_FULL + sv], pv)) {
                                tc_fence_after();
                                if (hp < 8) S4_T(8 + 2 * hp);
                                const uint64_t dP = desc_add(dP0, g * a.p_bytes), dV = desc_add(dV0, sv * a.v_tile_bytes);
                                const uint32_t tO = tmem_O + g * 2 * a.dv_pad;
                                mma_chain(tO, dP, dV, 2 * 2048, 2 * a.dv_pad * 16, LT_PAD / 16, idesc_v, 0);
                                mma_chain(tO + a.dv_pad, desc_add(dP, (LT_PAD / 8) * 2048), desc_add(dV, (LT_PAD / 8) * (a.dv_pad * 16)), 2 * 2048,
                                          2 * a.dv_pad * 16, LA_PAD / 16, idesc_v, 0);
                                umma_commit(&bars[B_O_FULL + g]);
                                umma_commit(&bars[B_V_EMPTY + sv]);
                                if (++sv == a.nv) { sv = 0; pv ^= 1; }
                                if (hp < 8) S4_T(9 + 2 * hp);
                                ++hp;
                                polls = 0;
                            }
                        }
                    }
                }
                umma_commit(&bars[B_PV_ALL]);
                // ---- phase C: Y = O · Wo^T
                mbar_wait_spin(&bars[B_SO_ALL], 0);   // every O_h is in sQO and every accumulator has been read
                tc_fence_after();
                S4_T(4);
                const uint32_t idesc_y = umma_idesc_f16(FMT, 128, a.C);
#pragma unroll 1
                for (int i = 0; i < a.n_wo; ++i) {
                    const int ks = min(WO_KSTEPS_PER_STAGE, n_ko - i * WO_KSTEPS_PER_STAGE);
                    mbar_wait_spin(&bars[B_WO_FULL + i], 0);
                    tc_fence_after();
                    if (i == 0) S4_T(5);
                    mma_chain(tmem_base, desc_add(dQO, i * WO_KSTEPS_PER_STAGE * 2 * 2048), umma_desc(smem_u32(smem + a.wo_off[i]), a.C * 16, 128, kLayoutNone),
                              2 * 2048, a.wo_kstep_bytes, ks, idesc_y, static_cast<uint32_t>(i));
                }
                umma_commit(&bars[B_Y]);
                S4_T(6);
            }
        }
    } else {
        // ================================================================ the four softmax / conversion / epilogue warpgroups
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_WG));
        const int w4 = warp >> 2;               // warpgroup 0..3
        const int slot = w4 & 1;                // head slot this warpgroup serves (with warpgroup w4 ^ 2)
        const bool left = w4 < 2;               // left half: score columns [0, 80) = text group + 4 audio groups; right: [80, 144)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // tile row == TMEM lane
        const uint32_t lane_taddr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t sQO_row = smem_u32(sQO) + row * 16;   // element (row, c) of a K-major core-matrix tile: (c / 8) * 2048 + row * 16 + (c % 8) * 2
        const uint32_t sP_row = smem_u32(sP + slot * a.p_bytes) + row * 16;
        const float sl = a.scale_log2;
        const int la_end = LT_PAD + a.La;
        const bool t0 = threadIdx.x == 0;       // the thread that writes the warpgroup trace stamps (left half of slot 0)

        // ---- phase A: Q accumulator -> 16-bit operand tile; the four warpgroups split the 16-column groups
        {
            const int n16 = a.Nq / 16, per = (n16 + 3) / 4;
            const int c_lo = min(w4 * per, n16), c_hi = min(c_lo + per, n16);
            mbar_wait_nocall(&bars[B_Q], 0);
            tc_fence_after();
            if (t0) S4_T(96);
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 4) {
                uint32_t r[64];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c0 + j < c_hi) tmem_ld16(tmem_base + lane_taddr + (c0 + j) * 16, r + j * 16);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c0 + j < c_hi) { reg_tie16(r + j * 16); cvt_store16<FMT>(r + j * 16, sQO_row + (c0 + j) * 2 * 2048); }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&bars[B_SQ]);
            if (t0) S4_T(97);
        }

        // ---- phase B
        const uint32_t tO = tmem_O + lane_taddr + slot * 2 * a.dv_pad;
        const uint32_t tS = tmem_base + lane_taddr + slot * L_PAD;
        const uint32_t tOt = a.Lt > 0 ? tO : tO + a.dv_pad, tOa = tO + a.dv_pad;
        // O_h = O_t / l_t + ap_scale * O_a / l_a -> 16 bit, over Q_h in sQO; l_t, l_a = accumulator column d_pad (ones row of V^T).
        // The pair splits the 16-column groups of the head: even groups left, odd groups right.
        auto convert_o = [&](int h, uint32_t ph) {
            mbar_wait_nocall(&bars[B_O_FULL + slot], ph);
            tc_fence_after();
            uint32_t lt, la;
            tmem_ld1(tOt + a.d_pad, lt);
            tmem_ld1(tOa + a.d_pad, la);
#pragma unroll 1
            for (int c0 = left ? 0 : 16; c0 < a.d_pad; c0 += 32) {
                uint32_t rt[16], ra[16];
                tmem_ld16(tOt + c0, rt);
                tmem_ld16(tOa + c0, ra);
                tmem_ld_wait();
                asm volatile("" : "+r"(lt), "+r"(la) : : "memory");
                const float w_t = a.Lt > 0 ? __fdividef(1.0f, __uint_as_float(lt)) : 0.f;
                const float w_a = __fdividef(a.ap_scale, __uint_as_float(la));
                reg_tie16(rt);
                reg_tie16(ra);
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(ra[i]), w_a, __uint_as_float(rt[i]) * w_t);
                store16<FMT>(sQO_row + (h * a.d_pad + c0) / 8 * 2048, v);
            }
            if (a.d_pad <= 16 && !left) tmem_ld_wait();   // the two single-column loads above must still be retired
        };
        // The exp2 passes of the two slots take turns (named barriers 8 / 9 are the token: "slot 0 / 1 may use the MUFU pipe"): a
        // slot's two warpgroups put two warps on every sub-partition's MUFU pipe, which saturates it (one warp alone reaches ~70 %);
        // the other slot's score loads, maxima, O conversion and barrier waits run underneath.  Every arrive has a taker.
        const bool tok = a.H > 2;
        if (tok && slot == 1) named_bar_arrive(8, 512);
#pragma unroll 1
        for (int it = 0;; ++it) {
            const int h = slot + 2 * it;
            const bool live = h < a.H;
            float mx_t = -INFINITY, mx_a = -INFINITY;
            uint32_t s[NL16 * 16];   // left: groups 0..4 = columns [0, 80); right: s[0..64) = groups 5..8 = columns [80, 144)
            const bool tr = t0 && it < 8;
            if (live) {
                if (tr) S4_T(32 + it * 8 + 0);
                mbar_wait_nocall(&bars[B_S_FULL + slot], it & 1);
                tc_fence_after();
                if (tr) S4_T(32 + it * 8 + 1);
                if (left) {
#pragma unroll
                    for (int c = 0; c < NL16; ++c) tmem_ld16(tS + c * 16, s + c * 16);
                } else {
#pragma unroll
                    for (int c = 0; c < NR16; ++c) tmem_ld16(tS + (NL16 + c) * 16, s + c * 16);
                }
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < NL16; ++c) reg_tie16(s + c * 16);
                tc_fence_before();
                mbar_arrive(&bars[B_S_FREE + slot]);   // the scores of this slot's next head may be issued now
                if (tr) S4_T(32 + it * 8 + 2);
                if (left) {
                    if (a.Lt < 16) mask16(s, a.Lt);
                    mx_t = max16(s);
#pragma unroll
                    for (int c = 1; c < NL16; ++c) mx_a = fmaxf(mx_a, max16(s + c * 16));
                } else {
                    if (la_end < L_PAD) mask16(s + (NR16 - 1) * 16, la_end - (L_PAD - 16));
#pragma unroll
                    for (int c = 0; c < NR16; ++c) mx_a = fmaxf(mx_a, max16(s + c * 16));
                }
                // the two halves of a row exchange their audio maxima (double-buffered by head parity: a thread can run at most one
                // barrier ahead of its partner)
                float* mine = sMax + (((it & 1) * 2 + slot) * 2 + (left ? 0 : 1)) * 128 + row;
                *mine = mx_a;
                named_bar_sync(1 + slot, 256);
                mx_a = fmaxf(mx_a, *(left ? mine + 128 : mine - 128));
                if (tr) S4_T(32 + it * 8 + 3);
            }
            // the O of this slot's previous head: its P·V ran under the code above; also frees this slot's P tile
            if (it > 0) convert_o(h - 2, (it - 1) & 1);
            if (!live) break;
            if (tr) S4_T(32 + it * 8 + 4);
            if (tok) named_bar_sync(8 + slot, 512);
            const float mxs_t = mx_t * sl, mxs_a = mx_a * sl;
            if (left) {
                exp_store16<FMT>(s, sl, mxs_t, sP_row);
#pragma unroll
                for (int c = 1; c < NL16; ++c) exp_store16<FMT>(s + c * 16, sl, mxs_a, sP_row + c * 2 * 2048);
            } else {
#pragma unroll
                for (int c = 0; c < NR16; ++c) exp_store16<FMT>(s + c * 16, sl, mxs_a, sP_row + (NL16 + c) * 2 * 2048);
            }
            if (tok && h + 1 < a.H) named_bar_arrive(9 - slot, 512);   // the other slot's turn
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&bars[B_P_FULL + slot]);
            if (tr) S4_T(32 + it * 8 + 5);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[B_SO_ALL]);
        if (t0) S4_T(98);

        // ---- phase C epilogue: Y accumulator + bias -> 16 bit -> SWIZZLE_128B staging boxes -> TMA store; one 64-column box per warpgroup
        mbar_wait_nocall(&bars[B_Y], 0);
        tc_fence_after();
        if (t0) S4_T(99);
        {
            const uint32_t stage_row = smem_u32(sP) + row * 128;   // staging boxes lie over the dead P / K / V regions
            const int sw = row & 7;
#pragma unroll 1
            for (int kb = w4; kb < a.C / 64; kb += 4) {
                uint32_t r[64];
#pragma unroll
                for (int j = 0; j < 4; ++j) tmem_ld16(tmem_base + lane_taddr + kb * 64 + j * 16, r + j * 16);
                tmem_ld_wait();
                const uint32_t box = stage_row + kb * X_CHUNK_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    reg_tie16(r + j * 16);
                    uint32_t p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 bb = *reinterpret_cast<const float2*>(sBias + kb * 64 + j * 16 + 2 * i);
                        p[i] = pack2<FMT>(__uint_as_float(r[j * 16 + 2 * i]) + bb.x, __uint_as_float(r[j * 16 + 2 * i + 1]) + bb.y);
                    }
                    st_shared_v4(box + (((2 * j) ^ sw) << 4), p[0], p[1], p[2], p[3]);
                    st_shared_v4(box + (((2 * j + 1) ^ sw) << 4), p[4], p[5], p[6], p[7]);
                }
                fence_proxy_async_smem();
                named_bar_sync(3 + w4, 128);
                if (q == 0 && lane == 0) {
                    if (a.accumulate) tma_reduce_add_3d(&tmY, sP + kb * X_CHUNK_BYTES, kb * 64, m0, b);
                    else tma_store_3d(&tmY, sP + kb * X_CHUNK_BYTES, kb * 64, m0, b);
                    tma_store_commit();
                }
            }
            if (t0) S4_T(100);
            if (q == 0 && lane == 0) tma_store_wait_read();
            if (t0) S4_T(101);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc<512>(tmem_base);
}

}  // namespace

bool ipattn_site4_supported(int C, int H, int Lt, int La) {
    const SiteGeom g = site_geom(C, H);
    if (!g.ok) return false;
    const KvPackGeom kg = kv_pack_geom(Lt, La, C, H);
    if (kg.Lt_pad != LT_PAD || kg.La_pad != LA_PAD) return false;
    if (2 * L_PAD + 4 * kg.dv_pad > 512) return false;
    const SitePlan p = site_plan(g, Lt, La, 1);
    return p.ok && p.off_bar + TAIL_BYTES + 1024 <= 227u * 1024u;
}

int ipattn_site4(const void* X, const void* wpack, const void* cache_tiles, const float* bo, void* Y, int Bc, int S, int C, int H,
                 int Lt, int La, float ap_scale, int dtype, bool accumulate, cudaStream_t st) {
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "ipattn_site4: dtype %d", dtype);
    APX_REQUIRE(ipattn_site4_supported(C, H, Lt, La), APX_EUNSUPPORTED, "ipattn_site4: C=%d H=%d context %d+%d is not covered", C, H, Lt, La);
    APX_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wpack) & 15) == 0 && (reinterpret_cast<uintptr_t>(cache_tiles) & 15) == 0,
                APX_EINVAL, "ipattn_site4: X / Y / pack / cache must be 16-byte aligned");
    APX_REQUIRE(Bc <= 65535, APX_EUNSUPPORTED, "ipattn_site4: Bc=%d", Bc);
    const SiteGeom g = site_geom(C, H);
    const SitePlan p = site_plan(g, Lt, La, 1);
    const KvPackGeom kg = kv_pack_geom(Lt, La, C, H);

    CUtensorMap tmX, tmY;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(S), static_cast<uint64_t>(Bc)};
        uint64_t strides[2] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(S) * C * 2};
        uint32_t box[3] = {64, 128, 1};
        int rc = encode_tmap(&tmX, dtype, 3, X, dims, strides, box, 1);
        if (rc) return rc;
        rc = encode_tmap(&tmY, dtype, 3, Y, dims, strides, box, 1);
        if (rc) return rc;
    }
    Site4Args a;
    memset(&a, 0, sizeof(a));
    const uint8_t* pb = static_cast<const uint8_t*>(wpack);
    a.wq = pb + g.wq_off[0];
    a.wo = pb + g.wo_off[0];
    a.tiles = static_cast<const uint8_t*>(cache_tiles);
    a.bo = bo;
    a.S = S; a.C = C; a.H = H; a.d_pad = g.d_pad; a.dv_pad = kg.dv_pad; a.Nq = g.Nq; a.n_kb = g.n_kb;
    a.Lt = Lt; a.La = La;
    a.ap_scale = ap_scale;
    a.accumulate = accumulate ? 1 : 0;
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(g.d));
    a.kv_bytes = static_cast<uint32_t>(kg.bytes_per_bh);
    a.k_tile_bytes = static_cast<uint32_t>(kg.k_tile_bytes);
    a.v_tile_bytes = static_cast<uint32_t>(kg.vx_tile_bytes);
    a.vx_off = static_cast<uint32_t>(kg.k_tile_bytes + kg.v_tile_bytes);
    a.wq_chunk_bytes = static_cast<uint32_t>(g.Nq) * 128;
    a.wo_kstep_bytes = static_cast<uint32_t>(C) * 32;
    a.a_stage_bytes = p.a_stage_bytes;
    a.p_bytes = p.p_bytes;
    a.nk = p.nk; a.nv = p.nv; a.n_wo = p.n_wo; a.wo_early = p.wo_early;
    a.off_p = p.off_p; a.off_k = p.off_k; a.off_v = p.off_v; a.off_bar = p.off_bar;
    for (int i = 0; i < 4; ++i) a.wo_off[i] = p.wo_off[i];
    const uint32_t smem_bytes = p.off_bar + TAIL_BYTES + 1024;
    // diagnostic timeline (APX_SITE_TRACE=1): synchronises and prints, never enabled on the product path
    static const bool trace_on = [] { const char* e = getenv("APX_SITE_TRACE"); return e && e[0] == '1'; }();
    static long long* trace_buf = nullptr;
    if (trace_on) {
        if (!trace_buf) APX_CUDA(cudaMalloc(&trace_buf, 128 * sizeof(long long)));
        APX_CUDA(cudaMemsetAsync(trace_buf, 0, 128 * sizeof(long long), st));
        a.trace = trace_buf;
    }

    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, Site4Args);
    static const KernelFn table[2] = {ipattn_site4_kernel<0>, ipattn_site4_kernel<1>};
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        for (int i = 0; i < 2; ++i)
            APX_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(table[i]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    table[dtype == APX_F16 ? 0 : 1]<<<dim3(ceil_div(S, 128), Bc), NTHREADS, smem_bytes, st>>>(tmX, tmY, a);
    APX_LAUNCH_CHECK("ipattn_site4_kernel");
    if (trace_on) {
        static int printed = 0;
        long long t[128];
        APX_CUDA(cudaStreamSynchronize(st));
        APX_CUDA(cudaMemcpy(t, trace_buf, sizeof(t), cudaMemcpyDeviceToHost));
        if (printed++ < 2) {
            const long long t0 = t[0];
            auto rel = [&](int slot) { return t[slot] ? t[slot] - t0 : -1; };
            printf("[apx site4 trace] Bc=%d S=%d C=%d H=%d Lt=%d La=%d ring=%d n_wo=%d early=%d (cycles since the MMA thread started; CTA 0)\n", Bc, S, C, H, Lt, La,
                   p.nk, p.n_wo, p.wo_early);
            printf("  mma : a_full0 %lld q_committed %lld sq_seen %lld | so_all_seen %lld wo_full0 %lld y_committed %lld\n", rel(1), rel(2), rel(3), rel(4),
                   rel(5), rel(6));
            for (int h = 0; h < H && h < 8; ++h) printf("  mma h%d: p_full_seen %lld pv_issued %lld\n", h, rel(8 + 2 * h), rel(9 + 2 * h));
            printf("  wg0 : q_seen %lld sq_arrived %lld | so_arrived %lld y_seen %lld staged %lld stored %lld\n", rel(96), rel(97), rel(98), rel(99), rel(100), rel(101));
            for (int i = 0; 2 * i < H && i < 8; ++i)
                printf("  wg0 h%d: top %lld s_seen %lld ld_done %lld max_done %lld oprev_done %lld p_arrived %lld\n", 2 * i, rel(32 + i * 8), rel(33 + i * 8),
                       rel(34 + i * 8), rel(35 + i * 8), rel(36 + i * 8), rel(37 + i * 8));
        }
    }
    return APX_OK;
}

}  // namespace apx
