// apx_flash_attn.cu — O = softmax(Q K^T / sqrt(d)) V for the plain attention sites of the UNet (AttnProcessor2_0,
// APadapter/ap_adapter/attention_processor.py:267-274: F.scaled_dot_product_attention without a mask) on tcgen05, streaming the
// keys in blocks of 128 with the running maximum / sum in registers (the self-attention sites have 1000 keys: the score row of
// apx_attn_tc.cu's core kernel would not fit the 512 tensor-memory columns).
//
// One CTA = one (128-query tile, head, batch element); 192 threads:
//   warps 0-3  one query row (= TMEM lane) per thread: two passes over the 128-column score block in tensor memory (row maximum,
//              then exp2 -> unnormalised 16-bit P tile in shared memory), and the running output row in registers
//   warp 4     TMEM allocation + the single MMA-issuing thread
//   warp 5     TMA producer: the Q tile once, then a ring of K / V blocks — 4-D tensor maps over the caller's [B, rows, C] views land
//              each block directly in UMMA core-matrix order ([d/8][128][8]); rows past the end are zero-filled
// Per key block j:   S_j = Q K_j^T (128 x 128 x d, score buffers double-buffered in TMEM)      -> softmax threads
//                    PV_j = P_j V_j (128 x d x 128; V_j is read MN-major: no transposed copy)   -> softmax threads, one block later:
//                    O <- (O + PV_{j-1}) * exp2(m_{j-1} - m_j)
// so the tensor core computes S_{j+1} and P_{j-1} V_{j-1} while the softmax threads work on block j, and the accumulator never
// needs a read-modify-write in tensor memory.
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include "apx_site_common.cuh"
#include <cstdlib>
#include <cstring>

#ifndef APX_FLASH_POLY
#define APX_FLASH_POLY 0
#endif

namespace apx {
namespace {

constexpr int NT = 384;      // warps 0-3 / 4-7: softmax warpgroups A / B; warp 8: TMEM + MMA issue; warp 9: TMA; warps 10-11 idle (they complete
                             // the third warpgroup so that setmaxnreg can move its registers to the softmax warpgroups)
constexpr int REGS_WG = 232; // setmaxnreg: (232 - 168) * 256 registers claimed = (168 - 40) * 128 released; the kernel must not contain a call
constexpr int REGS_AUX = 40;
constexpr int BKV = 128;
constexpr int MAX_ST = 5;     // K / V ring: deep enough that a block's loads start two blocks before its scores are wanted

struct FaArgs {
    uint16_t* O;
    int64_t o_bs, o_rs;      // element strides of the output view
    int S, L, H, d;          // queries, keys, heads, head dim (d % 16 == 0, d <= 80)
    int nblk, stages;
    float scale_log2;        // log2(e) / sqrt(d)
    uint32_t tile_bytes;     // 128 * d * 2: a Q tile, a K block, a V block
    uint32_t off_k, off_v, off_p, off_bar;
    int swz;                 // 0: tiles in no-swizzle core-matrix order (4-D maps, 16-byte pieces); 64 / 128: whole 64- / 128-byte rows, TMA-swizzled
    long long* trace;        // diagnostic (APX_FLASH_TRACE=1): clock64 stamps of thread 0 of CTA (0,0,0), 4 per key block
};

#define FA_TRACE(slot) \
    do { if (a.trace && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) a.trace[(slot)] = clock64(); } while (0)
#define FA_TRACE_MMA(slot) \
    do { if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) a.trace[(slot)] = clock64(); } while (0)

enum { B_Q = 0, B_KV_FULL = 1, B_KV_FREE = 6, B_S_FULL = 11, B_S_FREE = 13, B_P_FULL = 15, B_PV_FULL = 17, B_PV_FREE = 19, B_COUNT = 21 };

__device__ __forceinline__ uint32_t idesc_major(int fmt, int M, int N, int a_mn, int b_mn) {
    return umma_idesc_f16(fmt, M, N) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16);
}

// 2^x for x <= 0 on the FMA pipe (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-4 polynomial (truncation
// 4e-5 relative, far below the 16-bit rounding of P); 2^n by adding n to the exponent field (the low mantissa bits of x + 1.5 * 2^23
// hold n).  Build option (-DAPX_FLASH_POLY=n: every n-th exponential goes this way).  Measured on B200 with n = 4: 12.1 us against
// 11.3 us with all exponentials on MUFU (B = 2, S = 1000, d = 32) — with two warps per scheduler the exp2 pass is bound by issue
// slots, not by the MUFU pipe, and the polynomial costs ten of them per element.  Off by default.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float y = x + 12582912.0f;
    const float f = x - (y - 12582912.0f);
    float p = fmaf(f, 0.0096181291f, 0.0555041087f);
    p = fmaf(p, f, 0.2402265070f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(y) << 23));
}
// Packed fp32 pairs (FFMA2 / FADD2 on sm_100): the exp2 pass is bound by issue slots, and these halve the scale-and-shift and the
// row-sum instructions around the MUFU operations.
__device__ __forceinline__ void fma2(float& y0, float& y1, float x0, float x1, float s, float c) {
    uint64_t xx, ss, cc, rr;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(xx), "l"(ss), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(rr));
}
__device__ __forceinline__ void add2(uint64_t& acc, float a, float b) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(v));
}
// 16 scores -> exp2(s * sl - m) in e[], accumulated into two packed partial sums
__device__ __forceinline__ void exp16(const uint32_t* r, float sl, float neg_m, float* e, uint64_t& acc0, uint64_t& acc1) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        float t0, t1;
        fma2(t0, t1, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), sl, neg_m);
        e[i] = ex2_approx(t0);
        e[i + 1] = ex2_approx(t1);
        if (i & 2) add2(acc1, e[i], e[i + 1]);
        else       add2(acc0, e[i], e[i + 1]);
    }
}

template <int POLY_EVERY>
__device__ __forceinline__ float ex2_mixed(float x, int i) {   // i: compile-time element index
    return (POLY_EVERY > 0 && (i % (POLY_EVERY > 0 ? POLY_EVERY : 1)) == (POLY_EVERY > 0 ? POLY_EVERY : 1) - 1) ? ex2_poly(x) : ex2_approx(x);
}
constexpr int POLY_EVERY = APX_FLASH_POLY;   // 0: all on MUFU; n: every n-th exponential on the FMA pipe

template <int FMT, int NG>
__global__ void __launch_bounds__(NT, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                  FaArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                      // [d/8][128][8], or [128][d] swizzled
    uint8_t* sK = smem + a.off_k;            // stages x [d/8][128 keys][8]
    uint8_t* sV = smem + a.off_v;            // stages x [d/8][128 keys][8]   (read MN-major as B[N = dims][K = keys])
    uint8_t* sP = smem + a.off_p;            // 2 x [128 keys / 8][128 rows][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

    FA_TRACE(63);
    if (threadIdx.x == 256) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        for (int i = 0; i < B_COUNT; ++i) {
            const bool from_threads = (i >= B_S_FREE && i < B_PV_FULL) || i >= B_PV_FREE;   // S_FREE, P_FULL, PV_FREE: the 128 softmax threads
            mbar_init(&bars[i], from_threads ? 128u : 1u);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_slot);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    pdl_launch_dependents();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tensor memory: score buffers at columns 0 and 128, P V accumulators at 256 and 384
    const uint32_t tS = tmem_base, tPV = tmem_base + 256;

    if (warp >= 8) {   // the register re-balancing must dominate each role's code: one branch per register budget
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
    if (warp == 9) {
        if (elect_one()) {
            mbar_expect_tx(&bars[B_Q], a.tile_bytes);
            if (a.swz) tma_load_3d(sQ, &tmQ, h * a.d, m0, b, &bars[B_Q]);
            else       tma_load_4d(sQ, &tmQ, 0, m0, h * (a.d / 8), b, &bars[B_Q]);
            for (int j = 0; j < a.nblk; ++j) {
                const int st = j % a.stages;
                if (j >= a.stages) mbar_wait_nocall(&bars[B_KV_FREE + st], ((j / a.stages) - 1) & 1);
                mbar_expect_tx(&bars[B_KV_FULL + st], 2 * a.tile_bytes);
                if (a.swz) {
                    tma_load_3d(sK + st * a.tile_bytes, &tmK, h * a.d, j * BKV, b, &bars[B_KV_FULL + st]);
                    tma_load_3d(sV + st * a.tile_bytes, &tmV, h * a.d, j * BKV, b, &bars[B_KV_FULL + st]);
                } else {
                    tma_load_4d(sK + st * a.tile_bytes, &tmK, 0, j * BKV, h * (a.d / 8), b, &bars[B_KV_FULL + st]);
                    tma_load_4d(sV + st * a.tile_bytes, &tmV, 0, j * BKV, h * (a.d / 8), b, &bars[B_KV_FULL + st]);
                }
            }
        }
    } else if (warp == 8) {
        if (elect_one()) {
            const uint32_t id_s = idesc_major(FMT, 128, BKV, 0, 0);
            const uint32_t id_pv = idesc_major(FMT, 128, NG * 16, 0, 1);
            // Operand tiles.  swz == 0: [d/8][128][8] core matrices — K-major: k-group stride 2048 B, 8-row stride 128 B, a 16-element
            // k-step is two k-groups; V read MN-major: next 8 keys (K group) 128 B on = "leading" offset, next 8 dims 2048 B on = "stride"
            // offset.  swz == 64 / 128: [128 rows][d] with rows of 64 / 128 bytes in the TMA swizzle pattern — K-major: 8-row stride
            // 8 x row bytes, a k-step is 32 B along the row; V read MN-major: the d contiguous elements ARE the swizzle row, the next
            // 8 keys are 8 rows on ("stride" offset), a 16-key k-step is 16 rows.
            const uint32_t rowb = static_cast<uint32_t>(a.d) * 2;
            const uint64_t lay = a.swz == 128 ? kLayoutSW128 : a.swz == 64 ? kLayoutSW64 : kLayoutNone;
            const uint32_t k_lbo = a.swz ? 16u : 2048u, k_sbo = a.swz ? 8 * rowb : 128u, k_step = a.swz ? 32u : 2 * 2048u;
            const uint32_t v_lbo = a.swz ? a.tile_bytes : 128u, v_sbo = a.swz ? 8 * rowb : 2048u, v_step = a.swz ? 16 * rowb : 2 * 128u;
            const uint64_t dQ = umma_desc(smem_u32(sQ), k_lbo, k_sbo, lay);
            // P_i V_i: A = the P tile (K-major over the keys); B[N = dims][K = keys] = the V block read MN-major (8 dims contiguous; the
            // next 8 dims 2048 B on; the next 8 keys 128 B on); 8 k-steps of 16 keys
            // Event-driven issue: the score product of the next block goes out as soon as its K block has landed and its score buffer
            // has been read (it does not wait behind a P V whose P tile is still being computed), the P V of a block as soon as its P
            // tile is complete and its accumulator has been folded.  try_wait on a completed phase stays true until this thread acts
            // (the non-blocking mbarrier.test_wait measured the same here: profiles/r2w_flash_poll_ab.txt).
            mbar_wait_spin(&bars[B_Q], 0);
            int js = 0, jp = 0;
            while (jp < a.nblk) {
                if (js < a.nblk) {
                    const int sb = js & 1, st = js % a.stages;
                    if (mbar_try_wait(&bars[B_KV_FULL + st], (js / a.stages) & 1) &&
                        (js < 2 || mbar_try_wait(&bars[B_S_FREE + sb], ((js >> 1) - 1) & 1))) {
                        tc_fence_after();
                        mma_chain(tS + sb * 128, dQ, umma_desc(smem_u32(sK + st * a.tile_bytes), k_lbo, k_sbo, lay), k_step, k_step, a.d / 16, id_s, 0);
                        umma_commit(&bars[B_S_FULL + sb]);
                        if (js < 12) FA_TRACE_MMA(64 + js * 6 + 2);
                        ++js;
                    }
                }
                if (jp < js) {
                    const int pb = jp & 1, st = jp % a.stages;
                    if (mbar_try_wait(&bars[B_P_FULL + pb], (jp >> 1) & 1) && (jp < 2 || mbar_try_wait(&bars[B_PV_FREE + pb], ((jp >> 1) - 1) & 1))) {
                        tc_fence_after();
                        // P V: A = the P tile (K-major over the keys), B = the V block read MN-major, 8 k-steps of 16 keys
                        mma_chain(tPV + pb * 128, umma_desc(smem_u32(sP + pb * 32768), 2048, 128, kLayoutNone),
                                  umma_desc(smem_u32(sV + st * a.tile_bytes), v_lbo, v_sbo, lay), 2 * 2048, v_step, BKV / 16, id_pv, 0);
                        umma_commit(&bars[B_PV_FULL + pb]);
                        umma_commit(&bars[B_KV_FREE + st]);
                        if (jp < 12) FA_TRACE_MMA(64 + jp * 6 + 4);
                        ++jp;
                    }
                }
            }
        }
    }   // warps 10, 11: idle
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_WG));
        // ---- softmax warpgroups.  Warpgroup g owns the key blocks j = g, g + 2, ... with its OWN running (m, l, O): block j's score
        // buffer, P tile and P V accumulator are all indexed by j & 1 = g, so the two warpgroups never synchronise inside the loop — one
        // computes exp2 (MUFU-bound) while the other waits on tensor memory / folds its P V — and their partial results are merged once.
        const int g = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_taddr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const float sl = a.scale_log2;
        constexpr int DR = NG * 16;         // head dim (compile time: the output row lives in registers)
        constexpr bool ONE_PASS = NG <= 2;  // d <= 32: the whole 128-column score row fits in the 232 registers next to the output row
        float m = -INFINITY, l = 0.f;
        float o[DR];
#pragma unroll
        for (int i = 0; i < DR; ++i) o[i] = 0.f;
        const uint32_t ts = tS + g * 128 + lane_taddr;
        const uint32_t tp = tPV + g * 128 + lane_taddr;
        const uint32_t p_row = smem_u32(sP + g * 32768) + row * 16;

        // The exp2 passes of the two warpgroups take turns (named barriers 2 / 3 as a token): run at the same time they only split
        // the MUFU pipe between them and then idle together; alternating, one warpgroup's row maxima / P V fold / barrier waits hide
        // behind the other's exp2 pass.
        if (g == 1) named_bar_arrive(2, 256);
        FA_TRACE(0);
        int it = 0;                         // how many blocks this warpgroup has processed: barrier phases
        for (int j = g; j < a.nblk; j += 2, ++it) {
            const int valid = min(BKV, a.L - j * BKV);     // keys of this block that exist
            mbar_wait_spin(&bars[B_S_FULL + g], it & 1);
            tc_fence_after();
            if (it < 15) FA_TRACE(1 + it * 4);
            float m_new, alpha;
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
            uint64_t acc0 = 0, acc1 = 0;            // two packed pairs of partial row sums
            if (ONE_PASS) {
                // the score row: one trip to tensor memory, after which the buffer goes straight back to the MMA thread
                uint32_t r[128];
#pragma unroll
                for (int c = 0; c < 8; ++c) tmem_ld16(ts + c * 16, r + c * 16);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) reg_tie16(r + c * 16);
                tc_fence_before();
                mbar_arrive(&bars[B_S_FREE + g]);
                if (valid < BKV) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (valid - c * 16 < 16) mask16(r + c * 16, valid - c * 16);
                }
                float mx8[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) mx8[c] = max16(r + c * 16);
                const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
                m_new = fmaxf(m, mx * sl);
                alpha = ex2_approx(m - m_new);   // 0 for the first block (m = -inf)
                if (it < 15) FA_TRACE(2 + it * 4);
                uint32_t pv[DR];
                if (it >= 1) {   // this warpgroup's previous P V has long completed: start its trip out of tensor memory under the exp2 pass
                    mbar_wait_spin(&bars[B_PV_FULL + g], (it - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < NG; ++c) tmem_ld16(tp + c * 16, pv + c * 16);
                }
                named_bar_sync(2 + g, 256);               // my turn on the MUFU pipe
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float e[16];
                    if (POLY_EVERY == 0) {
                        exp16(r + c * 16, sl, -m_new, e, acc0, acc1);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            e[i] = ex2_mixed<POLY_EVERY>(fmaf(__uint_as_float(r[c * 16 + i]), sl, -m_new), i);
                            sum4[i & 3] += e[i];
                        }
                    }
                    store16<FMT>(p_row + c * 2 * 2048, e);
                }
                named_bar_arrive(3 - g, 256);             // the other warpgroup's turn
                fence_proxy_async_smem();                     // P tile -> visible to the tensor core
                mbar_arrive(&bars[B_P_FULL + g]);
                if (it < 15) FA_TRACE(3 + it * 4);
                if (it >= 1) {
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < NG; ++c) {
                        reg_tie16(pv + c * 16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[c * 16 + i] = (o[c * 16 + i] + __uint_as_float(pv[c * 16 + i])) * alpha;
                    }
                    tc_fence_before();
                    mbar_arrive(&bars[B_PV_FREE + g]);
                }
            } else {
                // two passes over tensor memory (row maximum, then exp2): 64 columns in registers at a time
                float mx = -INFINITY;
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t r[64];
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld16(ts + hf * 64 + c * 16, r + c * 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        reg_tie16(r + c * 16);
                        const int nv = valid - (hf * 64 + c * 16);
                        if (nv < 16) mask16(r + c * 16, nv);
                        mx = fmaxf(mx, max16(r + c * 16));
                    }
                }
                m_new = fmaxf(m, mx * sl);
                alpha = ex2_approx(m - m_new);
                if (it < 15) FA_TRACE(2 + it * 4);
                if (it >= 1) {   // fold in this warpgroup's previous P V (16 columns at a time): it has also finished reading the P tile
                    mbar_wait_spin(&bars[B_PV_FULL + g], (it - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < NG; ++c) {
                        uint32_t pv[16];
                        tmem_ld16(tp + c * 16, pv);
                        tmem_ld_wait16(pv);
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[c * 16 + i] = (o[c * 16 + i] + __uint_as_float(pv[i])) * alpha;
                    }
                    tc_fence_before();
                    mbar_arrive(&bars[B_PV_FREE + g]);
                }
                named_bar_sync(2 + g, 256);               // my turn on the MUFU pipe
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t r[64];
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld16(ts + hf * 64 + c * 16, r + c * 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        reg_tie16(r + c * 16);
                        const int nv = valid - (hf * 64 + c * 16);
                        if (nv < 16) mask16(r + c * 16, nv);
                        float e[16];
                        if (POLY_EVERY == 0) {
                            exp16(r + c * 16, sl, -m_new, e, acc0, acc1);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                e[i] = ex2_mixed<POLY_EVERY>(fmaf(__uint_as_float(r[c * 16 + i]), sl, -m_new), i);
                                sum4[i & 3] += e[i];
                            }
                        }
                        store16<FMT>(p_row + (hf * 8 + c * 2) * 2048, e);
                    }
                }
                named_bar_arrive(3 - g, 256);             // the other warpgroup's turn
                tc_fence_before();
                mbar_arrive(&bars[B_S_FREE + g]);         // the score buffer may take this warpgroup's next block
                fence_proxy_async_smem();
                mbar_arrive(&bars[B_P_FULL + g]);
                if (it < 15) FA_TRACE(3 + it * 4);
            }
            {
                float a0, a1, a2, a3;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc0));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(acc1));
                l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3])) + ((a0 + a1) + (a2 + a3));
            }
            m = m_new;
            if (it < 15) FA_TRACE(4 + it * 4);
        }
        if ((a.nblk & 1) == g) named_bar_sync(2 + g, 256);   // the token's last hand-over has no taker: consume it (barriers end balanced)
        if (it >= 1) {   // this warpgroup's last P V
            mbar_wait_spin(&bars[B_PV_FULL + g], (it - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < NG; ++c) {
                uint32_t pv[16];
                tmem_ld16(tp + c * 16, pv);
                tmem_ld_wait16(pv);
#pragma unroll
                for (int i = 0; i < 16; ++i) o[c * 16 + i] += __uint_as_float(pv[i]);
            }
        }
        // ---- merge: warpgroup B hands (m, l, O) of every row to warpgroup A through the (now dead) P tiles
        float* xch = reinterpret_cast<float*>(sP) + row;     // [2 + DR][128] floats, column = row: conflict-free
        named_bar_sync(1, 256);                               // both warpgroups have seen their last P V: no MMA reads the P tiles any more
        if (g == 1) {
            xch[0] = m;
            xch[128] = l;
#pragma unroll
            for (int i = 0; i < DR; ++i) xch[(2 + i) * 128] = o[i];
        }
        named_bar_sync(1, 256);
        if (g == 0) {
            const float mb = xch[0], lb = xch[128];
            const float mm = fmaxf(m, mb);
            const float fa = ex2_approx(m - mm), fb = ex2_approx(mb - mm);   // warpgroup B without blocks: mb = -inf -> fb = 0
            const float inv = __fdividef(1.0f, l * fa + lb * fb);   // (a precise division would bring a call into a setmaxnreg kernel)
            const int q = m0 + row;
            uint16_t* op = a.O + static_cast<int64_t>(b) * a.o_bs + static_cast<int64_t>(q) * a.o_rs + h * DR;
#pragma unroll
            for (int c = 0; c < NG; ++c) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int e0 = c * 16 + 2 * i;
                    w[i] = pack2<FMT>((o[e0] * fa + xch[(2 + e0) * 128] * fb) * inv, (o[e0 + 1] * fa + xch[(3 + e0) * 128] * fb) * inv);
                }
                if (q < a.S) {
                    *reinterpret_cast<uint4*>(op + c * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(op + c * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
            FA_TRACE(62);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem_base);
}

}  // namespace

bool flash_attn_supported(int C, int H) {
    if (H <= 0 || C % H) return false;
    const int d = C / H;
    return d % 16 == 0 && d >= 16 && d <= 80;
}

// Q / O: [B, S, C] views, K / V: [B, L, C] views (element strides *_bs per batch element, *_rs per row; columns contiguous)
int flash_attn(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs, const void* V, int64_t v_bs,
               int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S, int L, int C, int H, int dtype, cudaStream_t st) {
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "flash_attn: dtype %d", dtype);
    APX_REQUIRE(flash_attn_supported(C, H), APX_EUNSUPPORTED, "flash_attn: head dim %d (need d %% 16 == 0, 16 <= d <= 80)", H > 0 ? C / H : 0);
    APX_REQUIRE(B > 0 && S > 0 && L > 0, APX_EINVAL, "flash_attn: B=%d S=%d L=%d", B, S, L);
    const int64_t strides_all[8] = {q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs};
    for (int i = 0; i < 8; ++i) APX_REQUIRE(strides_all[i] % 8 == 0, APX_EINVAL, "flash_attn: strides must be multiples of 8 elements");
    APX_REQUIRE(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(O)) & 15) == 0,
                APX_EINVAL, "flash_attn: Q / K / V / O must be 16-byte aligned");
    const int d = C / H;
    FaArgs a;
    a.O = static_cast<uint16_t*>(O);
    a.o_bs = o_bs; a.o_rs = o_rs;
    a.S = S; a.L = L; a.H = H; a.d = d;
    a.nblk = ceil_div(L, BKV);
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(d));
    a.tile_bytes = 128u * d * 2;
    a.stages = std::min(a.nblk, std::min(MAX_ST, static_cast<int>((220 * 1024 - a.tile_bytes - 2 * 32768) / (2 * a.tile_bytes))));
    a.off_k = a.tile_bytes;
    a.off_v = a.off_k + a.stages * a.tile_bytes;
    a.off_p = a.off_v + a.stages * a.tile_bytes;
    a.off_bar = a.off_p + 2 * 32768;
    static const bool swz_off = [] { const char* e = getenv("APX_FLASH_SWIZZLE"); return e && e[0] == '0'; }();
    a.swz = swz_off ? 0 : d == 32 ? 64 : d == 64 ? 128 : 0;   // head rows of exactly one swizzle span
    const size_t smem = 1024 + a.off_bar + B_COUNT * 8 + 16;
    APX_REQUIRE(smem <= 227 * 1024, APX_EUNSUPPORTED, "flash_attn: %zu B of shared memory", smem);

    CUtensorMap tmQ, tmK, tmV;
    auto make = [&](CUtensorMap* tm, const void* p, int64_t bs, int64_t rs, int rows) {
        // a [B][C/8][rows][8] view: the box {8, 128, d/8, 1} lands as [d/8][128][8] = UMMA core-matrix order
        uint64_t dims[4] = {8, static_cast<uint64_t>(rows), static_cast<uint64_t>(C / 8), static_cast<uint64_t>(B)};
        uint64_t strides[3] = {static_cast<uint64_t>(rs) * 2, 16, static_cast<uint64_t>(bs) * 2};
        uint32_t box[4] = {8, 128, static_cast<uint32_t>(d / 8), 1};
        if (a.swz) {   // [B][rows][C]: the box {d, 128, 1} lands as 128 rows of d elements in the TMA swizzle pattern
            uint64_t dims3[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(rows), static_cast<uint64_t>(B)};
            uint64_t strides3[2] = {static_cast<uint64_t>(rs) * 2, static_cast<uint64_t>(bs) * 2};
            uint32_t box3[3] = {static_cast<uint32_t>(d), 128, 1};
            return encode_tmap(tm, dtype, 3, p, dims3, strides3, box3, a.swz == 128 ? 1 : 2);
        }
        return encode_tmap(tm, dtype, 4, p, dims, strides, box, 0);
    };
    int rc = make(&tmQ, Q, q_bs, q_rs, S);
    if (rc) return rc;
    rc = make(&tmK, K, k_bs, k_rs, L);
    if (rc) return rc;
    rc = make(&tmV, V, v_bs, v_rs, L);
    if (rc) return rc;

    typedef void (*Fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, FaArgs);
    static const Fn table[2][5] = {
        {flash_attn_kernel<0, 1>, flash_attn_kernel<0, 2>, flash_attn_kernel<0, 3>, flash_attn_kernel<0, 4>, flash_attn_kernel<0, 5>},
        {flash_attn_kernel<1, 1>, flash_attn_kernel<1, 2>, flash_attn_kernel<1, 3>, flash_attn_kernel<1, 4>, flash_attn_kernel<1, 5>}};
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        for (int f = 0; f < 2; ++f)
            for (int n = 0; n < 5; ++n)
                APX_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(table[f][n]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const Fn fn = table[dtype == APX_F16 ? 0 : 1][d / 16 - 1];
    static const bool trace_on = [] { const char* e = getenv("APX_FLASH_TRACE"); return e && e[0] == '1'; }();
    static long long* trace_buf = nullptr;
    a.trace = nullptr;
    if (trace_on) {
        if (!trace_buf) APX_CUDA(cudaMalloc(&trace_buf, 160 * sizeof(long long)));
        APX_CUDA(cudaMemsetAsync(trace_buf, 0, 160 * sizeof(long long), st));
        a.trace = trace_buf;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(S, 128), H, B);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    cfg.attrs = attrs;
    cfg.numAttrs = 0;
    add_pdl_attr(attrs, &cfg.numAttrs);
    APX_CUDA(cudaLaunchKernelEx(&cfg, fn, tmQ, tmK, tmV, a));
    APX_LAUNCH_CHECK("flash_attn_kernel");
    if (trace_on) {
        long long t[160];
        APX_CUDA(cudaStreamSynchronize(st));
        APX_CUDA(cudaMemcpy(t, trace_buf, sizeof(t), cudaMemcpyDeviceToHost));
        printf("[apx flash trace] B=%d S=%d L=%d C=%d H=%d (cycles since kernel entry, thread 0 of CTA 0): loop start %lld\n", B, S, L, C, H, t[0] - t[63]);
        for (int j = 0; j < a.nblk && j < 15; ++j)
            printf("  block %d: s_seen %lld max_done %lld p_arrived %lld folded %lld\n", j, t[1 + j * 4] - t[63], t[2 + j * 4] - t[63],
                   t[3 + j * 4] - t[63], t[4 + j * 4] - t[63]);
        printf("  stored %lld\n", t[62] - t[63]);
        for (int j = 0; j < a.nblk && j < 12; ++j)
            printf("  mma block %d: kv_full %lld s_free %lld s_issued %lld | p_full %lld pv_free %lld\n", j, t[64 + j * 6] - t[63], t[64 + j * 6 + 1] - t[63],
                   t[64 + j * 6 + 2] - t[63], t[64 + j * 6 + 3] - t[63], t[64 + j * 6 + 4] - t[63]);
        fflush(stdout);
    }
    return APX_OK;
}

}  // namespace apx
