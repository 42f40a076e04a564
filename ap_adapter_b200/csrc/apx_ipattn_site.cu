// apx_ipattn_site.cu — one adapter site in ONE tcgen05 kernel, second generation ("site kernel"):
//
//   Y = ( softmax(Q K_t^T / sqrt(d)) V_t + ap_scale * softmax(Q K_a^T / sqrt(d)) V_a ) Wo^T + bo,   Q = X Wq^T
//
// (to_q, both SDPAs, the ap_scale add and to_out of APadapter/ap_adapter/attention_processor.py:387-459.)
//
// What the first fused kernel (apx_ipattn_fused.cu) taught us on the B200 (profiles/r1_fused_v2_trace_*.txt): with one
// head at a time the kernel is a 7 k-cycle latency chain per head, a tcgen05.mma costs ~57 cycles to ISSUE whatever its
// N, so N = d = 32 projections are issue-bound, and a row-per-thread global epilogue costs 5.7 k cycles.  Hence:
//
//   phase A   Q for ALL heads at once:  Q[128, H*d_pad] = X[128, C] · Wq^T   as N = H*d_pad wide MMAs (tensor-bound),
//             X chunks (SWIZZLE_128B tensor TMA) and packed Wq chunks (1-D bulk TMA) through a 2-stage mbarrier ring;
//             both warpgroups convert the fp32 accumulator to the 16-bit K-major operand tile sQO (all heads side by side).
//   phase B   heads alternate between two softmax warpgroups (WG0: 0,2,4.. WG1: 1,3,5..), each with its own S / O_t / O_a
//             accumulators in tensor memory and its own P tile in shared memory.  The single MMA thread issues event-driven
//             (two in-order queues, S and P·V, polled with mbarrier.test_wait): S(h+2) goes out as soon as its warpgroup has
//             moved S(h) to registers, P·V(h) as soon as P(h) is stored; the exp2 passes of the two warpgroups take turns
//             (named-barrier token), so the MUFU pipe — the real bound of this kernel at d = 32 (one ex2 per score, 16/clk/SM)
//             — is fed by one warpgroup while the other loads scores, takes maxima and converts its previous head's output.
//             O_h = O_t / l_t + ap_scale · O_a / l_a is formed in registers and written over Q_h in sQO.
//   phase C   Y[128, C] = O[128, H*d_pad] · Wo^T as C-wide MMAs over the gathered O tile (Wo chunks stream through the
//             shared memory the P tiles occupied), + bias, 16-bit, swizzled staging tile, tensor-TMA store (rows past S clipped).
//
// Warp roles (384 threads): warps 0-3 = WG0, warps 4-7 = WG1 (TMEM lane quarter = warp % 4), warp 8 = TMEM allocator +
// single-thread MMA issuer (elect.sync, so that the descriptors stay in uniform registers), warp 9 = TMA producer,
// warps 10-11 idle (they only complete the third warpgroup so that setmaxnreg can move its registers to the other two).
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include "apx_site_common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace apx {

SiteGeom site_geom(int C, int H) {
    SiteGeom g;
    memset(&g, 0, sizeof(g));
    if (C <= 0 || H <= 0 || C % H || C % 64 || C > 256) return g;
    const int d = C / H;
    if (d % 8) return g;
    const int d_pad = (d + 15) / 16 * 16;
    const int Nq = H * d_pad;
    if (Nq > 256) return g;
    g.ok = 1; g.C = C; g.H = H; g.d = d; g.d_pad = d_pad; g.Nq = Nq;
    g.n_kb = C / 64;
    // One pair of sections (Wq, Wo) per supported cluster size cs = 1, 2, 4: the rows of Wq (heads) / the columns of Wo are grouped
    // by the CTA rank that owns them, so that every CTA streams only its own slice with 1-D bulk copies.
    size_t off = 0;
    for (int i = 0; i < 3; ++i) {
        const int cs = 1 << i;
        g.cs_ok[i] = (H % cs == 0) && ((C / cs) % 64 == 0) && ((Nq / cs) % 16 == 0);
        g.wq_off[i] = g.wo_off[i] = 0;
        if (!g.cs_ok[i]) continue;
        g.wq_off[i] = off; off += static_cast<size_t>(Nq) * C * 2;   // [cs][C/64][8][Nq/cs][8]
        g.wo_off[i] = off; off += static_cast<size_t>(Nq) * C * 2;   // [cs][Nq/8][C/cs][8]
    }
    g.total = off;
    return g;
}

namespace {

constexpr int X_CHUNK_BYTES = 128 * 64 * 2;   // one SWIZZLE_128B box: 128 rows x 64 channels
constexpr int WO_KSTEPS_PER_STAGE = 4;
constexpr int NTHREADS = 384;          // 3 warpgroups: two softmax warpgroups + {MMA warp, TMA warp, two idle warps}
constexpr int REGS_WG = 216;           // setmaxnreg: the softmax warpgroups keep a 144-column score row in registers,
constexpr int REGS_AUX = 72;           // the single-thread roles give theirs up: (168 - 72) * 128 freed = (216 - 168) * 256 claimed
                                       // (224 / 56 and 232 / 40 measured no faster: profiles/r2r_site_ab.txt)

// Per cluster size cs (G = Nq / cs rows of Wq, cn = C / cs columns of Wo per CTA):
// WQ section: [cs][C/64][8][G][8]: element (rank r, chunk kb, k-group kg, row n, e) = Wq[h*d + j, kb*64 + kg*8 + e] with
//             r*G + n = h*d_pad + j (rows with j >= d are zero).
// WO section: [cs][Nq/8][cn][8]:   element (rank r, k-group kg, column n, e) = Wo[r*cn + n, h*d + j], kg*8 + e = h*d_pad + j
//             (zero for j >= d).
__global__ void site_pack_kernel(const uint16_t* __restrict__ Wq, const uint16_t* __restrict__ Wo, uint16_t* __restrict__ out,
                                 SiteGeom g) {
    const size_t n_sec = static_cast<size_t>(g.Nq) * g.C;   // elements per section
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < 6 * n_sec;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int sec = static_cast<int>(idx / n_sec);   // 0: Wq cs=1, 1: Wo cs=1, 2: Wq cs=2, ...
        const int i = sec >> 1, cs = 1 << i;
        if (!g.cs_ok[i]) continue;
        size_t r = idx % n_sec;
        const int e = r % 8; r /= 8;
        if ((sec & 1) == 0) {
            const int G = g.Nq / cs;
            const int n = r % G; r /= G;
            const int kg = r % 8; r /= 8;
            const int kb = r % (g.C / 64);
            const int rank = r / (g.C / 64);
            const int row = rank * G + n;
            const int h = row / g.d_pad, j = row % g.d_pad;
            const int k = kb * 64 + kg * 8 + e;
            out[g.wq_off[i] / 2 + idx % n_sec] = j < g.d ? Wq[static_cast<size_t>(h * g.d + j) * g.C + k] : uint16_t(0);
        } else {
            const int cn = g.C / cs;
            const int n = r % cn; r /= cn;
            const int kg = r % (g.Nq / 8);
            const int rank = r / (g.Nq / 8);
            const int kk = kg * 8 + e;
            const int h = kk / g.d_pad, j = kk % g.d_pad;
            out[g.wo_off[i] / 2 + idx % n_sec] = j < g.d ? Wo[static_cast<size_t>(rank * cn + n) * g.C + h * g.d + j] : uint16_t(0);
        }
    }
}

struct SiteArgs {
    const uint8_t* wq;     // packed WQ section of this cluster size (rank-major)
    const uint8_t* wo;     // packed WO section of this cluster size (rank-major)
    const uint8_t* tiles;  // packed context K / V^T tiles, [Bc][H]{K tile, V^T tile}
    const float* bo;       // fp32[C] or null
    int S, C, H, d_pad, dv_pad, Nq, n_kb;
    int Lt, La, Lt_pad, La_pad, L_pad;
    float ap_scale, scale_log2;
    uint32_t kv_bytes, k_tile_bytes, v_tile_bytes, vx_off, wq_chunk_bytes, wo_kstep_bytes, a_stage_bytes, p_bytes;
    int nk, nv, n_wo, wo_early;                // K ring, V ring depths; Wo stages / how many have their own region
    int cs, hg, cn;                            // cluster size; heads and output columns per CTA (H / cs, C / cs)
    uint32_t off_p, off_k, off_v, off_bar;     // shared-memory offsets from the 1024-aligned base (sQO is at 0)
    uint32_t wo_off[4];                        // where each Wo stage lands
    int accumulate;                            // 1: Y += result (TMA reduce-add into the residual stream), 0: Y = result
    const float* ln_c1;                        // norm2 folded into the q projection (null: X is already normalised): c1[n] = sum_k Wq'[n, k],
    const float* ln_c2;                        // c2[n] = sum_k beta_k Wq[n, k]; the pack then holds Wq' = Wq * gamma
    float ln_eps;
    long long* trace;                          // diagnostic (APX_SITE_TRACE=1): clock64 stamps of CTA (0,0)
};



// trace slots: [0,32) MMA thread, [32,64) producer, [64,192) WG0 thread 0 (8 per head), [192,320) WG1 thread 128
#define SITE_T(slot)                                                                             \
    do {                                                                                         \
        if (TRACE && a.trace && blockIdx.x == 0 && blockIdx.y == 0) a.trace[(slot)] = clock64(); \
    } while (0)

enum {  // mbarrier slots (40 x 8 bytes reserved; the TMEM slot and the bias slice follow)
    B_A_FULL = 0, B_K_FULL = 4, B_K_EMPTY = 7, B_V_FULL = 10, B_V_EMPTY = 13, B_WO_FULL = 16, B_Q = 20, B_SQ = 21, B_S_FULL = 22,
    B_S_FREE = 24, B_P_FULL = 26, B_O_FULL = 28, B_PV_ALL = 30, B_SO_ALL = 31, B_Y = 32, B_OALL = 33, B_DONE = 34, B_STAT = 35, B_COUNT = 36
};
constexpr uint32_t BAR_REGION_BYTES = 384;   // barriers [0, 320) | TMEM slot at 320 | bias slice from 384 | folded-LayerNorm c1, c2 behind it

// O_h = O_t / l_t + ap_scale * O_a / l_a -> 16 bit, over Q_h in sQO (dst = this thread's row of head h's slot); l_t, l_a = accumulator
// column d_pad (the ones row of the V^T tile).  One 16-column group of each accumulator per round trip: with two, the 64 staging
// registers next to the 144-register score row spilled to local memory inside the head loop (-1.2 us per call at Bc = 16).
template <int FMT>
__device__ __forceinline__ void site_convert_head(const SiteArgs& a, uint32_t tOt, uint32_t tOa, uint32_t dst) {
    uint32_t lt, la;
    tmem_ld1(tOt + a.d_pad, lt);
    tmem_ld1(tOa + a.d_pad, la);
    float w_t = 0.f, w_a = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < a.d_pad; c0 += 16) {
        uint32_t rt[16], ra[16];
        tmem_ld16(tOt + c0, rt);
        tmem_ld16(tOa + c0, ra);
        tmem_ld_wait();
        if (c0 == 0) {
            asm volatile("" : "+r"(lt), "+r"(la) : : "memory");
            w_t = a.Lt > 0 ? __fdividef(1.0f, __uint_as_float(lt)) : 0.f;
            w_a = a.La > 0 ? __fdividef(a.ap_scale, __uint_as_float(la)) : 0.f;
        }
        reg_tie16(rt);
        reg_tie16(ra);
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(ra[i]), w_a, __uint_as_float(rt[i]) * w_t);
        store16<FMT>(dst + c0 / 8 * 2048, v);
    }
}

// N16 > 0: the score row is N16 16-column groups, the first LT16 of them text, and it is read from tensor memory ONCE and kept
//          in registers for both softmax passes (the score buffer is handed back to the MMA thread right after the load).
// N16 == 0: any context shape; two passes over tensor memory (maxima, exponentials).
// The code is kept compact on purpose: the SM's instruction cache holds 32 KB, and code that falls out of it runs at ~10 cycles
// per instruction (measured: profiles/r1_site_v3_trace_*.txt).
template <int FMT, int N16, int LT16, bool TRACE>
__global__ void __launch_bounds__(NTHREADS, 1)
ipattn_site_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, SiteArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQO = smem;                 // [Nq/8][128][8] 16 bit: Q of every head, overwritten head by head with O; then Y staging
    uint8_t* sP = smem + a.off_p;        // P tiles of the two warpgroups
    uint8_t* sK = smem + a.off_k;        // K tile ring
    uint8_t* sV = smem + a.off_v;        // V^T tile ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.off_bar + 320);
    float* sBias = reinterpret_cast<float*>(smem + a.off_bar + BAR_REGION_BYTES);
    // phase A stages (X chunk + Wq chunk, one per 64 input channels) lie over everything below the barriers: nothing else
    // is live until the Q accumulator is complete.

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // cluster of cs CTAs per tile: CTA `rank` owns heads [rank*hg, (rank+1)*hg) and output columns [rank*cn, (rank+1)*cn)
    const int rank = a.cs > 1 ? static_cast<int>(cluster_ctarank()) : 0;
    const int h_first = rank * a.hg;
    const int m0 = (blockIdx.x / a.cs) * 128;
    const int b = blockIdx.y;

    if (threadIdx.x == 288) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
#pragma unroll 1
        for (int i = 0; i < B_COUNT; ++i) {
            const uint32_t cnt = (i == B_SQ || i == B_SO_ALL || i == B_STAT) ? 256u : (i >= B_S_FREE && i < B_O_FULL) ? 128u
                                 : i == B_DONE ? static_cast<uint32_t>(a.cs > 1 ? a.cs - 1 : 1) : 1u;
            mbar_init(&bars[i], cnt);   // warpgroup -> MMA barriers count every arriving thread
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_slot);
    pdl_wait();   // programmatic dependent launch: barrier init / tensor-memory allocation above overlap the predecessor's tail
    float* sLn = sBias + a.C;            // [2][Nq]: c1, c2 of the folded LayerNorm in the padded head layout
    if (threadIdx.x < 256) {
        for (int i = threadIdx.x; i < a.C; i += 256) sBias[i] = a.bo ? a.bo[i] : 0.f;
        if (a.ln_c1) {
            const int d = a.C / a.H;
            for (int i = threadIdx.x; i < a.Nq; i += 256) {
                const int hh = i / a.d_pad, j = i - hh * a.d_pad;
                sLn[i] = j < d ? a.ln_c1[hh * d + j] : 0.f;
                sLn[a.Nq + i] = j < d ? a.ln_c2[hh * d + j] : 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    pdl_launch_dependents();
    tc_fence_after();
    // Cluster handshake before the first remote write into a peer: every CTA's barriers must exist AND its phase-A stages (which
    // overlay sQO) must be dead.  The warpgroup warps arrive once they have seen the Q accumulator complete, the other warps
    // here; every warp waits (aligned) before this CTA pushes / before it leaves.
    if (a.cs > 1 && warp >= 8) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // phase A: Q at [0, Nq).  phase B: S(g) at g * L_pad, then {O_t(g), O_a(g)}, dv_pad columns each (column d_pad = row sum).
    // phase C: Y at [0, C).
    const uint32_t tmem_O = tmem_base + 2 * a.L_pad;
    const int n_ko = a.Nq / 16;   // k-steps of the out-projection

    if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
    if (warp == 9) {
        // ================================================================ TMA producer (one thread)
        if (elect_one()) {
            SITE_T(32);
#pragma unroll 1
            for (int kb = 0; kb < a.n_kb; ++kb) {   // every phase-A chunk has a stage of its own: all loads are in flight at once
                uint8_t* dst = smem + kb * a.a_stage_bytes;
                mbar_expect_tx(&bars[B_A_FULL + kb], X_CHUNK_BYTES + a.wq_chunk_bytes);
                tma_load_3d(dst, &tmX, kb * 64, m0, b, &bars[B_A_FULL + kb]);
                tma_load_1d(dst + X_CHUNK_BYTES, a.wq + (static_cast<size_t>(rank) * a.n_kb + kb) * a.wq_chunk_bytes, a.wq_chunk_bytes,
                            &bars[B_A_FULL + kb]);
            }
            // the K / V rings and the early Wo stages may be filled at once when the phase-A stages end below them (cluster
            // launches: small Wq slices); otherwise only when the Q accumulator is complete and the stages are dead
            if (static_cast<uint32_t>(a.n_kb) * a.a_stage_bytes > a.off_k) {
                mbar_wait_nocall(&bars[B_Q], 0);
                if (a.ln_c1) mbar_wait_nocall(&bars[B_STAT], 0);   // the warpgroups have finished reading the X chunks (row statistics)
            }
            SITE_T(33);
            const uint8_t* kv_src = a.tiles + (static_cast<size_t>(b) * a.H + h_first) * a.kv_bytes;
            auto load_wo = [&](int w) {
                const int ks = min(WO_KSTEPS_PER_STAGE, n_ko - w * WO_KSTEPS_PER_STAGE);
                mbar_expect_tx(&bars[B_WO_FULL + w], ks * a.wo_kstep_bytes);
                tma_load_1d(smem + a.wo_off[w], a.wo + (static_cast<size_t>(rank) * n_ko + w * WO_KSTEPS_PER_STAGE) * a.wo_kstep_bytes,
                            ks * a.wo_kstep_bytes, &bars[B_WO_FULL + w]);
            };
            // Issue order = the MMA thread's consumption order: K(2p), K(2p+1), then V(2p), V(2p+1) (a different order can
            // deadlock: the ring waits of one tile kind would sit in front of a tile the MMA thread needs first).
            int sk = 0, sv = 0;               // ring slots of the next K / V^T tile
            uint32_t pk = 1, pv = 1;          // "empty" parities of those slots (first pass: free)
#pragma unroll 1
            for (int h0 = 0; h0 < a.hg; h0 += 2) {
#pragma unroll 1
                for (int h = h0; h < h0 + 2 && h < a.hg; ++h) {
                    mbar_wait_nocall(&bars[B_K_EMPTY + sk], pk);
                    mbar_expect_tx(&bars[B_K_FULL + sk], a.k_tile_bytes);
                    tma_load_1d(sK + sk * a.k_tile_bytes, kv_src + static_cast<size_t>(h) * a.kv_bytes, a.k_tile_bytes, &bars[B_K_FULL + sk]);
                    if (++sk == a.nk) { sk = 0; pk ^= 1; }
                }
#pragma unroll 1
                for (int h = h0; h < h0 + 2 && h < a.hg; ++h) {
                    mbar_wait_nocall(&bars[B_V_EMPTY + sv], pv);
                    mbar_expect_tx(&bars[B_V_FULL + sv], a.v_tile_bytes);
                    tma_load_1d(sV + sv * a.v_tile_bytes, kv_src + static_cast<size_t>(h) * a.kv_bytes + a.vx_off, a.v_tile_bytes,
                                &bars[B_V_FULL + sv]);
                    if (++sv == a.nv) { sv = 0; pv ^= 1; }
                }
                if (h0 == 0) {                // the early Wo stage(s) go out behind the first pair of tiles
#pragma unroll 1
                    for (int w = 0; w < a.wo_early; ++w) load_wo(w);
                }
            }
            SITE_T(34);
            mbar_wait_nocall(&bars[B_PV_ALL], 0);   // the P tiles and the K / V rings are dead once every P·V has completed
            SITE_T(35);
#pragma unroll 1
            for (int w = a.wo_early; w < a.n_wo; ++w) load_wo(w);
            SITE_T(36);
        }
        if (a.cs > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else if (warp == 8) {
        // ================================================================ MMA issuer (one thread)
        if (elect_one()) {
            const uint32_t idesc_q = umma_idesc_f16(FMT, 128, a.hg * a.d_pad);   // own heads only
            const uint32_t idesc_s = umma_idesc_f16(FMT, 128, a.L_pad);
            // ---- phase A: Q = X · Wq^T, all heads
            SITE_T(0);
#pragma unroll 1
            for (int kb = 0; kb < a.n_kb; ++kb) {
                mbar_wait_nocall(&bars[B_A_FULL + kb], 0);
                tc_fence_after();
                if (kb == 0) SITE_T(1);
                const uint32_t base = smem_u32(smem + kb * a.a_stage_bytes);
                mma_chain(tmem_base, umma_desc(base, 16, 1024, kLayoutSW128),
                          umma_desc(base + X_CHUNK_BYTES, a.hg * a.d_pad * 16, 128, kLayoutNone), 32, 2 * a.hg * a.d_pad * 16, 4, idesc_q,
                          static_cast<uint32_t>(kb));
            }
            umma_commit(&bars[B_Q]);
            SITE_T(2);
            // ---- phase B.  As soon as a warpgroup has moved its score row to registers (S_FREE) the scores of its NEXT head are
            //      issued, so they are complete long before it needs them; P·V of a head is issued when its P tile is ready
            //      (P_FULL) and runs under the other warpgroup's / the next head's softmax.
            const uint64_t dQO = umma_desc(smem_u32(sQO), 2048, 128, kLayoutNone);
            const uint32_t head_bytes = (a.d_pad / 8) * 2048;   // one head's slice of sQO
            const uint32_t idesc_v = umma_idesc_f16(FMT, 128, a.dv_pad);
            const uint64_t dP0 = umma_desc(smem_u32(sP), 2048, 128, kLayoutNone);
            const uint64_t dK0 = umma_desc(smem_u32(sK), a.L_pad * 16, 128, kLayoutNone);
            const uint64_t dV0 = umma_desc(smem_u32(sV), a.dv_pad * 16, 128, kLayoutNone);
            int sk = 0, sv = 0;               // K ring slot of the next S; V ring slot of the next P·V
            uint32_t pk = 0, pv = 0;          // "full" parities of those ring slots
            mbar_wait_nocall(&bars[B_SQ], 0);   // every Q column is in sQO; the Q accumulator columns are free
            tc_fence_after();
            SITE_T(3);
            {
                // Event-driven issue (as in apx_flash_attn.cu): two in-order queues — the scores of head hs go out as soon as their K
                // tile has landed and warpgroup hs & 1 has moved its previous score row to registers; the P·V of head hp as soon as
                // its P tile is complete.  Neither queue waits behind the other, so one warpgroup's P·V is never held up by the
                // other warpgroup's score buffer (round 1 issued S(2p), S(2p+1), PV(2p-2), PV(2p-1) in order with blocking waits, which
                // pulled the warpgroups into lockstep: -0.5 us per call at Bc = 2, profiles/r2q_site_ab.txt).  test_wait, not try_wait:
                // try_wait may park the thread when the phase is incomplete, and this thread polls several barriers.
                int hs = 0, hp = 0;
                uint32_t polls = 0;
                while (hp < a.hg) {
                    if (++polls > (1u << 30)) asm volatile("trap;");   // a protocol bug becomes a CUDA error (after ~30 s), not a hung GPU
                    if (hs < a.hg) {
                        const int g = hs & 1;
                        if (mbar_test_wait(&bars[B_K_FULL + sk], pk) && (hs < 2 || mbar_test_wait(&bars[B_S_FREE + g], ((hs >> 1) - 1) & 1))) {
                            tc_fence_after();
                            mma_chain(tmem_base + g * a.L_pad, desc_add(dQO, (h_first + hs) * head_bytes), desc_add(dK0, sk * a.k_tile_bytes),
                                      2 * 2048, 2 * a.L_pad * 16, a.d_pad / 16, idesc_s, 0);
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
                            if (hp < 8) SITE_T(8 + 2 * hp);
                            const uint64_t dP = desc_add(dP0, g * a.p_bytes), dV = desc_add(dV0, sv * a.v_tile_bytes);
                            const uint32_t tO = tmem_O + g * 2 * a.dv_pad;
                            mma_chain(tO, dP, dV, 2 * 2048, 2 * a.dv_pad * 16, a.Lt_pad / 16, idesc_v, 0);
                            mma_chain(tO + a.dv_pad, desc_add(dP, (a.Lt_pad / 8) * 2048), desc_add(dV, (a.Lt_pad / 8) * (a.dv_pad * 16)),
                                      2 * 2048, 2 * a.dv_pad * 16, a.La_pad / 16, idesc_v, 0);
                            umma_commit(&bars[B_O_FULL + g]);
                            umma_commit(&bars[B_V_EMPTY + sv]);
                            if (++sv == a.nv) { sv = 0; pv ^= 1; }
                            if (hp < 8) SITE_T(9 + 2 * hp);
                            ++hp;
                            polls = 0;
                        }
                    }
                }
            }
            umma_commit(&bars[B_PV_ALL]);
            // ---- phase C: Y = O · Wo^T
            mbar_wait_nocall(&bars[B_SO_ALL], 0);   // every own O_h is in sQO and every accumulator has been read
            tc_fence_after();
            SITE_T(5);
        }
        if (a.cs > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // whole warp; see the handshake note above
        if (elect_one()) {
            const uint32_t idesc_y = umma_idesc_f16(FMT, 128, a.cn);             // own output columns only
            const uint64_t dQO = umma_desc(smem_u32(sQO), 2048, 128, kLayoutNone);
            const uint32_t head_bytes = (a.d_pad / 8) * 2048;
            if (a.cs > 1) {
                // the peers' O slices arrive by bulk copies over distributed shared memory (pushed head by head by their warpgroups)
                mbar_expect_tx(&bars[B_OALL], static_cast<uint32_t>(a.cs - 1) * a.hg * head_bytes);
                mbar_wait_nocall(&bars[B_OALL], 0);   // every peer's slices have landed here
                tc_fence_after();
            }
#pragma unroll 1
            for (int i = 0; i < a.n_wo; ++i) {
                const int ks = min(WO_KSTEPS_PER_STAGE, n_ko - i * WO_KSTEPS_PER_STAGE);
                mbar_wait_nocall(&bars[B_WO_FULL + i], 0);
                tc_fence_after();
                if (i == 0) SITE_T(6);
                mma_chain(tmem_base, desc_add(dQO, i * WO_KSTEPS_PER_STAGE * 2 * 2048),
                          umma_desc(smem_u32(smem + a.wo_off[i]), a.cn * 16, 128, kLayoutNone), 2 * 2048, a.wo_kstep_bytes, ks,
                          idesc_y, static_cast<uint32_t>(i));
            }
            umma_commit(&bars[B_Y]);
            SITE_T(7);
            if (a.cs > 1) {
                // tell every peer that its slices have landed here; leave only when every peer has said the same (then no bulk
                // copy can still be reading this CTA's shared memory)
                const uint32_t done = smem_u32(&bars[B_DONE]);
#pragma unroll 1
                for (int r = 0; r < a.cs; ++r)
                    if (r != rank) mbar_arrive_remote(map_to_cta(done, r));
                mbar_wait_nocall(&bars[B_DONE], 0);
            }
        }
    } else {   // warps 10, 11: idle
        if (a.cs > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    } else {
        // ================================================================ the two softmax / conversion / epilogue warpgroups
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_WG));
        const int g = warp >> 2;                // warpgroup
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // tile row == TMEM lane
        const uint32_t lane_taddr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t sQO_row = smem_u32(sQO) + row * 16;   // element (row, c) of a K-major core-matrix tile:
        const uint32_t sP_row = smem_u32(sP + g * a.p_bytes) + row * 16;   //   (c / 8) * 2048 + row * 16 + (c % 8) * 2
        const float sl = a.scale_log2;
        const int la_end = a.Lt_pad + a.La;
        const bool t0 = lane == 0 && q == 0;    // the thread of this warpgroup that writes trace stamps
        const int tb = 64 + g * 128;

        // ---- phase A: Q accumulator -> 16-bit operand tile; the warpgroups split the 16-column groups
        {
            const int n16 = a.hg * a.d_pad / 16, half = (n16 + 1) / 2;   // own heads' columns of the Q accumulator
            const int c_lo = g == 0 ? 0 : half, c_hi = g == 0 ? half : n16;
            const uint32_t dst0 = sQO_row + (h_first * a.d_pad / 8) * 2048;   // own heads' slot of sQO
            if (t0) SITE_T(tb + 120);
            // norm2 folded into the q projection: mean / rstd of this thread's row from the X chunks in shared memory, while the
            // tensor core multiplies them (both warpgroups, each for all 128 rows: no exchange); applied in the conversion below
            float ln_mean = 0.f, ln_rstd = 1.f;
            if (a.ln_c1) {
                float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int kb = 0; kb < a.n_kb; ++kb) {
                    mbar_wait_nocall(&bars[B_A_FULL + kb], 0);
                    const uint4* rowp = reinterpret_cast<const uint4*>(smem + kb * a.a_stage_bytes + row * 128);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 w = rowp[j ^ (row & 7)];   // SWIZZLE_128B: chunk j of the row sits at j ^ (row & 7)
                        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float lo, hi;
                            unpack16(ww[k], FMT, lo, hi);
                            if (kb == 0 && j == 0 && k == 0) shift = lo;
                            lo -= shift; hi -= shift;
                            s1 += lo + hi;
                            s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                        }
                    }
                }
                const float inv = 1.0f / static_cast<float>(a.C), m = s1 * inv;
                ln_mean = shift + m;
                ln_rstd = rsqrtf(fmaxf(s2 * inv - m * m, 0.f) + a.ln_eps);
                asm volatile("fence.acq_rel.cta;" : "+f"(s1), "+f"(s2) : : "memory");   // the loads have RETURNED before the stages are released (measured in apx_gemm16.cu)
                mbar_arrive(&bars[B_STAT]);
                mbar_wait_nocall(&bars[B_STAT], 0);   // the Q tile below is written OVER the phase-A stages: every row must have been read
            }
            mbar_wait_nocall(&bars[B_Q], 0);
            tc_fence_after();
            if (a.cs > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // phase-A stages are dead here
            if (t0) SITE_T(tb + 121);
            // A compact loop (two 16-column groups per round trip): this code runs ONCE per CTA, so what it costs is mostly its
            // instruction fetch — the 8-group unrolled form below is 10 KB of straight-line code (measured: 1400-1900 cycles cold).
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 2) {
                uint32_t r[32];
                const bool two = c0 + 1 < c_hi;
                tmem_ld16(tmem_base + lane_taddr + c0 * 16, r);
                if (two) tmem_ld16(tmem_base + lane_taddr + (c0 + 1) * 16, r + 16);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (j == 0 || two) {
                        reg_tie16(r + j * 16);
                        if (a.ln_c1) {   // q = rstd (x Wq'^T - mean c1) + c2
                            const float* c1p = sLn + (h_first * a.d_pad) + (c0 + j) * 16;
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                v[i] = fmaf(ln_rstd, __uint_as_float(r[j * 16 + i]) - ln_mean * c1p[i], c1p[a.Nq + i]);
                            store16<FMT>(dst0 + (c0 + j) * 2 * 2048, v);
                        } else {
                            cvt_store16<FMT>(r + j * 16, dst0 + (c0 + j) * 2 * 2048);
                        }
                    }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&bars[B_SQ]);
            if (t0) SITE_T(tb + 122);
        }

        // ---- phase B: this warpgroup's heads.  Order per head: score row -> registers (the buffer goes straight back to the MMA
        //      thread), maxima, convert the O of this warpgroup's PREVIOUS head (its P·V ran meanwhile), exponentials -> P tile.
        //      Nothing in this loop waits for work that was issued after the loop iteration began.
        const uint32_t tO = tmem_O + lane_taddr + g * 2 * a.dv_pad;
        const uint32_t tS = tmem_base + lane_taddr + g * a.L_pad;
        // O_h = O_t / l_t + ap_scale * O_a / l_a -> 16 bit, over Q_h in sQO; l_t, l_a = accumulator column d_pad (ones row of V^T).
        // A segment that does not exist reads the other accumulator with weight 0 (its own columns were never written).
        const uint32_t tOt = a.Lt > 0 ? tO : tO + a.dv_pad, tOa = a.La > 0 ? tO + a.dv_pad : tO;
        auto convert_o = [&](int h, uint32_t ph, int tslot) {
            mbar_wait_nocall(&bars[B_O_FULL + g], ph);
            tc_fence_after();
            if (tslot >= 0) SITE_T(tslot);
            site_convert_head<FMT>(a, tOt, tOa, sQO_row + ((h_first + h) * a.d_pad) / 8 * 2048);
            if (a.cs > 1) {
                // all-gather of the O slices over distributed shared memory, head by head: push this head's slice into every
                // peer's sQO as soon as it exists; each copy completes its bytes on the peer's B_OALL barrier
                fence_proxy_async_smem();
                named_bar_sync(5 + g, 128);
                if (t0) {
                    const uint32_t head_bytes = (a.d_pad / 8) * 2048;
                    const uint32_t slot = smem_u32(sQO) + (h_first + h) * head_bytes, bar = smem_u32(&bars[B_OALL]);
#pragma unroll 1
                    for (int r = 0; r < a.cs; ++r)
                        if (r != rank) dsmem_bulk_copy(map_to_cta(slot, r), sQO + (h_first + h) * head_bytes, head_bytes, map_to_cta(bar, r));
                }
            }
        };
        // The exp2 passes of the two warpgroups take turns (named barriers 7 / 8 are the token: "warpgroup 0 / 1 may use the MUFU
        // pipe"): run at the same time they only split the pipe and then idle together; alternating, one warpgroup's score load,
        // maxima, O conversion and barrier waits hide behind the other's exp2 pass.  Every arrive has a taker (barriers end balanced).
        // With one head per warpgroup (clusters of 4) there is nothing to hide behind: both go at once.
        const bool tok = a.hg > 2;
        if (tok && g == 1) named_bar_arrive(7, 256);
        bool cl_waited = false;
#pragma unroll 1
        for (int it = 0;; ++it) {
            const int h = g + 2 * it;   // local head index
            const bool live = h < a.hg;
            const int ts = tb + it * 8;
            const bool tr = t0 && it < 15 && live;
            float mx_t = -INFINITY, mx_a = -INFINITY;
            uint32_t s[(N16 > 0 ? N16 : 1) * 16];
            if (live) {
                if (tr) SITE_T(ts + 0);
                mbar_wait_nocall(&bars[B_S_FULL + g], it & 1);
                tc_fence_after();
                if (tr) SITE_T(ts + 1);
                // two softmaxes over one score row: text columns [0, Lt), audio columns [Lt_pad, Lt_pad + La); P stays unnormalised
                if (N16 > 0) {
#pragma unroll
                    for (int c = 0; c < N16; ++c) tmem_ld16(tS + c * 16, s + c * 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < N16; ++c) reg_tie16(s + c * 16);
                    tc_fence_before();
                    mbar_arrive(&bars[B_S_FREE + g]);   // the scores of this warpgroup's next head may be issued now
                    if (tr) SITE_T(ts + 6);
                    // padded columns (only the last group of a segment can have any) -> -inf
                    if (LT16 > 0 && a.Lt < LT16 * 16) mask16(s + (LT16 > 0 ? LT16 - 1 : 0) * 16, a.Lt - (LT16 - 1) * 16);
                    if (N16 > LT16 && la_end < N16 * 16) mask16(s + (N16 - 1) * 16, la_end - (N16 - 1) * 16);
#pragma unroll
                    for (int c = 0; c < N16; ++c) {
                        if (c < LT16) mx_t = fmaxf(mx_t, max16(s + c * 16)); else mx_a = fmaxf(mx_a, max16(s + c * 16));
                    }
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
                }
                if (tr) SITE_T(ts + 2);
            }
            // the O of this warpgroup's previous head: its P·V ran under the code above; also frees this warpgroup's P tile
            if (it > 0) {
                if (a.cs > 1 && !cl_waited) {   // every peer's barriers exist and its phase-A stages are dead (see the handshake note)
                    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
                    cl_waited = true;
                }
                convert_o(h - 2, (it - 1) & 1, tr ? ts + 5 : -1);
            }
            if (!live) break;
            if (tr) SITE_T(ts + 4);
            if (tok) named_bar_sync(7 + g, 256);
            const float mxs_t = mx_t * sl, mxs_a = mx_a * sl;
            if (N16 > 0) {
#pragma unroll
                for (int c = 0; c < N16; ++c) exp_store16<FMT>(s + c * 16, sl, c < LT16 ? mxs_t : mxs_a, sP_row + c * 2 * 2048);
            } else {
#pragma unroll 1
                for (int c = 0; c < a.L_pad / 16; ++c) {   // pass 2: exponentials
                    uint32_t r[16];
                    tmem_ld16(tS + c * 16, r);
                    tmem_ld_wait16(r);
                    const bool text = c * 16 < a.Lt_pad;
                    mask16(r, (text ? a.Lt : la_end) - c * 16);
                    exp_store16<FMT>(r, sl, text ? mxs_t : mxs_a, sP_row + c * 2 * 2048);
                }
                tc_fence_before();
                mbar_arrive(&bars[B_S_FREE + g]);
            }
            if (tok && h + 1 < a.hg) named_bar_arrive(8 - g, 256);   // the other warpgroup's turn
            if (tr) SITE_T(ts + 7);
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&bars[B_P_FULL + g]);
            if (tr) SITE_T(ts + 3);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[B_SO_ALL]);

        // ---- phase C epilogue: Y accumulator + bias -> 16 bit -> SWIZZLE_128B staging boxes -> TMA store
        if (a.cs > 1 && !cl_waited) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (t0) SITE_T(tb + 123);
        mbar_wait_nocall(&bars[B_Y], 0);
        tc_fence_after();
        if (t0) SITE_T(tb + 124);
        {
            // 64-column boxes alternate between the warpgroups; each box is converted by one warpgroup (4 loads, one round trip),
            // fenced, and stored by that warpgroup's first thread while the next box is being converted.
            // staging boxes lie over the dead P / K / V / Wo regions (sQO may still be the source of this CTA's pushes)
            const uint32_t stage_row = smem_u32(sP) + row * 128;
            const int sw = row & 7;
            const int kb0 = rank * (a.cn / 64);   // first own 64-column box of Y
#pragma unroll 1
            for (int kb = g; kb < a.cn / 64; kb += 2) {
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
                        const float2 bb = *reinterpret_cast<const float2*>(sBias + (kb0 + kb) * 64 + j * 16 + 2 * i);
                        p[i] = pack2<FMT>(__uint_as_float(r[j * 16 + 2 * i]) + bb.x, __uint_as_float(r[j * 16 + 2 * i + 1]) + bb.y);
                    }
                    // 16-byte chunks 2j, 2j+1 of the 128-byte row; chunk k lives at k ^ (row & 7)  (SWIZZLE_128B)
                    st_shared_v4(box + (((2 * j) ^ sw) << 4), p[0], p[1], p[2], p[3]);
                    st_shared_v4(box + (((2 * j + 1) ^ sw) << 4), p[4], p[5], p[6], p[7]);
                }
                fence_proxy_async_smem();
                named_bar_sync(3 + g, 128);
                if (t0) {
                    // accumulate: Y holds the block's residual stream, the box is added to it (one 16-bit rounding, as h + y)
                    if (a.accumulate) tma_reduce_add_3d(&tmY, sP + kb * X_CHUNK_BYTES, (kb0 + kb) * 64, m0, b);
                    else tma_store_3d(&tmY, sP + kb * X_CHUNK_BYTES, (kb0 + kb) * 64, m0, b);
                    tma_store_commit();
                }
            }
            if (t0) {
                SITE_T(tb + 125);
                tma_store_wait_read();
                SITE_T(tb + 126);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem_base);
}

}  // namespace

SitePlan site_plan(const SiteGeom& g, int Lt, int La, int cs) {
    SitePlan p;
    memset(&p, 0, sizeof(p));
    const int ci = cs == 1 ? 0 : cs == 2 ? 1 : cs == 4 ? 2 : -1;
    if (!g.ok || ci < 0 || !g.cs_ok[ci] || Lt < 0 || La < 0 || Lt + La <= 0) return p;
    const KvPackGeom kg = kv_pack_geom(Lt, La, g.C, g.H);
    if (kg.L_pad < 16 || kg.L_pad > 256) return p;
    // tensor-memory columns of phase B: S(0) S(1) | O_t O_a (0) | O_t O_a (1)
    if (2 * kg.L_pad + 4 * kg.dv_pad > 512) return p;
    if (g.n_kb > 4) return p;
    const uint32_t qo_bytes = static_cast<uint32_t>(std::max(g.Nq, g.C)) * 256;   // sQO: the full O tile of the query tile
    p.a_stage_bytes = X_CHUNK_BYTES + static_cast<uint32_t>(g.Nq / cs) * 128;     // X chunk + own rows of the Wq chunk
    p.p_bytes = 128u * kg.L_pad * 2;
    const uint32_t wo_stage = WO_KSTEPS_PER_STAGE * static_cast<uint32_t>(g.C / cs) * 32;
    p.n_wo = (g.Nq / 16 + WO_KSTEPS_PER_STAGE - 1) / WO_KSTEPS_PER_STAGE;
    if (p.n_wo > 4) return p;
    const uint32_t limit = 227u * 1024u - 1024u;   // minus the alignment slack of the dynamic shared-memory base
    const int ring_max = std::min(3, g.H / cs), ring_min = std::min(2, g.H / cs);
    const int earlies[3] = {p.n_wo, 1, 0};         // Wo stages with a region of their own (prefetched right after phase A)
    for (int ei = 0; ei < 3; ++ei)
        for (int ring = ring_max; ring >= ring_min; --ring) {
            const int early = earlies[ei];
            uint32_t off = qo_bytes;
            p.off_p = off; off += static_cast<uint32_t>(align_up(2 * p.p_bytes, 1024));
            p.off_k = off; off += static_cast<uint32_t>(align_up(ring * kg.k_tile_bytes, 128));
            p.off_v = off; off += static_cast<uint32_t>(align_up(ring * kg.vx_tile_bytes, 1024));
            // Wo stages that are loaded after the last P·V lie over the dead P tiles and K / V rings (extended if those are small)
            off = std::max(off, p.off_p + (p.n_wo - early) * wo_stage);
            for (int i = early; i < p.n_wo; ++i) p.wo_off[i] = p.off_p + (i - early) * wo_stage;
            for (int i = 0; i < early; ++i) { p.wo_off[i] = off; off += wo_stage; }
            off = std::max(off, static_cast<uint32_t>(g.n_kb) * p.a_stage_bytes);   // the phase-A stages overlay all of the above
            off = std::max(off, p.off_p + static_cast<uint32_t>(g.C / cs / 64) * X_CHUNK_BYTES);   // and so do the Y staging boxes
            p.off_bar = static_cast<uint32_t>(align_up(off, 128));
            const uint32_t total = p.off_bar + BAR_REGION_BYTES + static_cast<uint32_t>(g.C) * 4 + static_cast<uint32_t>(g.Nq) * 8;   // bias | c1, c2
            if (total <= limit) {
                p.ok = true; p.nk = ring; p.nv = ring; p.wo_early = early; p.smem_bytes = total + 1024;
                return p;
            }
        }
    return p;
}

bool ipattn_site_supported(int C, int H, int Lt, int La) {
    const SiteGeom g = site_geom(C, H);
    return g.ok && site_plan(g, Lt, La, 1).ok;
}

int ipattn_site_pack_weights(const void* Wq, const void* Wo, void* pack, int C, int H, int dtype, cudaStream_t st) {
    const SiteGeom g = site_geom(C, H);
    APX_REQUIRE(g.ok, APX_EUNSUPPORTED, "ipattn_site_pack: the site kernel does not cover C=%d H=%d", C, H);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "ipattn_site_pack: dtype %d", dtype);
    site_pack_kernel<<<148 * 8, 256, 0, st>>>(static_cast<const uint16_t*>(Wq), static_cast<const uint16_t*>(Wo),
                                              static_cast<uint16_t*>(pack), g);
    APX_LAUNCH_CHECK("site_pack_kernel");
    return APX_OK;
}

// Cluster size for a launch: the largest of {4, 2, 1} that divides the heads, leaves 64-column multiples per CTA and still fits
// the grid into one wave of the 148 SMs (one CTA per SM); more CTAs per tile = a shorter per-CTA head loop.
int ipattn_site_pick_cluster(int Bc, int S, int C, int H) {
    static const int forced = [] { const char* e = getenv("APX_SITE_CLUSTER"); return e ? atoi(e) : 0; }();
    const int tiles = ceil_div(S, 128) * Bc;
    for (int cs = 4; cs >= 2; cs >>= 1) {
        if (H % cs || (C / cs) % 64) continue;
        if (forced == cs) return cs;
        if (forced == 0 && tiles * cs <= 148) return cs;
    }
    return 1;
}

int ipattn_site(const void* X, const void* wpack, const void* cache_tiles, const float* bo, void* Y, int Bc, int S, int C,
                int H, int Lt, int La, float ap_scale, int dtype, bool accumulate, cudaStream_t st, const float* ln_c1, const float* ln_c2,
                float ln_eps) {
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "ipattn_site: dtype %d", dtype);
    int cs = ipattn_site_pick_cluster(Bc, S, C, H);
    const SiteGeom g = site_geom(C, H);
    APX_REQUIRE(g.ok, APX_EUNSUPPORTED, "ipattn_site: C=%d H=%d is not covered", C, H);
    SitePlan p = site_plan(g, Lt, La, cs);
    if (!p.ok && cs > 1) { cs = 1; p = site_plan(g, Lt, La, 1); }
    APX_REQUIRE(p.ok, APX_EUNSUPPORTED, "ipattn_site: context %d+%d does not fit the site kernel at C=%d", Lt, La, C);
    const int ci = cs == 1 ? 0 : cs == 2 ? 1 : 2;
    APX_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wpack) & 15) == 0 && (reinterpret_cast<uintptr_t>(cache_tiles) & 15) == 0,
                APX_EINVAL, "ipattn_site: X / Y / pack / cache must be 16-byte aligned");
    APX_REQUIRE(Bc <= 65535, APX_EUNSUPPORTED, "ipattn_site: Bc=%d", Bc);
    const KvPackGeom kg = kv_pack_geom(Lt, La, C, H);

    CUtensorMap tmX, tmY;
    {
        // [Bc][S][C] viewed as a 3-D tensor so that the rows of a tile past S are zero-filled (load) / clipped (store)
        uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(S), static_cast<uint64_t>(Bc)};
        uint64_t strides[2] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(S) * C * 2};
        uint32_t box[3] = {64, 128, 1};
        int rc = encode_tmap(&tmX, dtype, 3, X, dims, strides, box, 1);
        if (rc) return rc;
        rc = encode_tmap(&tmY, dtype, 3, Y, dims, strides, box, 1);
        if (rc) return rc;
    }
    SiteArgs a;
    const uint8_t* pb = static_cast<const uint8_t*>(wpack);
    a.wq = pb + g.wq_off[ci];
    a.wo = pb + g.wo_off[ci];
    a.tiles = static_cast<const uint8_t*>(cache_tiles);
    a.bo = bo;
    a.S = S; a.C = C; a.H = H; a.d_pad = g.d_pad; a.dv_pad = kg.dv_pad; a.Nq = g.Nq; a.n_kb = g.n_kb;
    a.Lt = Lt; a.La = La; a.Lt_pad = kg.Lt_pad; a.La_pad = kg.La_pad; a.L_pad = kg.L_pad;
    a.ap_scale = ap_scale;
    a.accumulate = accumulate ? 1 : 0;
    APX_REQUIRE((ln_c1 == nullptr) == (ln_c2 == nullptr), APX_EINVAL, "ipattn_site: c1 and c2 of the folded LayerNorm come together");
    a.ln_c1 = ln_c1; a.ln_c2 = ln_c2; a.ln_eps = ln_eps;
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(g.d));
    a.kv_bytes = static_cast<uint32_t>(kg.bytes_per_bh);
    a.k_tile_bytes = static_cast<uint32_t>(kg.k_tile_bytes);
    a.v_tile_bytes = static_cast<uint32_t>(kg.vx_tile_bytes);   // the site kernel streams the EXTENDED V^T tile
    a.vx_off = static_cast<uint32_t>(kg.k_tile_bytes + kg.v_tile_bytes);
    a.wq_chunk_bytes = static_cast<uint32_t>(g.Nq / cs) * 128;    // [8 k-groups][Nq / cs rows][8] x 2 B
    a.wo_kstep_bytes = static_cast<uint32_t>(C / cs) * 32;        // [2 k-groups][C / cs columns][8] x 2 B
    a.a_stage_bytes = p.a_stage_bytes;
    a.p_bytes = p.p_bytes;
    a.nk = p.nk; a.nv = p.nv; a.n_wo = p.n_wo; a.wo_early = p.wo_early;
    a.cs = cs; a.hg = H / cs; a.cn = C / cs;
    a.off_p = p.off_p; a.off_k = p.off_k; a.off_v = p.off_v; a.off_bar = p.off_bar;
    for (int i = 0; i < 4; ++i) a.wo_off[i] = p.wo_off[i];

    // diagnostic timeline (APX_SITE_TRACE=1): synchronises and prints, never enabled on the product path
    static const bool trace_on = [] { const char* e = getenv("APX_SITE_TRACE"); return e && e[0] == '1'; }();
    static long long* trace_buf = nullptr;
    a.trace = nullptr;
    if (trace_on) {
        if (!trace_buf) APX_CUDA(cudaMalloc(&trace_buf, 320 * sizeof(long long)));
        APX_CUDA(cudaMemsetAsync(trace_buf, 0, 320 * sizeof(long long), st));
        a.trace = trace_buf;
    }
    // kernel variant: score rows of 32 / 48 / 144 columns (La = 8 / 32 / 128 behind 1..16 text tokens) stay in registers
    const int n16 = kg.L_pad / 16;
    const int variant = kg.Lt_pad != 16 ? 0 : n16 == 2 ? 1 : n16 == 3 ? 2 : n16 == 9 ? 3 : 0;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, SiteArgs);
    static const KernelFn table[3][4] = {
        {ipattn_site_kernel<0, 0, 0, false>, ipattn_site_kernel<0, 2, 1, false>, ipattn_site_kernel<0, 3, 1, false>, ipattn_site_kernel<0, 9, 1, false>},
        {ipattn_site_kernel<1, 0, 0, false>, ipattn_site_kernel<1, 2, 1, false>, ipattn_site_kernel<1, 3, 1, false>, ipattn_site_kernel<1, 9, 1, false>},
        {ipattn_site_kernel<0, 0, 0, true>, ipattn_site_kernel<0, 2, 1, true>, ipattn_site_kernel<0, 3, 1, true>, ipattn_site_kernel<0, 9, 1, true>}};
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j)
                APX_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(table[i][j]), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              227 * 1024));
    }
    const KernelFn fn = table[(trace_on && dtype == APX_F16) ? 2 : (dtype == APX_F16 ? 0 : 1)][variant];
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(cs) * ceil_div(S, 128), Bc);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = static_cast<unsigned>(cs);
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = cs > 1 ? 1 : 0;
    // no programmatic-launch attribute on the site kernel itself: its predecessor in the step is a PyTorch kernel (nothing to overlap), and
    // back-to-back site launches overlapping each other would flatter the kernel's stand-alone timing.  It still TRIGGERS its dependents.
    APX_CUDA(cudaLaunchKernelEx(&cfg, fn, tmX, tmY, a));
    APX_LAUNCH_CHECK("ipattn_site_kernel");
    if (trace_on) {
        static int printed = 0;
        long long t[320];
        APX_CUDA(cudaStreamSynchronize(st));
        APX_CUDA(cudaMemcpy(t, trace_buf, sizeof(t), cudaMemcpyDeviceToHost));
        if (printed++ < 2) {
            const long long t0 = t[0];
            auto rel = [&](int slot) { return t[slot] ? t[slot] - t0 : -1; };
            printf("[apx site trace] Bc=%d S=%d C=%d H=%d Lt=%d La=%d ring=%d n_wo=%d cluster=%d (cycles since the MMA thread started; CTA 0)\n", Bc, S, C, H,
                   Lt, La, p.nk, p.n_wo, cs);
            printf("  mma : a_full0 %lld q_committed %lld sq_seen %lld (unused %lld) | so_all_seen %lld wo_full0 %lld y_committed %lld\n",
                   rel(1), rel(2), rel(3), rel(4), rel(5), rel(6), rel(7));
            for (int h = 0; h < H && h < 8; ++h) printf("  mma h%d: p_full_seen %lld issued %lld\n", h, rel(8 + 2 * h), rel(9 + 2 * h));
            printf("  tma : start %lld q_done_seen %lld kv_all_issued %lld pv_all_seen %lld wo_issued %lld\n", rel(32), rel(33), rel(34),
                   rel(35), rel(36));
            for (int g = 0; g < 2; ++g) {
                const int o = 64 + g * 128;
                printf("  wg%d : q_wait %lld q_seen %lld sq_arrived %lld | y_wait %lld y_seen %lld staged %lld stored %lld\n", g, rel(o + 120),
                       rel(o + 121), rel(o + 122), rel(o + 123), rel(o + 124), rel(o + 125), rel(o + 126));
                for (int i = 0; g + 2 * i < H && i < 15; ++i)
                    printf("  wg%d h%d: top %lld s_seen %lld ld_done %lld max_done %lld oprev_done %lld exp_done %lld p_arrived %lld (o_full_seen %lld)\n", g,
                           g + 2 * i, rel(o + i * 8), rel(o + i * 8 + 1), rel(o + i * 8 + 6), rel(o + i * 8 + 2), rel(o + i * 8 + 4),
                           rel(o + i * 8 + 7), rel(o + i * 8 + 3), rel(o + i * 8 + 5));
            }
        }
    }
    return APX_OK;
}

}  // namespace apx
