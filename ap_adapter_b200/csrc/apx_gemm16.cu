// apx_gemm16.cu — the 16-bit-output form of Y = act(LN?(A) · W^T + bias) (+ 16-bit residual): same TMA -> tcgen05 main loop as
// apx_gemm_tc.cu, with an epilogue written for the launch-bound token GEMMs of the UNet (M = 128 .. 2000 rows, K = 256 .. 2560).
//
// What the timeline of the first kernel showed on the B200 (APX_GEMM_TRACE=1, profiles/r2k_gemm_trace.txt): the main loop of
// these GEMMs ends 2.2 us after the launch, and the staged "8 columns of a row per thread" epilogue then took 0.3 - 0.55 us per
// 16-row step — 9 us of a 13 us kernel for the 2000 x 2048 x 256 GEGLU projection — because four warps (one per scheduler) ran a
// long dependent instruction stream full of run-time epilogue switches and generic-address loads.  Here:
//   * 8 epilogue warps (2 per TMEM lane quarter, each owning half of the tile's columns), epilogue options are template
//     parameters, every shared-memory access is an explicit ld/st.shared;
//   * splits == 1: a thread owns one accumulator ROW: tcgen05.ld -> LayerNorm correction / bias / GELU / GEGLU / residual in
//     registers -> 16-bit tile in shared memory (the TMA swizzle pattern, conflict-free) -> ONE TMA tensor store per 64 columns,
//     which also clips the rows beyond M.  No fp32 staging pass;
//   * splits > 1 (cluster split-K over distributed shared memory as before): slabs are indexed by sender rank so the sum over
//     ranks is one base + p * slab address; a folded LayerNorm now works WITH split-K — every CTA's (mean, M2) of its K slice
//     rides in the padding words of its slab rows and the finisher merges them (Chan's update), so the K = 640 projections
//     of the 128-token sites no longer run 10 k-blocks on one SM;
//   * the LayerNorm row statistics are taken by two threads per row (half of each 128-byte row chunk each).
//
// Replaces the nn.Linear / LayerNorm / GEGLU dispatches around the hot path (diffusers BasicTransformerBlock FeedForward and
// attention projections called from pipeline/modeling_audioldm2.py:1140-1156) and timm's Block linears
// (audio_encoder/models_mae.py:220-231).
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include <cstdlib>
#include <vector>

namespace apx {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int NT = 320;                      // warp 0 TMA, warp 1 TMEM + MMA, warps 2..9 epilogue
constexpr int NEPI = 256;
constexpr int LDT = 64 + 4;                  // fp32 row pitch (words) of a split-K partial tile; words 64, 65 carry (mean, M2)

template <int BN>
struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_REGION = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
    static constexpr int OFF_BAR = STAGE_REGION;       // 256 B of mbarriers + the TMEM address slot
    static constexpr int OFF_BIAS = OFF_BAR + 256;     // float[BN]
    static constexpr int OFF_C1 = OFF_BIAS + 512;      // float[BN]
    static constexpr int OFF_STAT = OFF_C1 + 512;      // float2[2][128]: (mean, M2) of each half row
    static constexpr int OFF_RECV = OFF_STAT + 2048;   // split-K: `splits` slabs of 128 / splits rows (own slot unused)
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + OFF_RECV;
};
constexpr int RECV_BYTES = BM * LDT * 4;

struct Args {
    const float* bias;       // fp32 [N] or null (c2 of a folded LayerNorm)
    const float* c1;         // LN: row sums of the gamma-folded weight
    const void* residual;    // 16-bit [M, Nout] or null
    int64_t ldr;
    void* out;               // 16-bit [M, Nout]
    int64_t ldo;
    int M, N, K;
    int splits;
    float eps;
    long long* trace;
};

#define G16_TRACE(slot) \
    do { if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g.trace[(slot)] = clock64(); } while (0)

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 ld_shared_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_shared_f1(uint32_t addr, float a) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
template <int FMT>
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
    if (FMT == 0) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        lo = f.x; hi = f.y;
    } else {
        lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
    }
}

// One output value group: LayerNorm correction, bias, activation.  `x` holds NA accumulator columns starting at tile column `tc`
// (bias / c1 come from shared memory at byte addresses sb / sc); the results (NA, or NA / 2 for GEGLU) are left in x[0 ..).
template <int NA, int ACT, bool LN>
__device__ __forceinline__ void act_group(float (&x)[NA], uint32_t sb, uint32_t sc, int tc, float mean, float rstd) {
#pragma unroll
    for (int i = 0; i < NA; i += 4) {
        const float4 b = ld_shared_f4(sb + (tc + i) * 4);
        if (LN) {
            const float4 c = ld_shared_f4(sc + (tc + i) * 4);
            x[i] = rstd * (x[i] - mean * c.x); x[i + 1] = rstd * (x[i + 1] - mean * c.y);
            x[i + 2] = rstd * (x[i + 2] - mean * c.z); x[i + 3] = rstd * (x[i + 3] - mean * c.w);
        }
        x[i] += b.x; x[i + 1] += b.y; x[i + 2] += b.z; x[i + 3] += b.w;
    }
    if (ACT == APX_EPI_GELU) {
#pragma unroll
        for (int i = 0; i < NA; ++i) x[i] = gelu_erf(x[i]);
    } else if (ACT == APX_EPI_GEGLU) {   // (hidden, gate) column pairs -> hidden * gelu(gate)
#pragma unroll
        for (int i = 0; i < NA / 2; ++i) x[i] = x[2 * i] * gelu_erf(x[2 * i + 1]);
    }
}

// 8 finished values -> 16-bit words; with a residual the value is rounded to the output type first and the sum once more (the
// caller's `y + h` on 16-bit tensors)
template <int FMT>
__device__ __forceinline__ uint4 pack8(const float* v, bool has_res, uint4 res) {
    uint32_t w[4];
    const uint32_t rw[4] = {res.x, res.y, res.z, res.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[i] = pack2<FMT>(v[2 * i], v[2 * i + 1]);
        if (has_res) {
            float a, b, c, d;
            unpack2<FMT>(w[i], a, b);
            unpack2<FMT>(rw[i], c, d);
            w[i] = pack2<FMT>(a + c, b + d);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int BN, int ACT, bool LN, int FMT>
__global__ void __launch_bounds__(NT, 2)
gemm16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg<BN>::OFF_BAR);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_bar = empty_bar + STAGES;
    uint64_t* recv_bar = acc_bar + 1;   // split-K: peers' partial slabs have landed
    uint64_t* done_bar = acc_bar + 2;   // split-K: every peer holds its slab of my tile (my outgoing copies are complete)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 3);
    const uint32_t s_bias = sbase + Cfg<BN>::OFF_BIAS, s_c1 = sbase + Cfg<BN>::OFF_C1, s_stat = sbase + Cfg<BN>::OFF_STAT;
    const uint32_t s_recv = sbase + Cfg<BN>::OFF_RECV;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int splits = g.splits;
    const int z = blockIdx.z;   // == %cluster_ctarank (cluster dims 1 x 1 x splits)
    const int total_kb = g.K / BK;
    const int kb0 = total_kb * z / splits;
    const int num_kb = total_kb * (z + 1) / splits - kb0;

    if (threadIdx.x == 0) G16_TRACE(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], LN ? 1 + NEPI : 1);   // + the statistics threads that read the A stage
        }
        mbar_init(acc_bar, 1);
        mbar_init(recv_bar, 1);
        mbar_init(done_bar, splits > 1 ? splits - 1 : 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    if (splits > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // "my mbarriers exist"
    pdl_wait();                 // everything above is on-chip set-up; from here on the predecessor's results are read
    __syncthreads();
    pdl_launch_dependents();    // this CTA holds its tensor memory: the next kernel may start its own set-up
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) G16_TRACE(1);

    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + Cfg<BN>::B_STAGE_BYTES);
                tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, (kb0 + i) * BK, m0, &full_bar[s]);
                tma_load_2d(sB + s * Cfg<BN>::B_STAGE_BYTES, &tmB, (kb0 + i) * BK, n0, &full_bar[s]);
                if (i == 0) G16_TRACE(2);
            }
            G16_TRACE(3);
        }
        __syncwarp();
        if (splits > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(FMT, BM, BN);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full_bar[s], (i / STAGES) & 1);
                tc_fence_after();
                if (i == 0) G16_TRACE(4);
                const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(sB + s * Cfg<BN>::B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t da = umma_desc(a_addr + k * 32, 16, 1024, kLayoutSW128);
                    const uint64_t db = umma_desc(b_addr + k * 32, 16, 1024, kLayoutSW128);
                    umma_ss(tmem_base, da, db, idesc, (i | k) != 0);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(acc_bar);
            G16_TRACE(5);
        }
        __syncwarp();
        if (splits > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        const int te = threadIdx.x - 64;     // 0 .. 255
        const int q = warp & 3;              // TMEM lane quarter this warp may read
        const int hf = te >> 7;              // which half of the tile's columns this thread owns
        const int trow = q * 32 + lane;      // accumulator row
        if (te < BN) {
            st_shared_f1(s_bias + te * 4, g.bias ? g.bias[n0 + te] : 0.f);
            if (LN) st_shared_f1(s_c1 + te * 4, g.c1[n0 + te]);
        }
        float mean = 0.f, rstd = 1.f, m_cta = 0.f, m2_cta = 0.f;
        if (LN) {
            // Row statistics from the pipeline stages while the tensor core consumes them: this thread reads 4 of the 8 16-byte
            // chunks of its row (SWIZZLE_128B: chunk j is stored at j ^ (row & 7)).  Shifted sums against cancellation.
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
            const uint32_t row_addr = sbase + trow * 128;
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full_bar[s], (i / STAGES) & 1);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const uint4 w = ld_shared_v4(row_addr + s * A_STAGE_BYTES + (((hf * 4 + jj) ^ (trow & 7)) << 4));
                    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float lo, hi;
                        unpack2<FMT>(ww[k], lo, hi);
                        if (i == 0 && jj == 0 && k == 0) shift = lo;
                        lo -= shift; hi -= shift;
                        s1 += lo + hi;
                        s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                    }
                }
                // The stage may only be released once this thread's loads from it have RETURNED.  Nothing ties a plain ld.shared to the
                // mbarrier arrive that follows it: ptxas schedules the arrive right behind the last LDS, and in the instantiation where
                // it did (<64, GEGLU, LN>) the row statistics were wrong in ~1 of 5 launches on the B200 (tools/stress_g16.py: 36/150
                // launches without the fence, 0/150 with it; an empty asm consuming the sums is not enough — ptxas hoists the arrive
                // over the arithmetic again).  The fence waits for the loads.
                asm volatile("fence.acq_rel.cta;" : "+f"(s1), "+f"(s2) : : "memory");
                mbar_arrive(&empty_bar[s]);
            }
            const float n_half = static_cast<float>(num_kb * (BK / 2));
            const float d = s1 / n_half;
            st_shared_f2(s_stat + (hf * 128 + trow) * 8, shift + d, fmaxf(s2 - s1 * d, 0.f));
        }
        // bias / c1 / half-row statistics are visible; every statistics thread is done reading the pipeline stages (the epilogue
        // stages its tile over them)
        named_bar_sync(1, NEPI);
        if (LN) {
            const float2 a = ld_shared_f2(s_stat + trow * 8), b = ld_shared_f2(s_stat + (128 + trow) * 8);
            const float n_half = static_cast<float>(num_kb * (BK / 2));
            const float dlt = b.x - a.x;
            m_cta = 0.5f * (a.x + b.x);
            m2_cta = a.y + b.y + dlt * dlt * (0.5f * n_half);
            mean = m_cta;
            rstd = rsqrtf(m2_cta / static_cast<float>(g.K) + g.eps);   // splits == 1: the CTA's slice is the whole row
            if (te == 0) G16_TRACE(6);
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        constexpr bool GEGLU = ACT == APX_EPI_GEGLU;

        if (splits == 1) {
            // ---------------- one thread per accumulator row and column half, straight from TMEM
            constexpr int OUT_COLS = GEGLU ? BN / 2 : BN;          // output columns of the tile
            constexpr int ROW_BYTES = OUT_COLS >= 64 ? 128 : 64;   // shared-memory row of one TMA store box
            constexpr int NCH = BN / 32;                           // 16-column accumulator chunks per thread
            constexpr int OPC = GEGLU ? 8 : 16;                    // outputs per chunk
            const int row = m0 + trow;
            const int ocol0 = (GEGLU ? n0 / 2 : n0);               // first output column of the tile
            const bool has_res = g.residual != nullptr;
            uint4 rr[NCH][OPC / 8];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int k = 0; k < OPC / 8; ++k) rr[c][k] = make_uint4(0, 0, 0, 0);
            if (has_res && row < g.M) {
                const uint16_t* rp = reinterpret_cast<const uint16_t*>(g.residual) + static_cast<int64_t>(row) * g.ldr + ocol0;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int oc = (hf * (BN / 2) + c * 16) / (GEGLU ? 2 : 1);
#pragma unroll
                    for (int k = 0; k < OPC / 8; ++k) rr[c][k] = *reinterpret_cast<const uint4*>(rp + oc + k * 8);
                }
            }
            mbar_wait(acc_bar, 0);
            tc_fence_after();
            if (te == 0) G16_TRACE(7);
            auto chunk = [&](int c, const uint4* rrc) {
                const int tc = hf * (BN / 2) + c * 16;   // tile column of the chunk
                uint32_t r[16];
                tmem_ld16(taddr + tc, r);
                tmem_ld_wait16(r);
                float x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(r[i]);
                act_group<16, ACT, LN>(x, s_bias, s_c1, tc, mean, rstd);
                const int oc = tc / (GEGLU ? 2 : 1);     // output column within the tile
#pragma unroll
                for (int k = 0; k < OPC / 8; ++k) {
                    const uint4 o = pack8<FMT>(x + k * 8, rrc != nullptr, rrc ? rrc[k] : make_uint4(0, 0, 0, 0));
                    const int occ = oc + k * 8;
                    const int box = occ >> 6, j = (occ & 63) >> 3;
                    const int js = ROW_BYTES == 128 ? (j ^ (trow & 7)) : j;   // 64-byte rows (GEGLU of a 64-column tile): plain layout
                    st_shared_v4(sbase + box * (BM * 128) + trow * ROW_BYTES + (js << 4), o.x, o.y, o.z, o.w);
                }
            };
            if (has_res) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) chunk(c, rr[c]);   // unrolled: the prefetched residual words stay in registers
            } else {
                // This code runs ONCE per CTA, so part of what it costs is its instruction fetch (apx_ipattn_site.cu: -0.5 us per call for
                // 800 fewer cold instructions): one chunk body instead of NCH copies of it (erf-GELU / GEGLU: ~350 instructions each).
                // Back-to-back replays of one GEMM keep the code warm and read the same either way (profiles/r3e_gemm16_epilogue_ab.txt:
                // 3.84 vs 3.87 ms over the step's GEMM mix).
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) chunk(c, nullptr);
            }
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA store
            named_bar_sync(1, NEPI);
            if (te == 0) {
                G16_TRACE(8);
#pragma unroll
                for (int b = 0; b < (OUT_COLS + 63) / 64; ++b) tma_store_2d(&tmO, sbase + b * (BM * 128), ocol0 + b * 64, m0);
                tma_store_commit();
                tma_store_wait_read();
                G16_TRACE(10);
            }
        } else {
            // ---------------- split-K: fp32 partial tile -> slab exchange over DSMEM -> 8 columns of a row per thread
            mbar_wait(acc_bar, 0);
            tc_fence_after();
            if (te == 0) G16_TRACE(7);
            const uint32_t sP = sbase;   // over the pipeline stages: every MMA that read them has completed
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                const int tc = hf * (BN / 2) + c * 16;
                uint32_t r[16];
                tmem_ld16(taddr + tc, r);
                tmem_ld_wait16(r);
#pragma unroll
                for (int i = 0; i < 16; i += 4) st_shared_v4(sP + (trow * LDT + tc + i) * 4, r[i], r[i + 1], r[i + 2], r[i + 3]);
            }
            if (LN && hf == 0) st_shared_f2(sP + (trow * LDT + 64) * 4, m_cta, m2_cta);
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the bulk-copy engine
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' mbarriers are initialised (long since)
            named_bar_sync(1, NEPI);
            if (te == 0) G16_TRACE(8);
            const int rows_per = BM / splits;    // rows of the tile this CTA finishes
            const uint32_t slab_bytes = static_cast<uint32_t>(rows_per) * LDT * 4;
            if (te == 0) mbar_expect_tx(recv_bar, (splits - 1) * slab_bytes);
            if (te < splits && te != z)   // peer `te` finishes rows [te * rows_per, ..): it gets them in its slot `z`
                dsmem_bulk_copy(map_to_cta(s_recv + z * slab_bytes, te), smem + static_cast<size_t>(te) * slab_bytes, slab_bytes,
                                map_to_cta(smem_u32(recv_bar), te));
            mbar_wait(recv_bar, 0);
            if (te < splits && te != z) mbar_arrive_remote(map_to_cta(smem_u32(done_bar), te));   // "I hold all my slabs"
            if (te == 0) G16_TRACE(9);

            const int cg = te & 7, r0 = te >> 3;   // 8 threads per row, 32 rows per step
            const int tc = cg * 8;
            const int ocol = GEGLU ? (n0 + tc) >> 1 : n0 + tc;
            constexpr int NOUT = GEGLU ? 4 : 8;
            const bool has_res = g.residual != nullptr;
            for (int rl = r0; rl < rows_per; rl += 32) {
                const int row = m0 + z * rows_per + rl;
                if (row >= g.M) break;
                uint4 res = make_uint4(0, 0, 0, 0);
                if (has_res) {
                    const uint16_t* rp = reinterpret_cast<const uint16_t*>(g.residual) + static_cast<int64_t>(row) * g.ldr + ocol;
                    if (NOUT == 8) res = *reinterpret_cast<const uint4*>(rp);
                    else { const uint2 t = *reinterpret_cast<const uint2*>(rp); res.x = t.x; res.y = t.y; }
                }
                // partials are summed in rank order whatever z is, so a row's value does not depend on where in a tile it sits
                float x[8];
                float cnt = 0.f, mu = 0.f, m2 = 0.f;
                for (int p = 0; p < splits; ++p) {
                    const uint32_t rp = (p == z ? sP : s_recv) + p * slab_bytes + rl * (LDT * 4);
                    const float4 b0 = ld_shared_f4(rp + tc * 4), b1 = ld_shared_f4(rp + tc * 4 + 16);
                    if (p == 0) {
                        x[0] = b0.x; x[1] = b0.y; x[2] = b0.z; x[3] = b0.w; x[4] = b1.x; x[5] = b1.y; x[6] = b1.z; x[7] = b1.w;
                    } else {
                        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
                    }
                    if (LN) {   // merge the K slices' (mean, M2)
                        const float2 st = ld_shared_f2(rp + 64 * 4);
                        const float np = static_cast<float>((total_kb * (p + 1) / splits - total_kb * p / splits) * BK);
                        const float nt = cnt + np, dlt = st.x - mu;
                        mu += dlt * (np / nt);
                        m2 += st.y + dlt * dlt * (cnt * np / nt);
                        cnt = nt;
                    }
                }
                float rs = 1.f;
                if (LN) rs = rsqrtf(m2 / cnt + g.eps);
                act_group<8, ACT, LN>(x, s_bias, s_c1, tc, mu, rs);
                uint16_t* op = reinterpret_cast<uint16_t*>(g.out) + static_cast<int64_t>(row) * g.ldo + ocol;
                if (NOUT == 8) {
                    *reinterpret_cast<uint4*>(op) = pack8<FMT>(x, has_res, res);
                } else {
                    x[4] = x[5] = x[6] = x[7] = 0.f;
                    const uint4 o = pack8<FMT>(x, has_res, res);
                    *reinterpret_cast<uint2*>(op) = make_uint2(o.x, o.y);
                }
            }
            if (te == 0) G16_TRACE(10);
            mbar_wait(done_bar, 0);   // nobody is still reading this CTA's shared memory
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) G16_TRACE(11);
    if (warp == 1) tmem_dealloc<BN>(tmem_base);
    if (threadIdx.x == 32) G16_TRACE(12);
}

typedef void (*Gemm16Fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, Args);

template <int BN, int ACT, bool LN>
constexpr Gemm16Fn pick_fmt(int fmt) { return fmt == 0 ? gemm16_kernel<BN, ACT, LN, 0> : gemm16_kernel<BN, ACT, LN, 1>; }
template <int BN, int ACT>
Gemm16Fn pick_ln(bool ln, int fmt) { return ln ? pick_fmt<BN, ACT, true>(fmt) : pick_fmt<BN, ACT, false>(fmt); }
template <int BN>
Gemm16Fn pick_act(int act, bool ln, int fmt) {
    return act == APX_EPI_GEGLU ? pick_ln<BN, APX_EPI_GEGLU>(ln, fmt) : act == APX_EPI_GELU ? pick_ln<BN, APX_EPI_GELU>(ln, fmt)
                                                                                               : pick_ln<BN, APX_EPI_NONE>(ln, fmt);
}

// Split-K factor, from the B200 timings of the UNet's token GEMMs (tools/time_gemm_step.py with APX_GEMM_SPLITK=1/2/4, profiles/
// r2k_gemm_splitk.txt): the slab exchange and the narrower finisher cost about 0.75 us, a k-block about 0.23 us, so splitting pays
// from 8 k-blocks (K >= 512) and only while every CTA of the cluster still gets an SM of its own.  APX_GEMM_SPLITK=n forces n.
int pick_splits16(int M, int N, int K) {
    static const int forced = [] { const char* e = getenv("APX_GEMM_SPLITK"); return e ? atoi(e) : 0; }();
    const int tiles = ceil_div(M, BM) * (N / 64), num_kb = K / BK;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return 2 * forced <= num_kb || forced == 1 ? forced : 1;
    if (num_kb < 8) return 1;
    for (int s = 8; s >= 2; s >>= 1)
        if (2 * s <= num_kb && tiles * s <= 148) return s;
    return 1;
}

}  // namespace

bool gemm16_supported(const float* residual, int out_dtype, int in_dtype, int epilogue, const void* out2) {
    static const bool off = [] { const char* e = getenv("APX_GEMM16"); return e && e[0] == '0'; }();
    if (off || out2 || out_dtype != in_dtype) return false;
    return !residual || (epilogue & APX_EPI_RES16);
}

// Same contract as gemm_tc for the cases gemm16_supported() accepts (the arguments were validated there).
int gemm16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* residual, int64_t ldr, void* out,
           int64_t ldo, int M, int N, int K, int dtype, int epilogue, cudaStream_t st, const float* ln_c1, float ln_eps) {
    const int act = epilogue & 3;
    const bool ln = (epilogue & APX_EPI_LNFOLD) != 0;
    const bool wide = (N % 128 == 0) && (static_cast<int64_t>(ceil_div(M, BM)) * (N / 128) >= 148);
    const int splits = (wide || (epilogue & APX_EPI_NOSPLIT)) ? 1 : pick_splits16(M, N, K);
    const int BN = wide ? 128 : 64;
    const int n_out = act == APX_EPI_GEGLU ? N / 2 : N;
    const int out_cols = act == APX_EPI_GEGLU ? BN / 2 : BN;

    CUtensorMap tmA, tmB, tmO;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
        uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
        uint32_t box[2] = {BK, BM};
        int rc = encode_tmap(&tmA, dtype, 2, A, dims, strides, box, 1);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
        uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
        uint32_t box[2] = {BK, static_cast<uint32_t>(BN)};
        int rc = encode_tmap(&tmB, dtype, 2, W, dims, strides, box, 1);
        if (rc) return rc;
    }
    {   // store boxes: 64 output columns (128-byte rows, SWIZZLE_128B) or the tile's 32 (GEGLU of a 64-column tile, plain rows)
        uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(M)};
        uint64_t strides[1] = {static_cast<uint64_t>(ldo) * 2};
        uint32_t box[2] = {static_cast<uint32_t>(out_cols >= 64 ? 64 : out_cols), BM};
        int rc = encode_tmap(&tmO, dtype, 2, out, dims, strides, box, out_cols >= 64 ? 1 : 0);
        if (rc) return rc;
    }
    const Gemm16Fn fn = wide ? pick_act<128>(act, ln, dtype == APX_F16 ? 0 : 1) : pick_act<64>(act, ln, dtype == APX_F16 ? 0 : 1);
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        for (int a = 0; a < 3; ++a)
            for (int l = 0; l < 2; ++l)
                for (int f = 0; f < 2; ++f) {
                    APX_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_act<64>(a, l != 0, f)),
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM_BYTES + RECV_BYTES));
                    APX_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_act<128>(a, l != 0, f)),
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES));
                }
    }
    Args g{bias, ln_c1, residual, ldr, out, ldo, M, N, K, splits, ln_eps, nullptr};
    // diagnostic timeline (APX_GEMM_TRACE=1): synchronises and prints, never enabled on the product path
    static const bool trace_on = [] { const char* e = getenv("APX_GEMM_TRACE"); return e && e[0] == '1'; }();
    static long long* trace_buf = nullptr;
    if (trace_on) {
        if (!trace_buf) APX_CUDA(cudaMalloc(&trace_buf, 16 * sizeof(long long)));
        APX_CUDA(cudaMemsetAsync(trace_buf, 0, 16 * sizeof(long long), st));
        g.trace = trace_buf;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(M, BM), N / BN, splits);
    cfg.blockDim = dim3(NT);
    static const bool one_cta = [] { const char* e = getenv("APX_G16_ONE_CTA"); return e && e[0] == '1'; }();   // diagnostic: one CTA per SM
    cfg.dynamicSmemBytes = (wide ? Cfg<128>::SMEM_BYTES : Cfg<64>::SMEM_BYTES) + ((splits > 1 || one_cta) ? RECV_BYTES : 0);
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 1;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = static_cast<unsigned>(splits);
    cfg.attrs = attrs;
    cfg.numAttrs = splits > 1 ? 1 : 0;
    add_pdl_attr(attrs, &cfg.numAttrs);
    APX_CUDA(cudaLaunchKernelEx(&cfg, fn, tmA, tmB, tmO, g));
    APX_LAUNCH_CHECK("gemm16_kernel");
    if (trace_on) {
        long long t[16];
        APX_CUDA(cudaStreamSynchronize(st));
        APX_CUDA(cudaMemcpy(t, trace_buf, sizeof(t), cudaMemcpyDeviceToHost));
        auto rel = [&](int slot) { return t[slot] ? t[slot] - t[0] : -1; };
        printf("[apx gemm16 trace] M=%d N=%d K=%d BN=%d splits=%d epi=%d | setup %lld | tma first %lld all %lld | mma first_full %lld committed %lld | "
               "epi stats %lld acc_seen %lld staged %lld exchanged %lld stored %lld | sync %lld dealloc %lld (cycles, CTA 0)\n",
               M, N, K, BN, splits, epilogue, rel(1), rel(2), rel(3), rel(4), rel(5), rel(6), rel(7), rel(8), rel(9), rel(10), rel(11), rel(12));
        fflush(stdout);
    }
    return APX_OK;
}

}  // namespace apx
