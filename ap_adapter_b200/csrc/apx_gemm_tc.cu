// apx_gemm_tc.cu — Y = act(A · W^T + bias) + residual on the 5th-gen tensor cores.
//
// One CTA owns a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: 128x64 A box + BNx64 W box per k-block, SWIZZLE_128B, 3-stage mbarrier ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accum in TMEM)
//   warps 2..5  epilogue: tcgen05.ld (lane quarter = warp_id % 4) -> bias / erf-GELU / fp32 residual -> store
//
// Split-K (small grids): `splits` CTAs of one thread-block cluster (cluster dims 1 x 1 x splits) each accumulate a slice of K for
// the same 128 x 64 tile in their own TMEM, dump the fp32 partial tile to their own shared memory, push each peer the 128 / splits
// rows that peer finishes (cp.async.bulk shared::cta -> shared::cluster, completing bytes on the peer's mbarrier) and then run the
// epilogue on their own row slice.  No workspace, no second kernel, and the epilogue is spread over the cluster as well.  A CTA
// streaming a long K on its own is bound by one SM's L2 port (~100 GB/s): at M <= 512 tokens this is what the GEMM time was.
//
// Replaces the nn.Linear dispatches of the reference hot path
// (APadapter/ap_adapter/attention_processor.py:387,406-407,435-436,457; timm Block linears used by
// audio_encoder/models_mae.py:220-231).
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include <cstdlib>

namespace apx {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct GemmCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_REGION = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGE_REGION + 256 + 1024;   // barriers | LayerNorm row statistics; + recv slabs when splits > 1
};

struct GemmArgs {
    const float* bias;
    const float* residual;
    int64_t ldr;
    void* out;
    int64_t ldo;
    int M, N, K;
    int in_fmt;     // 0 f16, 1 bf16
    int out_dtype;  // APX_F16 / APX_BF16 / APX_F32
    int epilogue;
    int splits;     // CTAs per cluster along K (1, 2, 4 or 8; > 1 only with BN = 64)
    void* out2;           // optional second output [M, N]: the same values in the 16-bit INPUT format (row pitch ldo2 elements) — the
    int64_t ldo2;         // A operand of the next GEMM when `out` is an fp32 residual stream
    const float* ln_c1;   // APX_EPI_LNFOLD: c1[n] = sum_k W'[n, k] of the gamma-folded weight (bias then holds c2[n] = sum_k beta_k W[n, k] + b[n])
    float ln_eps;
    long long* trace;     // diagnostic (APX_GEMM_TRACE=1): clock64 stamps of CTA (0, 0, 0)
};

#define GEMM_TRACE(slot) \
    do { if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g.trace[(slot)] = clock64(); } while (0)

constexpr int LDP = 64 + 4;                        // fp32 row pitch of a staged 64-column tile in shared memory (words)
constexpr int MAX_RECV_BYTES = 7 * 16 * LDP * 4;   // (splits - 1) slabs of 128 / splits rows, largest at splits = 8

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Raw residual words of one 8-column accumulator group (loaded one row ahead of their use).
struct Res8 { uint4 a, b; };

__device__ __forceinline__ Res8 load_res8(const GemmArgs& g, int row, int ocol, int nout) {
    Res8 r;
    r.a = make_uint4(0, 0, 0, 0); r.b = r.a;
    if (g.epilogue & APX_EPI_RES16) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(g.residual) + static_cast<int64_t>(row) * g.ldr + ocol;
        if (nout == 8) r.a = *reinterpret_cast<const uint4*>(p);
        else { const uint2 t = *reinterpret_cast<const uint2*>(p); r.a.x = t.x; r.a.y = t.y; }
    } else {
        const float* p = g.residual + static_cast<int64_t>(row) * g.ldr + ocol;
        r.a = *reinterpret_cast<const uint4*>(p);
        if (nout == 8) r.b = *reinterpret_cast<const uint4*>(p + 4);
    }
    return r;
}

// bias -> activation -> residual -> store for 8 consecutive accumulator columns of one row (col % 8 == 0)
__device__ __forceinline__ void finish8(const GemmArgs& g, int row, int ocol, int nout, float (&v)[8], const float (&bias)[8],
                                        const Res8& res, const float (&c1)[8], float2 stat) {
    if (g.epilogue & APX_EPI_LNFOLD) {
        // LayerNorm folded into the GEMM: the accumulator is x · (W gamma)^T over the RAW rows; (x - mean) rstd comes out as
        // rstd (acc - mean c1), and bias holds beta · W^T + b
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = stat.y * (v[i] - stat.x * c1[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += bias[i];
    const int epi = g.epilogue & 3;
    if (epi == APX_EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    } else if (epi == APX_EPI_GEGLU) {   // (hidden, gate) pairs -> hidden * gelu(gate): 4 outputs at column col / 2
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = v[2 * i] * gelu_erf(v[2 * i + 1]);
    }
    if (g.residual) {
        if (g.epilogue & APX_EPI_RES16) {
            // the 16-bit residual is the caller's `y + h`: y is rounded to the output type first, then the sum is rounded once more
#pragma unroll
            for (int i = 0; i < 8; i += 2) unpack16(pack16(v[i], v[i + 1], g.out_dtype), g.out_dtype, v[i], v[i + 1]);
            float lo, hi;
            unpack16(res.a.x, g.out_dtype, lo, hi); v[0] += lo; v[1] += hi;
            unpack16(res.a.y, g.out_dtype, lo, hi); v[2] += lo; v[3] += hi;
            unpack16(res.a.z, g.out_dtype, lo, hi); v[4] += lo; v[5] += hi;   // zero words when nout == 4
            unpack16(res.a.w, g.out_dtype, lo, hi); v[6] += lo; v[7] += hi;
        } else {
            v[0] += __uint_as_float(res.a.x); v[1] += __uint_as_float(res.a.y); v[2] += __uint_as_float(res.a.z); v[3] += __uint_as_float(res.a.w);
            v[4] += __uint_as_float(res.b.x); v[5] += __uint_as_float(res.b.y); v[6] += __uint_as_float(res.b.z); v[7] += __uint_as_float(res.b.w);
        }
    }
    if (g.out2) {
        uint16_t* op2 = reinterpret_cast<uint16_t*>(g.out2) + static_cast<int64_t>(row) * g.ldo2 + ocol;
        const uint32_t q0 = pack16(v[0], v[1], g.in_fmt), q1 = pack16(v[2], v[3], g.in_fmt);
        if (nout == 8) *reinterpret_cast<uint4*>(op2) = make_uint4(q0, q1, pack16(v[4], v[5], g.in_fmt), pack16(v[6], v[7], g.in_fmt));
        else *reinterpret_cast<uint2*>(op2) = make_uint2(q0, q1);
    }
    if (g.out_dtype == APX_F32) {
        float* op = reinterpret_cast<float*>(g.out) + static_cast<int64_t>(row) * g.ldo + ocol;
        *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
        if (nout == 8) *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint16_t* op = reinterpret_cast<uint16_t*>(g.out) + static_cast<int64_t>(row) * g.ldo + ocol;
        const uint32_t p0 = pack16(v[0], v[1], g.out_dtype), p1 = pack16(v[2], v[3], g.out_dtype);
        if (nout == 8) *reinterpret_cast<uint4*>(op) = make_uint4(p0, p1, pack16(v[4], v[5], g.out_dtype), pack16(v[6], v[7], g.out_dtype));
        else *reinterpret_cast<uint2*>(op) = make_uint2(p0, p1);
    }
}

template <int BN>
__global__ void __launch_bounds__(192, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * GemmCfg<BN>::B_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_bar = empty_bar + STAGES;
    uint64_t* recv_bar = acc_bar + 1;   // split-K: peers' partial slabs have landed in sRecv
    uint64_t* done_bar = acc_bar + 2;   // split-K: every peer has received everything (my outgoing copies are complete)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 3);
    float2* sStat = reinterpret_cast<float2*>(smem + GemmCfg<BN>::STAGE_REGION + 256);   // (mean, rstd) of the tile's 128 rows (APX_EPI_LNFOLD)
    float* sRecv = reinterpret_cast<float*>(smem + GemmCfg<BN>::STAGE_REGION + 256 + 1024);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int splits = g.splits;
    const int z = blockIdx.z;   // == %cluster_ctarank (cluster dims 1 x 1 x splits, gridDim.z == splits)
    const int kb0 = (g.K / BK) * z / splits;
    const int num_kb = (g.K / BK) * (z + 1) / splits - kb0;

    if (threadIdx.x == 0) GEMM_TRACE(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], (g.epilogue & APX_EPI_LNFOLD) ? 1 + 128 : 1);   // + the 128 statistics threads that read the A stage
        }
        mbar_init(acc_bar, 1);
        mbar_init(recv_bar, 1);
        mbar_init(done_bar, splits > 1 ? splits - 1 : 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    // cluster barrier phase 1: "my mbarriers exist".  The matching wait sits right before each warp's first remote operation / exit.
    if (splits > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    pdl_wait();                 // programmatic dependent launch: on-chip set-up above, the predecessor's results below
    __syncthreads();
    pdl_launch_dependents();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) GEMM_TRACE(1);

    if (warp == 0) {
        if (elect_one()) {   // elect.sync: the TMA / MMA operands stay in uniform registers (no per-instruction R2UR waterfall)
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + GemmCfg<BN>::B_STAGE_BYTES);
                tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, (kb0 + i) * BK, m0, &full_bar[s]);
                tma_load_2d(sB + s * GemmCfg<BN>::B_STAGE_BYTES, &tmB, (kb0 + i) * BK, n0, &full_bar[s]);
                if (i == 0) GEMM_TRACE(2);
            }
            GEMM_TRACE(3);
        }
        __syncwarp();
        if (splits > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(g.in_fmt, BM, BN);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (i == 0) GEMM_TRACE(4);
                const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(sB + s * GemmCfg<BN>::B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // SW128 K-major: 8-row groups are 1024 B apart; a 16-element k-step is +32 B.
                    const uint64_t da = umma_desc(a_addr + k * 32, 16, 1024, kLayoutSW128);
                    const uint64_t db = umma_desc(b_addr + k * 32, 16, 1024, kLayoutSW128);
                    umma_ss(tmem_base, da, db, idesc, (i | k) != 0);
                }
                umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
            }
            umma_commit(acc_bar);  // accumulator complete
            GEMM_TRACE(5);
        }
        __syncwarp();
        if (splits > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        // ---- epilogue warps.  (1) TMEM lane quarter (warp % 4), one accumulator row per thread -> fp32 tile in shared memory;
        // (2) split-K only: slab exchange with the peers; (3) BN / 8 threads per row, 8 columns each: bias / activation / residual /
        // store with every global access a contiguous row segment (the residual is fetched one row ahead).
        const int q = warp & 3;
        const int trow = q * 32 + lane;
        const int te = threadIdx.x - 64;
        constexpr int LDT = BN + 4;          // fp32 row pitch of the staged tile (words)
        constexpr int TPR = BN / 8;          // threads per row in (3)
        constexpr int ROWS_STEP = 128 / TPR;
        const int cg = te % TPR, r0 = te / TPR;
        const int col = n0 + cg * 8;
        const bool geglu = (g.epilogue & 3) == APX_EPI_GEGLU;
        const int ocol = geglu ? col >> 1 : col, nout = geglu ? 4 : 8;
        const int rows_per = BM / splits;    // rows of the tile this CTA finishes
        const int row_base = z * rows_per;
        float bias[8];
        {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (g.bias) { b0 = *reinterpret_cast<const float4*>(g.bias + col); b1 = *reinterpret_cast<const float4*>(g.bias + col + 4); }
            bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w; bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
        }
        float c1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (g.epilogue & APX_EPI_LNFOLD) {
            const float4 a0 = *reinterpret_cast<const float4*>(g.ln_c1 + col), a1 = *reinterpret_cast<const float4*>(g.ln_c1 + col + 4);
            c1[0] = a0.x; c1[1] = a0.y; c1[2] = a0.z; c1[3] = a0.w; c1[4] = a1.x; c1[5] = a1.y; c1[6] = a1.z; c1[7] = a1.w;
            // Row statistics of the A tile, read from the pipeline stages while the tensor core consumes them (one row per thread; the
            // SWIZZLE_128B chunk order does not matter for sums).  Shifted by the row's first element against cancellation.
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full_bar[s], (i / STAGES) & 1);
                const uint4* rowp = reinterpret_cast<const uint4*>(sA + s * A_STAGE_BYTES + trow * 128);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 w = rowp[j ^ (trow & 7)];   // chunk j of the row (SWIZZLE_128B: stored at j ^ (row & 7))
                    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float lo, hi;
                        unpack16(ww[k], g.in_fmt, lo, hi);
                        if (i == 0 && j == 0 && k == 0) shift = lo;
                        lo -= shift; hi -= shift;
                        s1 += lo + hi;
                        s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                    }
                }
                asm volatile("fence.acq_rel.cta;" : "+f"(s1), "+f"(s2) : : "memory");   // the loads above have RETURNED before the stage is released (see apx_gemm16.cu)
                mbar_arrive(&empty_bar[s]);
            }
            const float inv_k = 1.0f / static_cast<float>(g.K);
            const float m = s1 * inv_k;
            sStat[trow] = make_float2(shift + m, rsqrtf(fmaxf(s2 * inv_k - m * m, 0.f) + g.ln_eps));
            // the accumulator tile is staged OVER the pipeline stages below: every statistics thread must be done reading them
            named_bar_sync(1, 128);
            if (te == 0) GEMM_TRACE(6);
        }
        Res8 cur;
        cur.a = make_uint4(0, 0, 0, 0); cur.b = cur.a;
        if (g.residual && m0 + row_base + r0 < g.M) cur = load_res8(g, m0 + row_base + r0, ocol, nout);

        mbar_wait(acc_bar, 0);
        tc_fence_after();
        if (te == 0) GEMM_TRACE(7);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float* sP = reinterpret_cast<float*>(smem);   // over the pipeline stages: every MMA that read them has completed
#pragma unroll 2
        for (int c = 0; c < BN; c += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c, r);
            tmem_ld_wait();
            const uint32_t dst = smem_u32(sP + trow * LDT + c);
#pragma unroll
            for (int i = 0; i < 16; i += 4) st_shared_v4(dst + i * 4, r[i], r[i + 1], r[i + 2], r[i + 3]);
        }
        if (splits > 1) {
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the bulk-copy engine
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' mbarriers are initialised (long since)
        }
        named_bar_sync(1, 128);
        if (te == 0) GEMM_TRACE(8);
        if (splits > 1) {
            // every peer gets the rows it finishes; slot of sender z in receiver p: z < p ? z : z - 1
            const uint32_t slab_bytes = static_cast<uint32_t>(rows_per) * LDT * 4;
            if (te == 0) mbar_expect_tx(recv_bar, (splits - 1) * slab_bytes);
            if (te < splits && te != z) {
                const int slot = z < te ? z : z - 1;
                dsmem_bulk_copy(map_to_cta(smem_u32(sRecv) + slot * slab_bytes, te), sP + te * rows_per * LDT, slab_bytes,
                                map_to_cta(smem_u32(recv_bar), te));
            }
            mbar_wait(recv_bar, 0);
            if (te < splits && te != z) mbar_arrive_remote(map_to_cta(smem_u32(done_bar), te));   // "I hold all my slabs"
            if (te == 0) GEMM_TRACE(9);
        }
#pragma unroll 1
        for (int rl = r0; rl < rows_per; rl += ROWS_STEP) {
            const int row = m0 + row_base + rl;
            Res8 nxt;
            nxt.a = make_uint4(0, 0, 0, 0); nxt.b = nxt.a;
            if (g.residual && rl + ROWS_STEP < rows_per && row + ROWS_STEP < g.M) nxt = load_res8(g, row + ROWS_STEP, ocol, nout);
            if (row < g.M) {
                // partials are summed in rank order whatever z is, so a row's value does not depend on where in a tile it sits
                float v[8];
                for (int p = 0; p < splits; ++p) {
                    const float* rp = p == z ? sP + (row_base + rl) * LDT + cg * 8
                                             : sRecv + ((p < z ? p : p - 1) * rows_per + rl) * LDT + cg * 8;
                    const float4 b0 = *reinterpret_cast<const float4*>(rp), b1 = *reinterpret_cast<const float4*>(rp + 4);
                    if (p == 0) {
                        v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
                    } else {
                        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                    }
                }
                finish8(g, row, ocol, nout, v, bias, cur, c1, (g.epilogue & APX_EPI_LNFOLD) ? sStat[row_base + rl] : make_float2(0.f, 1.f));
            }
            cur = nxt;
        }
        if (te == 0) GEMM_TRACE(10);
        if (splits > 1) mbar_wait(done_bar, 0);   // nobody is still reading this CTA's shared memory
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) GEMM_TRACE(11);
    if (warp == 1) tmem_dealloc<BN>(tmem_base);
    if (threadIdx.x == 32) GEMM_TRACE(12);
}

template <int BN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t st) {
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        APX_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmCfg<BN>::SMEM_BYTES + MAX_RECV_BYTES));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(g.M, BM), g.N / BN, g.splits);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = GemmCfg<BN>::SMEM_BYTES + (g.splits > 1 ? (g.splits - 1) * (BM / g.splits) * LDP * 4 : 0);
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 1;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = static_cast<unsigned>(g.splits);
    cfg.attrs = attrs;
    cfg.numAttrs = g.splits > 1 ? 1 : 0;
    add_pdl_attr(attrs, &cfg.numAttrs);
    APX_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN>, tmA, tmB, g));
    APX_LAUNCH_CHECK("gemm_tc_kernel");
    return APX_OK;
}

// Split-K factor: only when the 128 x 64 tiling leaves most of the 148 SMs idle, at least two k-blocks per CTA, and the grid stays
// within two CTAs per SM.  APX_GEMM_SPLITK=n forces n (1 turns it off).
int pick_splits(int M, int N, int K) {
    static const int forced = [] { const char* e = getenv("APX_GEMM_SPLITK"); return e ? atoi(e) : 0; }();
    const int tiles = ceil_div(M, BM) * (N / 64), num_kb = K / BK;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced <= num_kb ? forced : 1;
    if (tiles >= 148) return 1;
    for (int s = 8; s >= 2; s >>= 1)
        if (2 * s <= num_kb && tiles * s <= 2 * 148) return s;
    return 1;
}

}  // namespace

int gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
            int64_t ldr, void* out, int64_t ldo, int M, int N, int K, int in_dtype, int out_dtype, int epilogue,
            cudaStream_t st, const float* ln_c1, float ln_eps, void* out2, int64_t ldo2) {
    APX_REQUIRE(in_dtype == APX_F16 || in_dtype == APX_BF16, APX_EUNSUPPORTED, "gemm_tc: in_dtype %d", in_dtype);
    APX_REQUIRE(K % BK == 0 && K > 0, APX_EUNSUPPORTED, "gemm_tc: K=%d must be a multiple of %d", K, BK);
    APX_REQUIRE(N % 64 == 0 && N > 0, APX_EUNSUPPORTED, "gemm_tc: N=%d must be a multiple of 64", N);
    APX_REQUIRE(M > 0, APX_EINVAL, "gemm_tc: M=%d", M);
    APX_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, APX_EUNSUPPORTED, "gemm_tc: lda/ldw must be multiples of 8 elements");
    APX_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                APX_EINVAL, "gemm_tc: A/W/out must be 16-byte aligned");
    APX_REQUIRE(ldo % 8 == 0, APX_EUNSUPPORTED, "gemm_tc: ldo must be a multiple of 8 elements");
    APX_REQUIRE(!residual || (ldr % ((epilogue & APX_EPI_RES16) ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
                APX_EINVAL, "gemm_tc: residual rows must be 16-byte aligned");
    APX_REQUIRE(!(epilogue & APX_EPI_RES16) || out_dtype != APX_F32, APX_EINVAL, "gemm_tc: APX_EPI_RES16 needs a 16-bit out_dtype");
    APX_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, APX_EINVAL, "gemm_tc: bias alignment");

    APX_REQUIRE(!out2 || (ldo2 % 8 == 0 && (reinterpret_cast<uintptr_t>(out2) & 15) == 0), APX_EINVAL, "gemm_tc: out2 alignment");
    // Wide tiles once the grid already fills the machine, narrow ones otherwise; split-K clusters when even those leave SMs idle.
    const bool wide = (N % 128 == 0) && (static_cast<int64_t>(ceil_div(M, BM)) * (N / 128) >= 148);
    APX_REQUIRE(!(epilogue & APX_EPI_LNFOLD) || (ln_c1 && bias && (reinterpret_cast<uintptr_t>(ln_c1) & 15) == 0), APX_EINVAL,
                "gemm_tc: APX_EPI_LNFOLD needs c1 and c2 (16-byte aligned)");
    if (gemm16_supported(residual, out_dtype, in_dtype, epilogue, out2))
        return gemm16(A, lda, W, ldw, bias, residual, ldr, out, ldo, M, N, K, in_dtype, epilogue, st, ln_c1, ln_eps);
    // the row statistics of a folded LayerNorm are taken over the CTA's whole K range: no split-K
    const int splits = (wide || (epilogue & (APX_EPI_LNFOLD | APX_EPI_NOSPLIT))) ? 1 : pick_splits(M, N, K);
    const int BN = wide ? 128 : 64;

    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
        uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
        uint32_t box[2] = {BK, BM};
        int rc = encode_tmap(&tmA, in_dtype, 2, A, dims, strides, box, 1);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
        uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
        uint32_t box[2] = {BK, static_cast<uint32_t>(BN)};
        int rc = encode_tmap(&tmB, in_dtype, 2, W, dims, strides, box, 1);
        if (rc) return rc;
    }
    GemmArgs g{bias, residual, ldr, out, ldo, M, N, K, in_dtype == APX_F16 ? 0 : 1, out_dtype, epilogue, splits, out2, ldo2, ln_c1, ln_eps, nullptr};
    // diagnostic timeline (APX_GEMM_TRACE=1): synchronises and prints, never enabled on the product path
    static const bool trace_on = [] { const char* e = getenv("APX_GEMM_TRACE"); return e && e[0] == '1'; }();
    static long long* trace_buf = nullptr;
    if (trace_on) {
        if (!trace_buf) APX_CUDA(cudaMalloc(&trace_buf, 16 * sizeof(long long)));
        APX_CUDA(cudaMemsetAsync(trace_buf, 0, 16 * sizeof(long long), st));
        g.trace = trace_buf;
    }
    const int rc = wide ? launch_gemm<128>(tmA, tmB, g, st) : launch_gemm<64>(tmA, tmB, g, st);
    if (trace_on && rc == APX_OK) {
        long long t[16];
        APX_CUDA(cudaStreamSynchronize(st));
        APX_CUDA(cudaMemcpy(t, trace_buf, sizeof(t), cudaMemcpyDeviceToHost));
        auto rel = [&](int slot) { return t[slot] ? t[slot] - t[0] : -1; };
        printf("[apx gemm trace] M=%d N=%d K=%d BN=%d splits=%d epi=%d | setup %lld | tma first %lld all %lld | mma first_full %lld committed %lld | "
               "epi stats %lld acc_seen %lld staged %lld exchanged %lld stored %lld | sync %lld dealloc %lld (cycles, CTA 0)\n",
               M, N, K, BN, splits, epilogue, rel(1), rel(2), rel(3), rel(4), rel(5), rel(6), rel(7), rel(8), rel(9), rel(10), rel(11), rel(12));
        fflush(stdout);
    }
    return rc;
}

}  // namespace apx
