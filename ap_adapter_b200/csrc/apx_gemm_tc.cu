// apx_gemm_tc.cu — Y = act(A · W^T + bias) + residual on the 5th-gen tensor cores.
//
// One CTA owns a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: 128x64 A box + BNx64 W box per k-block, SWIZZLE_128B, 3-stage mbarrier ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accum in TMEM)
//   warps 2..5  epilogue: tcgen05.ld (lane quarter = warp_id % 4) -> bias / erf-GELU / fp32 residual -> store
//
// Replaces the nn.Linear dispatches of the reference hot path
// (APadapter/ap_adapter/attention_processor.py:387,406-407,435-436,457; timm Block linears used by
// audio_encoder/models_mae.py:220-231).
#include "apx_internal.h"
#include "apx_ptx.cuh"

namespace apx {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct GemmCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;
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
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int num_kb = g.K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {   // elect.sync: the TMA / MMA operands stay in uniform registers (no per-instruction R2UR waterfall)
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], A_STAGE_BYTES + GemmCfg<BN>::B_STAGE_BYTES);
                tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, kb * BK, m0, &full_bar[s]);
                tma_load_2d(sB + s * GemmCfg<BN>::B_STAGE_BYTES, &tmB, kb * BK, n0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(g.in_fmt, BM, BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(sB + s * GemmCfg<BN>::B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // SW128 K-major: 8-row groups are 1024 B apart; a 16-element k-step is +32 B.
                    const uint64_t da = umma_desc(a_addr + k * 32, 16, 1024, kLayoutSW128);
                    const uint64_t db = umma_desc(b_addr + k * 32, 16, 1024, kLayoutSW128);
                    umma_ss(tmem_base, da, db, idesc, (kb | k) != 0);
                }
                umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
            }
            umma_commit(acc_bar);  // accumulator complete
        }
    } else {
        // ---- epilogue warps: TMEM lane quarter (warp % 4), one output row per thread
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const bool row_ok = row < g.M;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c, r);
            tmem_ld_wait();
            if (row_ok) {
                float v[16];
                const int col = n0 + c;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                if (g.bias) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(g.bias + col + i);
                        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                    }
                }
                if (g.epilogue == APX_EPI_GELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
                }
                if (g.residual) {
                    const float* rp = g.residual + static_cast<int64_t>(row) * g.ldr + col;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(rp + i);
                        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                    }
                }
                if (g.out_dtype == APX_F32) {
                    float* op = reinterpret_cast<float*>(g.out) + static_cast<int64_t>(row) * g.ldo + col;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
                    uint16_t* op = reinterpret_cast<uint16_t*>(g.out) + static_cast<int64_t>(row) * g.ldo + col;
                    uint32_t p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) p[i] = pack16(v[2 * i], v[2 * i + 1], g.out_dtype);
                    *reinterpret_cast<uint4*>(op) = make_uint4(p[0], p[1], p[2], p[3]);
                    *reinterpret_cast<uint4*>(op + 8) = make_uint4(p[4], p[5], p[6], p[7]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

template <int BN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        APX_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmCfg<BN>::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(ceil_div(g.M, BM), g.N / BN);
    gemm_tc_kernel<BN><<<grid, 192, GemmCfg<BN>::SMEM_BYTES, st>>>(tmA, tmB, g);
    APX_LAUNCH_CHECK("gemm_tc_kernel");
    return APX_OK;
}

}  // namespace

int gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
            int64_t ldr, void* out, int64_t ldo, int M, int N, int K, int in_dtype, int out_dtype, int epilogue,
            cudaStream_t st) {
    APX_REQUIRE(in_dtype == APX_F16 || in_dtype == APX_BF16, APX_EUNSUPPORTED, "gemm_tc: in_dtype %d", in_dtype);
    APX_REQUIRE(K % BK == 0 && K > 0, APX_EUNSUPPORTED, "gemm_tc: K=%d must be a multiple of %d", K, BK);
    APX_REQUIRE(N % 64 == 0 && N > 0, APX_EUNSUPPORTED, "gemm_tc: N=%d must be a multiple of 64", N);
    APX_REQUIRE(M > 0, APX_EINVAL, "gemm_tc: M=%d", M);
    APX_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, APX_EUNSUPPORTED, "gemm_tc: lda/ldw must be multiples of 8 elements");
    APX_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                APX_EINVAL, "gemm_tc: A/W/out must be 16-byte aligned");
    APX_REQUIRE(ldo % 8 == 0, APX_EUNSUPPORTED, "gemm_tc: ldo must be a multiple of 8 elements");
    APX_REQUIRE(!residual || (ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0), APX_EINVAL,
                "gemm_tc: residual must be 16-byte aligned with ldr %% 4 == 0");
    APX_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, APX_EINVAL, "gemm_tc: bias alignment");

    // Wide tiles once the grid already fills the machine, narrow ones otherwise.
    const bool wide = (N % 128 == 0) && (static_cast<int64_t>(ceil_div(M, BM)) * (N / 128) >= 148);
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
    GemmArgs g{bias, residual, ldr, out, ldo, M, N, K, in_dtype == APX_F16 ? 0 : 1, out_dtype, epilogue};
    return wide ? launch_gemm<128>(tmA, tmB, g, st) : launch_gemm<64>(tmA, tmB, g, st);
}

}  // namespace apx
