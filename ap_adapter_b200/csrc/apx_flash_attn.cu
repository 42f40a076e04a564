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

namespace apx {
namespace {

constexpr int NT = 192;
constexpr int BKV = 128;
constexpr int MAX_ST = 3;

struct FaArgs {
    uint16_t* O;
    int64_t o_bs, o_rs;      // element strides of the output view
    int S, L, H, d;          // queries, keys, heads, head dim (d % 16 == 0, d <= 80)
    int nblk, stages;
    float scale_log2;        // log2(e) / sqrt(d)
    uint32_t tile_bytes;     // 128 * d * 2: a Q tile, a K block, a V block
    uint32_t off_k, off_v, off_p, off_bar;
};

enum { B_Q = 0, B_KV_FULL = 1, B_KV_FREE = 4, B_S_FULL = 7, B_S_FREE = 9, B_P_FULL = 11, B_PV_FULL = 13, B_PV_FREE = 15, B_COUNT = 17 };

__device__ __forceinline__ uint32_t idesc_major(int fmt, int M, int N, int a_mn, int b_mn) {
    return umma_idesc_f16(fmt, M, N) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16);
}

template <int FMT>
__global__ void __launch_bounds__(NT, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                  FaArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* sQ = smem;                      // [d/8][128][8]
    uint8_t* sK = smem + a.off_k;            // stages x [d/8][128 keys][8]
    uint8_t* sV = smem + a.off_v;            // stages x [d/8][128 keys][8]   (read MN-major as B[N = dims][K = keys])
    uint8_t* sP = smem + a.off_p;            // 2 x [128 keys / 8][128 rows][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

    if (threadIdx.x == 160) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        for (int i = 0; i < B_COUNT; ++i) {
            const bool from_threads = (i >= B_S_FREE && i < B_PV_FULL) || i >= B_PV_FREE;   // S_FREE, P_FULL, PV_FREE: the 128 softmax threads
            mbar_init(&bars[i], from_threads ? 128u : 1u);
        }
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    pdl_launch_dependents();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tensor memory: score buffers at columns 0 and 128, P V accumulators at 256 and 384
    const uint32_t tS = tmem_base, tPV = tmem_base + 256;

    if (warp == 5) {
        if (elect_one()) {
            mbar_expect_tx(&bars[B_Q], a.tile_bytes);
            tma_load_4d(sQ, &tmQ, 0, m0, h * (a.d / 8), b, &bars[B_Q]);
            for (int j = 0; j < a.nblk; ++j) {
                const int st = j % a.stages;
                if (j >= a.stages) mbar_wait(&bars[B_KV_FREE + st], ((j / a.stages) - 1) & 1);
                mbar_expect_tx(&bars[B_KV_FULL + st], 2 * a.tile_bytes);
                tma_load_4d(sK + st * a.tile_bytes, &tmK, 0, j * BKV, h * (a.d / 8), b, &bars[B_KV_FULL + st]);
                tma_load_4d(sV + st * a.tile_bytes, &tmV, 0, j * BKV, h * (a.d / 8), b, &bars[B_KV_FULL + st]);
            }
        }
    } else if (warp == 4) {
        if (elect_one()) {
            const uint32_t id_s = idesc_major(FMT, 128, BKV, 0, 0);
            const uint32_t id_pv = idesc_major(FMT, 128, a.d, 0, 1);
            const uint64_t dQ = umma_desc(smem_u32(sQ), 2048, 128, kLayoutNone);   // K-major: k-group stride 128 rows x 16 B, 8-row stride 128 B
            // P_i V_i: A = the P tile (K-major over the keys); B[N = dims][K = keys] = the V block read MN-major (8 dims contiguous; the
            // next 8 dims 2048 B on; the next 8 keys 128 B on); 8 k-steps of 16 keys
            auto issue_pv = [&](int i) {
                const int pb = i & 1, st = i % a.stages;
                mbar_wait_spin(&bars[B_P_FULL + pb], (i >> 1) & 1);
                if (i >= 2) mbar_wait_spin(&bars[B_PV_FREE + pb], ((i >> 1) - 1) & 1);
                tc_fence_after();
                mma_chain(tPV + pb * 128, umma_desc(smem_u32(sP + pb * 32768), 2048, 128, kLayoutNone),
                          umma_desc(smem_u32(sV + st * a.tile_bytes), 128, 2048, kLayoutNone), 2 * 2048, 2 * 128, BKV / 16, id_pv, 0);
                umma_commit(&bars[B_PV_FULL + pb]);
                umma_commit(&bars[B_KV_FREE + st]);
            };
            mbar_wait_spin(&bars[B_Q], 0);
            for (int j = 0; j < a.nblk; ++j) {
                const int sb = j & 1, st = j % a.stages;
                mbar_wait_spin(&bars[B_KV_FULL + st], (j / a.stages) & 1);
                if (j >= 2) mbar_wait_spin(&bars[B_S_FREE + sb], ((j >> 1) - 1) & 1);
                tc_fence_after();
                mma_chain(tS + sb * 128, dQ, umma_desc(smem_u32(sK + st * a.tile_bytes), 2048, 128, kLayoutNone), 2 * 2048, 2 * 2048,
                          a.d / 16, id_s, 0);
                umma_commit(&bars[B_S_FULL + sb]);
                if (j >= 1) issue_pv(j - 1);
            }
            issue_pv(a.nblk - 1);
        }
    } else {
        const int row = warp * 32 + lane;
        const uint32_t lane_taddr = static_cast<uint32_t>(warp * 32) << 16;
        const float sl = a.scale_log2;
        const int ng = a.d / 16;            // 16-column groups of an output row
        float m = -INFINITY, l = 0.f;
        float o[80];
#pragma unroll
        for (int i = 0; i < 80; ++i) o[i] = 0.f;

        for (int j = 0; j < a.nblk; ++j) {
            const int sb = j & 1;
            const int valid = min(BKV, a.L - j * BKV);     // keys of this block that exist
            const uint32_t ts = tS + sb * 128 + lane_taddr;
            mbar_wait_nocall(&bars[B_S_FULL + sb], (j >> 1) & 1);
            tc_fence_after();
            // pass 1: the row maximum of the block
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
            const float m_new = fmaxf(m, mx * sl);
            const float alpha = ex2_approx(m - m_new);   // 0 for the first block (m = -inf)
            // pass 2: P = exp2(s * sl - m_new) -> 16-bit tile (K-major core matrices over the keys), row sum in fp32
            float sum = 0.f;
            const uint32_t p_row = smem_u32(sP + sb * 32768) + row * 16;
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
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        e[i] = ex2_approx(fmaf(__uint_as_float(r[c * 16 + i]), sl, -m_new));
                        sum += e[i];
                    }
                    store16<FMT>(p_row + (hf * 8 + c * 2) * 2048, e);
                }
            }
            tc_fence_before();
            mbar_arrive(&bars[B_S_FREE + sb]);            // the score buffer may take block j + 2
            fence_proxy_async_smem();                     // P tile -> visible to the tensor core
            mbar_arrive(&bars[B_P_FULL + sb]);
            l = l * alpha + sum;
            // fold in the previous block's P V (it ran while this block's softmax did), then move everything to the new maximum
            if (j >= 1) {
                const int pb = (j - 1) & 1;
                mbar_wait_nocall(&bars[B_PV_FULL + pb], ((j - 1) >> 1) & 1);
                tc_fence_after();
                const uint32_t tp = tPV + pb * 128 + lane_taddr;
                uint32_t r[80];
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    if (c < ng) tmem_ld16(tp + c * 16, r + c * 16);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    if (c < ng) {
                        reg_tie16(r + c * 16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[c * 16 + i] = (o[c * 16 + i] + __uint_as_float(r[c * 16 + i])) * alpha;
                    }
                tc_fence_before();
                mbar_arrive(&bars[B_PV_FREE + pb]);
            }
            m = m_new;
        }
        {   // the last block's P V, normalisation, store
            const int pb = (a.nblk - 1) & 1;
            mbar_wait_nocall(&bars[B_PV_FULL + pb], ((a.nblk - 1) >> 1) & 1);
            tc_fence_after();
            const uint32_t tp = tPV + pb * 128 + lane_taddr;
            uint32_t r[80];
#pragma unroll
            for (int c = 0; c < 5; ++c)
                if (c < ng) tmem_ld16(tp + c * 16, r + c * 16);
            tmem_ld_wait();
            const float inv = 1.0f / l;
            const int q = m0 + row;
            uint16_t* op = a.O + static_cast<int64_t>(b) * a.o_bs + static_cast<int64_t>(q) * a.o_rs + h * a.d;
#pragma unroll
            for (int c = 0; c < 5; ++c)
                if (c < ng) {
                    reg_tie16(r + c * 16);
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        w[i] = pack2<FMT>((o[c * 16 + 2 * i] + __uint_as_float(r[c * 16 + 2 * i])) * inv,
                                          (o[c * 16 + 2 * i + 1] + __uint_as_float(r[c * 16 + 2 * i + 1])) * inv);
                    if (q < a.S) {
                        *reinterpret_cast<uint4*>(op + c * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(op + c * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem_base);
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
    a.stages = std::min(a.nblk, MAX_ST);
    a.off_k = a.tile_bytes;
    a.off_v = a.off_k + a.stages * a.tile_bytes;
    a.off_p = a.off_v + a.stages * a.tile_bytes;
    a.off_bar = a.off_p + 2 * 32768;
    const size_t smem = 128 + a.off_bar + B_COUNT * 8 + 16;
    APX_REQUIRE(smem <= 227 * 1024, APX_EUNSUPPORTED, "flash_attn: %zu B of shared memory", smem);

    CUtensorMap tmQ, tmK, tmV;
    auto make = [&](CUtensorMap* tm, const void* p, int64_t bs, int64_t rs, int rows) {
        // a [B][C/8][rows][8] view: the box {8, 128, d/8, 1} lands as [d/8][128][8] = UMMA core-matrix order
        uint64_t dims[4] = {8, static_cast<uint64_t>(rows), static_cast<uint64_t>(C / 8), static_cast<uint64_t>(B)};
        uint64_t strides[3] = {static_cast<uint64_t>(rs) * 2, 16, static_cast<uint64_t>(bs) * 2};
        uint32_t box[4] = {8, 128, static_cast<uint32_t>(d / 8), 1};
        return encode_tmap(tm, dtype, 4, p, dims, strides, box, 0);
    };
    int rc = make(&tmQ, Q, q_bs, q_rs, S);
    if (rc) return rc;
    rc = make(&tmK, K, k_bs, k_rs, L);
    if (rc) return rc;
    rc = make(&tmV, V, v_bs, v_rs, L);
    if (rc) return rc;

    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        APX_CUDA(cudaFuncSetAttribute(flash_attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        APX_CUDA(cudaFuncSetAttribute(flash_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
    if (dtype == APX_F16) APX_CUDA(cudaLaunchKernelEx(&cfg, flash_attn_kernel<0>, tmQ, tmK, tmV, a));
    else                  APX_CUDA(cudaLaunchKernelEx(&cfg, flash_attn_kernel<1>, tmQ, tmK, tmV, a));
    APX_LAUNCH_CHECK("flash_attn_kernel");
    return APX_OK;
}

}  // namespace apx
