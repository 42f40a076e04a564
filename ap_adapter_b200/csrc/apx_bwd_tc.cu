// apx_bwd_tc.cu — the attention part of the adapter-site backward on the tensor cores (tcgen05), for contexts of up to 16 text + 128
// audio tokens (time/freq pooling 2, 4, 8: three of the four pooling rates train_apadapter_v2.py:445-446 draws).
//
// Given Q = X Wq^T and dO = dY Wo (both [Bc, S, C], from the tcgen05 GEMM) and the forward's packed context tiles, one CTA per
// (128-query tile, head, batch element) computes (attention_processor.py:429-454 differentiated; tau = 1/sqrt(d), ap = ap_scale):
//
//   S  = Q K^T                 tcgen05.mma, K tile K-major            P_seg = softmax_seg(tau S)   (recomputed, two segments)
//   dP = dO V^T                tcgen05.mma, the V^T tile read MN-major (no second copy of V)
//   dS = P o (w dP - delta),   delta_seg = sum_j P w dP,  w = 1 (text) | ap (audio)                 (registers, one query row per thread)
//   dQ   = tau dS K            A = dS tile (K-major), B = the K tile read MN-major
//   dK_a = tau dS_a^T Q        A = the dS tile read MN-major (transposed for free), B = the Q tile read MN-major;  M = 128 audio keys
//   dV_a = ap  P_a^T dO        A = the P tile read MN-major, B = the dO tile read MN-major
//
// dK_a / dV_a leave the CTA as fp32 atomics into [Bc, La, C] (summed over the query tiles), dQ as 16-bit rows.  The five contractions
// use four shared-memory tiles, each written once: "MN-major" descriptors (instruction-descriptor bits 15 / 16) let the tensor core read
// a [k-group][row][8] tile with rows and k swapped, so no transposed copy of P, dS, Q, dO, K or V is ever built.
// The CUDA-core kernel (apx_bwd.cu) remains the path for every other shape (the un-pooled 512-token context, d = 40).
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include "apx_site_common.cuh"
#include <cstdlib>
#include <cstring>

namespace apx {
namespace {

constexpr int NTHREADS = 192;   // warps 0-3: one query row (= TMEM lane) per thread; warp 4: TMEM allocator + MMA issuer; warp 5: TMA
constexpr int MAX16 = 9;        // 16-column groups of a score row held in registers (16 text-padded + 128 audio-padded columns)

struct BwdTcArgs {
    const uint8_t* tiles;   // packed context tiles [Bc][H]{K tile, V^T tile, extended V^T tile}
    uint16_t* dQ;           // [Bc, S, C]
    float* dKa;             // [Bc, La, C] fp32, zero-initialised
    float* dVa;
    int S, C, H, d, Lt, La, Lt_pad, La_pad, L_pad;
    float tau, ap_scale, scale_log2;
    uint32_t kv_bytes, k_tile_bytes, v_tile_bytes;
    uint32_t off_do, off_k, off_v, off_p, off_ds, off_bar;   // shared-memory offsets (the Q tile is at 0)
};

enum { B_LOAD = 0, B_SP, B_DS, B_G, B_COUNT };

// kind::f16 instruction descriptor with explicit operand majors (bit 15: A, bit 16: B; 0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t idesc_major(int fmt, int M, int N, int a_mn, int b_mn) {
    return umma_idesc_f16(fmt, M, N) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16);
}

// Descriptor of a no-swizzle tile read MN-major: `kg` = byte stride between groups of 8 along K, `mg` = between groups of 8 along M / N.
// As for K-major tiles, the "leading" offset is the K-group stride and the "stride" offset the M/N-group stride (checked on the B200:
// with the two exchanged every gradient is wrong by O(1), tests/test_gpu_ipattn.py::test_adapter_backward_matches_autograd_of_the_oracle).
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t kg, uint32_t mg) { return umma_desc(addr, kg, mg, kLayoutNone); }

template <int FMT>
__global__ void __launch_bounds__(NTHREADS, 1)
ipattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO, BwdTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* sQ = smem;                  // [d/8][128][8]   query tile   (rows = queries, k = head dims)
    uint8_t* sDO = smem + a.off_do;      // [d/8][128][8]   dO tile
    uint8_t* sK = smem + a.off_k;        // [d/8][L_pad][8] K tile       (rows = keys, k = head dims)
    uint8_t* sV = smem + a.off_v;        // [L_pad/8][d][8] V^T tile     (rows = head dims, k = keys)
    uint8_t* sP = smem + a.off_p;        // [(Lt_pad+128)/8][128][8]  w * P   (rows = queries, k = keys)
    uint8_t* sDS = smem + a.off_ds;      // same shape: tau * dS
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const uint32_t q_bytes = 128u * a.d * 2;

    if (threadIdx.x == 160) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmDO);
        mbar_init(&bars[B_LOAD], 1);
        mbar_init(&bars[B_SP], 1);
        mbar_init(&bars[B_DS], 128);
        mbar_init(&bars[B_G], 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tensor memory: S at [0, L_pad), dP at [L_pad, 2 L_pad); the three gradient accumulators reuse the same columns afterwards
    const uint32_t tS = tmem_base, tDP = tmem_base + a.L_pad;
    const uint32_t tDQ = tmem_base, tDK = tmem_base + a.d, tDV = tmem_base + 2 * a.d;

    if (warp == 5) {
        if (elect_one()) {
            const uint8_t* tile = a.tiles + (static_cast<size_t>(b) * a.H + h) * a.kv_bytes;
            mbar_expect_tx(&bars[B_LOAD], 2 * q_bytes + a.k_tile_bytes + a.v_tile_bytes);
            tma_load_4d(sQ, &tmQ, 0, m0, h * (a.d / 8), b, &bars[B_LOAD]);     // rows past S zero-filled
            tma_load_4d(sDO, &tmDO, 0, m0, h * (a.d / 8), b, &bars[B_LOAD]);
            tma_load_1d(sK, tile, a.k_tile_bytes, &bars[B_LOAD]);
            tma_load_1d(sV, tile + a.k_tile_bytes, a.v_tile_bytes, &bars[B_LOAD]);
        }
    } else if (warp == 4) {
        if (elect_one()) {
            mbar_wait_spin(&bars[B_LOAD], 0);
            tc_fence_after();
            const uint64_t dQt = umma_desc(smem_u32(sQ), 2048, 128, kLayoutNone);     // K-major A: k-group stride 2048, 8-row stride 128
            const uint64_t dDOt = umma_desc(smem_u32(sDO), 2048, 128, kLayoutNone);
            // S = Q K^T: B = K tile, K-major (k-group stride L_pad*16, 8-row stride 128)
            mma_chain(tS, dQt, umma_desc(smem_u32(sK), a.L_pad * 16, 128, kLayoutNone), 2 * 2048, 2 * a.L_pad * 16, a.d / 16,
                      idesc_major(FMT, 128, a.L_pad, 0, 0), 0);
            // dP = dO V^T: B[N = keys][K = dims] = the V^T tile read MN-major: 8 keys are contiguous (16 B), the next 8 keys (N group) are
            // d*16 bytes on, the next 8 dims (K group) 128 bytes on.  SBO = N-group stride, LBO = K-group stride; a k-step is 2 K groups.
            mma_chain(tDP, dDOt, desc_mn(smem_u32(sV), 128, a.d * 16), 2 * 2048, 2 * 128, a.d / 16,
                      idesc_major(FMT, 128, a.L_pad, 0, 1), 0);
            umma_commit(&bars[B_SP]);
            mbar_wait_spin(&bars[B_DS], 0);   // w P and tau dS are in shared memory; S / dP have been read
            tc_fence_after();
            // dQ[128 q, d] = (tau dS) K: A = dS tile K-major over the keys; B[N = dims][K = keys] = the K tile read MN-major
            // (8 dims contiguous; next 8 dims L_pad*16 bytes on; next 8 keys 128 bytes on)
            mma_chain(tDQ, umma_desc(smem_u32(sDS), 2048, 128, kLayoutNone), desc_mn(smem_u32(sK), 128, a.L_pad * 16), 2 * 2048,
                      2 * 128, a.L_pad / 16, idesc_major(FMT, 128, a.d, 0, 1), 0);
            // dK_a[128 keys, d] = (tau dS_a)^T Q and dV_a = (ap P_a)^T dO: A[M = keys][K = queries] = the dS / P tile read MN-major from its
            // first audio column (8 keys contiguous; next 8 keys 2048 bytes on; next 8 queries 128 bytes on); B[N = dims][K = queries] = the
            // Q / dO tile read MN-major (8 dims contiguous; next 8 dims 2048 bytes on; next 8 queries 128 bytes on); 8 k-steps of 16 queries
            const uint32_t a_off = (a.Lt_pad / 8) * 2048;
            const uint32_t idg = idesc_major(FMT, 128, a.d, 1, 1);
            mma_chain(tDK, desc_mn(smem_u32(sDS) + a_off, 128, 2048), desc_mn(smem_u32(sQ), 128, 2048), 2 * 128,
                      2 * 128, 8, idg, 0);
            mma_chain(tDV, desc_mn(smem_u32(sP) + a_off, 128, 2048), desc_mn(smem_u32(sDO), 128, 2048), 2 * 128,
                      2 * 128, 8, idg, 0);
            umma_commit(&bars[B_G]);
        }
    } else {
        const int row = warp * 32 + lane;
        const uint32_t lane_taddr = static_cast<uint32_t>(warp * 32) << 16;
        const int n16 = a.L_pad / 16, lt16 = a.Lt_pad / 16;
        const int la_end = a.Lt_pad + a.La;
        const float sl = a.scale_log2;
        mbar_wait_nocall(&bars[B_SP], 0);
        tc_fence_after();
        // ---- P = softmax of both segments, kept in registers (fp32)
        uint32_t s[MAX16 * 16];
#pragma unroll
        for (int c = 0; c < MAX16; ++c)
            if (c < n16) tmem_ld16(tS + lane_taddr + c * 16, s + c * 16);
        tmem_ld_wait();
        float mx_t = -INFINITY, mx_a = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAX16; ++c)
            if (c < n16) {
                reg_tie16(s + c * 16);
                const bool text = c < lt16;
                const int valid = (text ? a.Lt : la_end) - c * 16;
                if (valid < 16) mask16(s + c * 16, valid);
                const float m = max16(s + c * 16);
                if (text) mx_t = fmaxf(mx_t, m); else mx_a = fmaxf(mx_a, m);
            }
        float l_t = 0.f, l_a = 0.f;
#pragma unroll
        for (int c = 0; c < MAX16; ++c)
            if (c < n16) {
                const bool text = c < lt16;
                const float mxs = (text ? mx_t : mx_a) * sl;
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float e = ex2_approx(fmaf(__uint_as_float(s[c * 16 + i]), sl, -mxs));
                    s[c * 16 + i] = __float_as_uint(e);
                    acc += e;
                }
                if (text) l_t += acc; else l_a += acc;
            }
        const float il_t = a.Lt > 0 ? 1.0f / l_t : 0.f, il_a = 1.0f / l_a;
        // ---- delta_seg = sum_j P_j w dP_j (first pass over dP in tensor memory)
        float dl_t = 0.f, dl_a = 0.f;
#pragma unroll
        for (int c = 0; c < MAX16; ++c)
            if (c < n16) {
                uint32_t r[16];
                tmem_ld16(tDP + lane_taddr + c * 16, r);
                tmem_ld_wait16(r);
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) acc = fmaf(__uint_as_float(s[c * 16 + i]), __uint_as_float(r[i]), acc);
                if (c < lt16) dl_t += acc; else dl_a += acc;
            }
        dl_t *= il_t;                    // sum_j p_j dp_j            (text: w = 1)
        dl_a *= il_a * a.ap_scale;       // sum_j p_j ap dp_j         (audio)
        // ---- second pass: tau dS = tau p (w dp - delta) and w p, both 16 bit, into the K-major tiles
        const uint32_t sP_row = smem_u32(sP) + row * 16, sDS_row = smem_u32(sDS) + row * 16;
#pragma unroll
        for (int c = 0; c < MAX16; ++c)
            if (c < n16) {
                uint32_t r[16];
                tmem_ld16(tDP + lane_taddr + c * 16, r);
                tmem_ld_wait16(r);
                const bool text = c < lt16;
                const float il = text ? il_t : il_a, w = text ? 1.0f : a.ap_scale, dl = text ? dl_t : dl_a;
                float pv[16], dsv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p = __uint_as_float(s[c * 16 + i]) * il;
                    pv[i] = w * p;
                    dsv[i] = a.tau * p * (w * __uint_as_float(r[i]) - dl);
                }
                store16<FMT>(sP_row + c * 2 * 2048, pv);
                store16<FMT>(sDS_row + c * 2 * 2048, dsv);
            }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[B_DS]);
        // ---- gradients out: dQ row (16 bit) and, for the lanes that are audio keys, dK_a / dV_a rows (fp32 atomics)
        mbar_wait_nocall(&bars[B_G], 0);
        tc_fence_after();
        const int srow = m0 + row;
        uint16_t* dq = a.dQ + (static_cast<size_t>(b) * a.S + srow) * a.C + h * a.d;
        float* dk = a.dKa + (static_cast<size_t>(b) * a.La + row) * a.C + h * a.d;
        float* dv = a.dVa + (static_cast<size_t>(b) * a.La + row) * a.C + h * a.d;
#pragma unroll 1
        for (int c = 0; c < a.d; c += 16) {
            uint32_t rq[16], rk[16], rv[16];
            tmem_ld16(tDQ + lane_taddr + c, rq);
            tmem_ld16(tDK + lane_taddr + c, rk);
            tmem_ld16(tDV + lane_taddr + c, rv);
            tmem_ld_wait();
            reg_tie16(rq); reg_tie16(rk); reg_tie16(rv);
            if (srow < a.S) {
                uint32_t p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = pack2<FMT>(__uint_as_float(rq[2 * i]), __uint_as_float(rq[2 * i + 1]));
                *reinterpret_cast<uint4*>(dq + c) = make_uint4(p[0], p[1], p[2], p[3]);
                *reinterpret_cast<uint4*>(dq + c + 8) = make_uint4(p[4], p[5], p[6], p[7]);
            }
            if (row < a.La) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    atomicAdd(dk + c + i, __uint_as_float(rk[i]));
                    atomicAdd(dv + c + i, __uint_as_float(rv[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem_base);
}

}  // namespace

bool ipattn_bwd_tc_supported(int C, int H, int Lt, int La) {
    static const bool on = [] { const char* e = getenv("APX_BWD_TC"); return !(e && e[0] == '0'); }();
    if (!on || C <= 0 || H <= 0 || C % H) return false;
    const int d = C / H;
    const KvPackGeom g = kv_pack_geom(Lt, La, C, H);
    return d % 16 == 0 && d <= 80 && La > 0 && g.La_pad <= 128 && g.Lt_pad <= 16 && g.L_pad <= MAX16 * 16 && C % 8 == 0;
}

int ipattn_bwd_tc(const void* Q, const void* dO, const void* cache_tiles, void* dQ, float* dKa, float* dVa, int Bc, int S, int C, int H, int Lt,
                  int La, float ap_scale, int dtype, cudaStream_t st) {
    APX_REQUIRE(ipattn_bwd_tc_supported(C, H, Lt, La), APX_EUNSUPPORTED, "ipattn_bwd_tc: C=%d H=%d context %d+%d is not covered", C, H, Lt, La);
    const KvPackGeom g = kv_pack_geom(Lt, La, C, H);
    CUtensorMap tmQ, tmDO;
    {
        // [Bc][C/8][S][8] view: the box {8, 128, d/8, 1} lands as [d/8][128][8] = K-major core-matrix order
        uint64_t dims[4] = {8, static_cast<uint64_t>(S), static_cast<uint64_t>(C / 8), static_cast<uint64_t>(Bc)};
        uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, 16, static_cast<uint64_t>(S) * C * 2};
        uint32_t box[4] = {8, 128, static_cast<uint32_t>(g.d / 8), 1};
        int rc = encode_tmap(&tmQ, dtype, 4, Q, dims, strides, box, 0);
        if (rc) return rc;
        rc = encode_tmap(&tmDO, dtype, 4, dO, dims, strides, box, 0);
        if (rc) return rc;
    }
    BwdTcArgs a;
    memset(&a, 0, sizeof(a));
    a.tiles = static_cast<const uint8_t*>(cache_tiles);
    a.dQ = static_cast<uint16_t*>(dQ);
    a.dKa = dKa; a.dVa = dVa;
    a.S = S; a.C = C; a.H = H; a.d = g.d; a.Lt = Lt; a.La = La; a.Lt_pad = g.Lt_pad; a.La_pad = g.La_pad; a.L_pad = g.L_pad;
    a.tau = 1.0f / sqrtf(static_cast<float>(g.d));
    a.ap_scale = ap_scale;
    a.scale_log2 = 1.4426950408889634f * a.tau;
    a.kv_bytes = static_cast<uint32_t>(g.bytes_per_bh);
    a.k_tile_bytes = static_cast<uint32_t>(g.k_tile_bytes);
    a.v_tile_bytes = static_cast<uint32_t>(g.v_tile_bytes);
    const uint32_t q_bytes = 128u * g.d * 2, pd_bytes = static_cast<uint32_t>((g.Lt_pad + 128) / 8) * 2048;
    uint32_t off = q_bytes;
    a.off_do = off; off += q_bytes;
    a.off_k = off; off += static_cast<uint32_t>(align_up(g.k_tile_bytes, 128));
    a.off_v = off; off += static_cast<uint32_t>(align_up(g.v_tile_bytes, 128));
    a.off_p = off; off += pd_bytes;
    a.off_ds = off; off += pd_bytes;
    a.off_bar = off;
    const uint32_t smem_bytes = off + 64 + 128;
    APX_REQUIRE(smem_bytes <= 227 * 1024, APX_EUNSUPPORTED, "ipattn_bwd_tc: %u B of shared memory", smem_bytes);
    typedef void (*Fn)(const CUtensorMap, const CUtensorMap, BwdTcArgs);
    static const Fn table[2] = {ipattn_bwd_tc_kernel<0>, ipattn_bwd_tc_kernel<1>};
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask))
        for (int i = 0; i < 2; ++i) APX_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(table[i]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    table[dtype == APX_F16 ? 0 : 1]<<<dim3(ceil_div(S, 128), H, Bc), NTHREADS, smem_bytes, st>>>(tmQ, tmDO, a);
    APX_LAUNCH_CHECK("ipattn_bwd_tc_kernel");
    return APX_OK;
}

}  // namespace apx
