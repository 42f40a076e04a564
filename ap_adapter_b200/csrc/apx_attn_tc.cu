// apx_attn_tc.cu — the decoupled (text + ap_scale * audio) cross-attention core on tcgen05.
//
//   O[b, s, h, :] = softmax(q K_t^T / sqrt(d)) V_t + ap_scale * softmax(q K_a^T / sqrt(d)) V_a
//
// One CTA = one (128-query tile, head, batch element).  Both contexts' K and V^T tiles of the head are
// staged ONCE by TMA bulk copies from the packed context cache (built by kv_pack, below), the Q tile by a
// 4-D tensor-map TMA that lands directly in the UMMA no-swizzle K-major core-matrix order.
//   S  = Q K^T      tcgen05.mma  128 x N x 16   (N <= 128 per chunk), fp32 in TMEM
//   P  = exp2(...)  one query row per thread: tcgen05.ld -> registers -> 16-bit P tile in shared memory
//   O_seg += P V    tcgen05.mma  128 x d_pad x 16, one TMEM accumulator per segment
// and the ap_scale-weighted add of the two normalised accumulators happens in registers in the epilogue —
// the two SDPA calls + scale-add of APadapter/ap_adapter/attention_processor.py:429-454 in one kernel.
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace apx {

KvPackGeom kv_pack_geom(int Lt, int La, int C, int H) {
    KvPackGeom g;
    g.Lt = Lt; g.La = La; g.H = H; g.C = C;
    g.Lt_pad = (Lt + 15) / 16 * 16;
    g.La_pad = (La + 15) / 16 * 16;
    g.L_pad = g.Lt_pad + g.La_pad;
    g.d = C / H;
    g.d_pad = (g.d + 15) / 16 * 16;
    g.dv_pad = g.d_pad + 16;
    g.k_tile_bytes = static_cast<size_t>(g.d_pad) * g.L_pad * 2;
    g.v_tile_bytes = g.k_tile_bytes;
    g.vx_tile_bytes = static_cast<size_t>(g.dv_pad) * g.L_pad * 2;
    g.bytes_per_bh = g.k_tile_bytes + g.v_tile_bytes + g.vx_tile_bytes;
    return g;
}

namespace {

constexpr int CHUNK = 128;  // KV columns per S tile

struct AttnArgs {
    const uint8_t* tiles;  // packed K / V^T tiles, [Bc][H]{K tile, V^T tile}
    void* O;               // [Bc, S, H*d] view: element strides o_bs (batch), o_rs (row)
    int64_t o_bs, o_rs;
    int64_t q_bs, q_rs;    // element strides of the Q view (only used by the manual staging path)
    int Bc, S, C, H;
    int Lt, La, Lt_pad, La_pad, L_pad, d, d_pad;
    float ap_scale;
    float scale_log2;      // log2(e) / sqrt(d)
    int fmt;               // 0 f16, 1 bf16
    uint32_t k_tile_bytes, v_tile_bytes, bytes_per_bh;
    int dv_pad;
    const uint16_t* Q;     // only read when q_manual != 0
    int q_manual;          // debug hedge: stage the Q tile with ld.global/st.shared instead of the 4-D TMA
};

template <int TMEM_COLS>
__global__ void __launch_bounds__(128)
ipattn_core_kernel(const __grid_constant__ CUtensorMap tmQ, AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const uint32_t q_bytes = 128u * a.d_pad * 2;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + q_bytes;
    uint8_t* sV = sK + a.k_tile_bytes;
    uint8_t* sP = sV + a.v_tile_bytes;
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(sP + 128 * CHUNK * 2);
    uint64_t* bar_s = bar_load + 1;   // S = Q K^T complete
    uint64_t* bar_pv = bar_load + 2;  // O += P V complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 3);

    const int t = threadIdx.x;
    const int warp = t >> 5;
    const int m0 = blockIdx.x * 128;
    const int h = blockIdx.y;
    const int b = blockIdx.z;

    if (t == 0) {
        tma_prefetch_desc(&tmQ);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_pv, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O0 = tmem_base + CHUNK;
    const uint32_t tmem_O1 = tmem_base + CHUNK + a.d_pad;

    if (t == 0) {
        const uint8_t* tile = a.tiles + (static_cast<size_t>(b) * a.H + h) * a.bytes_per_bh;
        mbar_expect_tx(bar_load, (a.q_manual ? 0 : q_bytes) + a.k_tile_bytes + a.v_tile_bytes);
        if (!a.q_manual) tma_load_4d(sQ, &tmQ, 0, m0, h * (a.d / 8), b, bar_load);
        tma_load_1d(sK, tile, a.k_tile_bytes, bar_load);
        tma_load_1d(sV, tile + a.k_tile_bytes, a.v_tile_bytes, bar_load);
    }
    if (a.q_manual) {
        const int srow_q = m0 + t;
        const uint16_t* qrow = a.Q + static_cast<int64_t>(b) * a.q_bs + static_cast<int64_t>(srow_q) * a.q_rs + h * a.d;
        for (int k = 0; k < a.d_pad / 8; ++k) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (srow_q < a.S && h * a.d + k * 8 + 8 <= a.C) v = *reinterpret_cast<const uint4*>(qrow + k * 8);
            *reinterpret_cast<uint4*>(sQ + k * 2048 + t * 16) = v;
        }
        fence_proxy_async_smem();
    }
    mbar_wait(bar_load, 0);
    __syncthreads();

    const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV), sP_u = smem_u32(sP);
    const uint32_t lane_taddr = static_cast<uint32_t>(warp * 32) << 16;  // this warp's TMEM lane quarter
    // Two barriers so that no waiter can fall two phases behind: consecutive commits on the same barrier
    // are always separated by a __syncthreads() that every waiter of the earlier phase has passed.
    uint32_t s_phase = 0, pv_phase = 0;
    float inv_l[2] = {0.f, 0.f};

#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
        const int Lseg = seg == 0 ? a.Lt : a.La;
        if (Lseg == 0) continue;
        const int col0 = seg == 0 ? 0 : a.Lt_pad;
        const int Lseg_pad = seg == 0 ? a.Lt_pad : a.La_pad;
        const int nchunks = (Lseg_pad + CHUNK - 1) / CHUNK;
        const uint32_t tmem_O = seg == 0 ? tmem_O0 : tmem_O1;
        float mx = -INFINITY;

        // issue S = Q K_chunk^T for the KV columns [col0 + n0, col0 + n0 + N)
        auto issue_s = [&](int n0, int N) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_f16(a.fmt, 128, N);
            for (int kk = 0; kk < a.d_pad / 16; ++kk) {
                const uint64_t da = umma_desc(sQ_u + kk * 2 * 2048, 2048, 128, kLayoutNone);
                const uint64_t db = umma_desc(sK_u + (col0 + n0) * 16 + kk * 2 * (a.L_pad * 16), a.L_pad * 16, 128,
                                              kLayoutNone);
                umma_ss(tmem_S, da, db, idesc, kk > 0);
            }
            umma_commit(bar_s);
        };
        // row maximum over the valid columns of the S tile currently in TMEM
        auto row_max = [&](int N, int valid) {
            for (int c = 0; c < N; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_S + lane_taddr + c, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
        };

        if (nchunks > 1) {  // pass 1: global row maximum of the segment
            for (int c = 0; c < nchunks; ++c) {
                const int n0 = c * CHUNK;
                const int N = min(CHUNK, Lseg_pad - n0);
                if (warp == 0 && elect_one()) issue_s(n0, N);
                mbar_wait(bar_s, s_phase);
                s_phase ^= 1;
                tc_fence_after();
                row_max(N, min(N, Lseg - n0));
                tc_fence_before();
                __syncthreads();
            }
        }

        float lsum = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const int n0 = c * CHUNK;
            const int N = min(CHUNK, Lseg_pad - n0);
            const int valid = min(N, Lseg - n0);
            if (warp == 0 && elect_one()) issue_s(n0, N);
            mbar_wait(bar_s, s_phase);
            s_phase ^= 1;
            tc_fence_after();
            if (nchunks == 1) row_max(N, valid);
            const float mxs = mx * a.scale_log2;
            // P = exp2(S * scale_log2 - max * scale_log2), unnormalised, 16-bit, K-major core-matrix order:
            // element (row, col) at  (col / 8) * 2048 + row * 16 + (col % 8) * 2
            for (int cc = 0; cc < N; cc += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_S + lane_taddr + cc, r);
                tmem_ld_wait();
                float p[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float e = exp2f(fmaf(__uint_as_float(r[i]), a.scale_log2, -mxs));
                    p[i] = (cc + i < valid) ? e : 0.f;
                }
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    w[i] = pack16(p[2 * i], p[2 * i + 1], a.fmt);
                    float lo, hi;
                    unpack16(w[i], a.fmt, lo, hi);  // sum what the tensor core will actually see
                    lsum += lo + hi;
                }
                uint8_t* dst = sP + (cc / 8) * 2048 + t * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + 2048) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncthreads();
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                const uint32_t idesc = umma_idesc_f16(a.fmt, 128, a.d_pad);
                for (int kk = 0; kk < N / 16; ++kk) {
                    const uint64_t da = umma_desc(sP_u + kk * 2 * 2048, 2048, 128, kLayoutNone);
                    const uint64_t db = umma_desc(sV_u + ((col0 + n0) / 8 + kk * 2) * (a.d_pad * 16), a.d_pad * 16,
                                                  128, kLayoutNone);
                    umma_ss(tmem_O, da, db, idesc, (c | kk) != 0);
                }
                umma_commit(bar_pv);
            }
            mbar_wait(bar_pv, pv_phase);  // P tile and S tile are free again after this
            pv_phase ^= 1;
            tc_fence_after();
        }
        inv_l[seg] = 1.0f / lsum;
    }

    // ---- epilogue: O = O_t / l_t + ap_scale * O_a / l_a, one row per thread
    const int srow = m0 + t;
    const float w0 = inv_l[0];
    const float w1 = a.ap_scale * inv_l[1];
    uint16_t* orow = reinterpret_cast<uint16_t*>(a.O) + static_cast<int64_t>(b) * a.o_bs + static_cast<int64_t>(srow) * a.o_rs + h * a.d;
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
        if (srow < a.S) {
            uint32_t p[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = pack16(v[2 * i], v[2 * i + 1], a.fmt);
            if (c < a.d) *reinterpret_cast<uint4*>(orow + c) = make_uint4(p[0], p[1], p[2], p[3]);
            if (c + 8 < a.d) *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(p[4], p[5], p[6], p[7]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- context-cache packing ---------------------------------------------------------------------
// K tile  (B operand of S = Q K^T, N = L_pad rows, K = d_pad):   [d_pad/8][L_pad][8]
// V^T tile (B operand of O = P V,  N = d_pad rows, K = L_pad):   [L_pad/8][d_pad][8]
// extended V^T tile (N = dv_pad rows):                           [L_pad/8][dv_pad][8]; row d_pad is all ones, rows above it zero
// Text tokens occupy KV columns [0, Lt), audio tokens [Lt_pad, Lt_pad + La); everything else is zero.
__global__ void kv_pack_kernel(const uint16_t* __restrict__ Kt, const uint16_t* __restrict__ Vt,
                               const uint16_t* __restrict__ Ka, const uint16_t* __restrict__ Va,
                               uint16_t* __restrict__ tiles, KvPackGeom g, int Bc, uint16_t one) {
    const int L = g.Lt + g.La;
    const size_t k_elems = static_cast<size_t>(g.d_pad) * g.L_pad, vx_elems = static_cast<size_t>(g.dv_pad) * g.L_pad;
    const size_t per_bh = 2 * k_elems + vx_elems;
    const size_t total = static_cast<size_t>(Bc) * g.H * per_bh;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t bh = idx / per_bh;
        size_t r = idx % per_bh;
        const bool is_v = r >= k_elems;                 // V^T or extended V^T
        const bool is_x = r >= 2 * k_elems;             // extended V^T
        const int vrows = is_x ? g.dv_pad : g.d_pad;
        if (is_x) r -= 2 * k_elems; else if (is_v) r -= k_elems;
        const int e = r % 8;
        const size_t in_tile = r / 8;
        const int hh = bh % g.H;
        const int bb = bh / g.H;
        int l, dim;  // KV column (padded index space) and head dimension
        if (!is_v) { dim = (in_tile / g.L_pad) * 8 + e; l = in_tile % g.L_pad; }
        else       { l = (in_tile / vrows) * 8 + e; dim = in_tile % vrows; }
        uint16_t v = 0;
        if (is_v && dim == g.d_pad) {
            v = one;
        } else if (dim < g.d) {
            const int col = hh * g.d + dim;
            if (l < g.Lt) {
                const uint16_t* src = is_v ? Vt : Kt;  // [Bc, L, C]; text rows come from the to_k/to_v product
                v = src[(static_cast<size_t>(bb) * L + l) * g.C + col];
            } else if (l >= g.Lt_pad && l - g.Lt_pad < g.La) {
                const uint16_t* src = is_v ? Va : Ka;  // audio rows come from the to_k_ip/to_v_ip product
                v = src[(static_cast<size_t>(bb) * L + g.Lt + (l - g.Lt_pad)) * g.C + col];
            }
        }
        tiles[idx] = v;
    }
}

}  // namespace

int kv_pack(const void* Kt, const void* Vt, const void* Ka, const void* Va, void* cache, int Bc, int Lt, int La,
            int C, int H, int dtype, cudaStream_t st) {
    const KvPackGeom g = kv_pack_geom(Lt, La, C, H);
    const size_t total = static_cast<size_t>(Bc) * H * g.bytes_per_bh / 2;
    const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 8));
    const uint16_t one = dtype == APX_F16 ? 0x3C00 : 0x3F80;   // 1.0 in fp16 / bf16
    kv_pack_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(Kt), static_cast<const uint16_t*>(Vt),
                                           static_cast<const uint16_t*>(Ka), static_cast<const uint16_t*>(Va),
                                           static_cast<uint16_t*>(cache), g, Bc, one);
    APX_LAUNCH_CHECK("kv_pack_kernel");
    return APX_OK;
}

// Q / O are [Bc, S, H*d] VIEWS (element strides q_bs / q_rs, o_bs / o_rs; head h at column h*d); C = H*d.
static int attention_core_launch(const void* Q, int64_t q_bs, int64_t q_rs, const void* cache_tiles, void* O, int64_t o_bs,
                                 int64_t o_rs, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype,
                                 cudaStream_t st) {
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "ipattn_core: dtype %d", dtype);
    APX_REQUIRE(H > 0 && C % H == 0, APX_EINVAL, "ipattn_core: C=%d not divisible by H=%d", C, H);
    const KvPackGeom g = kv_pack_geom(Lt, La, C, H);
    APX_REQUIRE(g.d % 8 == 0 && g.d_pad <= 80, APX_EUNSUPPORTED, "ipattn_core: head dim %d (need d %% 8 == 0, d <= 80)",
                g.d);
    APX_REQUIRE(Lt + La > 0, APX_EINVAL, "ipattn_core: empty context");
    APX_REQUIRE(q_rs % 8 == 0 && q_bs % 8 == 0 && o_rs % 8 == 0 && o_bs % 8 == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(O) & 15) == 0,
                APX_EINVAL, "ipattn_core: Q / O rows must be 16-byte aligned");
    const size_t smem = 128 + 128u * g.d_pad * 2 + g.k_tile_bytes + g.v_tile_bytes + 128 * CHUNK * 2 + 64;
    APX_REQUIRE(smem <= 227 * 1024, APX_EUNSUPPORTED, "ipattn_core: context of %d+%d tokens needs %zu B of shared memory",
                Lt, La, smem);

    CUtensorMap tmQ;
    {
        // Q viewed as [Bc][C/8][S][8]: the box {8, 128, d_pad/8, 1} lands as [d_pad/8][128][8] = core-matrix order.
        uint64_t dims[4] = {8, static_cast<uint64_t>(S), static_cast<uint64_t>(C / 8), static_cast<uint64_t>(Bc)};
        uint64_t strides[3] = {static_cast<uint64_t>(q_rs) * 2, 16, static_cast<uint64_t>(q_bs) * 2};
        uint32_t box[4] = {8, 128, static_cast<uint32_t>(g.d_pad / 8), 1};
        int rc = encode_tmap(&tmQ, dtype, 4, Q, dims, strides, box, 0);
        if (rc) return rc;
    }
    AttnArgs a;
    a.tiles = static_cast<const uint8_t*>(cache_tiles);
    a.O = O;
    a.o_bs = o_bs; a.o_rs = o_rs; a.q_bs = q_bs; a.q_rs = q_rs;
    a.Bc = Bc; a.S = S; a.C = C; a.H = H;
    a.Lt = Lt; a.La = La; a.Lt_pad = g.Lt_pad; a.La_pad = g.La_pad; a.L_pad = g.L_pad; a.d = g.d; a.d_pad = g.d_pad;
    a.ap_scale = ap_scale;
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(g.d));
    a.fmt = dtype == APX_F16 ? 0 : 1;
    a.k_tile_bytes = static_cast<uint32_t>(g.k_tile_bytes);
    a.v_tile_bytes = static_cast<uint32_t>(g.v_tile_bytes);
    a.dv_pad = g.dv_pad;
    a.bytes_per_bh = static_cast<uint32_t>(g.bytes_per_bh);
    a.Q = static_cast<const uint16_t*>(Q);
    static const int q_manual = [] { const char* e = getenv("APX_Q_STAGE"); return e && !strcmp(e, "manual") ? 1 : 0; }();
    a.q_manual = q_manual;

    dim3 grid(ceil_div(S, 128), H, Bc);
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        APX_CUDA(cudaFuncSetAttribute(ipattn_core_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        APX_CUDA(cudaFuncSetAttribute(ipattn_core_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (g.d_pad <= 64) ipattn_core_kernel<256><<<grid, 128, smem, st>>>(tmQ, a);
    else               ipattn_core_kernel<512><<<grid, 128, smem, st>>>(tmQ, a);
    APX_LAUNCH_CHECK("ipattn_core_kernel");
    return APX_OK;
}

int ipattn_core_tc(const void* Q, const void* cache_tiles, void* O, int Bc, int S, int C, int H, int Lt, int La,
                   float ap_scale, int dtype, cudaStream_t st) {
    return attention_core_launch(Q, static_cast<int64_t>(S) * C, C, cache_tiles, O, static_cast<int64_t>(S) * C, C, Bc, S, C, H,
                                 Lt, La, ap_scale, dtype, st);
}

// ---- plain multi-head self-attention on the same kernel (the AudioMAE ViT blocks: timm Attention, N = 513, d = 64) ----------
// K / V are [B, L, H*d] views; they are re-laid into K / V^T operand tiles (one "audio" segment of L keys) in `tiles`.
namespace {
__global__ void kv_pack_views_kernel(const uint16_t* __restrict__ K, int64_t k_bs, int64_t k_rs, const uint16_t* __restrict__ V,
                                     int64_t v_bs, int64_t v_rs, uint16_t* __restrict__ tiles, KvPackGeom g, int B) {
    // one block per (batch, head, 8-key group): stage the 8 x d slab of K and of V, write both tile orders coalesced
    __shared__ uint16_t sk[8][96 + 2], sv[8][96 + 2];
    const int lg = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int l0 = lg * 8;
    for (int i = threadIdx.x; i < 8 * g.d_pad; i += blockDim.x) {
        const int r = i / g.d_pad, c = i % g.d_pad;
        const int l = l0 + r;
        const bool ok = l < g.La && c < g.d;
        sk[r][c] = ok ? K[b * k_bs + static_cast<int64_t>(l) * k_rs + h * g.d + c] : uint16_t(0);
        sv[r][c] = ok ? V[b * v_bs + static_cast<int64_t>(l) * v_rs + h * g.d + c] : uint16_t(0);
    }
    __syncthreads();
    uint16_t* kt = tiles + (static_cast<size_t>(b) * g.H + h) * (g.bytes_per_bh / 2);
    uint16_t* vt = kt + g.k_tile_bytes / 2;
    for (int i = threadIdx.x; i < 8 * g.d_pad; i += blockDim.x) {
        // K tile [d_pad/8][L_pad][8]: (dim group dg, key l, e) ; this block owns keys l0..l0+7
        const int dg = i / 64, rem = i % 64, r = rem / 8, e = rem % 8;
        kt[(static_cast<size_t>(dg) * g.L_pad + l0 + r) * 8 + e] = sk[r][dg * 8 + e];
        // V^T tile [L_pad/8][d_pad][8]: (key group lg, dim, e = key within the group)
        const int dim = i / 8, ee = i % 8;
        vt[(static_cast<size_t>(lg) * g.d_pad + dim) * 8 + ee] = sv[ee][dim];
    }
}
}  // namespace

size_t attention_tc_tiles_bytes(int B, int H, int d, int L) { return static_cast<size_t>(B) * H * kv_pack_geom(0, L, H * d, H).bytes_per_bh; }

int attention_tc(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs, const void* V, int64_t v_bs,
                 int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S, int H, int d, int L, int dtype, void* tiles,
                 cudaStream_t st) {
    APX_REQUIRE(d % 8 == 0 && d <= 80, APX_EUNSUPPORTED, "attention_tc: head dim %d", d);
    const KvPackGeom g = kv_pack_geom(0, L, H * d, H);
    kv_pack_views_kernel<<<dim3(g.L_pad / 8, H, B), 128, 0, st>>>(static_cast<const uint16_t*>(K), k_bs, k_rs,
                                                                  static_cast<const uint16_t*>(V), v_bs, v_rs,
                                                                  static_cast<uint16_t*>(tiles), g, B);
    APX_LAUNCH_CHECK("kv_pack_views_kernel");
    return attention_core_launch(Q, q_bs, q_rs, tiles, O, o_bs, o_rs, B, S, H * d, H, 0, L, 1.0f, dtype, st);
}

}  // namespace apx
