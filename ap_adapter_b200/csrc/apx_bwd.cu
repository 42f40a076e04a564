// apx_bwd.cu — backward of one adapter site for adapter-only training (train_apadapter_v2.py:892-979, config 5):
//
//   Y = (softmax(Q K_t^T / sqrt(d)) V_t + ap_scale * softmax(Q K_a^T / sqrt(d)) V_a) Wo^T + bo,  Q = X Wq^T,
//   K_a = E_a Wk_ip^T,  V_a = E_a Wv_ip^T          (attention_processor.py:387-459)
//
// Given dY it produces dX (so that the loss reaches the adapter sites further up the UNet), dWk_ip and dWv_ip (the only
// trainable tensors; everything else is frozen).  Sequence: dO = dY Wo and Q = X Wq^T on the tcgen05 GEMM, the attention
// backward on CUDA cores (scores and probabilities are recomputed per query row; dK_a / dV_a accumulate in shared memory
// per block and are flushed with atomics), dX = dQ Wq on the tcgen05 GEMM, and the two weight gradients as
// sum-over-tokens outer products  dW[c, k] = sum_m dK_a[m, c] E_a[m, k].
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include <algorithm>

namespace apx {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

constexpr int BWD_MAXJ = 17;   // keys per lane -> Lt + La <= 544 (the un-pooled AudioMAE grid: 8 + 512, train_apadapter_v2.py:445-446)

struct BwdArgs {
    const uint16_t* Q;    // [Bc, S, C]
    const uint16_t* dO;   // [Bc, S, C]
    const uint16_t* K;    // [Bc, L, C] natural layout of the context cache (text rows first)
    const uint16_t* V;
    uint16_t* dQ;         // [Bc, S, C]
    float* dKa;           // [Bc, La, C] fp32, zero-initialised
    float* dVa;
    int S, C, H, d, L, Lt, La;
    float tau, ap_scale;
    int fmt, pitch;       // K / V row pitch in shared memory (16-bit elements)
};

// NJ = key groups of 32 a lane owns one key of (ceil(L / 32)), R = query rows per warp (4 warps per block), NJC = key groups
// whose dK_a / dV_a partial sums are held in registers at once in phase 2 (NJ = 17 walks its keys in chunks of 5 groups).
// Phase 1, per row: recompute scores and probabilities (two softmax segments), write dq, and KEEP for the lane's keys the two
// coefficients the K / V gradients need (tau * ds_j and ap * p_j, zero for text keys) in registers.
// Phase 2, per 8-dim chunk of the head: dK_a[j, c] += sum_rows coef_k[row][j] * q[row][c] (same for dV_a with dO) accumulates
// in REGISTERS over the warp's rows, is reduced over the block's warps in shared memory and flushed with one global atomic
// per (key, dim) per block.  (The first version did one shared-memory atomic per key, dim AND row: 2.3 ms per call.)
template <int NJ, int R, int NJC>
__global__ void __launch_bounds__(128) ipattn_bwd_kernel(BwdArgs a) {
    extern __shared__ __align__(16) uint8_t smem_bwd[];
    uint16_t* sK = reinterpret_cast<uint16_t*>(smem_bwd);
    uint16_t* sV = sK + static_cast<size_t>(a.L) * a.pitch;
    float* sq = reinterpret_cast<float*>(sV + static_cast<size_t>(a.L) * a.pitch);   // [4 warps][R][2][d]: q and dO of the warp's rows
    float* sred = sq + 4 * R * 2 * a.d;                                               // [2][NJC * 32][8]: block reduction of one chunk

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * (4 * R);
    const int dv = a.d / 8;
    const uint16_t* Kg = a.K + (static_cast<size_t>(b) * a.L) * a.C + h * a.d;
    const uint16_t* Vg = a.V + (static_cast<size_t>(b) * a.L) * a.C + h * a.d;
    for (int i = threadIdx.x; i < a.L * dv; i += 128) {
        const int l = i / dv, c = i % dv;
        *reinterpret_cast<uint4*>(sK + l * a.pitch + c * 8) = *reinterpret_cast<const uint4*>(Kg + static_cast<size_t>(l) * a.C + c * 8);
        *reinterpret_cast<uint4*>(sV + l * a.pitch + c * 8) = *reinterpret_cast<const uint4*>(Vg + static_cast<size_t>(l) * a.C + c * 8);
    }
    __syncthreads();

    float ck[R][NJ], cv[R][NJ];   // tau * ds_j and ap * p_j of the lane's keys, per row (zero for text keys / missing rows)
    float* wq = sq + warp * R * 2 * a.d;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int s = q0 + warp * R + r;
        float* myq = wq + r * 2 * a.d;
        float* mydo = myq + a.d;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) { ck[r][jj] = 0.f; cv[r][jj] = 0.f; }
        if (s >= a.S) {   // warp-uniform; the row contributes nothing
            for (int i = lane; i < 2 * a.d; i += 32) myq[i] = 0.f;
            continue;
        }
        const size_t rowoff = (static_cast<size_t>(b) * a.S + s) * a.C + h * a.d;
        for (int i = lane; i < a.d; i += 32) {
            myq[i] = ld16f(a.Q + rowoff, i, a.fmt);
            mydo[i] = ld16f(a.dO + rowoff, i, a.fmt);
        }
        __syncwarp();
        float sc[NJ], dp[NJ];
        float m_t = -INFINITY, m_a = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int j = jj * 32 + lane;
            float accs = -INFINITY, accd = 0.f;
            if (j < a.L) {
                accs = 0.f;
                const uint16_t* kr = sK + j * a.pitch;
                const uint16_t* vr = sV + j * a.pitch;
#pragma unroll 1
                for (int c = 0; c < a.d; c += 8) {   // 128-bit K / V row reads (odd 16-byte pitch: conflict-free), q / dO broadcast
                    const uint4 kv = *reinterpret_cast<const uint4*>(kr + c);
                    const uint4 vv = *reinterpret_cast<const uint4*>(vr + c);
                    const float4 qa = *reinterpret_cast<const float4*>(myq + c), qb = *reinterpret_cast<const float4*>(myq + c + 4);
                    const float4 da = *reinterpret_cast<const float4*>(mydo + c), db = *reinterpret_cast<const float4*>(mydo + c + 4);
                    float x0, x1;
                    unpack16(kv.x, a.fmt, x0, x1); accs = fmaf(qa.x, x0, accs); accs = fmaf(qa.y, x1, accs);
                    unpack16(kv.y, a.fmt, x0, x1); accs = fmaf(qa.z, x0, accs); accs = fmaf(qa.w, x1, accs);
                    unpack16(kv.z, a.fmt, x0, x1); accs = fmaf(qb.x, x0, accs); accs = fmaf(qb.y, x1, accs);
                    unpack16(kv.w, a.fmt, x0, x1); accs = fmaf(qb.z, x0, accs); accs = fmaf(qb.w, x1, accs);
                    unpack16(vv.x, a.fmt, x0, x1); accd = fmaf(da.x, x0, accd); accd = fmaf(da.y, x1, accd);
                    unpack16(vv.y, a.fmt, x0, x1); accd = fmaf(da.z, x0, accd); accd = fmaf(da.w, x1, accd);
                    unpack16(vv.z, a.fmt, x0, x1); accd = fmaf(db.x, x0, accd); accd = fmaf(db.y, x1, accd);
                    unpack16(vv.w, a.fmt, x0, x1); accd = fmaf(db.z, x0, accd); accd = fmaf(db.w, x1, accd);
                }
                accs *= a.tau;
                if (j < a.Lt) m_t = fmaxf(m_t, accs); else { m_a = fmaxf(m_a, accs); accd *= a.ap_scale; }
            }
            sc[jj] = accs; dp[jj] = accd;
        }
        m_t = warp_max_f(m_t);
        m_a = warp_max_f(m_a);
        float l_t = 0.f, l_a = 0.f;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int j = jj * 32 + lane;
            float p = 0.f;
            if (j < a.L) {
                if (j < a.Lt) { p = __expf(sc[jj] - m_t); l_t += p; }
                else          { p = __expf(sc[jj] - m_a); l_a += p; }
            }
            sc[jj] = p;
        }
        l_t = warp_sum_f(l_t);
        l_a = warp_sum_f(l_a);
        const float il_t = a.Lt > 0 ? 1.f / l_t : 0.f, il_a = a.La > 0 ? 1.f / l_a : 0.f;
        float dl_t = 0.f, dl_a = 0.f;   // delta_seg = sum_j p_j dp_j
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const int j = jj * 32 + lane;
            if (j < a.L) {
                const float p = sc[jj] * (j < a.Lt ? il_t : il_a);
                sc[jj] = p;
                if (j < a.Lt) dl_t += p * dp[jj]; else dl_a += p * dp[jj];
            }
        }
        dl_t = warp_sum_f(dl_t);
        dl_a = warp_sum_f(dl_a);
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {   // ds_j = p_j (dp_j - delta)
            const int j = jj * 32 + lane;
            float ds = 0.f;
            if (j < a.L) {
                ds = sc[jj] * (dp[jj] - (j < a.Lt ? dl_t : dl_a));
                if (j >= a.Lt) { ck[r][jj] = a.tau * ds; cv[r][jj] = a.ap_scale * sc[jj]; }
            }
            dp[jj] = ds;
        }
        // dq = tau sum_j ds_j k_j : the ds of the row go through shared memory, then lanes own PAIRS of head dims
        float* sds = sq + 4 * R * 2 * a.d + 2 * NJC * 32 * 8 + warp * NJ * 32;
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) sds[jj * 32 + lane] = dp[jj];
        __syncwarp();
        uint16_t* dqg = a.dQ + rowoff;
        for (int dim = 2 * lane; dim < a.d; dim += 64) {
            float d0 = 0.f, d1 = 0.f;
#pragma unroll 4
            for (int j = 0; j < a.L; ++j) {
                const float ds = sds[j];
                float k0, k1;
                unpack16(*reinterpret_cast<const uint32_t*>(sK + j * a.pitch + dim), a.fmt, k0, k1);
                d0 = fmaf(ds, k0, d0);
                d1 = fmaf(ds, k1, d1);
            }
            *reinterpret_cast<uint32_t*>(dqg + dim) = pack16(d0 * a.tau, d1 * a.tau, a.fmt);
        }
        __syncwarp();
    }
    __syncwarp();

    // ---- phase 2: dK_a / dV_a, one 8-dim chunk of the head and NJC key groups at a time
    for (int c0 = 0; c0 < a.d; c0 += 8) {
#pragma unroll
      for (int j0 = 0; j0 < NJ; j0 += NJC) {
        float ak[NJC][8], av[NJC][8];
#pragma unroll
        for (int jj = 0; jj < NJC; ++jj)
#pragma unroll
            for (int c = 0; c < 8; ++c) { ak[jj][c] = 0.f; av[jj][c] = 0.f; }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 qa = *reinterpret_cast<const float4*>(wq + r * 2 * a.d + c0);
            const float4 qb = *reinterpret_cast<const float4*>(wq + r * 2 * a.d + c0 + 4);
            const float4 da = *reinterpret_cast<const float4*>(wq + r * 2 * a.d + a.d + c0);
            const float4 db = *reinterpret_cast<const float4*>(wq + r * 2 * a.d + a.d + c0 + 4);
            const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const float dov[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
            for (int jj = 0; jj < NJC; ++jj)
                if (j0 + jj < NJ) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        ak[jj][c] = fmaf(ck[r][j0 + jj < NJ ? j0 + jj : 0], qv[c], ak[jj][c]);
                        av[jj][c] = fmaf(cv[r][j0 + jj < NJ ? j0 + jj : 0], dov[c], av[jj][c]);
                    }
                }
        }
        // reduce over the 4 warps in shared memory, then one global atomic per (key, dim)
        __syncthreads();   // previous chunk's flush has finished reading sred
        for (int i = threadIdx.x; i < 2 * NJC * 32 * 8; i += 128) sred[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < NJC; ++jj)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // [key][8]: the 8 dims of a key are contiguous; consecutive lanes = consecutive keys -> stride 8 words: 8-way
                // bank conflict on plain stores, harmless for the few atomics left here
                atomicAdd(&sred[(jj * 32 + lane) * 8 + c], ak[jj][c]);
                atomicAdd(&sred[NJC * 32 * 8 + (jj * 32 + lane) * 8 + c], av[jj][c]);
            }
        __syncthreads();
        for (int i = threadIdx.x; i < NJC * 32 * 8; i += 128) {
            const int j = j0 * 32 + i / 8, c = i % 8;
            if (j >= a.Lt && j < a.L && c0 + c < a.d) {
                const size_t g = (static_cast<size_t>(b) * a.La + (j - a.Lt)) * a.C + h * a.d + c0 + c;
                atomicAdd(&a.dKa[g], sred[i]);
                atomicAdd(&a.dVa[g], sred[NJC * 32 * 8 + i]);
            }
        }
      }
    }
}

// out[N1, N2] (fp32) += sum over this block's slice of m of A[m, N1] (fp32) * B[row(m), N2] (16 bit),
// row(m) = (m / La) * L + Lt + m % La  (audio rows of ctx).  The token dimension is split over blockIdx.z (the output is
// only C x 768: without the split 48 blocks would walk all tokens); `out` must be zeroed.
constexpr int WGRAD_MSPLIT = 128;   // tokens per block
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ A, const uint16_t* __restrict__ Bm, float* __restrict__ out,
                                                   int M, int N1, int N2, int La, int L, int Lt, int fmt) {
    __shared__ float sA[16][64 + 1];
    __shared__ float sB[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int n1_0 = blockIdx.x * 64, n2_0 = blockIdx.y * 64;
    const int m_lo = blockIdx.z * WGRAD_MSPLIT, m_hi = min(M, m_lo + WGRAD_MSPLIT);
    float acc[4][4] = {};
    for (int m0 = m_lo; m0 < m_hi; m0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            const int m = m0 + r;
            float va = 0.f, vb = 0.f;
            if (m < m_hi) {
                if (n1_0 + c < N1) va = A[static_cast<size_t>(m) * N1 + n1_0 + c];
                const size_t brow = static_cast<size_t>(m / La) * L + Lt + m % La;
                if (n2_0 + c < N2) vb = ld16f(Bm, brow * N2 + n2_0 + c, fmt);
            }
            sA[r][c] = va; sB[r][c] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float x[4], y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { x[i] = sA[k][ty * 4 + i]; y[i] = sB[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x[i], y[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n1 = n1_0 + ty * 4 + i, n2 = n2_0 + tx * 4 + j;
            if (n1 < N1 && n2 < N2) atomicAdd(&out[static_cast<size_t>(n1) * N2 + n2], acc[i][j]);
        }
}

__global__ void transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int n) {   // out = in^T, n x n
    __shared__ uint16_t t[32][33];
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8)
        if (x < n && y0 + j < n) t[j][threadIdx.x] = in[static_cast<size_t>(y0 + j) * n + x];
    __syncthreads();
    const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += 8)
        if (ox < n && oy0 + j < n) out[static_cast<size_t>(oy0 + j) * n + ox] = t[threadIdx.x][j];
}

}  // namespace
struct BwdWs { size_t wqT, woT, q, dO, dQ, dKa, dVa, total; };
int bwd_tail(uint8_t* ws, const BwdWs& w, const void* ctx, void* dX, float* dWk_ip, float* dWv_ip, int Bc, int S, int C, int H, int Lt, int La, int Dctx,
             int dtype, void* stream);
namespace {
BwdWs bwd_ws(int Bc, int S, int C, int La) {
    BwdWs w;
    size_t off = 0;
    const size_t act = align_up(static_cast<size_t>(Bc) * S * C * 2, 256), ww = align_up(static_cast<size_t>(C) * C * 2, 256);
    const size_t g = align_up(static_cast<size_t>(Bc) * std::max(La, 1) * C * 4, 256);
    w.wqT = off; off += ww;
    w.woT = off; off += ww;
    w.q = off; off += act;
    w.dO = off; off += act;
    w.dQ = off; off += act;
    w.dKa = off; off += g;
    w.dVa = off; off += g;
    w.total = off;
    return w;
}

}  // namespace
}  // namespace apx

using namespace apx;

extern "C" {

size_t apx_ipattn_bwd_workspace_bytes(int Bc, int S, int C, int La) {
    if (Bc <= 0 || S <= 0 || C <= 0 || La < 0) return 0;
    return bwd_ws(Bc, S, C, La).total;
}

int apx_ipattn_bwd(const void* dY, const void* X, const void* Wq, const void* Wo, const void* cache, const void* ctx, void* dX,
                   float* dWk_ip, float* dWv_ip, int Bc, int S, int C, int H, int Lt, int La, int Dctx, float ap_scale, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream) {
    APX_REQUIRE(dY && X && Wq && Wo && cache && ctx && dX && dWk_ip && dWv_ip && workspace, APX_EINVAL, "apx_ipattn_bwd: null pointer");
    APX_REQUIRE(Bc > 0 && S > 0 && C > 0 && H > 0 && C % H == 0 && Lt >= 0 && La > 0 && Dctx > 0, APX_EINVAL,
                "apx_ipattn_bwd: bad shape Bc=%d S=%d C=%d H=%d Lt=%d La=%d Dctx=%d (the adapter branch needs La > 0)", Bc, S, C, H, Lt,
                La, Dctx);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "apx_ipattn_bwd: dtype %d", dtype);
    const int d = C / H, L = Lt + La;
    APX_REQUIRE(d % 8 == 0 && d <= 96 && C % 64 == 0, APX_EUNSUPPORTED, "apx_ipattn_bwd: head dim %d / C=%d", d, C);
    APX_REQUIRE(L <= BWD_MAXJ * 32, APX_EUNSUPPORTED, "apx_ipattn_bwd: context of %d tokens (the backward kernel takes <= %d)", L,
                BWD_MAXJ * 32);
    const BwdWs w = bwd_ws(Bc, S, C, La);
    APX_REQUIRE(workspace_bytes >= w.total, APX_EWORKSPACE, "apx_ipattn_bwd: workspace %zu < %zu", workspace_bytes, w.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    const int M = Bc * S;
    int rc;
    // transposed weights: dO = dY Wo = dY (Wo^T)^T and dX = dQ Wq = dQ (Wq^T)^T on the A W^T GEMM
    dim3 tg((C + 31) / 32, (C + 31) / 32), tb(32, 8);
    transpose16_kernel<<<tg, tb, 0, st>>>(static_cast<const uint16_t*>(Wq), reinterpret_cast<uint16_t*>(ws + w.wqT), C);
    APX_LAUNCH_CHECK("transpose16_kernel");
    transpose16_kernel<<<tg, tb, 0, st>>>(static_cast<const uint16_t*>(Wo), reinterpret_cast<uint16_t*>(ws + w.woT), C);
    APX_LAUNCH_CHECK("transpose16_kernel");
    rc = apx_linear(dY, C, ws + w.woT, C, nullptr, nullptr, 0, ws + w.dO, C, M, C, C, dtype, dtype, APX_EPI_NONE, stream);
    if (rc) return rc;
    rc = apx_linear(X, C, Wq, C, nullptr, nullptr, 0, ws + w.q, C, M, C, C, dtype, dtype, APX_EPI_NONE, stream);
    if (rc) return rc;
    APX_CUDA(cudaMemsetAsync(ws + w.dKa, 0, static_cast<size_t>(Bc) * La * C * 4, st));
    APX_CUDA(cudaMemsetAsync(ws + w.dVa, 0, static_cast<size_t>(Bc) * La * C * 4, st));

    if (ipattn_bwd_tc_supported(C, H, Lt, La)) {
        // tensor-core attention backward (apx_bwd_tc.cu): contexts of up to 16 + 128 padded tokens, d % 16 == 0
        const size_t tiles_off = 2 * align_up(static_cast<size_t>(Bc) * L * C * 2, 256);   // [K natural][V natural][tiles], see apx_api.cu
        rc = ipattn_bwd_tc(ws + w.q, ws + w.dO, static_cast<const uint8_t*>(cache) + tiles_off, ws + w.dQ, reinterpret_cast<float*>(ws + w.dKa),
                           reinterpret_cast<float*>(ws + w.dVa), Bc, S, C, H, Lt, La, ap_scale, dtype, st);
        if (rc) return rc;
        return bwd_tail(ws, w, ctx, dX, dWk_ip, dWv_ip, Bc, S, C, H, Lt, La, Dctx, dtype, stream);
    }
    BwdArgs a;
    a.Q = reinterpret_cast<const uint16_t*>(ws + w.q);
    a.dO = reinterpret_cast<const uint16_t*>(ws + w.dO);
    a.K = static_cast<const uint16_t*>(cache);   // [K natural][V natural][tiles]: see ctx_cache_layout in apx_api.cu
    a.V = reinterpret_cast<const uint16_t*>(static_cast<const uint8_t*>(cache) + align_up(static_cast<size_t>(Bc) * L * C * 2, 256));
    a.dQ = reinterpret_cast<uint16_t*>(ws + w.dQ);
    a.dKa = reinterpret_cast<float*>(ws + w.dKa);
    a.dVa = reinterpret_cast<float*>(ws + w.dVa);
    a.S = S; a.C = C; a.H = H; a.d = d; a.L = L; a.Lt = Lt; a.La = La;
    a.tau = 1.0f / sqrtf(static_cast<float>(d));
    a.ap_scale = ap_scale;
    a.fmt = dtype == APX_F16 ? 0 : 1;
    int p16 = d / 8 + 1;
    if (p16 % 2 == 0) ++p16;
    a.pitch = p16 * 8;
    const int nj = ceil_div(L, 32);
    const int njt = nj <= 1 ? 1 : nj <= 2 ? 2 : nj <= 5 ? 5 : nj <= 9 ? 9 : 17;   // instantiated key-group counts
    const int rows = njt <= 5 ? 8 : njt <= 9 ? 4 : 2;              // query rows per warp (register budget of the coefficient cache)
    const int njc = njt <= 9 ? njt : 5;                            // key groups per phase-2 pass
    const size_t smem = static_cast<size_t>(L) * a.pitch * 2 * 2 + static_cast<size_t>(4) * rows * 2 * d * 4 +
                        static_cast<size_t>(2) * njc * 32 * 8 * 4 + static_cast<size_t>(4) * njt * 32 * 4;
    APX_REQUIRE(smem <= 227 * 1024, APX_EUNSUPPORTED, "apx_ipattn_bwd: context %d+%d at d=%d needs %zu B of shared memory", Lt, La, d, smem);
    typedef void (*BwdFn)(BwdArgs);
    static const BwdFn fns[5] = {ipattn_bwd_kernel<1, 8, 1>, ipattn_bwd_kernel<2, 8, 2>, ipattn_bwd_kernel<5, 8, 5>, ipattn_bwd_kernel<9, 4, 9>,
                                 ipattn_bwd_kernel<17, 2, 5>};
    const BwdFn fn = fns[njt == 1 ? 0 : njt == 2 ? 1 : njt == 5 ? 2 : njt == 9 ? 3 : 4];
    for (int i = 0; i < 5; ++i)   // per-device attribute; cheap enough to set on every call (a static flag would be per-process)
        APX_CUDA(cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    fn<<<dim3(ceil_div(S, 4 * rows), H, Bc), 128, smem, st>>>(a);
    APX_LAUNCH_CHECK("ipattn_bwd_kernel");
    return bwd_tail(ws, w, ctx, dX, dWk_ip, dWv_ip, Bc, S, C, H, Lt, La, Dctx, dtype, stream);
}

}  // extern "C"

namespace apx {
// dX = dQ Wq on the tcgen05 GEMM, then the two weight gradients dW[c, k] = sum_m dK_a[m, c] E_a[m, k]
int bwd_tail(uint8_t* ws, const BwdWs& w, const void* ctx, void* dX, float* dWk_ip, float* dWv_ip, int Bc, int S, int C, int H, int Lt, int La, int Dctx,
             int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int M = Bc * S, L = Lt + La;
    int rc = apx_linear(ws + w.dQ, C, ws + w.wqT, C, nullptr, nullptr, 0, dX, C, M, C, C, dtype, dtype, APX_EPI_NONE, stream);
    if (rc) return rc;
    dim3 wg(ceil_div(C, 64), ceil_div(Dctx, 64), ceil_div(Bc * La, WGRAD_MSPLIT));
    APX_CUDA(cudaMemsetAsync(dWk_ip, 0, static_cast<size_t>(C) * Dctx * 4, st));
    APX_CUDA(cudaMemsetAsync(dWv_ip, 0, static_cast<size_t>(C) * Dctx * 4, st));
    const int fmt = dtype == APX_F16 ? 0 : 1;
    wgrad_kernel<<<wg, 256, 0, st>>>(reinterpret_cast<const float*>(ws + w.dKa), static_cast<const uint16_t*>(ctx), dWk_ip, Bc * La, C, Dctx, La, L, Lt, fmt);
    APX_LAUNCH_CHECK("wgrad_kernel");
    wgrad_kernel<<<wg, 256, 0, st>>>(reinterpret_cast<const float*>(ws + w.dVa), static_cast<const uint16_t*>(ctx), dWv_ip, Bc * La, C, Dctx, La, L, Lt, fmt);
    APX_LAUNCH_CHECK("wgrad_kernel");
    (void)H;
    return APX_OK;
}
}  // namespace apx
