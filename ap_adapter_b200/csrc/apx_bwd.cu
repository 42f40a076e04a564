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

constexpr int BWD_MAXJ = 9;    // keys per lane -> Lt + La <= 288
constexpr int BWD_QPB = 32;    // query rows per block

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
    int la_pitch;         // pitch of the [d][La] fp32 accumulators
};

__global__ void __launch_bounds__(128) ipattn_bwd_kernel(BwdArgs a) {
    extern __shared__ __align__(16) uint8_t smem_bwd[];
    uint16_t* sK = reinterpret_cast<uint16_t*>(smem_bwd);
    uint16_t* sV = sK + static_cast<size_t>(a.L) * a.pitch;
    float* sdK = reinterpret_cast<float*>(sV + static_cast<size_t>(a.L) * a.pitch);   // [d][la_pitch]: conflict-free per-key atomics
    float* sdV = sdK + static_cast<size_t>(a.d) * a.la_pitch;
    float* sq = sdV + static_cast<size_t>(a.d) * a.la_pitch;                           // [4 warps][2][d]: q and dO of the current row

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * BWD_QPB;
    const int dv = a.d / 8;
    const uint16_t* Kg = a.K + (static_cast<size_t>(b) * a.L) * a.C + h * a.d;
    const uint16_t* Vg = a.V + (static_cast<size_t>(b) * a.L) * a.C + h * a.d;
    for (int i = threadIdx.x; i < a.L * dv; i += 128) {
        const int l = i / dv, c = i % dv;
        *reinterpret_cast<uint4*>(sK + l * a.pitch + c * 8) = *reinterpret_cast<const uint4*>(Kg + static_cast<size_t>(l) * a.C + c * 8);
        *reinterpret_cast<uint4*>(sV + l * a.pitch + c * 8) = *reinterpret_cast<const uint4*>(Vg + static_cast<size_t>(l) * a.C + c * 8);
    }
    for (int i = threadIdx.x; i < 2 * a.d * a.la_pitch; i += 128) sdK[i] = 0.f;
    __syncthreads();

    float* myq = sq + warp * 2 * a.d;
    float* mydo = myq + a.d;
    for (int qi = warp; qi < BWD_QPB; qi += 4) {
        const int s = q0 + qi;
        if (s >= a.S) break;   // warp-uniform
        const size_t rowoff = (static_cast<size_t>(b) * a.S + s) * a.C + h * a.d;
        for (int i = lane; i < a.d; i += 32) {
            myq[i] = ld16f(a.Q + rowoff, i, a.fmt);
            mydo[i] = ld16f(a.dO + rowoff, i, a.fmt);
        }
        __syncwarp();
        // scores s_j = tau q.k_j and dp_j = w_seg * dO.v_j for the keys this lane owns (j = lane + 32 jj)
        float sc[BWD_MAXJ], dp[BWD_MAXJ];
        float m_t = -INFINITY, m_a = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < BWD_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            float accs = -INFINITY, accd = 0.f;
            if (j < a.L) {
                accs = 0.f;
                const uint16_t* kr = sK + j * a.pitch;
                const uint16_t* vr = sV + j * a.pitch;
                for (int c = 0; c < a.d; c += 2) {
                    float k0, k1, v0, v1;
                    unpack16(*reinterpret_cast<const uint32_t*>(kr + c), a.fmt, k0, k1);
                    unpack16(*reinterpret_cast<const uint32_t*>(vr + c), a.fmt, v0, v1);
                    accs = fmaf(myq[c], k0, accs); accs = fmaf(myq[c + 1], k1, accs);
                    accd = fmaf(mydo[c], v0, accd); accd = fmaf(mydo[c + 1], v1, accd);
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
        for (int jj = 0; jj < BWD_MAXJ; ++jj) {
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
        for (int jj = 0; jj < BWD_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            if (j < a.L) {
                const float p = sc[jj] * (j < a.Lt ? il_t : il_a);
                sc[jj] = p;
                if (j < a.Lt) dl_t += p * dp[jj]; else dl_a += p * dp[jj];
            }
        }
        dl_t = warp_sum_f(dl_t);
        dl_a = warp_sum_f(dl_a);
        // ds_j = p_j (dp_j - delta); audio keys also feed dK_a += tau ds_j q and dV_a += ap p_j dO (this lane owns key j)
#pragma unroll
        for (int jj = 0; jj < BWD_MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            float ds = 0.f;
            if (j < a.L) {
                const float p = sc[jj];
                ds = p * (dp[jj] - (j < a.Lt ? dl_t : dl_a));
                if (j >= a.Lt) {
                    const int ja = j - a.Lt;
                    const float wk = a.tau * ds, wv = a.ap_scale * p;
                    for (int c = 0; c < a.d; ++c) {
                        atomicAdd(&sdK[c * a.la_pitch + ja], wk * myq[c]);
                        atomicAdd(&sdV[c * a.la_pitch + ja], wv * mydo[c]);
                    }
                }
            }
            dp[jj] = ds;
        }
        // dq = tau sum_j ds_j k_j : lanes over head dims, ds broadcast by shuffles
        float dq[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < BWD_MAXJ; ++jj) {
            if (jj * 32 >= a.L) break;
            const int lim = min(32, a.L - jj * 32);
            for (int src = 0; src < lim; ++src) {
                const float ds = __shfl_sync(0xffffffffu, dp[jj], src);
                const uint16_t* kr = sK + (jj * 32 + src) * a.pitch;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int dim = lane + 32 * k;
                    if (dim < a.d) dq[k] = fmaf(ds, ld16f(kr, dim, a.fmt), dq[k]);
                }
            }
        }
        uint16_t* dqg = a.dQ + rowoff;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int dim = lane + 32 * k;
            if (dim < a.d) {
                const float v = dq[k] * a.tau;
                if (a.fmt == 0) reinterpret_cast<__half*>(dqg)[dim] = __float2half_rn(v);
                else reinterpret_cast<__nv_bfloat16*>(dqg)[dim] = __float2bfloat16_rn(v);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // flush the block's dK_a / dV_a partial sums
    for (int i = threadIdx.x; i < a.La * a.d; i += 128) {
        const int ja = i / a.d, c = i % a.d;
        const size_t g = (static_cast<size_t>(b) * a.La + ja) * a.C + h * a.d + c;
        atomicAdd(&a.dKa[g], sdK[c * a.la_pitch + ja]);
        atomicAdd(&a.dVa[g], sdV[c * a.la_pitch + ja]);
    }
}

// out[N1, N2] (fp32) = sum_m A[m, N1] (fp32) * B[row(m), N2] (16 bit), row(m) = (m / La) * L + Lt + m % La  (audio rows of ctx)
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ A, const uint16_t* __restrict__ Bm, float* __restrict__ out,
                                                   int M, int N1, int N2, int La, int L, int Lt, int fmt) {
    __shared__ float sA[16][64 + 1];
    __shared__ float sB[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int n1_0 = blockIdx.x * 64, n2_0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int m0 = 0; m0 < M; m0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            const int m = m0 + r;
            float va = 0.f, vb = 0.f;
            if (m < M) {
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
            if (n1 < N1 && n2 < N2) out[static_cast<size_t>(n1) * N2 + n2] = acc[i][j];
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

struct BwdWs { size_t wqT, woT, q, dO, dQ, dKa, dVa, total; };
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
    a.la_pitch = La + 1;
    const size_t smem = static_cast<size_t>(L) * a.pitch * 2 * 2 + static_cast<size_t>(2) * d * a.la_pitch * 4 + 4 * 2 * d * 4;
    APX_REQUIRE(smem <= 227 * 1024, APX_EUNSUPPORTED, "apx_ipattn_bwd: context %d+%d at d=%d needs %zu B of shared memory", Lt, La, d, smem);
    static bool attr_set = false;
    if (!attr_set) {
        APX_CUDA(cudaFuncSetAttribute(ipattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    ipattn_bwd_kernel<<<dim3(ceil_div(S, BWD_QPB), H, Bc), 128, smem, st>>>(a);
    APX_LAUNCH_CHECK("ipattn_bwd_kernel");
    rc = apx_linear(ws + w.dQ, C, ws + w.wqT, C, nullptr, nullptr, 0, dX, C, M, C, C, dtype, dtype, APX_EPI_NONE, stream);
    if (rc) return rc;
    dim3 wg(ceil_div(C, 64), ceil_div(Dctx, 64));
    wgrad_kernel<<<wg, 256, 0, st>>>(a.dKa, static_cast<const uint16_t*>(ctx), dWk_ip, Bc * La, C, Dctx, La, L, Lt, a.fmt);
    APX_LAUNCH_CHECK("wgrad_kernel");
    wgrad_kernel<<<wg, 256, 0, st>>>(a.dVa, static_cast<const uint16_t*>(ctx), dWv_ip, Bc * La, C, Dctx, La, L, Lt, a.fmt);
    APX_LAUNCH_CHECK("wgrad_kernel");
    return APX_OK;
}

}  // extern "C"
