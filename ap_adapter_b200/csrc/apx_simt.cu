// apx_simt.cu — CUDA-core kernels.
//
// (1) The HBM-bound stages of the AudioMAE encoder: patch-embed (+bias +pos-embed +CLS), LayerNorm,
//     CLS-drop + (avg+max)/2 token pooling, and the 513-token ViT self-attention with K/V of a head staged
//     in shared memory and warp-shuffle softmax  (audio_encoder/models_mae.py:19-44,548-570;
//     audio_encoder/AudioMAE.py:148-182).
// (2) Plain GEMM / attention kernels with the same contract as the tcgen05 kernels.  They are the
//     device-side cross-check the parity tests run next to the tensor-core path (APX_IMPL_SIMT).
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include <algorithm>

namespace apx {

namespace {

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// GEMM: 64x64 tile, 256 threads, 4x4 micro-tile, BK = 16, fp32 accumulate.
// ------------------------------------------------------------------------------------------------
struct SimtGemmArgs {
    const void* A; int64_t lda;
    const void* W; int64_t ldw;
    const float* bias; const void* residual; int64_t ldr;
    void* out; int64_t ldo;
    int M, N, K, in_fmt, out_dtype, epilogue;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtGemmArgs g) {
    __shared__ float sA[16][64 + 4];
    __shared__ float sW[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < g.K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, c = i & 15;
            const int gm = m0 + r, gn = n0 + r, gk = k0 + c;
            sA[c][r] = (gm < g.M && gk < g.K) ? ld16f(g.A, static_cast<int64_t>(gm) * g.lda + gk, g.in_fmt) : 0.f;
            sW[c][r] = (gn < g.N && gk < g.K) ? ld16f(g.W, static_cast<int64_t>(gn) * g.ldw + gk, g.in_fmt) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; w[i] = sW[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int epi = g.epilogue & 3;
    auto put = [&](int m, int n, float v) {   // n = OUTPUT column
        if (g.residual) {
            const int64_t ro = static_cast<int64_t>(m) * g.ldr + n;
            if (g.epilogue & APX_EPI_RES16) {   // y rounded to the output type, then y + h (see apx.h)
                v = g.out_dtype == APX_F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v));
                v += ld16f(g.residual, ro, g.out_dtype);
            } else {
                v += reinterpret_cast<const float*>(g.residual)[ro];
            }
        }
        const int64_t o = static_cast<int64_t>(m) * g.ldo + n;
        if (g.out_dtype == APX_F32) reinterpret_cast<float*>(g.out)[o] = v;
        else if (g.out_dtype == APX_F16) reinterpret_cast<__half*>(g.out)[o] = __float2half_rn(v);
        else reinterpret_cast<__nv_bfloat16*>(g.out)[o] = __float2bfloat16_rn(v);
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        if (epi == APX_EPI_GEGLU) {   // columns (2j, 2j+1) = (hidden, gate)
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                const int n = n0 + tx * 4 + j;
                if (n + 1 >= g.N) continue;
                const float a = acc[i][j] + (g.bias ? g.bias[n] : 0.f), gt = acc[i][j + 1] + (g.bias ? g.bias[n + 1] : 0.f);
                put(m, n >> 1, a * gelu_erf_f(gt));
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[n];
            if (epi == APX_EPI_GELU) v = gelu_erf_f(v);
            put(m, n, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Segmented attention.  Block = 4 warps, QPB queries per block; K and V of one (batch, head) are staged in
// shared memory with a conflict-free row pitch; each lane owns keys lane, lane+32, ... for the scores and
// output dims lane, lane+32, lane+64 for P·V; softmax statistics by warp shuffles.
// ------------------------------------------------------------------------------------------------
constexpr int MAXJ = 17;  // keys per lane -> L <= 544
constexpr int QPB = 64;

struct SimtAttnArgs {
    const uint16_t* Q; int64_t q_bs, q_rs;
    const uint16_t* K; int64_t k_bs, k_rs;
    const uint16_t* V; int64_t v_bs, v_rs;
    uint16_t* O; int64_t o_bs, o_rs;
    int B, S, H, d, L, L0;
    float w0, w1, scale;
    int fmt, pitch;  // pitch of the K / V rows in shared memory, in 16-bit elements
};

__global__ void __launch_bounds__(128) attention_simt_kernel(SimtAttnArgs a) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    uint16_t* sK = reinterpret_cast<uint16_t*>(smem_attn);
    uint16_t* sV = sK + static_cast<size_t>(a.L) * a.pitch;
    float* sQ = reinterpret_cast<float*>(sV + static_cast<size_t>(a.L) * a.pitch);  // [4 warps][d]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * QPB;
    const int dv = a.d / 8;  // 16-byte vectors per row

    const uint16_t* Kg = a.K + b * a.k_bs + h * a.d;
    const uint16_t* Vg = a.V + b * a.v_bs + h * a.d;
    for (int i = threadIdx.x; i < a.L * dv; i += 128) {
        const int l = i / dv, c = i % dv;
        *reinterpret_cast<uint4*>(sK + l * a.pitch + c * 8) = *reinterpret_cast<const uint4*>(Kg + l * a.k_rs + c * 8);
        *reinterpret_cast<uint4*>(sV + l * a.pitch + c * 8) = *reinterpret_cast<const uint4*>(Vg + l * a.v_rs + c * 8);
    }
    __syncthreads();

    float* myq = sQ + warp * a.d;
    for (int qi = warp; qi < QPB; qi += 4) {
        const int s = q0 + qi;
        if (s >= a.S) break;  // warp-uniform
        const uint16_t* qg = a.Q + b * a.q_bs + static_cast<int64_t>(s) * a.q_rs + h * a.d;
        for (int i = lane; i < a.d; i += 32) myq[i] = ld16f(qg, i, a.fmt) * a.scale;
        __syncwarp();

        float sc[MAXJ];
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            float acc = -INFINITY;
            if (j < a.L) {
                acc = 0.f;
                const uint16_t* kr = sK + j * a.pitch;
                for (int c = 0; c < dv; ++c) {
                    const uint4 kv = *reinterpret_cast<const uint4*>(kr + c * 8);
                    const float4 qa = *reinterpret_cast<const float4*>(myq + c * 8);
                    const float4 qb = *reinterpret_cast<const float4*>(myq + c * 8 + 4);
                    float k0, k1;
                    unpack16(kv.x, a.fmt, k0, k1); acc = fmaf(qa.x, k0, acc); acc = fmaf(qa.y, k1, acc);
                    unpack16(kv.y, a.fmt, k0, k1); acc = fmaf(qa.z, k0, acc); acc = fmaf(qa.w, k1, acc);
                    unpack16(kv.z, a.fmt, k0, k1); acc = fmaf(qb.x, k0, acc); acc = fmaf(qb.y, k1, acc);
                    unpack16(kv.w, a.fmt, k0, k1); acc = fmaf(qb.z, k0, acc); acc = fmaf(qb.w, k1, acc);
                }
                if (j < a.L0) m0 = fmaxf(m0, acc); else m1 = fmaxf(m1, acc);
            }
            sc[jj] = acc;
        }
        m0 = warp_max(m0);
        m1 = warp_max(m1);
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            float p = 0.f;
            if (j < a.L) {
                if (j < a.L0) { p = __expf(sc[jj] - m0); l0 += p; }
                else          { p = __expf(sc[jj] - m1); l1 += p; }
            }
            sc[jj] = p;
        }
        l0 = warp_sum(l0);
        l1 = warp_sum(l1);
        const float f0 = a.L0 > 0 ? a.w0 / l0 : 0.f;
        const float f1 = a.L > a.L0 ? a.w1 / l1 : 0.f;

        float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            if (jj * 32 >= a.L) break;
            const float mine = sc[jj] * ((jj * 32 + lane) < a.L0 ? f0 : f1);
            const int lim = min(32, a.L - jj * 32);
            for (int src = 0; src < lim; ++src) {
                const float p = __shfl_sync(0xffffffffu, mine, src);
                const uint16_t* vr = sV + (jj * 32 + src) * a.pitch;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int dim = lane + 32 * k;
                    if (dim < a.d) o[k] = fmaf(p, ld16f(vr, dim, a.fmt), o[k]);
                }
            }
        }
        uint16_t* og = a.O + b * a.o_bs + static_cast<int64_t>(s) * a.o_rs + h * a.d;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int dim = lane + 32 * k;
            if (dim < a.d) {
                if (a.fmt == 0) reinterpret_cast<__half*>(og)[dim] = __float2half_rn(o[k]);
                else reinterpret_cast<__nv_bfloat16*>(og)[dim] = __float2bfloat16_rn(o[k]);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Patch embed: Conv2d(1, 768, k=16, s=16) as a per-time-step mini GEMM, fused with +bias, +pos_embed and
// the CLS row.  One block = one (clip, time step): the 16 x 128 mel strip it needs is contiguous (8 KB).
// Weight is passed transposed ([256][768]) so the 768 output channels are the coalesced dimension.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ mel, const float* __restrict__ wT,
                                                         const float* __restrict__ pb, const float* __restrict__ pos,
                                                         const float* __restrict__ cls, float* __restrict__ x) {
    __shared__ float strip[16][128];
    const int t = blockIdx.x, b = blockIdx.y;
    const float* src = mel + (static_cast<size_t>(b) * 1024 + t * 16) * 128;
    for (int i = threadIdx.x; i < 16 * 128 / 4; i += 256)
        reinterpret_cast<float4*>(&strip[0][0])[i] = reinterpret_cast<const float4*>(src)[i];
    __syncthreads();
    float* xb = x + static_cast<size_t>(b) * 513 * 768;
    if (t == 0)
        for (int e = threadIdx.x; e < 768; e += 256) xb[e] = cls[e] + pos[e];
    for (int e = threadIdx.x; e < 768; e += 256) {
        float acc[8] = {};
        for (int k = 0; k < 256; ++k) {
            const float w = wT[k * 768 + e];
            const int i = k >> 4, j = k & 15;
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[f] = fmaf(strip[i][f * 16 + j], w, acc[f]);
        }
        const float bias = pb[e];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            const int tok = 1 + t * 8 + f;
            xb[static_cast<size_t>(tok) * 768 + e] = acc[f] + bias + pos[static_cast<size_t>(tok) * 768 + e];
        }
    }
}

// Patch-embed as a GEMM (models_mae.py:19-44: Conv2d(1, 768, 16, stride 16) == [patches, 256] x [768, 256]^T): one block per
// (clip, time step) turns its contiguous 16 x 128 mel strip into 8 patch rows of the 16-bit A operand (k = ky * 16 + kx, the order of
// the flattened conv weight) and writes the matching rows of the fp32 residual operand (pos_embed; the CLS row carries cls + pos[0]
// - bias behind an all-zero patch, so that the GEMM's bias epilogue leaves exactly cls + pos[0]).
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ mel, const float* __restrict__ pos, const float* __restrict__ cls,
                                                      const float* __restrict__ pb, uint16_t* __restrict__ a16, float* __restrict__ res, int fmt) {
    const int t = blockIdx.x, b = blockIdx.y;
    const float* src = mel + (static_cast<size_t>(b) * 1024 + t * 16) * 128;
    const size_t row0 = static_cast<size_t>(b) * 513 + 1 + t * 8;
    // strip element (ky, f * 16 + kx) -> A[row0 + f][ky * 16 + kx]; two adjacent kx per thread
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
        const int ky = i >> 6, c2 = (i & 63) * 2, f = c2 >> 4, kx = c2 & 15;
        const float2 v = *reinterpret_cast<const float2*>(src + ky * 128 + c2);
        *reinterpret_cast<uint32_t*>(a16 + (row0 + f) * 256 + ky * 16 + kx) = pack16(v.x, v.y, fmt);
    }
    for (int i = threadIdx.x; i < 8 * 768 / 4; i += 256)
        reinterpret_cast<float4*>(res + row0 * 768)[i] = reinterpret_cast<const float4*>(pos + static_cast<size_t>(1 + t * 8) * 768)[i];
    if (t == 0) {
        const size_t r = static_cast<size_t>(b) * 513;
        for (int e = threadIdx.x; e < 768; e += 256) res[r * 768 + e] = cls[e] + pos[e] - pb[e];
        for (int e = threadIdx.x; e < 128; e += 256) reinterpret_cast<uint32_t*>(a16 + r * 256)[e] = 0u;
    }
}
__global__ void __launch_bounds__(256) patch_weight16_kernel(const float* __restrict__ wT /* [256][768] */, uint16_t* __restrict__ w16 /* [768][256] */, int fmt) {
    const int i = blockIdx.x * 256 + threadIdx.x;   // over 768 * 128 pairs
    if (i >= 768 * 128) return;
    const int n = i >> 7, k = (i & 127) * 2;
    reinterpret_cast<uint32_t*>(w16)[i] = pack16(wT[static_cast<size_t>(k) * 768 + n], wT[static_cast<size_t>(k + 1) * 768 + n], fmt);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two-pass statistics in registers, float4 I/O.  cols % 128 == 0, cols <= 1024.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bsh, void* __restrict__ y, int rows,
                                                       int cols, float eps, int out_dtype) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * cols);
    const int nv = cols / 128;  // float4 per lane
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nv) { v[i] = xr[i * 32 + lane]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
    const float mean = warp_sum(s) / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += a * a + b * b + c * c + d * d;
        }
    const float rstd = rsqrtf(warp_sum(q) / cols + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nv) {
            const int c4 = i * 32 + lane;
            const float4 g = reinterpret_cast<const float4*>(w)[c4];
            const float4 bb = reinterpret_cast<const float4*>(bsh)[c4];
            const float o0 = (v[i].x - mean) * rstd * g.x + bb.x;
            const float o1 = (v[i].y - mean) * rstd * g.y + bb.y;
            const float o2 = (v[i].z - mean) * rstd * g.z + bb.z;
            const float o3 = (v[i].w - mean) * rstd * g.w + bb.w;
            if (out_dtype == APX_F32) {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + static_cast<size_t>(row) * cols)[c4] =
                    make_float4(o0, o1, o2, o3);
            } else {
                uint2 p;
                p.x = pack16(o0, o1, out_dtype);
                p.y = pack16(o2, o3, out_dtype);
                reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(y) + static_cast<size_t>(row) * cols)[c4] = p;
            }
        }
}

// ------------------------------------------------------------------------------------------------
// Token pooling: drop CLS, view the 512 tokens as a 64 (time) x 8 (freq) grid, (AvgPool + MaxPool) / 2 with
// kernel = stride = (tp, fp), emit [B, (64/tp)*(8/fp), 768].   (audio_encoder/AudioMAE.py:148-182)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192) pool_tokens_kernel(const float* __restrict__ xn, float* __restrict__ emb, int tp,
                                                         int fp) {
    const int fo_n = 8 / fp;
    const int tok = blockIdx.x, b = blockIdx.y;
    const int to = tok / fo_n, fo = tok % fo_n;
    const float4* xb = reinterpret_cast<const float4*>(xn + static_cast<size_t>(b) * 513 * 768);
    const int c4 = threadIdx.x;  // 192 float4 = 768 channels
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int i = 0; i < tp; ++i)
        for (int j = 0; j < fp; ++j) {
            const int src = 1 + (to * tp + i) * 8 + (fo * fp + j);
            const float4 v = xb[static_cast<size_t>(src) * 192 + c4];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    const float inv = 1.0f / (tp * fp);
    float4 o;
    o.x = (s.x * inv + m.x) * 0.5f; o.y = (s.y * inv + m.y) * 0.5f;
    o.z = (s.z * inv + m.z) * 0.5f; o.w = (s.w * inv + m.w) * 0.5f;
    reinterpret_cast<float4*>(emb + (static_cast<size_t>(b) * gridDim.x + tok) * 768)[c4] = o;
}

// ------------------------------------------------------------------------------------------------
// Kaldi-compatible log-mel filterbank front-end (audio_encoder/AudioMAE.py:356-394 calls torchaudio.compliance.kaldi.fbank with
// htk_compat, hanning window, 25 ms / 10 ms frames at 16 kHz, 128 mel bins, no dither): one block per output frame.
//   frame = wave[i*160 : i*160+400] - global mean;  remove the frame's DC;  pre-emphasis 0.97 (first sample replicated);
//   Hann window (non-periodic);  zero-pad to 512;  |DFT|^2 (direct summation with a shared twiddle table: 257 x 400 MACs);
//   mel = fb[128 x 257] . power;  out = (log(max(mel, eps)) - norm_mean) / (2 norm_std);  frames past the signal are zero-padded
//   BEFORE the normalisation, as the reference does.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wave_mean_kernel(const float* __restrict__ w, int n, float* __restrict__ mean) {
    __shared__ float red[8];
    w += static_cast<size_t>(blockIdx.x) * n;      // one block per clip
    mean += blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += w[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        *mean = t / n;
    }
}

__global__ void __launch_bounds__(256) fbank_kernel(const float* __restrict__ wave, size_t wave_stride, int n_frames, const float* __restrict__ mean,
                                                   const float* __restrict__ melT /* [257][128] */, float* __restrict__ out,
                                                   float norm_mean, float norm_std2) {
    __shared__ float x[512], y[512], ct[512], st[512], pw[257 + 3], red[8];
    const int f = blockIdx.x, t = threadIdx.x;
    wave += static_cast<size_t>(blockIdx.y) * wave_stride;                 // blockIdx.y = clip
    mean += blockIdx.y;
    float* orow = out + (static_cast<size_t>(blockIdx.y) * gridDim.x + f) * 128;
    if (f >= n_frames) {   // zero-padded frame
        if (t < 128) orow[t] = (0.f - norm_mean) / norm_std2;
        return;
    }
    const float gmean = *mean;
    for (int i = t; i < 512; i += 256) {
        x[i] = i < 400 ? wave[static_cast<size_t>(f) * 160 + i] - gmean : 0.f;
        float sn, cs;
        sincospif(static_cast<float>(i) / 256.0f, &sn, &cs);   // angle 2 pi i / 512
        ct[i] = cs; st[i] = sn;
    }
    __syncthreads();
    float s = 0.f;
    for (int i = t; i < 400; i += 256) s += x[i];
    s = warp_sum(s);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    float fm = 0.f;
    for (int i = 0; i < 8; ++i) fm += red[i];
    fm *= (1.0f / 400.0f);
    for (int i = t; i < 512; i += 256) {
        float v = 0.f;
        if (i < 400) {
            const float cur = x[i] - fm, prev = x[i > 0 ? i - 1 : 0] - fm;
            const float win = 0.5f - 0.5f * cospif(2.0f * i / 399.0f);   // torch.hann_window(400, periodic=False)
            v = (cur - 0.97f * prev) * win;
        }
        y[i] = v;
    }
    __syncthreads();
    for (int k = t; k <= 256; k += 256) {
        float re = 0.f, im = 0.f;
#pragma unroll 4
        for (int n = 0; n < 400; ++n) {
            const int idx = (k * n) & 511;
            re = fmaf(y[n], ct[idx], re);
            im = fmaf(y[n], st[idx], im);
        }
        pw[k] = re * re + im * im;
    }
    __syncthreads();
    if (t < 128) {
        float m = 0.f;
        for (int k = 0; k < 257; ++k) m = fmaf(melT[k * 128 + t], pw[k], m);
        orow[t] = (logf(fmaxf(m, 1.1920928955078125e-07f)) - norm_mean) / norm_std2;
    }
}

}  // namespace

int gemm_simt(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
              int64_t ldr, void* out, int64_t ldo, int M, int N, int K, int in_dtype, int out_dtype, int epilogue,
              cudaStream_t st) {
    APX_REQUIRE(in_dtype == APX_F16 || in_dtype == APX_BF16, APX_EUNSUPPORTED, "gemm_simt: in_dtype %d", in_dtype);
    APX_REQUIRE(M > 0 && N > 0 && K > 0, APX_EINVAL, "gemm_simt: bad shape %dx%dx%d", M, N, K);
    SimtGemmArgs g{A, lda, W, ldw, bias, residual, ldr, out, ldo, M, N, K, in_dtype == APX_F16 ? 0 : 1, out_dtype, epilogue};
    dim3 grid(ceil_div(M, 64), ceil_div(N, 64));
    gemm_simt_kernel<<<grid, 256, 0, st>>>(g);
    APX_LAUNCH_CHECK("gemm_simt_kernel");
    return APX_OK;
}

int attention_simt(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs,
                   const void* V, int64_t v_bs, int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S,
                   int H, int d, int L, int L0, float w0, float w1, int dtype, cudaStream_t st) {
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "attention: dtype %d", dtype);
    APX_REQUIRE(d % 8 == 0 && d >= 8 && d <= 96, APX_EUNSUPPORTED, "attention: head dim %d (need d %% 8 == 0, d <= 96)", d);
    APX_REQUIRE(L >= 1 && L <= MAXJ * 32, APX_EUNSUPPORTED, "attention: context length %d (need 1..%d)", L, MAXJ * 32);
    APX_REQUIRE(L0 >= 0 && L0 <= L, APX_EINVAL, "attention: L0=%d outside [0, %d]", L0, L);
    APX_REQUIRE(B > 0 && S > 0 && H > 0, APX_EINVAL, "attention: bad shape");
    APX_REQUIRE(k_rs % 8 == 0 && v_rs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(K) & 15) == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0,
                APX_EINVAL, "attention: K/V rows must be 16-byte aligned");
    int p16 = d / 8 + 1;  // row pitch in 16-byte units, odd -> conflict-free 128-bit row reads
    if (p16 % 2 == 0) ++p16;
    SimtAttnArgs a{static_cast<const uint16_t*>(Q), q_bs, q_rs, static_cast<const uint16_t*>(K), k_bs, k_rs,
                   static_cast<const uint16_t*>(V), v_bs, v_rs, static_cast<uint16_t*>(O), o_bs, o_rs,
                   B, S, H, d, L, L0, w0, w1, 1.0f / sqrtf(static_cast<float>(d)), dtype == APX_F16 ? 0 : 1, p16 * 8};
    const size_t smem = static_cast<size_t>(L) * a.pitch * 2 * 2 + 4 * d * sizeof(float);
    APX_REQUIRE(smem <= 227 * 1024, APX_EUNSUPPORTED, "attention: L=%d d=%d needs %zu B of shared memory", L, d, smem);
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask)) {
        APX_CUDA(cudaFuncSetAttribute(attention_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    dim3 grid(ceil_div(S, QPB), H, B);
    attention_simt_kernel<<<grid, 128, smem, st>>>(a);
    APX_LAUNCH_CHECK("attention_simt_kernel");
    return APX_OK;
}

int patch_embed(const float* mel, const float* pwT, const float* pb, const float* pos, const float* cls, float* x,
                int B, cudaStream_t st) {
    APX_REQUIRE(B > 0, APX_EINVAL, "patch_embed: B=%d", B);
    patch_embed_kernel<<<dim3(64, B), 256, 0, st>>>(mel, pwT, pb, pos, cls, x);
    APX_LAUNCH_CHECK("patch_embed_kernel");
    return APX_OK;
}

int patchify(const float* mel, const float* pos, const float* cls, const float* pb, void* a16, float* res, int B, int fmt, cudaStream_t st) {
    APX_REQUIRE(B > 0 && B <= 65535, APX_EINVAL, "patchify: B=%d", B);
    patchify_kernel<<<dim3(64, B), 256, 0, st>>>(mel, pos, cls, pb, static_cast<uint16_t*>(a16), res, fmt);
    APX_LAUNCH_CHECK("patchify_kernel");
    return APX_OK;
}
int patch_weight16(const float* pwT, void* w16, int fmt, cudaStream_t st) {
    patch_weight16_kernel<<<768 * 128 / 256, 256, 0, st>>>(pwT, static_cast<uint16_t*>(w16), fmt);
    APX_LAUNCH_CHECK("patch_weight16_kernel");
    return APX_OK;
}

int layernorm(const float* x, const float* w, const float* b, void* y, int rows, int cols, float eps, int out_dtype,
              cudaStream_t st) {
    APX_REQUIRE(cols % 128 == 0 && cols <= 1024 && cols > 0, APX_EUNSUPPORTED, "layernorm: cols=%d", cols);
    APX_REQUIRE(rows > 0, APX_EINVAL, "layernorm: rows=%d", rows);
    layernorm_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(x, w, b, y, rows, cols, eps, out_dtype);
    APX_LAUNCH_CHECK("layernorm_kernel");
    return APX_OK;
}

int pool_tokens(const float* xn, float* emb, int B, int tp, int fp, cudaStream_t st) {
    APX_REQUIRE(tp >= 1 && tp <= 64 && 64 % tp == 0, APX_EINVAL, "pool: time_pool=%d must divide 64", tp);
    APX_REQUIRE(fp >= 1 && fp <= 8 && 8 % fp == 0, APX_EINVAL, "pool: freq_pool=%d must divide 8", fp);
    const int ntok = (64 / tp) * (8 / fp);
    pool_tokens_kernel<<<dim3(ntok, B), 192, 0, st>>>(xn, emb, tp, fp);
    APX_LAUNCH_CHECK("pool_tokens_kernel");
    return APX_OK;
}

int fbank(const float* wave16k, int B, int n_samples, const float* melT, float* out, int target_len, float* scratch, cudaStream_t st) {
    APX_REQUIRE(n_samples >= 400, APX_EINVAL, "fbank: %d samples (one 25 ms frame needs 400)", n_samples);
    APX_REQUIRE(target_len > 0 && B > 0 && B <= 65535, APX_EINVAL, "fbank: target_len=%d B=%d", target_len, B);
    const int n_frames = 1 + (n_samples - 400) / 160;   // snip_edges
    wave_mean_kernel<<<B, 256, 0, st>>>(wave16k, n_samples, scratch);
    APX_LAUNCH_CHECK("wave_mean_kernel");
    fbank_kernel<<<dim3(target_len, B), 256, 0, st>>>(wave16k, static_cast<size_t>(n_samples), std::min(n_frames, target_len), scratch, melT, out,
                                                      -4.2677393f, 2.0f * 4.5689974f);
    APX_LAUNCH_CHECK("fbank_kernel");
    return APX_OK;
}

// Polyphase FIR resampler (audio_encoder/AudioMAE.py:360-366 -> torchaudio.functional.resample, sinc interpolation with a Hann
// window): with orig / new reduced by their gcd, output sample j * new + p of a clip is  sum_t taps[p][t] * x[j * orig + t - width]
// (x zero outside the clip), for j * new + p < ceil(new * n / orig).  One thread per output sample, taps [new][kw] from L2.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, int n, const float* __restrict__ taps, int orig, int nw,
                                                      int width, int kw, float* __restrict__ y, int n_out) {
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= n_out) return;
    x += static_cast<size_t>(blockIdx.y) * n;
    y += static_cast<size_t>(blockIdx.y) * n_out;
    const int j = o / nw, p = o - j * nw;
    const float* tp = taps + static_cast<size_t>(p) * kw;
    const int i0 = j * orig - width;
    const int t_lo = max(0, -i0), t_hi = min(kw, n - i0);
    float acc = 0.f;
    for (int t = t_lo; t < t_hi; ++t) acc = fmaf(tp[t], x[i0 + t], acc);
    y[o] = acc;
}

int resample(const float* x, int B, int n, const float* taps, int orig, int nw, int width, float* y, cudaStream_t st) {
    APX_REQUIRE(B > 0 && B <= 65535 && n > 0 && orig > 0 && nw > 0 && width >= 0, APX_EINVAL, "resample: bad argument");
    const int kw = 2 * width + orig;
    const long long n_out = (static_cast<long long>(nw) * n + orig - 1) / orig;
    APX_REQUIRE(n_out > 0 && n_out < (1ll << 31), APX_EINVAL, "resample: output length %lld", n_out);
    resample_kernel<<<dim3(static_cast<unsigned>((n_out + 255) / 256), B), 256, 0, st>>>(x, n, taps, orig, nw, width, kw, y, static_cast<int>(n_out));
    APX_LAUNCH_CHECK("resample_kernel");
    return APX_OK;
}

}  // namespace apx
