// apx_norm.cu — the normalisation / elementwise neighbours of the adapter sites inside the AudioLDM2 UNet blocks (SURVEY.md §8f-2:
// pipeline/modeling_audioldm2.py:1140-1156, 1317-1333, 1492-1508 call ResnetBlock2D / Transformer2DModel around every attn2 site).
// HBM-bound CUDA-core kernels on channels-last (NHWC) 16-bit activations:
//
//   groupnorm_nhwc   y = [SiLU] GroupNorm(x + chan_bias[c] + bc_add[b, c])   one kernel for what the stock graph runs as conv-bias add,
//                    time-embedding add, three GroupNorm kernels and SiLU; reads and writes NHWC, so the cuDNN convolutions
//                    around it take their native layout and the token view [B, H*W, C] of a Transformer2DModel is free.
//   add_bias_nhwc    y = (a + b + bias[c]) * scale                           the ResnetBlock2D output: shortcut + conv2 + conv biases.
#include "apx_internal.h"
#include "apx_ptx.cuh"
#include <algorithm>
#include <cstdlib>

namespace apx {
namespace {

struct GnArgs {
    const uint16_t* x;        // [B, HW, C] (NHWC)
    uint16_t* y;              // [B, HW, C]
    const float* gamma;       // [C]
    const float* beta;        // [C]
    const float* chan_bias;   // [C] or null: added to x before the statistics (the bias of the convolution that produced x)
    const uint16_t* bc_add;   // [B, C] 16-bit (row stride bc_stride elements) or null: added to x before the statistics (time-embedding projection)
    int64_t bc_stride;
    int B, HW, C, G, cpg;
    float eps;
    int silu, fmt;
};

__device__ __forceinline__ float block_sum(float v, float* red) {   // blockDim.x <= 1024, every thread gets the total
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();                      // red may still be read from the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = lane < nw ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// One CLUSTER of `cs` CTAs per (group, batch element); CTA r of the cluster owns the pixel slice [r * HW / cs, (r + 1) * HW / cs).
// Its HW_r x cpg values are staged ONCE in shared memory as fp32 (with the two adds applied): global memory is read once and
// written once.  Statistics: sum and sum of squares of (v - shift), shift = the group's first value (exact enough in fp32: the
// shifted values are O(sigma)), block-reduced and exchanged between the CTAs of the cluster through distributed shared memory.
// Loads are issued four at a time per thread: with one 8-byte load in flight a thread only sees the memory latency.
__device__ __forceinline__ void gn_load4(const GnArgs& a, const uint16_t* xg, int b, int g, int q4, int item, float (&v)[4]) {
    const int p = item / q4, q = item - p * q4;
    const uint2 raw = *reinterpret_cast<const uint2*>(xg + static_cast<size_t>(p) * a.C + q * 4);
    unpack16(raw.x, a.fmt, v[0], v[1]);
    unpack16(raw.y, a.fmt, v[2], v[3]);
    const int c = g * a.cpg + q * 4;
    if (a.chan_bias) {
        const float4 cb = *reinterpret_cast<const float4*>(a.chan_bias + c);
        v[0] += cb.x; v[1] += cb.y; v[2] += cb.z; v[3] += cb.w;
    }
    if (a.bc_add) {
        const uint2 t = *reinterpret_cast<const uint2*>(a.bc_add + static_cast<size_t>(b) * a.bc_stride + c);
        float t0, t1, t2, t3;
        unpack16(t.x, a.fmt, t0, t1);
        unpack16(t.y, a.fmt, t2, t3);
        v[0] += t0; v[1] += t1; v[2] += t2; v[3] += t3;
    }
}

__global__ void __launch_bounds__(512) groupnorm_nhwc_kernel(GnArgs a, int cs) {
    extern __shared__ float sv[];           // [HW_r][cpg]
    __shared__ float red[32];
    __shared__ float part[8][2];            // partial (sum, sum of squares) of every CTA of the cluster, written by their owners
    const int g = blockIdx.x / cs, r = blockIdx.x - g * cs, b = blockIdx.y;
    const int q4 = a.cpg / 4;               // 4-channel (8-byte) pieces per pixel of this group
    const int p_lo = static_cast<int>(static_cast<long long>(a.HW) * r / cs), p_hi = static_cast<int>(static_cast<long long>(a.HW) * (r + 1) / cs);
    const int n4 = (p_hi - p_lo) * q4;
    const uint16_t* xg = a.x + (static_cast<size_t>(b) * a.HW + p_lo) * a.C + g * a.cpg;
    uint16_t* yg = a.y + (static_cast<size_t>(b) * a.HW + p_lo) * a.C + g * a.cpg;
    float shift;
    {   // the group's first value (pixel 0, first channel) with its adds: every CTA of the cluster computes the same number
        float v0[4];
        gn_load4(a, a.x + static_cast<size_t>(b) * a.HW * a.C + g * a.cpg, b, g, q4, 0, v0);
        shift = v0[0];
    }
    float s1 = 0.f, s2 = 0.f;
    const int step = blockDim.x;
    for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * step) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * step < n4) gn_load4(a, xg, b, g, q4, i0 + u * step, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * step < n4) {
                *reinterpret_cast<float4*>(sv + (i0 + u * step) * 4) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d = v[u][k] - shift;
                    s1 += d;
                    s2 = fmaf(d, d, s2);
                }
            }
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (cs > 1) {
        if (threadIdx.x < cs) {             // thread t publishes this CTA's partials in CTA t's table
            const uint32_t dst = map_to_cta(smem_u32(&part[r][0]), threadIdx.x);
            asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(s1), "f"(s2) : "memory");
        }
        cluster_sync_all();
        s1 = 0.f; s2 = 0.f;
        for (int i = 0; i < cs; ++i) { s1 += part[i][0]; s2 += part[i][1]; }
    }
    const float inv_n = 1.0f / (static_cast<float>(a.HW) * a.cpg);
    const float m = s1 * inv_n;
    const float mean = shift + m;
    const float rstd = rsqrtf(fmaxf(s2 * inv_n - m * m, 0.f) + a.eps);
    for (int i = threadIdx.x; i < n4; i += step) {
        const int p = i / q4, q = i - p * q4;
        const int c = g * a.cpg + q * 4;
        const float4 v = *reinterpret_cast<const float4*>(sv + i * 4);
        const float4 ga = *reinterpret_cast<const float4*>(a.gamma + c), be = *reinterpret_cast<const float4*>(a.beta + c);
        float o[4] = {(v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y, (v.z - mean) * rstd * ga.z + be.z,
                      (v.w - mean) * rstd * ga.w + be.w};
        if (a.silu) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = o[k] / (1.0f + __expf(-o[k]));
        }
        uint2 out;
        out.x = pack16(o[0], o[1], a.fmt);
        out.y = pack16(o[2], o[3], a.fmt);
        *reinterpret_cast<uint2*>(yg + static_cast<size_t>(p) * a.C + q * 4) = out;
    }
}

// Second form (round 2, second half): the same decomposition, with the latency chain of a small tensor cut down — most of the 109
// calls of a denoise step normalise 40 KB ... 1 MB and ran 7 - 8 us, of which the dependent global round trips were the bulk:
//   * the per-channel constants of the group (conv bias + time-embedding row, gamma, beta: cpg <= 64 values each) are loaded ONCE into
//     shared memory while the first batch of x loads is in flight, instead of three extra global loads per 4-channel item;
//   * no dependent "first value" load for the shift: the statistics are two passes over the values already staged in shared memory
//     (local mean, then local sum of squared deviations), merged across the cluster with Chan's update — one exchange, and better
//     conditioned than the shifted single pass.
__global__ void __launch_bounds__(512) groupnorm_nhwc_v2_kernel(GnArgs a, int cs) {
    extern __shared__ float sv[];           // [HW_r][cpg]
    __shared__ float red[32];
    __shared__ float4 part[8];              // (mean, M2, count) of every CTA of the cluster, written by their owners
    __shared__ float s_add[64], s_ga[64], s_be[64];
    const int g = blockIdx.x / cs, r = blockIdx.x - g * cs, b = blockIdx.y;
    const int q4 = a.cpg / 4;               // 4-channel (8-byte) pieces per pixel of this group
    const int p_lo = static_cast<int>(static_cast<uint32_t>(a.HW) * r / cs), p_hi = static_cast<int>(static_cast<uint32_t>(a.HW) * (r + 1) / cs);   // HW * 8 fits 32 bits
    const int n4 = (p_hi - p_lo) * q4;
    const uint16_t* xg = a.x + (static_cast<size_t>(b) * a.HW + p_lo) * a.C + g * a.cpg;
    uint16_t* yg = a.y + (static_cast<size_t>(b) * a.HW + p_lo) * a.C + g * a.cpg;
    const int step = blockDim.x;
    // first batch of x loads (up to four 8-byte pieces per thread) goes out before anything else
    uint2 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * step;
        raw[u] = make_uint2(0, 0);
        if (i < n4) { const int p = i / q4, q = i - p * q4; raw[u] = *reinterpret_cast<const uint2*>(xg + static_cast<size_t>(p) * a.C + q * 4); }
    }
    if (threadIdx.x < a.cpg) {
        const int c = g * a.cpg + threadIdx.x;
        float add = a.chan_bias ? a.chan_bias[c] : 0.f;
        if (a.bc_add) add += ld16f(a.bc_add, static_cast<size_t>(b) * a.bc_stride + c, a.fmt);
        s_add[threadIdx.x] = add;
        s_ga[threadIdx.x] = a.gamma[c];
        s_be[threadIdx.x] = a.beta[c];
    }
    __syncthreads();
    float s1 = 0.f;
    auto stage = [&](int i, uint2 w) {
        const int q = i % q4;
        float v0, v1, v2, v3;
        unpack16(w.x, a.fmt, v0, v1);
        unpack16(w.y, a.fmt, v2, v3);
        const float4 ad = *reinterpret_cast<const float4*>(s_add + q * 4);
        v0 += ad.x; v1 += ad.y; v2 += ad.z; v3 += ad.w;
        *reinterpret_cast<float4*>(sv + i * 4) = make_float4(v0, v1, v2, v3);
        s1 += (v0 + v1) + (v2 + v3);
    };
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (threadIdx.x + u * step < n4) stage(threadIdx.x + u * step, raw[u]);
    for (int i0 = threadIdx.x + 4 * step; i0 < n4; i0 += 4 * step) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * step;
            if (i < n4) { const int p = i / q4, q = i - p * q4; raw[u] = *reinterpret_cast<const uint2*>(xg + static_cast<size_t>(p) * a.C + q * 4); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * step < n4) stage(i0 + u * step, raw[u]);
    }
    // local statistics: mean of this CTA's values, then the sum of squared deviations from it (every thread re-reads its own items)
    const float n_loc = static_cast<float>(n4) * 4.f;
    const float mean_loc = block_sum(s1, red) / fmaxf(n_loc, 1.f);
    float m2 = 0.f;
    for (int i = threadIdx.x; i < n4; i += step) {
        const float4 v = *reinterpret_cast<const float4*>(sv + i * 4);
        const float d0 = v.x - mean_loc, d1 = v.y - mean_loc, d2 = v.z - mean_loc, d3 = v.w - mean_loc;
        m2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, m2))));
    }
    m2 = block_sum(m2, red);
    float mean = mean_loc;
    if (cs > 1) {
        if (threadIdx.x < cs) {             // thread t publishes this CTA's (mean, M2, count) in CTA t's table
            const uint32_t dst = map_to_cta(smem_u32(&part[r]), threadIdx.x);
            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(mean_loc), "f"(m2), "f"(n_loc), "f"(0.f) : "memory");
        }
        cluster_sync_all();
        // Chan's merge over the CTAs of the cluster
        float tot = 0.f, msum = 0.f;
        for (int i = 0; i < cs; ++i) { tot += part[i].z; msum += part[i].z * part[i].x; }
        mean = msum / tot;
        m2 = 0.f;
        for (int i = 0; i < cs; ++i) {
            const float d = part[i].x - mean;
            m2 += part[i].y + part[i].z * d * d;
        }
    }
    const float rstd = rsqrtf(m2 / (static_cast<float>(a.HW) * a.cpg) + a.eps);
    for (int i = threadIdx.x; i < n4; i += step) {
        const int p = i / q4, q = i - p * q4;
        const float4 v = *reinterpret_cast<const float4*>(sv + i * 4);
        const float4 ga = *reinterpret_cast<const float4*>(s_ga + q * 4), be = *reinterpret_cast<const float4*>(s_be + q * 4);
        float o[4] = {(v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y, (v.z - mean) * rstd * ga.z + be.z,
                      (v.w - mean) * rstd * ga.w + be.w};
        if (a.silu) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = o[k] / (1.0f + __expf(-o[k]));
        }
        uint2 out;
        out.x = pack16(o[0], o[1], a.fmt);
        out.y = pack16(o[2], o[3], a.fmt);
        *reinterpret_cast<uint2*>(yg + static_cast<size_t>(p) * a.C + q * 4) = out;
    }
}

// y = (a + b + bias[c]) * scale over [rows, C] 16-bit, 8 channels (16 bytes) per thread
__global__ void __launch_bounds__(256) add_bias_nhwc_kernel(const uint4* __restrict__ A, const uint4* __restrict__ Bm, const float* __restrict__ bias,
                                                           uint4* __restrict__ Y, size_t n8, int C8, float scale, int fmt) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 a = A[i], b = Bm[i];
        const int c = static_cast<int>(i % C8) * 8;
        float bs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + c), b1 = *reinterpret_cast<const float4*>(bias + c + 4);
            bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
        }
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a0, a1, b0, b1;
            unpack16(aw[k], fmt, a0, a1);
            unpack16(bw[k], fmt, b0, b1);
            ow[k] = pack16((a0 + b0 + bs[2 * k]) * scale, (a1 + b1 + bs[2 * k + 1]) * scale, fmt);
        }
        Y[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// h = a + b (16-bit, stored when b is given) and x = LayerNorm(h) * gamma + beta in ONE pass: the residual add of a transformer block and
// the LayerNorm of the next sub-block (diffusers' BasicTransformerBlock: "hidden_states = attn_output + hidden_states;
// norm_hidden_states = self.norm2(hidden_states)"), which the stock graph runs as two ATen kernels of 1.8 + 3.5 us at the UNet's small
// sites (128 x 640, 504 x 384).  One warp per row, every load of the row (a, b, gamma, beta) in flight before the first use, statistics in
// fp32 over the ROUNDED sum (what LayerNorm would read back), two exact passes in registers.  NP = 16-byte pieces per lane.
template <int NP>
__global__ void __launch_bounds__(256) add_layernorm16_kernel(const uint4* __restrict__ A, const uint4* __restrict__ Bm, const void* __restrict__ gamma,
                                                             const void* __restrict__ beta, int w_f32, uint4* __restrict__ Ysum,
                                                             uint4* __restrict__ Yln, int rows, int C8, float eps, int fmt) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // blockDim.x / 32 rows per CTA
    if (row >= rows) return;
    const size_t base = static_cast<size_t>(row) * C8;
    uint4 a[NP], b[NP];
    float g[NP][8], be[NP][8];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int c8 = i * 32 + lane;
        a[i] = make_uint4(0, 0, 0, 0); b[i] = make_uint4(0, 0, 0, 0);
        if (c8 < C8) {
            a[i] = A[base + c8];
            if (Bm) b[i] = Bm[base + c8];
        }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int c8 = i * 32 + lane;
        if (c8 < C8) {
            if (w_f32) {
                const float4 g0 = reinterpret_cast<const float4*>(gamma)[2 * c8], g1 = reinterpret_cast<const float4*>(gamma)[2 * c8 + 1];
                const float4 b0 = reinterpret_cast<const float4*>(beta)[2 * c8], b1 = reinterpret_cast<const float4*>(beta)[2 * c8 + 1];
                g[i][0] = g0.x; g[i][1] = g0.y; g[i][2] = g0.z; g[i][3] = g0.w; g[i][4] = g1.x; g[i][5] = g1.y; g[i][6] = g1.z; g[i][7] = g1.w;
                be[i][0] = b0.x; be[i][1] = b0.y; be[i][2] = b0.z; be[i][3] = b0.w; be[i][4] = b1.x; be[i][5] = b1.y; be[i][6] = b1.z; be[i][7] = b1.w;
            } else {
                const uint4 gw = reinterpret_cast<const uint4*>(gamma)[c8], bw = reinterpret_cast<const uint4*>(beta)[c8];
                unpack16(gw.x, fmt, g[i][0], g[i][1]); unpack16(gw.y, fmt, g[i][2], g[i][3]);
                unpack16(gw.z, fmt, g[i][4], g[i][5]); unpack16(gw.w, fmt, g[i][6], g[i][7]);
                unpack16(bw.x, fmt, be[i][0], be[i][1]); unpack16(bw.y, fmt, be[i][2], be[i][3]);
                unpack16(bw.z, fmt, be[i][4], be[i][5]); unpack16(bw.w, fmt, be[i][6], be[i][7]);
            }
        }
    }
    float v[NP][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int c8 = i * 32 + lane;
        const uint32_t aw[4] = {a[i].x, a[i].y, a[i].z, a[i].w}, bw[4] = {b[i].x, b[i].y, b[i].z, b[i].w};
        uint32_t hw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a0, a1, b0, b1;
            unpack16(aw[k], fmt, a0, a1);
            unpack16(bw[k], fmt, b0, b1);
            hw[k] = pack16(a0 + b0, a1 + b1, fmt);              // the 16-bit residual stream, rounded once
            unpack16(hw[k], fmt, v[i][2 * k], v[i][2 * k + 1]);  // LayerNorm reads the rounded value back
        }
        if (c8 < C8) {
            if (Bm) Ysum[base + c8] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[i][k];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float n = static_cast<float>(C8) * 8.f, mean = s / n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i * 32 + lane < C8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; q = fmaf(d, d, q); }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / n + eps);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int c8 = i * 32 + lane;
        if (c8 < C8) {
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ow[k] = pack16((v[i][2 * k] - mean) * rstd * g[i][2 * k] + be[i][2 * k],
                               (v[i][2 * k + 1] - mean) * rstd * g[i][2 * k + 1] + be[i][2 * k + 1], fmt);
            Yln[base + c8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
    }
}

}  // namespace
}  // namespace apx

using namespace apx;

extern "C" {

int apx_groupnorm_nhwc(const void* x, const float* gamma, const float* beta, const float* chan_bias, const void* bc_add, int64_t bc_stride,
                       void* y, int B, int HW, int C, int groups, float eps, int silu, int dtype, void* stream) {
    APX_REQUIRE(x && gamma && beta && y, APX_EINVAL, "apx_groupnorm_nhwc: null pointer");
    APX_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0, APX_EINVAL, "apx_groupnorm_nhwc: bad shape B=%d HW=%d C=%d G=%d", B, HW, C,
                groups);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "apx_groupnorm_nhwc: dtype %d", dtype);
    const int cpg = C / groups;
    APX_REQUIRE(cpg % 4 == 0, APX_EUNSUPPORTED, "apx_groupnorm_nhwc: %d channels per group (need a multiple of 4)", cpg);
    APX_REQUIRE(static_cast<size_t>(HW) * cpg * 4 <= 8 * 190 * 1024, APX_EUNSUPPORTED,
                "apx_groupnorm_nhwc: a group of %d x %d values does not fit the shared memory of a cluster of 8 CTAs", HW, cpg);
    APX_REQUIRE(B <= 65535, APX_EUNSUPPORTED, "apx_groupnorm_nhwc: B=%d", B);
    APX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(beta) & 15) == 0 && (reinterpret_cast<uintptr_t>(chan_bias) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bc_add) & 7) == 0 && (bc_add == nullptr || (bc_stride >= C && bc_stride % 4 == 0)),
                APX_EINVAL, "apx_groupnorm_nhwc: misaligned pointer / bad bc_add stride");
    static std::atomic<uint64_t> attr_mask{0};
    if (first_use_on_device(attr_mask))
    {
        APX_CUDA(cudaFuncSetAttribute(groupnorm_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        APX_CUDA(cudaFuncSetAttribute(groupnorm_nhwc_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    GnArgs a{static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), gamma, beta, chan_bias, static_cast<const uint16_t*>(bc_add), bc_stride,
             B, HW, C, groups, cpg, eps, silu ? 1 : 0, dtype == APX_F16 ? 0 : 1};
    // a group is split over a cluster of 1 / 2 / 4 / 8 CTAs so that a CTA owns ~<= 2048 four-channel items (8 per thread at 256 threads):
    // with one CTA per (group, clip) the B = 1 step ran 64 CTAs of up to 32 k values each, 26 us for the largest tensor
    const int n4 = HW * (cpg / 4);
    int cs = 1;
    while (cs < 8 && (n4 / cs > 2048 || static_cast<size_t>(HW) * cpg * 4 / cs > 200 * 1024) && HW / (2 * cs) >= 1) cs *= 2;
    const int per = (n4 + cs - 1) / cs;
    static const bool v1 = [] { const char* e = getenv("APX_GN_V1"); return e && e[0] == '1'; }();
    const bool use_v1 = v1 || cpg > 64;
    // second form: at most four 8-byte items per thread, so that every x load of a CTA is in flight at once
    const int threads = use_v1 ? (per >= 1024 ? 256 : per >= 256 ? 128 : 64) : (per > 1024 ? 512 : per > 512 ? 256 : per > 256 ? 128 : 64);
    const size_t smem = (static_cast<size_t>((HW + cs - 1) / cs) + 1) * cpg * 4;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(groups * cs), B);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = static_cast<unsigned>(cs);
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = cs > 1 ? 1 : 0;
    if (use_v1) APX_CUDA(cudaLaunchKernelEx(&cfg, groupnorm_nhwc_kernel, a, cs));
    else APX_CUDA(cudaLaunchKernelEx(&cfg, groupnorm_nhwc_v2_kernel, a, cs));
    APX_LAUNCH_CHECK("groupnorm_nhwc_kernel");
    return APX_OK;
}

int apx_add_bias_nhwc(const void* a, const void* b, const float* bias, void* y, int64_t rows, int C, float scale, int dtype, void* stream) {
    APX_REQUIRE(a && b && y, APX_EINVAL, "apx_add_bias_nhwc: null pointer");
    APX_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, APX_EINVAL, "apx_add_bias_nhwc: bad shape rows=%lld C=%d (C %% 8 == 0)", static_cast<long long>(rows), C);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "apx_add_bias_nhwc: dtype %d", dtype);
    APX_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                APX_EINVAL, "apx_add_bias_nhwc: misaligned pointer");
    const size_t n8 = static_cast<size_t>(rows) * C / 8;
    const int blocks = static_cast<int>(std::min<size_t>((n8 + 255) / 256, 148 * 8));
    add_bias_nhwc_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b), bias,
                                                                               static_cast<uint4*>(y), n8, C / 8, scale, dtype == APX_F16 ? 0 : 1);
    APX_LAUNCH_CHECK("add_bias_nhwc_kernel");
    return APX_OK;
}

int apx_add_layernorm(const void* a, const void* b, const void* gamma, const void* beta, int w_dtype, void* y_sum, void* y_ln, int64_t rows, int C,
                      float eps, int dtype, void* stream) {
    APX_REQUIRE(a && gamma && beta && y_ln && ((b == nullptr) == (y_sum == nullptr)), APX_EINVAL,
                "apx_add_layernorm: null pointer (b and y_sum come together)");
    APX_REQUIRE(rows > 0 && rows <= (1ll << 30) && C > 0 && C % 8 == 0 && C <= 2048, APX_EUNSUPPORTED,
                "apx_add_layernorm: rows=%lld C=%d (C %% 8 == 0, C <= 2048)", static_cast<long long>(rows), C);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "apx_add_layernorm: dtype %d", dtype);
    APX_REQUIRE(w_dtype == dtype || w_dtype == APX_F32, APX_EUNSUPPORTED, "apx_add_layernorm: gamma / beta dtype %d (the activations' or fp32)", w_dtype);
    APX_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                  reinterpret_cast<uintptr_t>(y_sum) | reinterpret_cast<uintptr_t>(y_ln)) & 15) == 0,
                APX_EINVAL, "apx_add_layernorm: pointers must be 16-byte aligned");
    const int C8 = C / 8, np = (C8 + 31) / 32;
    const int fmt = dtype == APX_F16 ? 0 : 1, wf = w_dtype == APX_F32 ? 1 : 0;
    const int rpc = 8;   // rows (warps) per CTA; 1 / 2 / 8 measure the same in the step (8.346 / 8.343 / 8.361 ms, profiles/r3p_bench_b1_addln_rows_*.json)
    const dim3 grid(static_cast<unsigned>((rows + rpc - 1) / rpc));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint4 *A = static_cast<const uint4*>(a), *Bm = static_cast<const uint4*>(b);
    uint4 *Ys = static_cast<uint4*>(y_sum), *Yl = static_cast<uint4*>(y_ln);
    const int r = static_cast<int>(rows);
    switch (np) {
        case 1: add_layernorm16_kernel<1><<<grid, 32 * rpc, 0, st>>>(A, Bm, gamma, beta, wf, Ys, Yl, r, C8, eps, fmt); break;
        case 2: add_layernorm16_kernel<2><<<grid, 32 * rpc, 0, st>>>(A, Bm, gamma, beta, wf, Ys, Yl, r, C8, eps, fmt); break;
        case 3: add_layernorm16_kernel<3><<<grid, 32 * rpc, 0, st>>>(A, Bm, gamma, beta, wf, Ys, Yl, r, C8, eps, fmt); break;
        case 4: add_layernorm16_kernel<4><<<grid, 32 * rpc, 0, st>>>(A, Bm, gamma, beta, wf, Ys, Yl, r, C8, eps, fmt); break;
        default: add_layernorm16_kernel<8><<<grid, 32 * rpc, 0, st>>>(A, Bm, gamma, beta, wf, Ys, Yl, r, C8, eps, fmt); break;
    }
    APX_LAUNCH_CHECK("add_layernorm16_kernel");
    return APX_OK;
}

}  // extern "C"
