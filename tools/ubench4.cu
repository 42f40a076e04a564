// ubench4.cu — how fast can ONE warp per SM sub-partition feed the MUFU pipe?  (test infrastructure)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench4 tools/ubench4.cu
// The exp2 pass of the softmax warpgroups (apx_ipattn_site.cu, apx_flash_attn.cu) measures 11.7 cycles per ex2 with one warp on a
// sub-partition and 8.9 with two, against the pipe's 8 (16 lanes/clk/SM).  This bench times, per warp, 144 exponentials of a
// register-resident row in the forms below, with W = 1 or 2 warps per sub-partition (blockDim = 128 * W), one CTA per SM:
//   0  ex2 only (independent, back to back)
//   1  fma + ex2
//   2  fma + ex2 + cvt.f16x2 of each PAIR right behind it (what ptxas emits for exp_store16)
//   3  as 2 + st.shared.v4 per 8 values (the full exp_store16)
//   4  as 3, but the cvt / st of group c - 1 placed behind the ex2 of group c (true dependences force the order)
//   5  fma + cvt.f16x2 of the ARGUMENTS + ex2.approx.f16x2 (packed result, no cvt behind the MUFU) + st.shared
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { uint32_t d; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// 2^x for x <= 0 on the FMA / ALU pipes: round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-4 minimax polynomial for 2^f
// (max rel. error 4e-6), exponent added in the integer domain.  x is clamped at -126 (the result then underflows towards 2^-126).
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.f);
    const float t = x + 12582912.f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float f = x - (t - 12582912.f);
    float p = 9.6181291e-3f;
    p = fmaf(p, f, 5.5504109e-2f);
    p = fmaf(p, f, 2.4022651e-1f);
    p = fmaf(p, f, 6.9314718e-1f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// the same for a PAIR of arguments with packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): 8 FMA-pipe instructions per 2 results
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b)); return v; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void ex2_poly2(float s0, float s1, float sl, float neg_mx, float& y0, float& y1) {
    float x0, x1;
    upk(fma2(pk(s0, s1), pk(sl, sl), pk(neg_mx, neg_mx)), x0, x1);
    x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
    const uint64_t x = pk(x0, x1);
    const uint64_t t = add2(x, pk(12582912.f, 12582912.f));
    const uint64_t n = add2(t, pk(-12582912.f, -12582912.f));
    const uint64_t f = fma2(n, pk(-1.f, -1.f), x);
    uint64_t p = fma2(pk(9.6181291e-3f, 9.6181291e-3f), f, pk(5.5504109e-2f, 5.5504109e-2f));
    p = fma2(p, f, pk(2.4022651e-1f, 2.4022651e-1f));
    p = fma2(p, f, pk(6.9314718e-1f, 6.9314718e-1f));
    p = fma2(p, f, pk(1.f, 1.f));
    float p0, p1, t0, t1;
    upk(p, p0, p1); upk(t, t0, t1);
    y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
__device__ __forceinline__ void sts4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* in, float* out, long long* cycles, float sl, float mx) {
    extern __shared__ uint8_t smem[];
    float s[144];
#pragma unroll
    for (int i = 0; i < 144; ++i) s[i] = in[(i * 257 + threadIdx.x) & 4095];
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + threadIdx.x * 16;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 8; ++rep) {
        sl = *reinterpret_cast<const volatile float*>(in + 4000 + rep);   // a fresh scale every pass: nothing can be hoisted out of the loop
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 144; ++i) acc += ex2f(s[i] + sl);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 144; ++i) acc += ex2f(fmaf(s[i], sl, -mx));
        } else if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack2(ex2f(fmaf(s[c * 16 + 2 * i], sl, -mx)), ex2f(fmaf(s[c * 16 + 2 * i + 1], sl, -mx)));
                if (MODE == 3) { sts4(base + c * 8192, w[0], w[1], w[2], w[3]); sts4(base + c * 8192 + 4096, w[4], w[5], w[6], w[7]); }
                else acc += __uint_as_float(w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7]);
            }
        } else if (MODE == 4) {
            float ep[16];
#pragma unroll
            for (int c = 0; c <= 9; ++c) {
                float e[16];
                if (c < 9) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) e[i] = ex2f(fmaf(s[c * 16 + i], sl, -mx));
                }
                if (c > 0) {
                    if (c < 9) {
#pragma unroll
                        for (int i = 0; i < 16; i += 2) ep[i] = fmaf(e[15], 0.f, ep[i]);
                    }
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = pack2(ep[2 * i], ep[2 * i + 1]);
                    sts4(base + (c - 1) * 8192, w[0], w[1], w[2], w[3]); sts4(base + (c - 1) * 8192 + 4096, w[4], w[5], w[6], w[7]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) ep[i] = e[i];
            }
        } else if (MODE >= 10) {
            // packed-pair polynomial: of every group of 4 (MODE 10), 6 (11) or 8 (12) exponentials the last PAIR goes to the FMA pipe; 13: all pairs
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pi = c * 8 + i;   // pair index
                    const int per = MODE == 10 ? 2 : MODE == 11 ? 3 : MODE == 12 ? 4 : 1;
                    float y0, y1;
                    if (pi % per == per - 1) ex2_poly2(s[2 * pi], s[2 * pi + 1], sl, -mx, y0, y1);
                    else { y0 = ex2f(fmaf(s[2 * pi], sl, -mx)); y1 = ex2f(fmaf(s[2 * pi + 1], sl, -mx)); }
                    w[i] = pack2(y0, y1);
                }
                sts4(base + c * 8192, w[0], w[1], w[2], w[3]); sts4(base + c * 8192 + 4096, w[4], w[5], w[6], w[7]);
            }
        } else if (MODE >= 6) {
            // MODE 6: every 2nd exponential on the FMA pipe; 7: every 3rd; 8: every 4th; 9: all of them
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int e0 = c * 16 + 2 * i, e1 = e0 + 1;
                    const int per = MODE == 6 ? 2 : MODE == 7 ? 3 : MODE == 8 ? 4 : 1;
                    const float x0 = fmaf(s[e0], sl, -mx), x1 = fmaf(s[e1], sl, -mx);
                    const float y0 = (e0 % per == per - 1) ? ex2_poly(x0) : ex2f(x0);
                    const float y1 = (e1 % per == per - 1) ? ex2_poly(x1) : ex2f(x1);
                    w[i] = pack2(y0, y1);
                }
                sts4(base + c * 8192, w[0], w[1], w[2], w[3]); sts4(base + c * 8192 + 4096, w[4], w[5], w[6], w[7]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = ex2h2(pack2(fmaf(s[c * 16 + 2 * i], sl, -mx), fmaf(s[c * 16 + 2 * i + 1], sl, -mx)));
                sts4(base + c * 8192, w[0], w[1], w[2], w[3]); sts4(base + c * 8192 + 4096, w[4], w[5], w[6], w[7]);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) cycles[blockIdx.x * 8 + threadIdx.x / 32] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + s[0] + reinterpret_cast<float*>(smem)[threadIdx.x];
}

template <int MODE>
static void run(const char* what, const float* in, float* out, long long* cyc) {
    for (int W = 1; W <= 2; ++W) {
        cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        k<MODE><<<148, 128 * W, 100 * 1024>>>(in, out, cyc, 0.25f, 1.5f);
        k<MODE><<<148, 128 * W, 100 * 1024>>>(in, out, cyc, 0.25f, 1.5f);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return; }
        long long h[8];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mxc = 0;
        for (int i = 0; i < 4 * W; ++i) mxc = h[i] > mxc ? h[i] : mxc;
        printf("  %-62s W=%d warps/sub-partition: %6.0f cycles per 144-exp pass = %5.2f cycles per ex2 per warp (%5.2f per sub-partition)\n", what, W,
               mxc / 8.0, mxc / 8.0 / 144.0, mxc / 8.0 / 144.0 / W);
    }
}

int main() {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = -0.01f * (i % 977) - 1e-4f * i;
    for (int i = 4000; i < 4096; ++i) h[i] = 0.25f + 0.001f * (i - 4000);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("ubench4: MUFU.EX2 feed rate of one / two warps per SM sub-partition, 148 CTAs (one per SM)\n");
    run<0>("ex2 only", in, out, cyc);
    run<1>("fma + ex2", in, out, cyc);
    run<2>("fma + ex2 + cvt.f16x2 per pair", in, out, cyc);
    run<3>("fma + ex2 + cvt.f16x2 + st.shared.v4 (exp_store16)", in, out, cyc);
    run<4>("as above, cvt / st of group c-1 behind the ex2 of group c", in, out, cyc);
    run<5>("fma + cvt.f16x2 of the arguments + ex2.f16x2 + st.shared.v4", in, out, cyc);
    run<6>("exp_store16 with every 2nd ex2 as an FMA-pipe polynomial", in, out, cyc);
    run<7>("exp_store16 with every 3rd ex2 as an FMA-pipe polynomial", in, out, cyc);
    run<8>("exp_store16 with every 4th ex2 as an FMA-pipe polynomial", in, out, cyc);
    run<9>("exp_store16 with EVERY ex2 as an FMA-pipe polynomial", in, out, cyc);
    run<10>("exp_store16, every 2nd PAIR as a packed (f32x2) polynomial", in, out, cyc);
    run<11>("exp_store16, every 3rd PAIR as a packed (f32x2) polynomial", in, out, cyc);
    run<12>("exp_store16, every 4th PAIR as a packed (f32x2) polynomial", in, out, cyc);
    run<13>("exp_store16, EVERY pair as a packed (f32x2) polynomial", in, out, cyc);
    return 0;
}
