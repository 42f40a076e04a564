// ubench2.cu — tcgen05.ld and MUFU.EX2 throughput as the site kernel uses them (test infrastructure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ap_adapter_b200/csrc -o tools/ubench2 tools/ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "apx_ptx.cuh"
using namespace apx;

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, "
        "%46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
          "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
          "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
          "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
}

struct Res { long long t[16][8]; };

// mode: number of warps that run the experiment concurrently (1, 4 or 8); the others wait at the barrier
__global__ void __launch_bounds__(256, 1) k(Res* out, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    int e = 0;
    for (int nw : {1, 4, 8}) {
        // (A) 9 x ld16 + one wait
        __syncthreads();
        if (warp < nw) {
            uint32_t s[144];
            long long t0 = clock64();
#pragma unroll
            for (int c = 0; c < 9; ++c) tmem_ld16(tm + c * 16, s + c * 16);
            tmem_ld_wait();
            long long t1 = clock64();
#pragma unroll
            for (int i = 0; i < 144; ++i) acc += __uint_as_float(s[i]);
            if (lane == 0 && warp == 0) out->t[e][0] = t1 - t0;
        }
        // (B) 2 x ld64 + ld16 + one wait
        __syncthreads();
        if (warp < nw) {
            uint32_t s[144];
            long long t0 = clock64();
            ld64(tm, s); ld64(tm + 64, s + 64); tmem_ld16(tm + 128, s + 128);
            tmem_ld_wait();
            long long t1 = clock64();
#pragma unroll
            for (int i = 0; i < 144; ++i) acc += __uint_as_float(s[i]);
            if (lane == 0 && warp == 0) out->t[e][1] = t1 - t0;
        }
        // (C) 9 x (ld16 + wait) serial
        __syncthreads();
        if (warp < nw) {
            long long t0 = clock64();
#pragma unroll 1
            for (int c = 0; c < 9; ++c) { uint32_t r[16]; tmem_ld16(tm + c * 16, r); tmem_ld_wait(); acc += __uint_as_float(r[c]); }
            long long t1 = clock64();
            if (lane == 0 && warp == 0) out->t[e][2] = t1 - t0;
        }
        // (D) 144 dependent-free ex2 per thread
        __syncthreads();
        if (warp < nw) {
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = -0.01f * (lane + i);
            long long t0 = clock64();
#pragma unroll
            for (int rep = 0; rep < 9; ++rep) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = ex2_approx(x[i] - 1.0f);
            }
            long long t1 = clock64();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += x[i];
            if (lane == 0 && warp == 0) out->t[e][3] = t1 - t0;
        }
        // (E) 144 x (ffma + ex2 + fadd) + 72 cvt + 18 st.shared.v4: the exp pass without tensor memory
        __syncthreads();
        if (warp < nw) {
            __shared__ __align__(16) uint8_t buf[8][2 * 512];
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = -0.01f * (lane + i);
            float l = 0.f;
            long long t0 = clock64();
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                float ee[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) ee[i] = ex2_approx(fmaf(x[i], 0.25f, -0.5f * c));
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack2<0>(ee[2 * i], ee[2 * i + 1]);
                st_shared_v4(smem_u32(&buf[warp][0]) + lane * 16, w[0], w[1], w[2], w[3]);
                st_shared_v4(smem_u32(&buf[warp][0]) + 512 + lane * 16, w[4], w[5], w[6], w[7]);
                l += (((ee[0] + ee[1]) + (ee[2] + ee[3])) + ((ee[4] + ee[5]) + (ee[6] + ee[7]))) +
                     (((ee[8] + ee[9]) + (ee[10] + ee[11])) + ((ee[12] + ee[13]) + (ee[14] + ee[15])));
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] += 0.001f;
            }
            long long t1 = clock64();
            acc += l;
            if (lane == 0 && warp == 0) out->t[e][4] = t1 - t0;
        }
        ++e;
    }
    sink[threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
    Res* d; float* sink;
    cudaMalloc(&d, sizeof(Res)); cudaMalloc(&sink, 1024 * 4); cudaMemset(d, 0, sizeof(Res));
    for (int rep = 0; rep < 2; ++rep) k<<<1, 256>>>(d, sink);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    Res h; cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int nws[3] = {1, 4, 8};
    for (int e = 0; e < 3; ++e)
        printf("%d warp(s): 9xld16+wait %lld | 2xld64+ld16+wait %lld | 9x(ld16+wait) %lld | 144 ex2 %lld | exp pass (ffma+ex2+add+cvt+sts) %lld cyc\n",
               nws[e], h.t[e][0], h.t[e][1], h.t[e][2], h.t[e][3], h.t[e][4]);
    return 0;
}
