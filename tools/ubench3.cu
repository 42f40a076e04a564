// ubench3.cu — does tcgen05.mma dispatch scale with the number of ISSUING warps?  (test infrastructure)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ap_adapter_b200/csrc -o tools/ubench3 tools/ubench3.cu
// One CTA; W in {1, 2, 4} warps each issue n MMAs (128 x N x 16, fp16) into their own TMEM columns from one elected lane and
// commit to their own mbarrier; the time from a common start to the LAST completion is reported.  If the ~65-cycle
// per-MMA floor seen with a single issuer (profiles/r1_ubench_tcgen05_latencies.txt) is a per-thread issue latency,
// W issuers finish W x n MMAs in about the time of n; if it is a property of the tensor pipe, in W times that.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "apx_ptx.cuh"

using namespace apx;

struct Res { long long t[3][4][3]; };   // [N index][issuer-count index][n index]

__global__ void __launch_bounds__(160, 1) ubench3_kernel(Res* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;              // 128 x 64 x 2 = 16 KB
    uint8_t* sB = smem + 16384;      // 256 x 64 x 2 = 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);   // 4 barriers
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
    __shared__ long long t_done[4];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1); fence_mbar_init(); }
    if (warp == 4) tmem_alloc<512>(slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    uint32_t phase = 0;
    const int Ns[3] = {32, 48, 96};
    const int Ws[4] = {1, 2, 4, 1};     // index 3: ONE issuer doing the work of 4 (the serial reference)
    const int ns[3] = {4, 9, 16};
    for (int ni = 0; ni < 3; ++ni)
        for (int wi = 0; wi < 4; ++wi)
            for (int ci = 0; ci < 3; ++ci) {
                const int N = Ns[ni], W = Ws[wi];
                const int n = wi == 3 ? 4 * ns[ci] : ns[ci];
                long long best = 1ll << 60;
                for (int rep = 0; rep < 4; ++rep) {
                    __syncthreads();
                    const long long t0 = clock64();
                    if (warp < W) {
                        if (elect_one()) {
                            const uint32_t idesc = umma_idesc_f16(0, 128, N);
                            for (int k = 0; k < n; ++k) {
                                const uint64_t da = umma_desc(smem_u32(sA) + (k & 3) * 2 * 2048, 2048, 128, kLayoutNone);
                                const uint64_t db = umma_desc(smem_u32(sB) + (k & 3) * 2 * (N * 16), N * 16, 128, kLayoutNone);
                                umma_ss(tmem + warp * 128, da, db, idesc, k > 0);
                            }
                            umma_commit(bar + warp);
                            while (!mbar_try_wait(bar + warp, phase)) {}
                            t_done[warp] = clock64();
                        }
                        __syncwarp();
                    }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        long long t1 = 0;
                        for (int w = 0; w < W; ++w) t1 = t_done[w] > t1 ? t_done[w] : t1;
                        if (t1 - t0 < best) best = t1 - t0;
                    }
                    // every barrier that was used flips phase; unused ones must be flipped too to stay in step
                    if (threadIdx.x == 0)
                        for (int w = W; w < 4; ++w) mbar_arrive(bar + w);
                    __syncthreads();
                    phase ^= 1;
                }
                if (threadIdx.x == 0) out->t[ni][wi][ci] = best;
            }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem);
}

int main() {
    Res* d; Res h;
    cudaMalloc(&d, sizeof(Res));
    cudaFuncSetAttribute(ubench3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    ubench3_kernel<<<1, 160, 64 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    const int Ns[3] = {32, 48, 96}, ns[3] = {4, 9, 16};
    const char* names[4] = {"1 issuer  x n", "2 issuers x n", "4 issuers x n", "1 issuer  x 4n"};
    for (int ni = 0; ni < 3; ++ni)
        for (int wi = 0; wi < 4; ++wi) {
            printf("mma 128x%dx16 %s:", Ns[ni], names[wi]);
            for (int ci = 0; ci < 3; ++ci) printf("  n=%2d -> %6lld cyc", ns[ci], h.t[ni][wi][ci]);
            printf("\n");
        }
    return 0;
}
