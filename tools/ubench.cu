// ubench.cu — sm_100a latency micro-benchmarks that the fused-kernel design depends on (test infrastructure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ap_adapter_b200/csrc -o tools/ubench tools/ubench.cu
// Prints cycles (SM clock) for:
//   mma   : issue n x tcgen05.mma (128 x N x 16) + tcgen05.commit -> mbarrier observed by the issuing thread
//   hand  : mbarrier.arrive by one warp -> try_wait success in another warp
//   ldtm  : tcgen05.ld x16 / x32 + wait::ld
//   fence : 2 x st.shared.v4 + fence.proxy.async
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "apx_ptx.cuh"

using namespace apx;

struct Res { long long x[8][8]; long long mma[2][4][4]; long long hand; long long ldtm16, ldtm32; long long fence; long long mma_wg[4]; };

__global__ void __launch_bounds__(192, 1) ubench_kernel(Res* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;              // 128 x 64 x 2 = 16 KB
    uint8_t* sB = smem + 16384;      // 256 x 64 x 2 = 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint64_t* bar2 = bar + 1;
    uint64_t* bar3 = bar + 2;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    volatile long long* t_arrive = reinterpret_cast<volatile long long*>(bar + 6);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar3, 128); fence_mbar_init(); }
    if (warp == 1) tmem_alloc<512>(slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    uint32_t phase = 0;
    if (warp == 1 && elect_one()) {
        const int Ns[4] = {16, 32, 144, 256};
        const int ns[4] = {1, 2, 4, 16};
        for (int layout = 0; layout < 2; ++layout)
            for (int ni = 0; ni < 4; ++ni)
                for (int ci = 0; ci < 4; ++ci) {
                    const int N = Ns[ni], n = ns[ci];
                    const uint32_t idesc = umma_idesc_f16(0, 128, N);
                    long long best = 1ll << 60;
                    for (int rep = 0; rep < 5; ++rep) {
                        const long long t0 = clock64();
                        for (int k = 0; k < n; ++k) {
                            uint64_t da, db;
                            if (layout == 0) {
                                da = umma_desc(smem_u32(sA) + (k & 3) * 2 * 2048, 2048, 128, kLayoutNone);
                                db = umma_desc(smem_u32(sB) + (k & 3) * 2 * (N * 16), N * 16, 128, kLayoutNone);
                            } else {
                                da = umma_desc(smem_u32(sA) + (k & 3) * 32, 16, 1024, kLayoutSW128);
                                db = umma_desc(smem_u32(sB) + (k & 3) * 32, 16, 1024, kLayoutSW128);
                            }
                            umma_ss(tmem, da, db, idesc, k > 0);
                        }
                        umma_commit(bar);
                        while (!mbar_try_wait(bar, phase)) {}
                        phase ^= 1;
                        const long long t1 = clock64();
                        if (t1 - t0 < best) best = t1 - t0;
                    }
                    out->mma[layout][ni][ci] = best;
                }
    }
    __syncthreads();
    // ---- issue-path costs in the MMA thread, with the warpgroup (warps 2..5) either parked or spinning on an mbarrier
    //      mode 0: parked at __syncthreads; mode 1: all 128 threads poll try_wait; mode 2: one lane per warp polls
    for (int mode = 0; mode < 3; ++mode) {
        uint64_t* bar_never = bar + 3;   // never completes during the experiment; completed at the end to release the pollers
        if (threadIdx.x == 0) { mbar_init(bar_never, 1); fence_mbar_init(); }
        __syncthreads();
        if (warp == 1 && elect_one()) {
            const uint32_t idesc = umma_idesc_f16(0, 128, 32);
            const uint64_t da0 = umma_desc(smem_u32(sA), 2048, 128, kLayoutNone);
            const uint64_t db0 = umma_desc(smem_u32(sB), 32 * 16, 128, kLayoutNone);
            for (volatile int i = 0; i < 500; ++i) {}
            // (a) 16 MMAs back to back, issue time only, then completion
            long long t0 = clock64();
            for (int k = 0; k < 16; ++k) umma_ss(tmem, da0 + (k & 3) * 256, db0 + (k & 3) * 64, idesc, k > 0);
            long long t1 = clock64();
            umma_commit(bar);
            long long t2 = clock64();
            while (!mbar_try_wait(bar, phase)) {}
            phase ^= 1;
            long long t3 = clock64();
            out->x[mode][0] = t1 - t0; out->x[mode][1] = t2 - t1; out->x[mode][2] = t3 - t2;
            // (b) 4 x (4 MMAs + commit to bar2-ish) then final commit + wait: the issue_q pattern
            t0 = clock64();
            for (int kb = 0; kb < 4; ++kb) {
                for (int k = 0; k < 4; ++k) umma_ss(tmem, da0 + k * 256, db0 + k * 64, idesc, (kb | k) > 0);
                umma_commit(bar2);
            }
            t1 = clock64();
            umma_commit(bar);
            while (!mbar_try_wait(bar, phase)) {}
            phase ^= 1;
            t2 = clock64();
            out->x[mode][3] = t1 - t0; out->x[mode][4] = t2 - t1;
            // (c) try_wait on an already completed phase (bar: previous phase) x 8
            t0 = clock64();
            for (int i = 0; i < 8; ++i) { while (!mbar_try_wait(bar, phase ^ 1)) {} }
            t1 = clock64();
            out->x[mode][5] = (t1 - t0) / 8;
            // (d) tcgen05.fence::after_thread_sync x 8
            t0 = clock64();
            for (int i = 0; i < 8; ++i) tc_fence_after();
            t1 = clock64();
            out->x[mode][6] = (t1 - t0) / 8;
            mbar_arrive(bar_never);   // release the pollers
        } else if (warp >= 2 && mode == 1) {
            while (!mbar_try_wait(bar_never, 0)) {}
        } else if (warp >= 2 && mode == 2) {
            if (lane == 0) { while (!mbar_try_wait(bar_never, 0)) {} }
            __syncwarp();
        }
        __syncthreads();
    }
    // bar2 was committed 12 times with count 1: re-init before the handoff test
    if (threadIdx.x == 0) { mbar_init(bar2, 1); fence_mbar_init(); }
    __syncthreads();
    // ---- handoff: warp 2 lane 0 arrives, warp 3 lane 0 spins
    if (warp == 2 && lane == 0) {
        for (volatile int i = 0; i < 2000; ++i) {}
        *t_arrive = clock64();
        mbar_arrive(bar2);
    }
    if (warp == 3 && lane == 0) {
        while (!mbar_try_wait(bar2, 0)) {}
        const long long t1 = clock64();
        out->hand = t1 - *t_arrive;
    }
    __syncthreads();
    // ---- the real thing: warpgroup (warps 2..5, 128 threads) arrives on bar3 -> MMA thread issues 2 MMAs N=144 -> commit
    //      -> warpgroup observes; measured by the warpgroup (arrive -> observed)
    uint32_t ph3 = 0, ph1 = phase;
    for (int it = 0; it < 4; ++it) {
        if (warp >= 2) {
            const long long t0 = clock64();
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar3);
            while (!mbar_try_wait(bar, ph1)) {}
            tc_fence_after();
            const long long t1 = clock64();
            if (threadIdx.x == 64) out->mma_wg[it] = t1 - t0;
        } else if (warp == 1 && lane == 0) {
            while (!mbar_try_wait(bar3, ph3)) {}
            tc_fence_after();
            const uint32_t idesc = umma_idesc_f16(0, 128, 144);
            for (int k = 0; k < 2; ++k) {
                const uint64_t da = umma_desc(smem_u32(sA) + k * 2 * 2048, 2048, 128, kLayoutNone);
                const uint64_t db = umma_desc(smem_u32(sB) + k * 2 * (144 * 16), 144 * 16, 128, kLayoutNone);
                umma_ss(tmem, da, db, idesc, k > 0);
            }
            umma_commit(bar);
        }
        ph3 ^= 1; ph1 ^= 1;
        __syncthreads();
    }
    // ---- tcgen05.ld
    if (warp == 2) {
        uint32_t r[32];
        long long t0 = clock64();
        tmem_ld16(tmem + (static_cast<uint32_t>(64) << 16), r);
        tmem_ld_wait();
        long long t1 = clock64();
        uint32_t acc = 0;
        for (int i = 0; i < 16; ++i) acc += r[i];
        const long long t2 = clock64();
        tmem_ld32(tmem + (static_cast<uint32_t>(64) << 16), r);
        tmem_ld_wait();
        const long long t3 = clock64();
        for (int i = 0; i < 32; ++i) acc += r[i];
        if (lane == 0) { out->ldtm16 = t1 - t0; out->ldtm32 = t3 - t2 + (acc == 0x12345 ? 1 : 0); }
        // ---- st.shared + fence.proxy.async
        t0 = clock64();
        st_shared_v4(smem_u32(sA) + threadIdx.x * 16, acc, 1, 2, 3);
        st_shared_v4(smem_u32(sA) + 2048 + threadIdx.x * 16, acc, 1, 2, 3);
        fence_proxy_async_smem();
        t1 = clock64();
        if (lane == 0) out->fence = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

int main() {
    Res* d;
    cudaMalloc(&d, sizeof(Res));
    cudaMemset(d, 0, sizeof(Res));
    cudaFuncSetAttribute(ubench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; ++rep) ubench_kernel<<<1, 192, 64 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    Res h;
    cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int Ns[4] = {16, 32, 144, 256};
    const int ns[4] = {1, 2, 4, 16};
    for (int l = 0; l < 2; ++l)
        for (int ni = 0; ni < 4; ++ni) {
            printf("mma %-8s N=%3d:", l ? "SW128" : "noswz", Ns[ni]);
            for (int ci = 0; ci < 4; ++ci) printf("  n=%2d -> %6lld cyc", ns[ci], h.mma[l][ni][ci]);
            printf("\n");
        }
    const char* modes[3] = {"warpgroup parked", "128 threads polling", "4 lanes polling"};
    for (int m = 0; m < 3; ++m)
        printf("issue path [%-19s]: 16 MMA issue %lld, commit %lld, completion wait %lld | 4x(4 MMA+commit) issue %lld, final commit+wait %lld | "
               "try_wait(done) %lld | fence::after %lld cyc\n", modes[m], h.x[m][0], h.x[m][1], h.x[m][2], h.x[m][3], h.x[m][4], h.x[m][5], h.x[m][6]);
    printf("handoff arrive->observe: %lld cyc\n", h.hand);
    printf("warpgroup arrive -> 2 MMA (N=144) -> commit -> warpgroup observes: %lld %lld %lld %lld cyc\n", h.mma_wg[0], h.mma_wg[1], h.mma_wg[2], h.mma_wg[3]);
    printf("tcgen05.ld x16 + wait: %lld cyc   x32 + wait: %lld cyc\n", h.ldtm16, h.ldtm32);
    printf("2 x st.shared.v4 + fence.proxy.async: %lld cyc\n", h.fence);
    return 0;
}
