// harness.cu — standalone kernel harness for libapx.so (test infrastructure, not product).
// Runs each case in a forked child (a device trap or hang in one case cannot take the others down),
// checks the tcgen05 kernels against the CUDA-core kernels and a scalar CPU computation, prints
// diagnostics that localise layout / descriptor mistakes, and times each kernel with CUDA events.
//
//   tools/apx_harness            run every case
//   tools/apx_harness gemm       only cases whose name contains "gemm"
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <string>
#include <vector>
#include "../../include/apx.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("  CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
#define AK(x) do { int r_ = (x); if (r_) { printf("  apx error %d: %s (%s:%d)\n", r_, apx_last_error(), __FILE__, __LINE__); return 1; } } while (0)

static uint16_t f2h(float f, int dt) {
    if (dt == APX_F16) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
    __nv_bfloat16 b = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &b, 2); return u;
}
static float h2f(uint16_t u, int dt) {
    if (dt == APX_F16) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
    __nv_bfloat16 b; memcpy(&b, &u, 2); return __bfloat162float(b);
}
struct Rng {
    std::mt19937 g; std::normal_distribution<float> n{0.f, 1.f};
    explicit Rng(uint32_t s) : g(s) {}
    float operator()() { return n(g); }
};
static std::vector<uint16_t> rand16(size_t n, float scale, int dt, Rng& r) {
    std::vector<uint16_t> v(n);
    for (auto& x : v) x = f2h(r() * scale, dt);
    return v;
}
static std::vector<float> tofloat(const std::vector<uint16_t>& v, int dt) {
    std::vector<float> f(v.size());
    for (size_t i = 0; i < v.size(); ++i) f[i] = h2f(v[i], dt);
    return f;
}
template <class T> static T* dev(const std::vector<T>& h) {
    T* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T) + 256) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}
template <class T> static std::vector<T> host(const T* d, size_t n) {
    std::vector<T> h(n);
    cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost);
    return h;
}

// compare got vs ref ([rows, cols]); tolerance relative to max |ref|
static int compare(const char* what, const std::vector<float>& got, const std::vector<float>& ref, int rows, int cols,
                   float tol) {
    double maxref = 0, maxerr = 0, se = 0, sr = 0;
    size_t bad = 0, nan = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        maxref = std::max(maxref, (double)std::fabs(ref[i]));
        if (!std::isfinite(got[i])) { ++nan; continue; }
        const double e = std::fabs((double)got[i] - ref[i]);
        maxerr = std::max(maxerr, e); se += e * e; sr += (double)ref[i] * ref[i];
    }
    const double thr = tol * std::max(maxref, 1e-30);
    std::vector<int> badrow(8, 0), badcol(16, 0);
    int shown = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        const double e = std::isfinite(got[i]) ? std::fabs((double)got[i] - ref[i]) : 1e30;
        if (e > thr) {
            ++bad;
            const int r = i / cols, c = i % cols;
            badrow[r % 8]++; badcol[c % 16]++;
            if (shown < 8) { printf("    mismatch [%d,%d] got % .6f ref % .6f\n", r, c, got[i], ref[i]); ++shown; }
        }
    }
    const bool ok = bad == 0 && nan == 0;
    printf("  %-28s %s  max_err/max_ref=%.3e  rel_l2=%.3e  bad=%zu/%zu nan=%zu\n", what, ok ? "OK  " : "FAIL", maxerr / std::max(maxref, 1e-30),
           std::sqrt(se / std::max(sr, 1e-30)), bad, ref.size(), nan);
    if (!ok) {
        printf("    bad by row%%8:"); for (int v : badrow) printf(" %d", v);
        printf("\n    bad by col%%16:"); for (int v : badcol) printf(" %d", v);
        printf("\n    (rows=%d cols=%d)\n", rows, cols);
    }
    return ok ? 0 : 1;
}

// Stream every apx call of a case goes to (null = legacy default stream); time_ms switches it to a capturing stream.
static cudaStream_t g_stream = nullptr;

// Device time per call: `iters` calls are captured into ONE CUDA graph and the graph is replayed, so that host-side
// submission cost (tensor-map encoding, launch) cannot bound the measurement of microsecond-scale kernels.
static float time_ms(const std::function<int()>& fn, int iters = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) if (fn()) return -1.f;
    cudaDeviceSynchronize();
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    g_stream = st;
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    for (int i = 0; ok && i < iters; ++i) ok = fn() == 0;
    g_stream = nullptr;
    if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !ok || !graph) { printf("  graph capture failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1.f; }
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { printf("  graph instantiate failed\n"); return -1.f; }
    cudaGraphLaunch(exec, st); cudaStreamSynchronize(st);
    const int reps = 5;
    cudaEventRecord(a, st);
    for (int i = 0; i < reps; ++i) cudaGraphLaunch(exec, st);
    cudaEventRecord(b, st); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); cudaStreamDestroy(st);
    return ms / (iters * reps);
}

// ---------------------------------------------------------------------------------------------- GEMM
static int case_gemm(int M, int N, int K, int dt, int out_dt, bool bias, int epi, bool resid, bool cpu_ref) {
    Rng r(1234 + M + N + K);
    auto A = rand16((size_t)M * K, 1.0f, dt, r);
    auto W = rand16((size_t)N * K, 1.0f / std::sqrt((float)K), dt, r);
    std::vector<float> b(N), res((size_t)M * N);
    for (auto& x : b) x = r();
    for (auto& x : res) x = r();
    uint16_t *dA = dev(A), *dW = dev(W);
    float *db = dev(b), *dres = dev(res);
    const size_t osz = (size_t)M * N * (out_dt == APX_F32 ? 4 : 2);
    void *o_tc, *o_simt; CK(cudaMalloc(&o_tc, osz + 256)); CK(cudaMalloc(&o_simt, osz + 256));
    CK(cudaMemset(o_tc, 0xFF, osz)); CK(cudaMemset(o_simt, 0xFF, osz));
    auto run = [&](int impl, void* out) -> int {
        AK(apx_set_impl(impl));
        AK(apx_linear(dA, K, dW, K, bias ? db : nullptr, resid ? dres : nullptr, N, out, N, M, N, K, dt, out_dt, epi, g_stream));
        return 0;
    };
    if (run(APX_IMPL_SIMT, o_simt)) return 1;
    CK(cudaDeviceSynchronize());
    if (run(APX_IMPL_TC, o_tc)) return 1;
    CK(cudaDeviceSynchronize());
    auto tof = [&](void* d) {
        if (out_dt == APX_F32) return host((float*)d, (size_t)M * N);
        return tofloat(host((uint16_t*)d, (size_t)M * N), out_dt);
    };
    auto g_tc = tof(o_tc), g_simt = tof(o_simt);
    const float tol = out_dt == APX_BF16 ? 8e-3f : 2e-3f;
    int fails = 0;
    if (cpu_ref) {
        auto Af = tofloat(A, dt), Wf = tofloat(W, dt);
        std::vector<float> ref((size_t)M * N);
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < K; ++k) acc += (double)Af[(size_t)m * K + k] * Wf[(size_t)n * K + k];
                if (bias) acc += b[n];
                if (epi == APX_EPI_GELU) acc = 0.5 * acc * (1.0 + std::erf(acc * 0.7071067811865476));
                if (resid) acc += res[(size_t)m * N + n];
                ref[(size_t)m * N + n] = (float)acc;
            }
        fails += compare("simt vs cpu", g_simt, ref, M, N, tol);
        fails += compare("tcgen05 vs cpu", g_tc, ref, M, N, tol);
    } else {
        fails += compare("tcgen05 vs simt", g_tc, g_simt, M, N, tol);
    }
    const float t_tc = time_ms([&] { return run(APX_IMPL_TC, o_tc); });
    const float t_simt = time_ms([&] { return run(APX_IMPL_SIMT, o_simt); }, 5);
    const double fl = 2.0 * M * N * K;
    printf("  time: tcgen05 %.1f us (%.1f TFLOP/s)   simt %.1f us (%.1f TFLOP/s)\n", t_tc * 1e3, fl / t_tc * 1e-9, t_simt * 1e3,
           fl / t_simt * 1e-9);
    return fails;
}

// ---------------------------------------------------------------------------------------------- adapter site
static void cpu_site(const std::vector<float>& X, const std::vector<float>& Wq, const std::vector<float>& ctx,
                     const std::vector<float>& Wk, const std::vector<float>& Wv, const std::vector<float>& Wki,
                     const std::vector<float>& Wvi, const std::vector<float>& Wo, const std::vector<float>& bo, int Bc,
                     int S, int C, int H, int Lt, int La, int D, float scale, std::vector<float>& Y) {
    const int L = Lt + La, d = C / H;
    std::vector<double> Kc((size_t)Bc * L * C), Vc((size_t)Bc * L * C);
    for (int b = 0; b < Bc; ++b)
        for (int l = 0; l < L; ++l)
            for (int c = 0; c < C; ++c) {
                const float* wk = l < Lt ? &Wk[(size_t)c * D] : &Wki[(size_t)c * D];
                const float* wv = l < Lt ? &Wv[(size_t)c * D] : &Wvi[(size_t)c * D];
                const float* e = &ctx[((size_t)b * L + l) * D];
                double ak = 0, av = 0;
                for (int k = 0; k < D; ++k) { ak += (double)e[k] * wk[k]; av += (double)e[k] * wv[k]; }
                Kc[((size_t)b * L + l) * C + c] = ak; Vc[((size_t)b * L + l) * C + c] = av;
            }
    Y.assign((size_t)Bc * S * C, 0.f);
    std::vector<double> q(C), o(C), p(L);
    for (int b = 0; b < Bc; ++b)
        for (int s = 0; s < S; ++s) {
            const float* x = &X[((size_t)b * S + s) * C];
            for (int c = 0; c < C; ++c) { double a = 0; for (int k = 0; k < C; ++k) a += (double)x[k] * Wq[(size_t)c * C + k]; q[c] = a; }
            for (int h = 0; h < H; ++h) {
                for (int l = 0; l < L; ++l) { double a = 0; for (int i = 0; i < d; ++i) a += q[h * d + i] * Kc[((size_t)b * L + l) * C + h * d + i]; p[l] = a / std::sqrt((double)d); }
                for (int i = 0; i < d; ++i) o[h * d + i] = 0;
                for (int seg = 0; seg < 2; ++seg) {
                    const int l0 = seg ? Lt : 0, l1 = seg ? L : Lt;
                    if (l1 <= l0) continue;
                    double mx = -1e300, sum = 0;
                    for (int l = l0; l < l1; ++l) mx = std::max(mx, p[l]);
                    for (int l = l0; l < l1; ++l) sum += std::exp(p[l] - mx);
                    const double w = seg ? scale : 1.0;
                    for (int l = l0; l < l1; ++l) { const double pl = std::exp(p[l] - mx) / sum * w; for (int i = 0; i < d; ++i) o[h * d + i] += pl * Vc[((size_t)b * L + l) * C + h * d + i]; }
                }
            }
            float* y = &Y[((size_t)b * S + s) * C];
            for (int c = 0; c < C; ++c) { double a = bo[c]; for (int k = 0; k < C; ++k) a += o[k] * Wo[(size_t)c * C + k]; y[c] = (float)a; }
        }
}

static int case_site(int Bc, int S, int C, int H, int Lt, int La, int dt, bool cpu_ref) {
    const int D = 768, L = Lt + La;
    Rng r(77 + S + C + La);
    auto X = rand16((size_t)Bc * S * C, 1.0f, dt, r);
    auto ctx = rand16((size_t)Bc * L * D, 1.0f, dt, r);
    auto Wq = rand16((size_t)C * C, 1.0f / std::sqrt((float)C), dt, r);
    auto Wo = rand16((size_t)C * C, 1.0f / std::sqrt((float)C), dt, r);
    auto Wk = rand16((size_t)C * D, 1.5f / std::sqrt((float)D), dt, r);
    auto Wv = rand16((size_t)C * D, 1.0f / std::sqrt((float)D), dt, r);
    auto Wki = rand16((size_t)C * D, 1.5f / std::sqrt((float)D), dt, r);
    auto Wvi = rand16((size_t)C * D, 1.0f / std::sqrt((float)D), dt, r);
    std::vector<float> bo(C);
    for (auto& x : bo) x = 0.1f * r();
    const float ap = 0.5f;
    uint16_t *dX = dev(X), *dctx = dev(ctx), *dWq = dev(Wq), *dWo = dev(Wo), *dWk = dev(Wk), *dWv = dev(Wv), *dWki = dev(Wki), *dWvi = dev(Wvi);
    float* dbo = dev(bo);
    const size_t cbytes = apx_ctx_kv_bytes(Bc, Lt, La, C, H), wsb = std::max(apx_ctx_kv_workspace_bytes(Bc, Lt, La, C), apx_ipattn_workspace_bytes(Bc, S, C));
    void *cache[3], *ws, *Y[3], *wpack = nullptr;
    for (int i = 0; i < 3; ++i) { CK(cudaMalloc(&cache[i], cbytes)); CK(cudaMalloc(&Y[i], (size_t)Bc * S * C * 2)); CK(cudaMemset(Y[i], 0xFF, (size_t)Bc * S * C * 2)); }
    CK(cudaMalloc(&ws, wsb));
    const bool fused = apx_ipattn_fused_ok(C, H, Lt, La) && apx_ipattn_pack_bytes(C, H) > 0;
    if (fused) {
        CK(cudaMalloc(&wpack, apx_ipattn_pack_bytes(C, H)));
        AK(apx_ipattn_pack(dWq, dWo, wpack, C, H, dt, nullptr));
    }
    // i = 0: product path (fused kernel when covered), 1: CUDA-core cross-check, 2: three tcgen05 kernels
    auto run = [&](int impl, int i, bool build_cache) -> int {
        AK(apx_set_impl(impl));
        if (build_cache) AK(apx_ctx_kv(dctx, dWk, dWv, dWki, dWvi, cache[i], Bc, Lt, La, D, C, H, dt, ws, wsb, g_stream));
        AK(apx_ipattn_fwd(dX, dWq, cache[i], dWo, impl == APX_IMPL_TC ? wpack : nullptr, dbo, Y[i], Bc, S, C, H, Lt, La, ap, dt, ws, wsb, g_stream));
        return 0;
    };
    if (run(APX_IMPL_SIMT, 1, true)) return 1;
    CK(cudaDeviceSynchronize());
    if (run(APX_IMPL_TC_UNFUSED, 2, true)) return 1;
    CK(cudaDeviceSynchronize());
    if (run(APX_IMPL_TC, 0, true)) return 1;
    CK(cudaDeviceSynchronize());
    auto y_tc = tofloat(host((uint16_t*)Y[0], (size_t)Bc * S * C), dt), y_simt = tofloat(host((uint16_t*)Y[1], (size_t)Bc * S * C), dt);
    auto y_un = tofloat(host((uint16_t*)Y[2], (size_t)Bc * S * C), dt);
    const float tol = dt == APX_BF16 ? 2e-2f : 4e-3f;
    int fails = 0;
    printf("  product path: %s\n", fused ? "FUSED kernel" : "three kernels (shape not covered by the fused kernel)");
    if (cpu_ref) {
        std::vector<float> ref;
        cpu_site(tofloat(X, dt), tofloat(Wq, dt), tofloat(ctx, dt), tofloat(Wk, dt), tofloat(Wv, dt), tofloat(Wki, dt), tofloat(Wvi, dt), tofloat(Wo, dt), bo, Bc, S, C, H, Lt, La, D, ap, ref);
        fails += compare("simt vs cpu", y_simt, ref, Bc * S, C, tol);
        fails += compare("tcgen05 unfused vs cpu", y_un, ref, Bc * S, C, tol);
        fails += compare("tcgen05 product vs cpu", y_tc, ref, Bc * S, C, tol);
    } else {
        fails += compare("tcgen05 unfused vs simt", y_un, y_simt, Bc * S, C, tol);
        fails += compare("tcgen05 product vs simt", y_tc, y_simt, Bc * S, C, tol);
    }
    // each fused kernel on its own (the product path picks between them by grid size)
    float t_var[4] = {0.f, 0.f, 0.f, 0.f};
    const char* vname[4] = {"", "all-heads site kernel", "one-head (cluster) site kernel", "all-heads site kernel, 4 softmax warpgroups"};
    for (int v = 1; v <= 3 && fused; ++v) {
        AK(apx_set_site_variant(v));
        if (apx_ipattn_site_kernel(Bc, S, C, H, Lt, La) == v) {
            CK(cudaMemset(Y[0], 0xFF, (size_t)Bc * S * C * 2));
            if (run(APX_IMPL_TC, 0, false)) return 1;
            CK(cudaDeviceSynchronize());
            auto y_v = tofloat(host((uint16_t*)Y[0], (size_t)Bc * S * C), dt);
            fails += compare(vname[v], y_v, y_simt, Bc * S, C, tol);
            t_var[v] = time_ms([&] { return run(APX_IMPL_TC, 0, false); });
        }
        AK(apx_set_site_variant(APX_SITE_AUTO));
    }
    const float t_tc = time_ms([&] { return run(APX_IMPL_TC, 0, false); });
    const float t_un = time_ms([&] { return run(APX_IMPL_TC_UNFUSED, 2, false); });
    const float t_simt = time_ms([&] { return run(APX_IMPL_SIMT, 1, false); }, 5);
    const double fl = 4.0 * Bc * S * (double)C * C + 4.0 * Bc * S * (double)L * C;
    printf("  time per site call: product %.1f us (%.1f TFLOP/s, kernel choice %d)   all-heads %.1f us   all-heads/4wg %.1f us   one-head %.1f us   unfused %.1f us   simt %.1f us\n",
           t_tc * 1e3, fl / t_tc * 1e-9, apx_ipattn_site_kernel(Bc, S, C, H, Lt, La), t_var[1] * 1e3, t_var[3] * 1e3, t_var[2] * 1e3, t_un * 1e3, t_simt * 1e3);
    AK(apx_set_impl(APX_IMPL_TC));
    return fails;
}

// ---------------------------------------------------------------------------------------------- driver
struct Case { std::string name; std::function<int()> fn; };

int main(int argc, char** argv) {
    const char* filt = argc > 1 ? argv[1] : "";
    std::vector<Case> cases;
    auto G = [&](const char* nm, int M, int N, int K, int dt, int odt, bool b, int e, bool r, bool cpu) {
        cases.push_back({nm, [=] { return case_gemm(M, N, K, dt, odt, b, e, r, cpu); }});
    };
    G("gemm_128x64x64_f16_cpu", 128, 64, 64, APX_F16, APX_F32, false, 0, false, true);
    G("gemm_128x128x128_bf16_cpu", 128, 128, 128, APX_BF16, APX_F32, false, 0, false, true);
    G("gemm_300x320x320_f16_bias_cpu", 300, 320, 320, APX_F16, APX_F16, true, 0, false, true);
    G("gemm_200x192x256_bf16_gelu_res_cpu", 200, 192, 256, APX_BF16, APX_F32, true, 1, true, true);
    G("gemm_100x64x1024_f16_bias_splitk8_cpu", 100, 64, 1024, APX_F16, APX_F16, true, 0, false, true);
    G("gemm_130x128x512_bf16_gelu_res_splitk4_cpu", 130, 128, 512, APX_BF16, APX_F32, true, 1, true, true);
    G("gemm_2000x256x256_f16", 2000, 256, 256, APX_F16, APX_F16, false, 0, false, false);
    G("gemm_2000x256x1024_f16_res_splitk4", 2000, 256, 1024, APX_F16, APX_F32, true, 0, true, false);
    G("gemm_500x384x1536_f16_bias_splitk8", 500, 384, 1536, APX_F16, APX_F16, true, 0, false, false);
    G("gemm_126x640x2560_bf16_bias_splitk8", 126, 640, 2560, APX_BF16, APX_BF16, true, 0, false, false);
    G("gemm_126x1920x640_f16_splitk4", 126, 1920, 640, APX_F16, APX_F16, false, 0, false, false);
    G("gemm_513x768x3072_f16_res_splitk4", 513, 768, 3072, APX_F16, APX_F32, true, 0, true, false);
    G("gemm_504x384x384_f16_bias", 504, 384, 384, APX_F16, APX_F16, true, 0, false, false);
    G("gemm_128x640x640_bf16_bias", 128, 640, 640, APX_BF16, APX_BF16, true, 0, false, false);
    G("gemm_16000x256x256_f16_bias", 16000, 256, 256, APX_F16, APX_F16, true, 0, false, false);
    G("gemm_1026x2304x768_f16_bias", 1026, 2304, 768, APX_F16, APX_F16, true, 0, false, false);
    G("gemm_4104x3072x768_bf16_gelu", 4104, 3072, 768, APX_BF16, APX_BF16, true, 1, false, false);
    G("gemm_4104x768x3072_f16_res", 4104, 768, 3072, APX_F16, APX_F32, true, 0, true, false);
    auto S = [&](const char* nm, int Bc, int Sq, int C, int H, int Lt, int La, int dt, bool cpu) {
        cases.push_back({nm, [=] { return case_site(Bc, Sq, C, H, Lt, La, dt, cpu); }});
    };
    S("site_cfg1_2x1024x320_77+8_f16_cpu", 2, 1024, 320, 8, 77, 8, APX_F16, true);
    S("site_1x200x256_8+32_f16_cpu", 1, 200, 256, 8, 8, 32, APX_F16, true);
    S("site_1x64x640_8+8_bf16_cpu", 1, 64, 640, 8, 8, 8, APX_BF16, true);
    S("site_1x130x384_8+128_f16_cpu", 1, 130, 384, 8, 8, 128, APX_F16, true);
    S("site_1x128x256_8+512_f16_cpu", 1, 128, 256, 8, 8, 512, APX_F16, true);
    S("site_1x128x256_textonly_f16_cpu", 1, 128, 256, 8, 8, 0, APX_F16, true);
    S("site_1x128x256_audioonly_f16_cpu", 1, 128, 256, 8, 0, 40, APX_F16, true);
    S("site_3x129x256_8+24_bf16_cpu", 3, 129, 256, 8, 8, 24, APX_BF16, true);
    S("site_2x96x512_16h_8+32_f16_cpu", 2, 96, 512, 16, 8, 32, APX_F16, true);
    S("site_2x1000x256_8+128_f16", 2, 1000, 256, 8, 8, 128, APX_F16, false);
    S("site_2x1000x256_8+32_f16", 2, 1000, 256, 8, 8, 32, APX_F16, false);
    S("site_2x1000x256_8+8_f16", 2, 1000, 256, 8, 8, 8, APX_F16, false);
    S("site_2x252x384_8+128_f16", 2, 252, 384, 8, 8, 128, APX_F16, false);
    S("site_2x64x640_8+128_f16", 2, 64, 640, 8, 8, 128, APX_F16, false);
    S("site_16x1000x256_8+128_f16", 16, 1000, 256, 8, 8, 128, APX_F16, false);
    S("site_16x252x384_8+128_bf16", 16, 252, 384, 8, 8, 128, APX_BF16, false);
    S("site_16x64x640_8+512_f16", 16, 64, 640, 8, 8, 512, APX_F16, false);
    S("site_16x64x640_8+128_f16", 16, 64, 640, 8, 8, 128, APX_F16, false);
    S("site_16x1000x256_8+120_bf16", 16, 1000, 256, 8, 8, 120, APX_BF16, false);
    S("site_3x200x256_1+128_f16_cpu", 3, 200, 256, 8, 1, 128, APX_F16, true);
    S("site_2x130x128_4h_16+113_f16_cpu", 2, 130, 128, 4, 16, 113, APX_F16, true);
    S("site_2x252x384_8+32_bf16", 2, 252, 384, 8, 8, 32, APX_BF16, false);
    S("site_2x64x640_8+8_f16", 2, 64, 640, 8, 8, 8, APX_F16, false);
    S("site_8x64x640_8+32_f16", 8, 64, 640, 8, 8, 32, APX_F16, false);
    S("site_2x300x256_8+200_f16_cpu", 2, 300, 256, 8, 8, 200, APX_F16, true);
    S("site_1x100x192_4h_8+128_f16_cpu", 1, 100, 192, 4, 8, 128, APX_F16, true);

    int total = 0, failed = 0;
    for (auto& c : cases) {
        if (*filt && c.name.find(filt) == std::string::npos) continue;
        if (getenv("APX_HARNESS_SKIP_BIG") && c.name.find("16x") != std::string::npos) continue;
        ++total;
        printf("[case] %s\n", c.name.c_str());
        fflush(stdout);
        pid_t pid = fork();
        if (pid == 0) {
            alarm(90);
            if (!apx_device_ok()) { printf("  %s\n", apx_last_error()); _exit(3); }
            int rc = c.fn();
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  trailing CUDA error: %s\n", cudaGetErrorString(e)); rc = 1; }
            fflush(stdout);
            _exit(rc ? 1 : 0);
        }
        int st = 0;
        waitpid(pid, &st, 0);
        const bool ok = WIFEXITED(st) && WEXITSTATUS(st) == 0;
        if (!ok) {
            ++failed;
            if (WIFSIGNALED(st)) printf("  child killed by signal %d\n", WTERMSIG(st));
        }
        printf("[%s] %s\n", ok ? "pass" : "FAIL", c.name.c_str());
        fflush(stdout);
    }
    printf("harness: %d/%d cases passed\n", total - failed, total);
    return failed ? 1 : 0;
}
