// apx_api.cu — the C ABI (include/apx.h): argument checking, workspace carving, TMA descriptor
// encoding and the host-side sequencing of the kernels for one adapter site and for the AudioMAE encoder.
#include "apx_internal.h"
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace apx {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_impl{APX_IMPL_TC};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }
int current_impl() { return g_impl.load(std::memory_order_relaxed); }
bool first_use_on_device(std::atomic<uint64_t>& mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const uint64_t bit = 1ull << dev;
    return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: libapx.so has no link-time dependency
// on libcuda.so, so it loads (and its symbols can be checked) on a machine without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

void add_pdl_attr(cudaLaunchAttribute* attrs, unsigned* n) {
    static const bool on = [] { const char* e = getenv("APX_PDL"); return !(e && e[0] == '0'); }();
    if (!on) return;
    attrs[*n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[*n].val.programmaticStreamSerializationAllowed = 1;
    ++*n;
}

int encode_tmap(CUtensorMap* tm, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    APX_REQUIRE(fn != nullptr, APX_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    const CUtensorMapDataType dt = dtype == APX_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = fn(tm, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    APX_REQUIRE(r == CUDA_SUCCESS, APX_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return APX_OK;
}

// ---- context cache layout: [K natural][V natural][packed K / V^T tiles] ---------------------------
struct CtxCacheLayout {
    size_t nat_bytes;    // one natural [Bc, L, C] tensor, 256-aligned
    size_t tiles_off;    // offset of the packed tiles
    size_t total;
};
static CtxCacheLayout ctx_cache_layout(int Bc, int Lt, int La, int C, int H) {
    CtxCacheLayout l;
    l.nat_bytes = align_up(static_cast<size_t>(Bc) * (Lt + La) * C * 2, 256);
    l.tiles_off = 2 * l.nat_bytes;
    const KvPackGeom g = kv_pack_geom(Lt, La, C, H);
    l.total = l.tiles_off + static_cast<size_t>(Bc) * H * g.bytes_per_bh;
    return l;
}

// rows < Lt of every batch element come from the text product, the others from the adapter product
__global__ void kv_merge_kernel(const uint4* __restrict__ kt, const uint4* __restrict__ vt, const uint4* __restrict__ ka,
                                const uint4* __restrict__ va, uint4* __restrict__ k, uint4* __restrict__ v, int L, int Lt,
                                int C8, size_t total) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int l = static_cast<int>((i / C8) % L);
        k[i] = l < Lt ? kt[i] : ka[i];
        v[i] = l < Lt ? vt[i] : va[i];
    }
}

}  // namespace apx

using namespace apx;

extern "C" {

const char* apx_last_error(void) { return g_err; }
int apx_version(void) { return 3; }
uint64_t apx_launch_count(void) { return g_launches.load(); }

int apx_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        set_error(APX_ECUDA, "no CUDA device");
        return 0;
    }
    if (p.major != 10) {
        set_error(APX_EUNSUPPORTED, "device '%s' is compute %d.%d; libapx is built for sm_100a only", p.name, p.major, p.minor);
        return 0;
    }
    return 1;
}

int apx_set_impl(int impl) {
    APX_REQUIRE(impl == APX_IMPL_TC || impl == APX_IMPL_SIMT || impl == APX_IMPL_TC_UNFUSED, APX_EINVAL,
                "apx_set_impl: unknown impl %d", impl);
    g_impl.store(impl);
    return APX_OK;
}
int apx_get_impl(void) { return current_impl(); }

int apx_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
               int64_t ldr, void* out, int64_t ldo, int M, int N, int K, int in_dtype, int out_dtype, int epilogue,
               void* stream) {
    APX_REQUIRE(A && W && out, APX_EINVAL, "apx_linear: null pointer");
    APX_REQUIRE(out_dtype == APX_F16 || out_dtype == APX_BF16 || out_dtype == APX_F32, APX_EINVAL,
                "apx_linear: out_dtype %d", out_dtype);
    APX_REQUIRE((epilogue & ~7) == 0 && (epilogue & 3) != 3, APX_EINVAL, "apx_linear: epilogue %d (LayerNorm folding: apx_linear_lnfold)", epilogue);
    APX_REQUIRE((epilogue & 3) != APX_EPI_GEGLU || N % 32 == 0, APX_EUNSUPPORTED, "apx_linear: GEGLU needs N %% 32 == 0 (N=%d)", N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (current_impl() == APX_IMPL_SIMT)
        return gemm_simt(A, lda, W, ldw, bias, residual, ldr, out, ldo, M, N, K, in_dtype, out_dtype, epilogue, st);
    return gemm_tc(A, lda, W, ldw, bias, residual, ldr, out, ldo, M, N, K, in_dtype, out_dtype, epilogue, st);
}

int apx_linear_lnfold(const void* A, int64_t lda, const void* Wg, int64_t ldw, const float* c1, const float* c2, float eps,
                      const float* residual, int64_t ldr, void* out, int64_t ldo, int M, int N, int K, int in_dtype, int out_dtype,
                      int epilogue, void* stream) {
    APX_REQUIRE(A && Wg && c1 && c2 && out, APX_EINVAL, "apx_linear_lnfold: null pointer");
    APX_REQUIRE(out_dtype == APX_F16 || out_dtype == APX_BF16 || out_dtype == APX_F32, APX_EINVAL, "apx_linear_lnfold: out_dtype %d", out_dtype);
    APX_REQUIRE((epilogue & ~7) == 0 && (epilogue & 3) != 3, APX_EINVAL, "apx_linear_lnfold: epilogue %d", epilogue);
    APX_REQUIRE((epilogue & 3) != APX_EPI_GEGLU || N % 32 == 0, APX_EUNSUPPORTED, "apx_linear_lnfold: GEGLU needs N %% 32 == 0 (N=%d)", N);
    APX_REQUIRE(eps > 0.f, APX_EINVAL, "apx_linear_lnfold: eps %g", eps);
    APX_REQUIRE(current_impl() != APX_IMPL_SIMT, APX_EUNSUPPORTED, "apx_linear_lnfold: tcgen05 path only");
    return gemm_tc(A, lda, Wg, ldw, c2, residual, ldr, out, ldo, M, N, K, in_dtype, out_dtype, epilogue | APX_EPI_LNFOLD,
                   static_cast<cudaStream_t>(stream), c1, eps);
}

int apx_flash_attention(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs, const void* V, int64_t v_bs,
                        int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S, int L, int C, int H, int dtype, void* stream) {
    APX_REQUIRE(Q && K && V && O, APX_EINVAL, "apx_flash_attention: null pointer");
    return flash_attn(Q, q_bs, q_rs, K, k_bs, k_rs, V, v_bs, v_rs, O, o_bs, o_rs, B, S, L, C, H, dtype, static_cast<cudaStream_t>(stream));
}

int apx_attention(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs, const void* V,
                  int64_t v_bs, int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S, int H, int d, int L,
                  int L0, float w0, float w1, int dtype, void* stream) {
    APX_REQUIRE(Q && K && V && O, APX_EINVAL, "apx_attention: null pointer");
    return attention_simt(Q, q_bs, q_rs, K, k_bs, k_rs, V, v_bs, v_rs, O, o_bs, o_rs, B, S, H, d, L, L0, w0, w1, dtype,
                          static_cast<cudaStream_t>(stream));
}

size_t apx_ctx_kv_bytes(int Bc, int Lt, int La, int C, int H) {
    if (Bc <= 0 || Lt < 0 || La < 0 || C <= 0 || H <= 0 || C % H) return 0;
    return ctx_cache_layout(Bc, Lt, La, C, H).total;
}
size_t apx_ctx_kv_workspace_bytes(int Bc, int Lt, int La, int C) {
    return 4 * align_up(static_cast<size_t>(Bc) * (Lt + La) * C * 2, 256);
}

int apx_ctx_kv(const void* ctx, const void* Wk, const void* Wv, const void* Wk_ip, const void* Wv_ip, void* cache,
               int Bc, int Lt, int La, int Dctx, int C, int H, int dtype, void* workspace, size_t workspace_bytes,
               void* stream) {
    APX_REQUIRE(ctx && Wk && Wv && Wk_ip && Wv_ip && cache && workspace, APX_EINVAL, "apx_ctx_kv: null pointer");
    APX_REQUIRE(Bc > 0 && Lt >= 0 && La >= 0 && Lt + La > 0 && C > 0 && H > 0 && C % H == 0, APX_EINVAL,
                "apx_ctx_kv: bad shape Bc=%d Lt=%d La=%d C=%d H=%d", Bc, Lt, La, C, H);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "apx_ctx_kv: dtype %d", dtype);
    APX_REQUIRE(C % 8 == 0, APX_EUNSUPPORTED, "apx_ctx_kv: C=%d must be a multiple of 8", C);
    APX_REQUIRE(workspace_bytes >= apx_ctx_kv_workspace_bytes(Bc, Lt, La, C), APX_EWORKSPACE,
                "apx_ctx_kv: workspace %zu < %zu", workspace_bytes, apx_ctx_kv_workspace_bytes(Bc, Lt, La, C));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int L = Lt + La, M = Bc * L;
    const CtxCacheLayout lay = ctx_cache_layout(Bc, Lt, La, C, H);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    void* kt = ws;
    void* vt = ws + lay.nat_bytes;
    void* ka = ws + 2 * lay.nat_bytes;
    void* va = ws + 3 * lay.nat_bytes;
    // Both weight sets are applied to every context row (one GEMM each over M = Bc*(Lt+La) rows); the merge /
    // pack kernels then pick text rows from the to_k/to_v products and audio rows from the to_k_ip/to_v_ip ones.
    const void* Ws[4] = {Wk, Wv, Wk_ip, Wv_ip};
    void* outs[4] = {kt, vt, ka, va};
    for (int i = 0; i < 4; ++i) {
        // never split along K: a context row's K / V must not depend on how many rows (clips) the call carries — the sharded run
        // reproduces the single-GPU run bit for bit (tests/test_gpu_ipattn.py::test_full_size_properties_batch8)
        int rc = current_impl() == APX_IMPL_SIMT
                     ? apx_linear(ctx, Dctx, Ws[i], Dctx, nullptr, nullptr, 0, outs[i], C, M, C, Dctx, dtype, dtype, APX_EPI_NONE, stream)
                     : gemm_tc(ctx, Dctx, Ws[i], Dctx, nullptr, nullptr, 0, outs[i], C, M, C, Dctx, dtype, dtype, APX_EPI_NOSPLIT, st);
        if (rc) return rc;
    }
    uint8_t* cb = static_cast<uint8_t*>(cache);
    const size_t total16 = static_cast<size_t>(M) * C / 8;
    const int blocks = static_cast<int>(std::min<size_t>((total16 + 255) / 256, 148 * 8));
    kv_merge_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint4*>(kt), static_cast<const uint4*>(vt),
                                            static_cast<const uint4*>(ka), static_cast<const uint4*>(va),
                                            reinterpret_cast<uint4*>(cb), reinterpret_cast<uint4*>(cb + lay.nat_bytes), L,
                                            Lt, C / 8, total16);
    APX_LAUNCH_CHECK("kv_merge_kernel");
    return kv_pack(kt, vt, ka, va, cb + lay.tiles_off, Bc, Lt, La, C, H, dtype, st);
}

size_t apx_ipattn_workspace_bytes(int Bc, int S, int C) {
    return 2 * align_up(static_cast<size_t>(Bc) * S * C * 2, 256);
}

// The pack blob holds the fused kernels' re-layouts of Wq / Wo: [all-heads kernel sections (apx_ipattn_site.cu)][one-head-per-CTA
// kernel sections (apx_ipattn_h1.cu)]; shapes neither covers take the three-kernel path, which reads the weights as they are.
struct PackLayout { size_t site_off, site_bytes, h1_off, h1_bytes, total; };
static PackLayout pack_layout(int C, int H) {
    PackLayout l;
    const SiteGeom sg = site_geom(C, H);
    const H1Geom hg = h1_geom(C, H);
    l.site_off = 0;
    l.site_bytes = sg.ok ? align_up(sg.total, 256) : 0;
    l.h1_off = l.site_bytes;
    l.h1_bytes = hg.ok ? align_up(hg.total, 256) : 0;
    l.total = l.site_bytes + l.h1_bytes;
    return l;
}
// APX_SITE_KERNEL=0 forces the three-kernel path (A/B measurements only)
static bool site_kernel_enabled() {
    static const bool on = [] { const char* e = getenv("APX_SITE_KERNEL"); return !(e && e[0] == '0'); }();
    return on;
}
static std::atomic<int> g_site_variant{[] { const char* e = getenv("APX_SITE_VARIANT"); return e ? atoi(e) : 0; }()};

// Which fused kernel runs a site call: 0 = none (three kernels), 1 = all heads per CTA, 2 = one head per CTA (cluster of H),
// 3 = all heads per CTA with four softmax warpgroups (full grids, 16 + 128 padded context columns).
// Measured on B200 (profiles/r2d_harness_sites.txt, us per call, CUDA-graph replay):
//   C=256 (2x1000, La=128): all-heads 12.6, one-head 27.3, three kernels 18.7     C=256 x16: 21.2 / 125 / 47
//   C=384 (2x252):          one-head 18.6, three kernels 19.4                      C=384 x16: 51 / 24.7
//   C=640 (2x64):           one-head 26.0, three kernels 18.8                      C=640 x16: 49 / 22.8
//   C=192, H=4 (1x100):     one-head 13.1, three kernels 19.1       C=256, La=200 (2x300): one-head 17.1, three kernels 20.9
// The all-heads kernel wins wherever it applies.  The one-head kernel pays for its decomposition — every CTA of a cluster streams
// the whole X tile, and the O all-gather runs at ~17 B/clk per SM over distributed shared memory — so it is picked only where the
// all-heads kernel does not apply, the gather is small (<= 90 KB per CTA) and the grid is a handful of clusters.
static int pick_site_kernel(int Bc, int S, int C, int H, int Lt, int La) {
    if (!site_kernel_enabled()) return 0;
    const bool all_ok = site_geom(C, H).ok && ipattn_site_supported(C, H, Lt, La);
    const bool h1_ok = ipattn_site_h1_supported(C, H, Lt, La);
    const int forced = g_site_variant.load(std::memory_order_relaxed);
    // four softmax warpgroups (score rows split over two threads): measured NO faster than two at Bc = 16 (23.0 vs 21.2 us,
    // profiles/r2h_site4_trace.txt: the exp2 passes of the two head slots saturate the MUFU pipe whenever they overlap, and the
    // per-head hand-off chain P -> P·V -> O -> next P bounds the rest) — kept as an opt-in variant (APX_SITE_ALLHEADS4), never auto-picked
    const bool four_ok = all_ok && ipattn_site4_supported(C, H, Lt, La);
    if (forced == APX_SITE_ALLHEADS) return all_ok ? 1 : 0;
    if (forced == APX_SITE_ONEHEAD) return h1_ok ? 2 : 0;
    if (forced == APX_SITE_ALLHEADS4) return four_ok ? 3 : 0;
    if (all_ok) return 1;
    if (!h1_ok) return 0;
    const H1Geom g = h1_geom(C, H);
    const int tiles = ceil_div(S, 128) * Bc;
    const size_t gather = static_cast<size_t>(H - 1) * g.d_pad * 256;
    // Round 2, second half: with the gemm16 epilogues the three-kernel form reads FASTER than the one-head kernel at C = 384 when the
    // calls run back to back (S = 252, Bc = 2: 16.0 vs 18.5 us at La = 128, profiles/r2x_harness.txt), but inside the denoise step the
    // two choices give the same step time (8.734 vs 8.735 ms, profiles/r2z_bench_h1max{384,320}.json: three launches behind PyTorch
    // kernels do not overlap the way back-to-back replays do) — so the single fused kernel stays the choice there.  C = 640 stays on
    // three kernels (25 vs 15 us).  APX_SITE_H1_MAXC moves the width limit.
    static const int h1_max_c = [] { const char* e = getenv("APX_SITE_H1_MAXC"); return e ? atoi(e) : 384; }();
    return (tiles * H <= 32 && gather <= 90 * 1024 && C <= h1_max_c) ? 2 : 0;
}

size_t apx_ipattn_pack_bytes(int C, int H) {
    if (C <= 0 || H <= 0 || C % H) return 0;
    return pack_layout(C, H).total;
}
int apx_ipattn_pack(const void* Wq, const void* Wo, void* wpack, int C, int H, int dtype, void* stream) {
    APX_REQUIRE(Wq && Wo && wpack, APX_EINVAL, "apx_ipattn_pack: null pointer");
    APX_REQUIRE(C > 0 && H > 0 && C % H == 0, APX_EINVAL, "apx_ipattn_pack: bad shape C=%d H=%d", C, H);
    const PackLayout l = pack_layout(C, H);
    APX_REQUIRE(l.total > 0, APX_EUNSUPPORTED, "apx_ipattn_pack: no single-kernel path covers C=%d H=%d", C, H);
    int rc = APX_OK;
    if (l.site_bytes) rc = ipattn_site_pack_weights(Wq, Wo, static_cast<uint8_t*>(wpack) + l.site_off, C, H, dtype, static_cast<cudaStream_t>(stream));
    if (rc == APX_OK && l.h1_bytes)
        rc = ipattn_site_h1_pack_weights(Wq, Wo, static_cast<uint8_t*>(wpack) + l.h1_off, C, H, dtype, static_cast<cudaStream_t>(stream));
    return rc;
}
int apx_ipattn_fused_ok(int C, int H, int Lt, int La) {
    if (C <= 0 || H <= 0 || C % H) return 0;
    return (ipattn_site_supported(C, H, Lt, La) || ipattn_site_h1_supported(C, H, Lt, La)) ? 1 : 0;
}
int apx_set_site_variant(int variant) {
    APX_REQUIRE(variant == APX_SITE_AUTO || variant == APX_SITE_ALLHEADS || variant == APX_SITE_ONEHEAD || variant == APX_SITE_ALLHEADS4, APX_EINVAL,
                "apx_set_site_variant: unknown variant %d", variant);
    g_site_variant.store(variant);
    return APX_OK;
}
int apx_ipattn_site_kernel(int Bc, int S, int C, int H, int Lt, int La) {
    if (Bc <= 0 || S <= 0 || C <= 0 || H <= 0 || C % H) return 0;
    return pick_site_kernel(Bc, S, C, H, Lt, La);
}

static int ipattn_fwd_impl(const void* X, const void* Wq, const void* cache, const void* Wo, const void* wpack, const float* bo,
                           void* Y, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype, void* workspace,
                           size_t workspace_bytes, void* stream, bool accumulate, const float* ln_c1 = nullptr, const float* ln_c2 = nullptr,
                           float ln_eps = 0.f) {
    APX_REQUIRE(X && cache && Y, APX_EINVAL, "apx_ipattn_fwd: null pointer");
    APX_REQUIRE(Bc > 0 && S > 0 && C > 0 && H > 0 && C % H == 0 && Lt >= 0 && La >= 0 && Lt + La > 0, APX_EINVAL,
                "apx_ipattn_fwd: bad shape Bc=%d S=%d C=%d H=%d Lt=%d La=%d", Bc, S, C, H, Lt, La);
    APX_REQUIRE(dtype == APX_F16 || dtype == APX_BF16, APX_EUNSUPPORTED, "apx_ipattn_fwd: dtype %d", dtype);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CtxCacheLayout lay = ctx_cache_layout(Bc, Lt, La, C, H);
    const uint8_t* cb = static_cast<const uint8_t*>(cache);
    // the product path: one fused tcgen05 kernel per site
    if (current_impl() == APX_IMPL_TC && wpack) {
        const PackLayout pl = pack_layout(C, H);
        const int which = pick_site_kernel(Bc, S, C, H, Lt, La);
        APX_REQUIRE(!ln_c1 || which == 1, APX_EUNSUPPORTED, "apx_ipattn_fwd_ln: only the all-heads site kernel folds the LayerNorm (kernel choice %d for "
                    "Bc=%d S=%d C=%d H=%d context %d+%d; query apx_ipattn_site_kernel first)", which, Bc, S, C, H, Lt, La);
        if (which == 1)
            return ipattn_site(X, static_cast<const uint8_t*>(wpack) + pl.site_off, cb + lay.tiles_off, bo, Y, Bc, S, C, H, Lt,
                               La, ap_scale, dtype, accumulate, st, ln_c1, ln_c2, ln_eps);
        if (which == 3)
            return ipattn_site4(X, static_cast<const uint8_t*>(wpack) + pl.site_off, cb + lay.tiles_off, bo, Y, Bc, S, C, H, Lt, La, ap_scale,
                                dtype, accumulate, st);
        if (which == 2)
            return ipattn_site_h1(X, static_cast<const uint8_t*>(wpack) + pl.h1_off, cb + lay.tiles_off, bo, Y, Bc, S, C, H, Lt,
                                  La, ap_scale, dtype, accumulate, st);
    }

    APX_REQUIRE(!ln_c1, APX_EUNSUPPORTED, "apx_ipattn_fwd_ln: needs the packed weights and a shape the all-heads site kernel runs");
    APX_REQUIRE(Wq && Wo && workspace, APX_EINVAL, "apx_ipattn_fwd: the three-kernel path needs Wq, Wo and a workspace");
    APX_REQUIRE((C / H) % 8 == 0 && (C / H) <= 80, APX_EUNSUPPORTED,
                "apx_ipattn_fwd: head dim %d (need d %% 8 == 0, d <= 80)", C / H);
    APX_REQUIRE(C % 64 == 0, APX_EUNSUPPORTED, "apx_ipattn_fwd: C=%d must be a multiple of 64", C);
    APX_REQUIRE(workspace_bytes >= apx_ipattn_workspace_bytes(Bc, S, C), APX_EWORKSPACE,
                "apx_ipattn_fwd: workspace %zu < %zu", workspace_bytes, apx_ipattn_workspace_bytes(Bc, S, C));
    const int M = Bc * S, d = C / H, L = Lt + La;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    void* q = ws;
    void* o = ws + align_up(static_cast<size_t>(M) * C * 2, 256);

    int rc = apx_linear(X, C, Wq, C, nullptr, nullptr, 0, q, C, M, C, C, dtype, dtype, APX_EPI_NONE, stream);
    if (rc) return rc;
    if (current_impl() == APX_IMPL_SIMT) {
        rc = attention_simt(q, static_cast<int64_t>(S) * C, C, cb, static_cast<int64_t>(L) * C, C, cb + lay.nat_bytes,
                            static_cast<int64_t>(L) * C, C, o, static_cast<int64_t>(S) * C, C, Bc, S, H, d, L, Lt, 1.0f,
                            ap_scale, dtype, st);
    } else {
        rc = ipattn_core_tc(q, cb + lay.tiles_off, o, Bc, S, C, H, Lt, La, ap_scale, dtype, st);
    }
    if (rc) return rc;
    // accumulate: the residual stream in Y is the 16-bit residual operand of the to_out GEMM and is overwritten in place
    return apx_linear(o, C, Wo, C, bo, accumulate ? static_cast<const float*>(Y) : nullptr, C, Y, C, M, C, C, dtype, dtype,
                      accumulate ? APX_EPI_RES16 : APX_EPI_NONE, stream);
}

int apx_ipattn_fwd(const void* X, const void* Wq, const void* cache, const void* Wo, const void* wpack, const float* bo,
                   void* Y, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype, void* workspace,
                   size_t workspace_bytes, void* stream) {
    return ipattn_fwd_impl(X, Wq, cache, Wo, wpack, bo, Y, Bc, S, C, H, Lt, La, ap_scale, dtype, workspace, workspace_bytes, stream,
                           false);
}
int apx_ipattn_fwd_add(const void* X, const void* Wq, const void* cache, const void* Wo, const void* wpack, const float* bo,
                       void* Y, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype, void* workspace,
                       size_t workspace_bytes, void* stream) {
    return ipattn_fwd_impl(X, Wq, cache, Wo, wpack, bo, Y, Bc, S, C, H, Lt, La, ap_scale, dtype, workspace, workspace_bytes, stream,
                           true);
}

int apx_ipattn_fwd_ln(const void* X, const void* cache, const void* wpack, const float* bo, const float* ln_c1, const float* ln_c2, float ln_eps,
                      void* Y, int Bc, int S, int C, int H, int Lt, int La, float ap_scale, int dtype, int accumulate, void* stream) {
    APX_REQUIRE(X && cache && wpack && ln_c1 && ln_c2 && Y, APX_EINVAL, "apx_ipattn_fwd_ln: null pointer");
    APX_REQUIRE(ln_eps > 0.f, APX_EINVAL, "apx_ipattn_fwd_ln: eps %g", ln_eps);
    return ipattn_fwd_impl(X, nullptr, cache, nullptr, wpack, bo, Y, Bc, S, C, H, Lt, La, ap_scale, dtype, nullptr, 0, stream, accumulate != 0,
                           ln_c1, ln_c2, ln_eps);
}

// ---- AudioMAE -----------------------------------------------------------------------------------
int apx_patch_embed(const float* mel, const float* patch_wT, const float* patch_b, const float* pos_embed,
                    const float* cls_token, float* x, int B, void* stream) {
    APX_REQUIRE(mel && patch_wT && patch_b && pos_embed && cls_token && x, APX_EINVAL, "apx_patch_embed: null pointer");
    return patch_embed(mel, patch_wT, patch_b, pos_embed, cls_token, x, B, static_cast<cudaStream_t>(stream));
}
int apx_layernorm(const float* x, const float* w, const float* b, void* y, int rows, int cols, float eps, int out_dtype,
                  void* stream) {
    APX_REQUIRE(x && w && b && y, APX_EINVAL, "apx_layernorm: null pointer");
    APX_REQUIRE(out_dtype == APX_F16 || out_dtype == APX_BF16 || out_dtype == APX_F32, APX_EINVAL,
                "apx_layernorm: out_dtype %d", out_dtype);
    return layernorm(x, w, b, y, rows, cols, eps, out_dtype, static_cast<cudaStream_t>(stream));
}
int apx_fbank(const float* wave16k, int n_samples, const float* mel_fbT, float* out, int target_len, float* scratch, void* stream) {
    APX_REQUIRE(wave16k && mel_fbT && out && scratch, APX_EINVAL, "apx_fbank: null pointer");
    return fbank(wave16k, 1, n_samples, mel_fbT, out, target_len, scratch, static_cast<cudaStream_t>(stream));
}
int apx_fbank_batch(const float* wave16k, int B, int n_samples, const float* mel_fbT, float* out, int target_len, float* scratch, void* stream) {
    APX_REQUIRE(wave16k && mel_fbT && out && scratch, APX_EINVAL, "apx_fbank_batch: null pointer");
    return fbank(wave16k, B, n_samples, mel_fbT, out, target_len, scratch, static_cast<cudaStream_t>(stream));
}
int apx_resample(const float* wave, int B, int n_samples, const float* taps, int orig, int new_, int width, float* out, void* stream) {
    APX_REQUIRE(wave && taps && out, APX_EINVAL, "apx_resample: null pointer");
    return resample(wave, B, n_samples, taps, orig, new_, width, out, static_cast<cudaStream_t>(stream));
}
int apx_pool_tokens(const float* xn, float* emb, int B, int time_pool, int freq_pool, void* stream) {
    APX_REQUIRE(xn && emb && B > 0, APX_EINVAL, "apx_pool_tokens: bad argument");
    return pool_tokens(xn, emb, B, time_pool, freq_pool, static_cast<cudaStream_t>(stream));
}

struct MaeWorkspace {
    size_t x, xn, qkv, att, hid, xf, tiles, a16, res0, w16, total;
};
static MaeWorkspace mae_ws(int B) {
    const size_t rows = static_cast<size_t>(B) * 513;
    MaeWorkspace w;
    size_t off = 0;
    w.x = off;   off += align_up(rows * 768 * 4, 256);
    w.xn = off;  off += align_up(rows * 768 * 2, 256);
    w.qkv = off; off += align_up(rows * 2304 * 2, 256);
    w.att = off; off += align_up(rows * 768 * 2, 256);
    w.hid = off; off += align_up(rows * 3072 * 2, 256);
    w.xf = off;  off += align_up(rows * 768 * 4, 256);
    w.tiles = off; off += align_up(attention_tc_tiles_bytes(B, 12, 64, 513), 256);   // K / V^T operand tiles of one ViT block
    w.a16 = off;  off += align_up(rows * 256 * 2, 256);    // patch matrix (A operand of the patch-embed GEMM)
    w.res0 = off; off += align_up(rows * 768 * 4, 256);    // its fp32 residual operand (pos_embed / CLS rows)
    w.w16 = off;  off += align_up(768 * 256 * 2, 256);     // conv weight as [768, 256] 16 bit
    w.total = off;
    return w;
}
size_t apx_audiomae_workspace_bytes(int B) { return B > 0 ? mae_ws(B).total : 0; }

// patch-embed -> depth x (LN, QKV, attention, proj+res, LN, fc1+GELU, fc2+res) -> final LN into `tokens`
static int audiomae_encode_into(const float* mel, const apx_vit_weights* w, float* tokens, int B, uint8_t* base,
                                const MaeWorkspace& ws, void* stream) {
    float* x = reinterpret_cast<float*>(base + ws.x);
    void* xn = base + ws.xn;
    uint16_t* qkv = reinterpret_cast<uint16_t*>(base + ws.qkv);
    void* att = base + ws.att;
    void* hid = base + ws.hid;
    const int rows = B * 513, dt = w->wdtype;
    const float eps = 1e-6f;  // partial(nn.LayerNorm, eps=1e-6): audio_encoder/models_mae.py:698
    int rc;
#define APX_STEP(call) do { rc = (call); if (rc) return rc; } while (0)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool tc = current_impl() != APX_IMPL_SIMT;
    const int fmt = dt == APX_F16 ? 0 : 1;
    // xn doubles as the 16-bit copy of the residual stream (A operand of the GEMMs with a folded LayerNorm)
    if (tc) {
        // patch-embed on the tensor cores: patch matrix + residual table, then ONE GEMM with bias + fp32 residual epilogue
        APX_STEP(patch_weight16(w->patch_wT, base + ws.w16, fmt, st));
        APX_STEP(patchify(mel, w->pos_embed, w->cls_token, w->patch_b, base + ws.a16, reinterpret_cast<float*>(base + ws.res0), B, fmt, st));
        APX_STEP(gemm_tc(base + ws.a16, 256, base + ws.w16, 256, w->patch_b, reinterpret_cast<const float*>(base + ws.res0), 768, x, 768, rows, 768,
                         256, dt, APX_F32, APX_EPI_NONE, st, nullptr, 0.f, xn, 768));
    } else {
        APX_STEP(apx_patch_embed(mel, w->patch_wT, w->patch_b, w->pos_embed, w->cls_token, x, B, stream));
    }
    static const bool mae_flash = [] { const char* e = getenv("APX_MAE_FLASH"); return !(e && e[0] == '0'); }();
    for (int l = 0; l < w->depth; ++l) {
        const apx_vit_block& k = w->blocks[l];
        const bool fold = tc && k.qkv_wg && k.qkv_c1 && k.qkv_c2 && k.fc1_wg && k.fc1_c1 && k.fc1_c2;
        // x += proj(attn(LN1(x)))
        if (fold) {
            APX_STEP(gemm_tc(xn, 768, k.qkv_wg, 768, k.qkv_c2, nullptr, 0, qkv, 2304, rows, 2304, 768, dt, dt, APX_EPI_LNFOLD, st, k.qkv_c1, eps));
        } else {
            APX_STEP(apx_layernorm(x, k.ln1_w, k.ln1_b, xn, rows, 768, eps, dt, stream));
            APX_STEP(apx_linear(xn, 768, k.qkv_w, 768, k.qkv_b, nullptr, 0, qkv, 2304, rows, 2304, 768, dt, dt, APX_EPI_NONE, stream));
        }
        if (!tc)   // CUDA-core cross-check (shared-memory K/V, warp-shuffle softmax)
            APX_STEP(apx_attention(qkv, 513LL * 2304, 2304, qkv + 768, 513LL * 2304, 2304, qkv + 1536, 513LL * 2304, 2304,
                                   att, 513LL * 768, 768, B, 513, 12, 64, 513, 513, 1.0f, 0.0f, dt, stream));
        else if (mae_flash)   // product path: flash-style tcgen05 attention straight from the q / k / v column slices of the qkv projection
            APX_STEP(flash_attn(qkv, 513LL * 2304, 2304, qkv + 768, 513LL * 2304, 2304, qkv + 1536, 513LL * 2304, 2304, att, 513LL * 768, 768, B,
                                513, 513, 768, 12, dt, st));
        else       // APX_MAE_FLASH=0: the round-1 core (K / V re-laid into operand tiles, the whole 513-key score row in tensor memory)
            APX_STEP(attention_tc(qkv, 513LL * 2304, 2304, qkv + 768, 513LL * 2304, 2304, qkv + 1536, 513LL * 2304, 2304, att,
                                  513LL * 768, 768, B, 513, 12, 64, 513, dt, base + ws.tiles, st));
        if (tc) APX_STEP(gemm_tc(att, 768, k.proj_w, 768, k.proj_b, x, 768, x, 768, rows, 768, 768, dt, APX_F32, APX_EPI_NONE, st, nullptr, 0.f,
                                 fold ? xn : nullptr, 768));
        else APX_STEP(apx_linear(att, 768, k.proj_w, 768, k.proj_b, x, 768, x, 768, rows, 768, 768, dt, APX_F32, APX_EPI_NONE, stream));
        // x += fc2(GELU(fc1(LN2(x))))
        if (fold) {
            APX_STEP(gemm_tc(xn, 768, k.fc1_wg, 768, k.fc1_c2, nullptr, 0, hid, 3072, rows, 3072, 768, dt, dt, APX_EPI_GELU | APX_EPI_LNFOLD, st,
                             k.fc1_c1, eps));
        } else {
            APX_STEP(apx_layernorm(x, k.ln2_w, k.ln2_b, xn, rows, 768, eps, dt, stream));
            APX_STEP(apx_linear(xn, 768, k.fc1_w, 768, k.fc1_b, nullptr, 0, hid, 3072, rows, 3072, 768, dt, dt, APX_EPI_GELU, stream));
        }
        const bool next_fold = fold && l + 1 < w->depth;
        if (tc) APX_STEP(gemm_tc(hid, 3072, k.fc2_w, 3072, k.fc2_b, x, 768, x, 768, rows, 768, 3072, dt, APX_F32, APX_EPI_NONE, st, nullptr, 0.f,
                                 next_fold ? xn : nullptr, 768));
        else APX_STEP(apx_linear(hid, 3072, k.fc2_w, 3072, k.fc2_b, x, 768, x, 768, rows, 768, 3072, dt, APX_F32, APX_EPI_NONE, stream));
    }
    APX_STEP(apx_layernorm(x, w->norm_w, w->norm_b, tokens, rows, 768, eps, APX_F32, stream));
#undef APX_STEP
    return APX_OK;
}

static int audiomae_check(const float* mel, const apx_vit_weights* w, const void* out, int B, const void* workspace,
                          size_t workspace_bytes) {
    APX_REQUIRE(mel && w && out && workspace && w->blocks, APX_EINVAL, "apx_audiomae: null pointer");
    APX_REQUIRE(B > 0, APX_EINVAL, "apx_audiomae: B=%d", B);
    APX_REQUIRE(w->embed == 768 && w->heads == 12 && w->mlp == 3072 && w->depth > 0, APX_EUNSUPPORTED,
                "apx_audiomae: only the ViT-B geometry (768/12/3072) is implemented, got %d/%d/%d", w->embed, w->heads,
                w->mlp);
    APX_REQUIRE(w->wdtype == APX_F16 || w->wdtype == APX_BF16, APX_EUNSUPPORTED, "apx_audiomae: wdtype %d", w->wdtype);
    APX_REQUIRE(workspace_bytes >= mae_ws(B).total, APX_EWORKSPACE, "apx_audiomae: workspace %zu < %zu", workspace_bytes,
                mae_ws(B).total);
    return APX_OK;
}

int apx_audiomae_encode(const float* mel, const apx_vit_weights* w, float* tokens, int B, void* workspace,
                        size_t workspace_bytes, void* stream) {
    int rc = audiomae_check(mel, w, tokens, B, workspace, workspace_bytes);
    if (rc) return rc;
    return audiomae_encode_into(mel, w, tokens, B, static_cast<uint8_t*>(workspace), mae_ws(B), stream);
}

int apx_audiomae_fwd(const float* mel, const apx_vit_weights* w, float* emb, int B, int time_pool, int freq_pool,
                     void* workspace, size_t workspace_bytes, void* stream) {
    int rc = audiomae_check(mel, w, emb, B, workspace, workspace_bytes);
    if (rc) return rc;
    const MaeWorkspace ws = mae_ws(B);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float* xf = reinterpret_cast<float*>(base + ws.xf);
    rc = audiomae_encode_into(mel, w, xf, B, base, ws, stream);
    if (rc) return rc;
    return apx_pool_tokens(xf, emb, B, time_pool, freq_pool, stream);
}

}  // extern "C"
