// apx_internal.h — declarations shared by the translation units of libapx.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "../../include/apx.h"

namespace apx {

// ---- error plumbing (thread-local message, integer codes across the ABI) -------------
int  set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define APX_REQUIRE(cond, code, ...)                                 \
    do {                                                             \
        if (!(cond)) return ::apx::set_error((code), __VA_ARGS__);   \
    } while (0)

#define APX_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::apx::set_error(APX_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                    __FILE__, __LINE__);                                            \
    } while (0)

#define APX_LAUNCH_CHECK(name)                                                                         \
    do {                                                                                               \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess)                                                                         \
            return ::apx::set_error(APX_ECUDA, "launch of %s failed: %s", (name), cudaGetErrorString(_e)); \
        ::apx::count_launch();                                                                         \
    } while (0)

// cudaFuncSetAttribute is PER DEVICE: `mask` (one per call site) has one bit per device ordinal; true on the first call on the
// current device (thread-safe; ordinals >= 64 simply set the attribute every time).
bool first_use_on_device(std::atomic<uint64_t>& mask);
// Appends the programmatic-dependent-launch attribute (APX_PDL=0 turns it off) to a launch whose kernel calls pdl_wait().
void add_pdl_attr(cudaLaunchAttribute* attrs, unsigned* n);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int    ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int    elem_size(int dtype) { return dtype == APX_F32 ? 4 : 2; }

// ---- TMA descriptor encoding (driver entry point resolved at run time; no -lcuda) ------
int encode_tmap(CUtensorMap* tm, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box,
                int swizzle128 /* 0 none, 1 SWIZZLE_128B, 2 SWIZZLE_64B */);

// ---- tcgen05 kernels --------------------------------------------------------------------
int gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
            int64_t ldr, void* out, int64_t ldo, int M, int N, int K, int in_dtype, int out_dtype, int epilogue,
            cudaStream_t st, const float* ln_c1 = nullptr, float ln_eps = 0.f, void* out2 = nullptr, int64_t ldo2 = 0);
// flash-style attention over any number of keys (apx_flash_attn.cu)
bool flash_attn_supported(int C, int H);
int flash_attn(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs, const void* V, int64_t v_bs,
               int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S, int L, int C, int H, int dtype, cudaStream_t st);

// internal epilogue flag: never split K (the context projections: a row's value must not depend on how many rows the call has)
#define APX_EPI_NOSPLIT 16
// 16-bit-output specialisation (apx_gemm16.cu); gemm_tc routes to it when gemm16_supported()
bool gemm16_supported(const float* residual, int out_dtype, int in_dtype, int epilogue, const void* out2);
int gemm16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* residual, int64_t ldr, void* out,
           int64_t ldo, int M, int N, int K, int dtype, int epilogue, cudaStream_t st, const float* ln_c1, float ln_eps);

// Packed per-(batch, head) K / V^T tiles in UMMA no-swizzle K-major core-matrix order.
// Per (batch, head) the cache holds three tiles: K, V^T, and V^T extended by 16 rows of which the first (row d_pad) is all ones
// (dv_pad = d_pad + 16 rows): with it P·[V | 1] also yields the softmax denominator of every row in accumulator column d_pad.
// The first-generation kernels read {K, V^T}, the site kernel {K, extended V^T}.
struct KvPackGeom {
    int Lt, La, Lt_pad, La_pad, L_pad, d, d_pad, dv_pad, H, C;
    size_t k_tile_bytes, v_tile_bytes, vx_tile_bytes, bytes_per_bh;
};
KvPackGeom kv_pack_geom(int Lt, int La, int C, int H);

int ipattn_core_tc(const void* Q, const void* cache, void* O, int Bc, int S, int C, int H, int Lt, int La,
                   float ap_scale, int dtype, cudaStream_t st);
// plain self-attention (one softmax over L keys) on the tcgen05 attention core; `tiles` = attention_tc_tiles_bytes(...) scratch
size_t attention_tc_tiles_bytes(int B, int H, int d, int L);
int attention_tc(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs, const void* V, int64_t v_bs,
                 int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S, int H, int d, int L, int dtype, void* tiles,
                 cudaStream_t st);
int kv_pack(const void* Kt, const void* Vt, const void* Ka, const void* Va, void* cache, int Bc, int Lt, int La,
            int C, int H, int dtype, cudaStream_t st);

// ---- adapter site in one kernel (apx_ipattn_site.cu): all-head projections + two softmax warpgroups ---------------
struct SiteGeom {
    int ok, C, H, d, d_pad, Nq, n_kb;
    int cs_ok[3];                       // cluster sizes 1, 2, 4 the shape supports
    size_t wq_off[3], wo_off[3], total; // weight-pack sections per cluster size
};
SiteGeom site_geom(int C, int H);
// shared-memory plan of the all-heads kernels for one (geometry, context, cluster size)
struct SitePlan {
    bool ok;
    int nk, nv, n_wo, wo_early;
    uint32_t off_p, off_k, off_v, off_bar, a_stage_bytes, p_bytes, smem_bytes, wo_off[4];
};
SitePlan site_plan(const SiteGeom& g, int Lt, int La, int cs);
bool ipattn_site_supported(int C, int H, int Lt, int La);
int ipattn_site_pack_weights(const void* Wq, const void* Wo, void* pack, int C, int H, int dtype, cudaStream_t st);
int ipattn_site(const void* X, const void* wpack, const void* cache_tiles, const float* bo, void* Y, int Bc, int S, int C,
                int H, int Lt, int La, float ap_scale, int dtype, bool accumulate, cudaStream_t st, const float* ln_c1 = nullptr,
                const float* ln_c2 = nullptr, float ln_eps = 0.f);

// ---- adapter site, one head per CTA + a cluster per query tile (apx_ipattn_h1.cu): any C = H * d with H <= 8 -------------------
struct H1Geom {
    int ok, C, H, d, d_pad, dv_pad, Nq, n_kb;
    size_t wq_off, wo_off, total;       // weight-pack sections
};
H1Geom h1_geom(int C, int H);
bool ipattn_site_h1_supported(int C, int H, int Lt, int La);
int ipattn_site_h1_pack_weights(const void* Wq, const void* Wo, void* pack, int C, int H, int dtype, cudaStream_t st);
int ipattn_site_h1(const void* X, const void* wpack, const void* cache_tiles, const float* bo, void* Y, int Bc, int S, int C, int H,
                   int Lt, int La, float ap_scale, int dtype, bool accumulate, cudaStream_t st);
int ipattn_site_pick_cluster(int Bc, int S, int C, int H);
// all-heads kernel with FOUR softmax warpgroups (apx_ipattn_site4.cu): 16 text-padded + 128 audio-padded columns, no cluster
bool ipattn_site4_supported(int C, int H, int Lt, int La);
int ipattn_site4(const void* X, const void* wpack, const void* cache_tiles, const float* bo, void* Y, int Bc, int S, int C, int H,
                 int Lt, int La, float ap_scale, int dtype, bool accumulate, cudaStream_t st);

// ---- tensor-core attention backward of the adapter site (apx_bwd_tc.cu) ----------------------------------------------------------
bool ipattn_bwd_tc_supported(int C, int H, int Lt, int La);
int ipattn_bwd_tc(const void* Q, const void* dO, const void* cache_tiles, void* dQ, float* dKa, float* dVa, int Bc, int S, int C, int H, int Lt,
                  int La, float ap_scale, int dtype, cudaStream_t st);

// ---- CUDA-core kernels --------------------------------------------------------------------
int gemm_simt(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
              int64_t ldr, void* out, int64_t ldo, int M, int N, int K, int in_dtype, int out_dtype, int epilogue,
              cudaStream_t st);
int attention_simt(const void* Q, int64_t q_bs, int64_t q_rs, const void* K, int64_t k_bs, int64_t k_rs,
                   const void* V, int64_t v_bs, int64_t v_rs, void* O, int64_t o_bs, int64_t o_rs, int B, int S,
                   int H, int d, int L, int L0, float w0, float w1, int dtype, cudaStream_t st);
int patch_embed(const float* mel, const float* pwT, const float* pb, const float* pos, const float* cls, float* x,
                int B, cudaStream_t st);
int layernorm(const float* x, const float* w, const float* b, void* y, int rows, int cols, float eps, int out_dtype,
              cudaStream_t st);
// patch-embed as a tensor-core GEMM: 16-bit patch matrix [B*513, 256] (row b*513 = zeros for CLS) + fp32 residual table [B*513, 768]
// (pos_embed rows; cls + pos[0] - bias for the CLS row) + the conv weight as [768, 256] 16-bit
int patchify(const float* mel, const float* pos, const float* cls, const float* pb, void* a16, float* res, int B, int fmt, cudaStream_t st);
int patch_weight16(const float* pwT, void* w16, int fmt, cudaStream_t st);
int pool_tokens(const float* xn, float* emb, int B, int tp, int fp, cudaStream_t st);
int fbank(const float* wave16k, int B, int n_samples, const float* melT, float* out, int target_len, float* scratch, cudaStream_t st);
int resample(const float* x, int B, int n, const float* taps, int orig, int nw, int width, float* y, cudaStream_t st);

int current_impl();

}  // namespace apx
