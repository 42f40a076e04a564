// apx_site_common.cuh — device helpers shared by the adapter-site kernels (apx_ipattn_site.cu: all heads per CTA;
// apx_ipattn_h1.cu: one head per CTA, a cluster per query tile): MMA issue chains, TMEM -> operand-tile conversion, softmax pieces.
#pragma once
#include "apx_ptx.cuh"

namespace apx {
namespace {

__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

// tcgen05.mma with the descriptors given as (lo, hi) words: along a K chain only the 14-bit start-address field in the low
// word moves (every operand lives below 256 KB, so it cannot carry out), which keeps the per-MMA issue cost at two 32-bit
// uniform adds.
__device__ __forceinline__ void umma_ss_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// n dependent tcgen05.mma steps along K: A advances by a_step bytes, B by b_step bytes per 16-element k-step.
__device__ __forceinline__ void mma_chain(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t a_step, uint32_t b_step, int n,
                                          uint32_t idesc, uint32_t acc_first) {
    uint32_t a_lo = static_cast<uint32_t>(da), b_lo = static_cast<uint32_t>(db);
    const uint32_t a_hi = static_cast<uint32_t>(da >> 32), b_hi = static_cast<uint32_t>(db >> 32);
    a_step >>= 4; b_step >>= 4;
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        umma_ss_lohi(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, acc_first | static_cast<uint32_t>(k));
        a_lo += a_step;
        b_lo += b_step;
    }
}

template <int FMT>
__device__ __forceinline__ void store16(uint32_t dst, const float* v) {   // 16 values -> two 16-byte core-matrix rows
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack2<FMT>(v[2 * i], v[2 * i + 1]);
    st_shared_v4(dst, w[0], w[1], w[2], w[3]);
    st_shared_v4(dst + 2048, w[4], w[5], w[6], w[7]);
}

// Orders every later use of r[0..16) after this point (placed after tcgen05.wait::ld; emits no instruction).
__device__ __forceinline__ void reg_tie16(uint32_t* r) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                      "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

__device__ __forceinline__ float max16(const uint32_t* r) {
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r[i]));
    return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
}
__device__ __forceinline__ void mask16(uint32_t* r, int nv) {   // columns >= nv of the group -> -inf (exp2 -> exactly 0)
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = i < nv ? r[i] : 0xff800000u;
}
// exp2(r * sl - mxs) of a 16-column group -> 16 bit -> two core-matrix rows of the P tile (unnormalised: the row sums come
// out of the tensor core, from the ones row of the V^T tile)
template <int FMT>
__device__ __forceinline__ void exp_store16(const uint32_t* r, float sl, float mxs, uint32_t dst) {
    float e[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(r[i]), sl, -mxs));
    store16<FMT>(dst, e);
}
template <int FMT>
__device__ __forceinline__ void cvt_store16(const uint32_t* r, uint32_t dst) {   // fp32 accumulator group -> 16 bit -> two core-matrix rows
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    store16<FMT>(dst, v);
}


}  // namespace
}  // namespace apx
