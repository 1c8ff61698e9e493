// Shared device helpers of the 3xTF32 tcgen05 GEMM kernels (gemm_3xtf32.cu, gemm_tma.cu).
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace nesie {
namespace {

// Dynamic shared memory limit the GEMM kernels opt in to (227 KB per CTA minus 1 KB for their static
// barriers).  Always this constant rather than the launch's own size: the attribute is per FUNCTION,
// and a CUDA graph replayed (or profiled node by node by ncu) after a later, smaller launch lowered
// it fails with a launch error.
constexpr int G_MAX_DYN_SMEM = 232448 - 1024;
constexpr int G_SMEM_BUDGET = 224 * 1024;          // stages (+ epilogue staging); + 1 KB alignment slack

// back-off between failed barrier polls (ns); NESIE_GEMM_BACKOFF_NS overrides it at the first launch
__device__ unsigned g_wait_backoff_ns = 64;
constexpr int G_TILE = 128;
constexpr int G_SLABK = 32;            // fp32 elements per 128-byte swizzle row
constexpr int G_MAXSTAGES = 4;
constexpr int G_LGROUPS = 4;                       // loader groups of 128 threads (stage it -> group it % 4)
constexpr int G_EPIW = 8;                          // NT epilogue warps: 2 per TMEM lane quarter
constexpr int G_THREADS = 32 * G_EPIW + 128 * G_LGROUPS + 32;  // epilogue + loaders + MMA warp
constexpr int G_ASLAB = G_TILE * 128;  // bytes of one A part-slab
constexpr int G_EPI_STAGE = G_EPIW * 4096;  // epilogue transpose staging: 32 rows x 128 B per warp

// cycle counters of CTA 0 (NESIE_GEMM_DBG bit 128), read back by nesie_gemm_debug_profile


struct GemmParams {
  int R, N, K;        // logical sizes
  int npad, nslab;    // N rounded up to 16, K slabs of 32
  int nstages;        // smem pipeline depth (2..4)
  long long lda, ldc;
  const float *A;
  const unsigned char *Bimg;  // [hi|lo][nslab][npad][128 B]
  float *C;
  int fast;           // rows 16-byte aligned and K % 4 == 0: float4 loads without per-element checks
  int dbg;            // NESIE_GEMM_DBG experiment bits: 1 no A loads, 2 no C stores, 4 no MMAs
};

__device__ __forceinline__ unsigned g_smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ unsigned long long g_desc(unsigned smem_addr) {
  return (unsigned long long)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}
// MN-major tf32 operands have exactly one legal shared-memory layout: 128-byte swizzle with 32-byte
// atomicity (layout type 1; CUTLASS calls it SW128_32B).  An atom is 4 reduction rows x 128 bytes
// (32 elements along M/N); byte-address bits [5,7) are XORed with bits [7,9).  LBO = stride between
// atoms along M/N, SBO = stride between atoms along K.
__device__ __forceinline__ unsigned long long g_desc_mn(unsigned smem_addr, unsigned lbo, unsigned sbo) {
  return (unsigned long long)((smem_addr & 0x3FFFF) >> 4) | ((unsigned long long)(lbo >> 4) << 16) |
         ((unsigned long long)(sbo >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// kind::tf32: D = f32 (c_format 1), A = B = TF32 (format 2), K-major, M x N
__host__ __device__ constexpr unsigned g_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void g_mma(unsigned d, unsigned long long a, unsigned long long b,
                                      unsigned idesc, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool g_elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void g_commit(unsigned mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar)
               : "memory");
}
// ... and on the barrier at the same offset in every CTA of `cta_mask` (CTA pairs that share operand
// slabs through multicast copies: a stage is free once BOTH CTAs' MMAs have read it)
__device__ __forceinline__ void g_commit_mc(unsigned mbar, unsigned short cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(mbar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ unsigned g_cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void g_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void g_mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(unsigned mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void g_mbar_wait(unsigned mbar, unsigned parity) {
  unsigned ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
    if (ok) break;
    // back off: ~20 warps of a CTA poll barriers most of the time, and a warp scheduler that keeps
    // re-issuing a failing poll starves a co-resident CTA of another kernel (measured: an FPS CTA
    // sharing the SM ran 2-40x slower without the sleep)
    if (const unsigned ns = g_wait_backoff_ns) __nanosleep(ns);
  }
}
__device__ __forceinline__ void g_bulk_g2s(unsigned dst, const void *src, unsigned bytes,
                                           unsigned mbar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar)
      : "memory");
}
// the same copy delivered to the same shared-memory offset (and signalled on the barrier at the same
// offset) of every CTA in cta_mask
__device__ __forceinline__ void g_bulk_g2s_mc(unsigned dst, const void *src, unsigned bytes,
                                              unsigned mbar, unsigned short cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "h"(cta_mask)
      : "memory");
}
// The load and its wait are ONE asm statement: the destination registers only become valid at the wait,
// and between two separate statements the compiler may already move or spill them.
__device__ __forceinline__ void g_tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// x = hi + lo with hi = x rounded to nearest TF32 and lo = (x - hi) rounded to nearest TF32.
// Round-to-nearest (cvt.rna) instead of clearing the low mantissa bits keeps the split errors
// sign-symmetric: with truncation every operand is biased towards zero and the bias accumulates
// over a million-row reduction (5e-5 measured in the weight gradient); the tensor core itself
// truncates operand bits beyond TF32, so lo is pre-rounded as well.
__device__ __forceinline__ float to_tf32_rn(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
  hi = to_tf32_rn(x);
  lo = to_tf32_rn(x - hi);
}
// The same split for the loaders' inner loops.  cvt.rna.tf32 compiles to four instructions (add half
// an ulp, Inf/NaN test, select, mask); the test is dropped here (Inf stays Inf, NaN stays NaN under
// add-and-mask), which makes the split 5 instructions per element instead of 9 -- the loaders'
// issue slots, not HBM, were the limit of these kernels.
__device__ __forceinline__ float tf32_rn_fast(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32_fast(float x, float &hi, float &lo) {
  hi = tf32_rn_fast(x);
  lo = tf32_rn_fast(x - hi);
}
__device__ __forceinline__ void g_sts128(unsigned addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 g_lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void g_sts32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}


}  // namespace
}  // namespace nesie
