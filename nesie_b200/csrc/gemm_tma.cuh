// TMA-fed variants of the 3xTF32 GEMM kernels (included by gemm_3xtf32.cu).
//
// The register-staged loaders of gemm_nt_3xtf32_kernel / gemm_wgrad_3xtf32_kernel pay one full
// memory round trip per pipeline stage with nothing else in flight for that stage (measured: 2700-
// 4800 cycles of "fill" per 16 KB slab, ~2.5 TB/s of reads chip-wide).  Here the fp32 operand tiles
// are brought in by the TMA engine (cp.async.bulk.tensor.2d) straight into the swizzled UMMA layout,
// so every free stage always has its load in flight and no thread holds data in registers:
//
//   producer (1 thread) : waits for a free stage, arms its mbarrier, issues the tensor copies
//   transform (8 warps) : the RAW fp32 tile is used as the `hi` operand as it is -- the tensor core
//                         reads the top 19 bits of each word, i.e. hi = x truncated to TF32 -- and
//                         lo = RN_tf32(x - trunc(x)) is written to a second tile of the same layout
//                         (LDS.128 + 16 ALU + STS.128 per four elements, half the old loader's work)
//   MMA (1 elected lane): hi*hi + hi*lo + lo*hi as before
//   epilogue            : unchanged
//
// Requirements (else the register-staged kernels run): 16-byte aligned base and row stride.
#pragma once
#include <cuda.h>

#include "gemm_common.cuh"

namespace nesie {
namespace {

constexpr int T_EPIW = 8;   // epilogue warps (NT); the weight-gradient kernel uses the first four
constexpr int T_XFW = 8;    // transform warps
constexpr int T_THREADS = 32 * (T_EPIW + T_XFW + 2);  // + producer warp + MMA warp
// A second build capped at 64 registers can share an SM with a CTA of the FPS kernel (256 threads x
// 112 registers) that the input pipeline runs on another stream.  Measured in bench.py (same box,
// alternating): capped 9.62 / 9.48 ms per step, uncapped 9.51 / 9.35 -- the cap costs the kernels
// more than co-residency returns, so the uncapped build is the default (NESIE_GEMM_REGS=64 selects
// the capped one).
constexpr int T_MAXREG = 64;
constexpr int T_CTRL_BYTES = 256;  // barriers + TMEM base of the NT kernel, behind its tiles
constexpr int T_XF0 = T_EPIW, T_PROD = T_EPIW + T_XFW, T_MMA = T_PROD + 1;

__device__ __forceinline__ void g_tma_2d(unsigned dst, const CUtensorMap *tm, int c0, int c1,
                                         unsigned mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(mbar)
      : "memory");
}
__device__ __forceinline__ void g_mbar_arm(unsigned mbar, unsigned bytes) { g_mbar_expect_tx(mbar, bytes); }

// lo part of a raw fp32 word whose top 19 bits the tensor core uses as hi
__device__ __forceinline__ float lo_of_raw(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return tf32_rn_fast(x - hi);
}
__device__ __forceinline__ float4 lo_of_raw4(float4 v) {
  return make_float4(lo_of_raw(v.x), lo_of_raw(v.y), lo_of_raw(v.z), lo_of_raw(v.w));
}

// Operand prologue: the operand is relu(y * scale + shift) of the stored tensor y (training BatchNorm
// + ReLU of the previous layer, applied on the fly so that the activation never goes to HBM).  The
// value differs from the raw word, so both hi and lo are written.
__device__ __forceinline__ void bn_relu_split4(float4 y, float4 sc, float4 sh, float4 &hi, float4 &lo) {
  const float a0 = fmaxf(fmaf(y.x, sc.x, sh.x), 0.f), a1 = fmaxf(fmaf(y.y, sc.y, sh.y), 0.f);
  const float a2 = fmaxf(fmaf(y.z, sc.z, sh.z), 0.f), a3 = fmaxf(fmaf(y.w, sc.w, sh.w), 0.f);
  split_tf32_fast(a0, hi.x, lo.x);
  split_tf32_fast(a1, hi.y, lo.y);
  split_tf32_fast(a2, hi.z, lo.z);
  split_tf32_fast(a3, hi.w, lo.w);
}
__device__ __forceinline__ float4 ldg4_guard(const float *p, int c, int n) {
  // four consecutive per-channel constants; channels >= n read as 0 (n is a multiple of 4 here)
  return c < n ? __ldg(reinterpret_cast<const float4 *>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ----------------------------------------------------------------------------------------------
// C[R x N] = A[R x K] * B[N x K]^T   (B pre-split / pre-swizzled image as in the register kernel)
// ----------------------------------------------------------------------------------------------
struct GemmTmaParams {
  int R, N, K;
  int npad, nslab, nstages;
  long long ldc;
  const unsigned char *Bimg;
  float *C;
  const float *pro_scale, *pro_shift;  // optional [K]: A operand = relu(a * scale + shift)
  float *col_stats;                    // optional [gridDim.x * 4][2][N]: column sums of C and C^2
  // optional BatchNorm-backward statistics of the OUTPUT (this GEMM is the data gradient that
  // produces g = dL/dA of the previous layer): with y = bn_y[row, col] the previous layer's
  // pre-activation and bn_stats = mean | invstd | scale | shift (4 x N), col_stats receives the
  // column sums of g * [y*scale+shift > 0] and of that times xhat = (y - mean) * invstd -- what
  // bn_colsum_kernel<1> computes in a separate sweep over g and y
  const float *bn_y;
  long long ldy;
  const float *bn_stats;
  // optional row-group extras of the OUTPUT (pooled layers of the MiniPointNets / SA modules):
  //  pool_max / pool_amax [R / pool_u][N]: per-column maximum over every unit of pool_u (16 or 32)
  //      consecutive rows and the first row within the unit that attains it (the way torch.max /
  //      F.max_pool2d route their gradient); pool_min / pool_amin likewise for the minimum (a
  //      BatchNorm that follows may have a negative scale).  Taken from the accumulator tile, so a
  //      max-pooled layer's pre-activation need not exist in HBM at all (store_c = 0).
  //  grp_bias [R >> grp_shift][N]: added to row r of C (and seen by the statistics and the pooling):
  //      the part of a layer's input that is constant over the rows of a group, multiplied through
  //      the weights once per group instead of once per row
  float *pool_max, *pool_min;
  unsigned char *pool_amax, *pool_amin;
  int pool_u;
  int store_c;
  const float *grp_bias;
  int grp_shift;
  int dbg;
};

// PAIR: launched as clusters of two CTAs that walk tiles 2i, 2i + 1 in lockstep; each CTA fetches half
// of every weight slab and multicasts it to both, so a slab crosses L2 -> shared memory once per PAIR
// of row tiles (at N x K >= 128 x 256 the weight stream, 2x the activation bytes per tile, is what
// bounds the kernel: profiles/r02_ncu_notes.md).  A stage is refilled when both CTAs have read it.
template <int MAXREG, bool PAIR>
__global__ void __maxnreg__(MAXREG)
gemm_nt_tma_kernel(const __grid_constant__ CUtensorMap tmA, GemmTmaParams p) {
  // No static shared memory and no alignment slack: the kernel's 224 KB + 256 B then leave room on
  // the SM for a small CTA of another kernel (the single-CTA-per-scene FPS launches of the input
  // pipeline), which otherwise pushes eight GEMM CTAs into a second wave.  The dynamic window of a
  // kernel without static shared memory starts 1 KB-aligned (checked below).
  extern __shared__ __align__(1024) unsigned char g_smem_al[];
  unsigned char *smem = g_smem_al;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bslab = p.npad * 128;                       // bytes of one B part-slab
  const int stage_bytes = 2 * G_ASLAB + 2 * bslab;      // A raw (= hi), A lo, B hi, B lo
  const int ntiles = (p.R + G_TILE - 1) / G_TILE;
  const unsigned smem_base = g_smem_u32(smem);
  if (smem_base & 1023u) __trap();
  const int crank = PAIR ? (int)g_cluster_ctarank() : 0;
  // first tile of this CTA (of its pair) and the tile stride; a pair's second CTA may run one tile past
  // the end (all-zero operand rows, nothing stored) so that both make the same number of steps
  const int tile0 = PAIR ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x;
  unsigned long long *ctrl = reinterpret_cast<unsigned long long *>(
      smem + (size_t)p.nstages * stage_bytes + (size_t)T_EPIW * 4096);
  unsigned long long *s_tma = ctrl, *s_full = ctrl + G_MAXSTAGES, *s_empty = ctrl + 2 * G_MAXSTAGES;
  unsigned long long *s_accf = ctrl + 3 * G_MAXSTAGES, *s_acce = s_accf + 2;
  unsigned &s_tmem = *reinterpret_cast<unsigned *>(s_acce + 2);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     g_smem_u32(&s_tmem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      g_mbar_init(g_smem_u32(&s_tma[s]), 1);
      g_mbar_init(g_smem_u32(&s_full[s]), 32 * T_XFW);
      g_mbar_init(g_smem_u32(&s_empty[s]), PAIR ? 2 : 1);
    }
    for (int a = 0; a < 2; ++a) {
      g_mbar_init(g_smem_u32(&s_accf[a]), 1);
      g_mbar_init(g_smem_u32(&s_acce[a]), 32 * T_EPIW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) g_cluster_sync();   // the peer's barriers exist before anything is multicast to them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;

  if (warp == T_PROD) {
    // ================================ producer ===============================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      unsigned it = 0;
      for (int tb = tile0; tb < ntiles; tb += gridDim.x) {
        const int tile = tb + crank;
        for (int ks = 0; ks < p.nslab; ++ks, ++it) {
          const int st = it % p.nstages;
          const unsigned ph = (it / p.nstages) & 1u;
          const unsigned sa = smem_base + (unsigned)(st * stage_bytes);
          const unsigned sb = sa + 2 * G_ASLAB;
          const unsigned bar = g_smem_u32(&s_tma[st]);
          g_mbar_wait(g_smem_u32(&s_empty[st]), ph ^ 1u);
          g_mbar_arm(bar, (unsigned)G_ASLAB + 2u * (unsigned)bslab);
          // box = 32 floats x 128 rows; columns >= K and rows >= R arrive as zeros
          g_tma_2d(sa, &tmA, ks * G_SLABK, tile * G_TILE, bar);
          if (PAIR) {
            const unsigned half = (unsigned)bslab >> 1, off = (unsigned)crank * half;
            g_bulk_g2s_mc(sb + off, p.Bimg + (size_t)ks * bslab + off, half, bar, 3);
            g_bulk_g2s_mc(sb + bslab + off, p.Bimg + (size_t)(p.nslab + ks) * bslab + off, half, bar, 3);
          } else {
            g_bulk_g2s(sb, p.Bimg + (size_t)ks * bslab, (unsigned)bslab, bar);
            g_bulk_g2s(sb + bslab, p.Bimg + (size_t)(p.nslab + ks) * bslab, (unsigned)bslab, bar);
          }
        }
      }
    }
  } else if (warp >= T_XF0 && warp < T_PROD) {
    // ================================ transform ==============================================
    const int xt = tid - 32 * T_XF0;  // 0..255
    unsigned it = 0;
    long long w_wait = 0, w_work = 0;
    for (int tb = tile0; tb < ntiles; tb += gridDim.x) {
      for (int ks = 0; ks < p.nslab; ++ks, ++it) {
        const int st = it % p.nstages;
        const unsigned ph = (it / p.nstages) & 1u;
        const unsigned sa = smem_base + (unsigned)(st * stage_bytes) + (unsigned)(xt * 16);
        const long long t0 = clock64();
        g_mbar_wait(g_smem_u32(&s_tma[st]), ph);
        const long long t1 = clock64();
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = g_lds128(sa + i * 4096);
        if (p.pro_scale) {
          // this thread's 16-byte position holds logical chunk (xt & 7) ^ (row & 7), and row & 7
          // = (xt >> 3) & 7 for all four of its rows: four fixed columns per slab
          const int col = ks * G_SLABK + 4 * ((xt & 7) ^ ((xt >> 3) & 7));
          const float4 sc = ldg4_guard(p.pro_scale, col, p.K), sh = ldg4_guard(p.pro_shift, col, p.K);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 hi, lo;
            bn_relu_split4(v[i], sc, sh, hi, lo);
            g_sts128(sa + i * 4096, hi);
            g_sts128(sa + G_ASLAB + i * 4096, lo);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) g_sts128(sa + G_ASLAB + i * 4096, lo_of_raw4(v[i]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        g_mbar_arrive(g_smem_u32(&s_full[st]));
        w_wait += t1 - t0;
        w_work += clock64() - t1;
      }
    }
    if ((p.dbg & 128) && blockIdx.x == 0 && xt == 0) {
      g_gemm_prof[0] = w_wait; g_gemm_prof[1] = w_work; g_gemm_prof[2] = 0; g_gemm_prof[3] = it;
    }
  } else if (warp == T_MMA) {
    // ================================ MMA issuer =============================================
    const unsigned idesc = g_idesc(128, p.npad);
    unsigned it = 0, tcount = 0;
    long long w_acce = 0, w_full = 0, w_issue = 0, t_begin = clock64();
    for (int tb = tile0; tb < ntiles; tb += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const unsigned d = tmem + (unsigned)(acc * 256);
      long long ta = clock64();
      g_mbar_wait(g_smem_u32(&s_acce[acc]), ((tcount >> 1) & 1u) ^ 1u);  // epilogue drained it
      w_acce += clock64() - ta;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < p.nslab; ++ks, ++it) {
        const int st = it % p.nstages;
        const unsigned ph = (it / p.nstages) & 1u;
        ta = clock64();
        g_mbar_wait(g_smem_u32(&s_full[st]), ph);
        const long long tb = clock64();
        w_full += tb - ta;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned a_hi = smem_base + (unsigned)(st * stage_bytes);
        const unsigned long long ah0 = g_desc(a_hi), al0 = ah0 + (G_ASLAB >> 4);
        const unsigned long long bh0 = al0 + (G_ASLAB >> 4), bl0 = bh0 + (unsigned long long)(bslab >> 4);
        if (g_elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // UMMA_K = 8 tf32 = 32 bytes = 2 descriptor units
            g_mma(d, ah0 + 2 * k, bh0 + 2 * k, idesc, (ks | k) ? 1u : 0u);
            g_mma(d, ah0 + 2 * k, bl0 + 2 * k, idesc, 1u);
            g_mma(d, al0 + 2 * k, bh0 + 2 * k, idesc, 1u);
          }
          // stage reusable once these MMAs have read it (in a pair: told to both CTAs)
          if (PAIR) g_commit_mc(g_smem_u32(&s_empty[st]), 3); else g_commit(g_smem_u32(&s_empty[st]));
          if (ks == p.nslab - 1) g_commit(g_smem_u32(&s_accf[acc]));
        }
        __syncwarp();
        w_issue += clock64() - tb;
      }
    }
    if ((p.dbg & 128) && blockIdx.x == 0 && lane == 0) {
      g_gemm_prof[4] = w_acce; g_gemm_prof[5] = w_full; g_gemm_prof[6] = w_issue;
      g_gemm_prof[7] = clock64() - t_begin;
    }
  } else if (warp < T_EPIW) {
    // ================================ epilogue ===============================================
    // thread = accumulator row; each warp transposes its 32 x 32 block through 4 KB of shared
    // memory so that a quarter-warp writes 128 contiguous bytes (see gemm_nt_3xtf32_kernel)
    unsigned tcount = 0;
    const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter (warp % 4), column interleave
    const bool vec = ((p.ldc & 3) == 0) && ((p.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    const unsigned stg = smem_base + (unsigned)(p.nstages * stage_bytes) + (unsigned)(warp * 4096);
    const int qr = lane >> 3, qc = lane & 7;   // read-back: row qr + 4i, 16-byte chunk qc
    // optional BatchNorm statistics of the OUTPUT: lane l sums column c0 + l of every 32 x 32 block
    // this warp drains (rows >= R are zero: TMA zero fill) -- up to four blocks (N = 256) per warp
    float cs1[4] = {0.f, 0.f, 0.f, 0.f}, cs2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tb = tile0; tb < ntiles; tb += gridDim.x, ++tcount) {
      const int tile = tb + crank;
      const int acc = tcount & 1;
      const long long gr = (long long)tile * G_TILE + q * 32 + lane;
      const long long te0 = clock64();
      g_mbar_wait(g_smem_u32(&s_accf[acc]), (tcount >> 1) & 1u);
      const long long te1 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        const int c0 = half * 32 + blk * 32 * (T_EPIW / 4);
        if (c0 >= p.npad) break;
        // rows >= R exist only in the last tile; with an operand prologue they are not zero
        const long long row0 = (long long)tile * G_TILE + q * 32;
        const long long left = (long long)p.R - row0;
        const int nvalid = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
        // group constants (grp_shift >= 4: rows 0-15 and 16-31 of the block lie in one group each): e0 / e1
        // for this lane's column in the statistics loop, el / eh for its 16-byte chunk in the store loop.
        // Requested BEFORE the accumulator load is issued: nothing may sit between tcgen05.ld and its
        // wait -- the compiler is free to move the asm statement's destination registers, which are not
        // valid until the wait (an earlier version did, and a few outputs per thousand runs were stale).
        float e0 = 0.f, e1 = 0.f;
        float4 el = make_float4(0.f, 0.f, 0.f, 0.f), eh = el;
        if (p.grp_bias) {
          const float *g0 = p.grp_bias + (row0 >> p.grp_shift) * p.N;
          const float *g1 = p.grp_bias + ((row0 + 16) >> p.grp_shift) * p.N;
          if (c0 + lane < p.N) {
            if (nvalid > 0) e0 = __ldg(g0 + c0 + lane);
            if (nvalid > 16) e1 = __ldg(g1 + c0 + lane);
          }
          if (c0 + qc * 4 < p.N) {
            if (nvalid > 0) el = __ldg(reinterpret_cast<const float4 *>(g0 + c0 + qc * 4));
            if (nvalid > 16) eh = __ldg(reinterpret_cast<const float4 *>(g1 + c0 + qc * 4));
          }
        }
        unsigned v[32];
        g_tmem_ld32(tmem + (unsigned)(acc * 256 + c0) + ((unsigned)(q * 32) << 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (vec) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            g_sts128(stg + (unsigned)(lane * 128 + ((j ^ (lane & 7)) << 4)),
                     make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
          __syncwarp();
          if ((p.col_stats && !p.bn_y) || p.pool_max) {
            float a1 = 0.f, a2 = 0.f;
            const unsigned wo = (unsigned)((lane & 3) << 2);
            const int ch = lane >> 2;
            const int col = c0 + lane;
            float mx0 = -INFINITY, mx1 = -INFINITY, mn0 = INFINITY, mn1 = INFINITY;
            int ix0 = 0, ix1 = 0, in0 = 0, in1 = 0;
            const bool want_min = p.pool_min != nullptr;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              float y;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y) : "r"(stg + (unsigned)(rr * 128 + ((ch ^ (rr & 7)) << 4)) + wo) : "memory");
              if (p.grp_bias) y += rr < 16 ? e0 : e1;
              if (rr < 16) {
                if (y > mx0) { mx0 = y; ix0 = rr; }
                if (want_min && y < mn0) { mn0 = y; in0 = rr; }
              } else {
                if (y > mx1) { mx1 = y; ix1 = rr - 16; }
                if (want_min && y < mn1) { mn1 = y; in1 = rr - 16; }
              }
              y = rr < nvalid ? y : 0.f;
              a1 += y;
              a2 = fmaf(y, y, a2);
            }
            cs1[blk] += a1;
            cs2[blk] += a2;
            if (p.pool_max && col < p.N) {
              if (p.pool_u == 16) {
                const long long u0 = (row0 >> 4) * p.N + col;
                if (nvalid >= 16) { p.pool_max[u0] = mx0; p.pool_amax[u0] = (unsigned char)ix0; }
                if (nvalid >= 32) { p.pool_max[u0 + p.N] = mx1; p.pool_amax[u0 + p.N] = (unsigned char)ix1; }
                if (want_min) {
                  if (nvalid >= 16) { p.pool_min[u0] = mn0; p.pool_amin[u0] = (unsigned char)in0; }
                  if (nvalid >= 32) { p.pool_min[u0 + p.N] = mn1; p.pool_amin[u0 + p.N] = (unsigned char)in1; }
                }
              } else if (nvalid >= 32) {   // one unit of 32 rows: strict comparisons keep the first row
                const long long u0 = (row0 >> 5) * p.N + col;
                if (mx1 > mx0) { mx0 = mx1; ix0 = ix1 + 16; }
                p.pool_max[u0] = mx0; p.pool_amax[u0] = (unsigned char)ix0;
                if (want_min) {
                  if (mn1 < mn0) { mn0 = mn1; in0 = in1 + 16; }
                  p.pool_min[u0] = mn0; p.pool_amin[u0] = (unsigned char)in0;
                }
              }
            }
          }
          const int cc = c0 + qc * 4;
          const long long grow = (long long)tile * G_TILE + q * 32 + qr;
          float *dst = p.C + grow * p.ldc + cc;
          if (p.bn_y) {
            // BatchNorm-backward statistics of the output g (see GemmTmaParams): the y tile is read in
            // the layout of the stores below (a quarter-warp covers 128 contiguous bytes of a row), a
            // lane sums its 8 rows of 4 columns, the 4 lanes that share those columns are combined by
            // two shuffles, and lane l keeps column c0 + 4 (l & 7) + (l >> 3)
            float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f), b2 = b1;
            const bool cok = cc < p.N;
            const float4 mu = ldg4_guard(p.bn_stats, cc, p.N), is = ldg4_guard(p.bn_stats + p.N, cc, p.N);
            const float4 sc = ldg4_guard(p.bn_stats + 2 * p.N, cc, p.N), sh = ldg4_guard(p.bn_stats + 3 * p.N, cc, p.N);
            const float *yb = p.bn_y + grow * p.ldy + cc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = qr + 4 * i;
              const float4 o = g_lds128(stg + (unsigned)(rr * 128 + ((qc ^ (rr & 7)) << 4)));
              if (grow + 4 * i < p.R && cok) {
                const float4 yv = __ldg(reinterpret_cast<const float4 *>(yb + (long long)(4 * i) * p.ldy));
                *reinterpret_cast<float4 *>(dst + (long long)(4 * i) * p.ldc) = o;
                const float gx = fmaf(yv.x, sc.x, sh.x) > 0.f ? o.x : 0.f, gy = fmaf(yv.y, sc.y, sh.y) > 0.f ? o.y : 0.f;
                const float gz = fmaf(yv.z, sc.z, sh.z) > 0.f ? o.z : 0.f, gw = fmaf(yv.w, sc.w, sh.w) > 0.f ? o.w : 0.f;
                b1.x += gx; b1.y += gy; b1.z += gz; b1.w += gw;
                b2.x = fmaf(gx, (yv.x - mu.x) * is.x, b2.x); b2.y = fmaf(gy, (yv.y - mu.y) * is.y, b2.y);
                b2.z = fmaf(gz, (yv.z - mu.z) * is.z, b2.z); b2.w = fmaf(gw, (yv.w - mu.w) * is.w, b2.w);
              }
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              b1.x += __shfl_xor_sync(0xffffffffu, b1.x, o); b1.y += __shfl_xor_sync(0xffffffffu, b1.y, o);
              b1.z += __shfl_xor_sync(0xffffffffu, b1.z, o); b1.w += __shfl_xor_sync(0xffffffffu, b1.w, o);
              b2.x += __shfl_xor_sync(0xffffffffu, b2.x, o); b2.y += __shfl_xor_sync(0xffffffffu, b2.y, o);
              b2.z += __shfl_xor_sync(0xffffffffu, b2.z, o); b2.w += __shfl_xor_sync(0xffffffffu, b2.w, o);
            }
            cs1[blk] += qr == 0 ? b1.x : (qr == 1 ? b1.y : (qr == 2 ? b1.z : b1.w));
            cs2[blk] += qr == 0 ? b2.x : (qr == 1 ? b2.y : (qr == 2 ? b2.z : b2.w));
          } else if (p.store_c) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = qr + 4 * i;
              float4 o = g_lds128(stg + (unsigned)(rr * 128 + ((qc ^ (rr & 7)) << 4)));
              if (grow + 4 * i < p.R && cc < p.N) {
                if (p.grp_bias) {
                  const float4 e = i < 4 ? el : eh;   // row qr + 4 i of the block: first / second 16 rows
                  o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
                }
                *reinterpret_cast<float4 *>(dst + (long long)(4 * i) * p.ldc) = o;
              }
            }
          }
          __syncwarp();
        } else if (gr < p.R) {
          float *crow = p.C + gr * p.ldc;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < p.N) crow[c0 + j] = __uint_as_float(v[j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      g_mbar_arrive(g_smem_u32(&s_acce[acc]));
      if ((p.dbg & 128) && blockIdx.x == 0 && tid == 0) {
        g_gemm_prof[8] += te1 - te0; g_gemm_prof[9] += clock64() - te1; g_gemm_prof[10] += 1;
      }
    }
    if (p.col_stats) {
      float *dst = p.col_stats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N;
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        // column owned by this lane: c0 + lane, or c0 + 4 (lane & 7) + (lane >> 3) in the bn_y mode
        const int c = half * 32 + blk * 32 * (T_EPIW / 4) + (p.bn_y ? 4 * (lane & 7) + (lane >> 3) : lane);
        if (c < p.N) {
          dst[c] = cs1[blk];
          dst[p.N + c] = cs2[blk];
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) g_cluster_sync();   // no CTA leaves while its peer may still signal its barriers
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                 : "memory");
}

// ----------------------------------------------------------------------------------------------
// Weight gradient  W'[n x k] = sum_r A[r, n] * B[r, k]  with MN-major (SW128_32B) operand tiles
// ----------------------------------------------------------------------------------------------
struct WgradTmaParams {
  int R, N, K;
  int kp;            // K rounded up to 32
  int kpart;         // B channels per CTA along gridDim.z (= kp when the K dimension is not split)
  int nstages;
  float *P;          // [gridDim.x][N][K] partial sums, one block per CTA
  int nchunks, chunk;
  const float *pro_scale, *pro_shift;  // optional [K]: B operand = relu(b * scale + shift)
  int dbg;
};

template <int MAXREG>
__global__ void __maxnreg__(MAXREG)
gemm_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      WgradTmaParams p) {
  extern __shared__ unsigned char g_smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(g_smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long s_tma[G_MAXSTAGES], s_full[G_MAXSTAGES], s_empty[G_MAXSTAGES],
      s_accf[2], s_acce[2];
  __shared__ unsigned s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // K split (gridDim.z > 1): a CTA owns `kpart` of the B channels.  Wide B tiles leave room for only two
  // pipeline stages next to the raw + lo copies (96 KB per stage at K = 256) and the TMA round trip is
  // then exposed; 128-channel parts fit three stages and two TMEM accumulators, at the price of loading
  // the A tile once per part.
  const int k0 = blockIdx.z * p.kpart;      // first channel of B handled by this CTA
  const int kpl = min(p.kpart, p.kp - k0);  // its (padded) channel count
  const int nacc = p.kpart <= 256 ? 2 : 1;  // TMEM accumulators (double-buffered when they fit)
  const int m0 = blockIdx.y * 128;          // first channel of A handled by this CTA
  const int ma = min(4, (p.N - m0 + 31) >> 5);  // 32-channel blocks of A that exist
  const int nb = kpl >> 5;                  // 32-channel blocks of B
  const int a_part = 128 * 128;             // A tile: 4 blocks x (32 rows x 128 B)
  const int b_part = p.kpart * 128;
  const int stage_bytes = 2 * a_part + 2 * b_part;   // A raw, A lo, B raw, B lo
  const int nslab = (p.R + 31) >> 5;        // reduction slabs of 32 rows
  const unsigned smem_base = g_smem_u32(smem);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     g_smem_u32(&s_tmem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      g_mbar_init(g_smem_u32(&s_tma[s]), 1);
      g_mbar_init(g_smem_u32(&s_full[s]), 32 * T_XFW);
      g_mbar_init(g_smem_u32(&s_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      g_mbar_init(g_smem_u32(&s_accf[a]), 1);
      g_mbar_init(g_smem_u32(&s_acce[a]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;

  if (warp == T_PROD) {
    // ================================ producer ===============================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      unsigned it = 0;
      for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
        const int slab_end = min(nslab, (chunk + 1) * p.chunk);
        for (int slab = chunk * p.chunk; slab < slab_end; ++slab, ++it) {
          const int st = it % p.nstages;
          const unsigned ph = (it / p.nstages) & 1u;
          const unsigned sa = smem_base + (unsigned)(st * stage_bytes);
          const unsigned sb = sa + 2u * (unsigned)a_part;
          const unsigned bar = g_smem_u32(&s_tma[st]);
          g_mbar_wait(g_smem_u32(&s_empty[st]), ph ^ 1u);
          g_mbar_arm(bar, (unsigned)(ma + nb) * 4096u);
          // box = 32 channels x 32 rows = one 4 KB column of SW128_32B atoms; rows >= R and
          // channels beyond the tensor arrive as zeros
          for (int m = 0; m < ma; ++m) g_tma_2d(sa + (unsigned)m * 4096u, &tmA, m0 + 32 * m, slab * 32, bar);
          for (int c = 0; c < nb; ++c) g_tma_2d(sb + (unsigned)c * 4096u, &tmB, k0 + 32 * c, slab * 32, bar);
        }
      }
    }
  } else if (warp >= T_XF0 && warp < T_PROD) {
    // ================================ transform ==============================================
    const int xt = tid - 32 * T_XF0;  // 0..255: one float4 of a 4 KB block
    unsigned it = 0;
    long long w_wait = 0, w_work = 0;
    for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
      const int slab_end = min(nslab, (chunk + 1) * p.chunk);
      for (int slab = chunk * p.chunk; slab < slab_end; ++slab, ++it) {
        const int st = it % p.nstages;
        const unsigned ph = (it / p.nstages) & 1u;
        const unsigned sa = smem_base + (unsigned)(st * stage_bytes) + (unsigned)(xt * 16);
        const unsigned sb = sa + 2u * (unsigned)a_part;
        const long long t0 = clock64();
        g_mbar_wait(g_smem_u32(&s_tma[st]), ph);
        const long long t1 = clock64();
        {
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < ma) v[i] = g_lds128(sa + i * 4096);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < ma) g_sts128(sa + (unsigned)a_part + i * 4096, lo_of_raw4(v[i]));
        }
        // position xt * 16 of a 4 KB block: row xt >> 3, 32-byte chunk ((xt >> 1) & 3) ^ (row & 3),
        // half xt & 1 -> the same four channels (relative to the block) for every block
        const int bch = (((xt >> 1) & 3) ^ ((xt >> 3) & 3)) * 8 + (xt & 1) * 4;
        for (int c = 0; c < nb; c += 4) {
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (c + i < nb) v[i] = g_lds128(sb + (c + i) * 4096);
          if (p.pro_scale) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (c + i < nb) {
                const int ch = k0 + (c + i) * 32 + bch;
                float4 hi, lo;
                bn_relu_split4(v[i], ldg4_guard(p.pro_scale, ch, p.K), ldg4_guard(p.pro_shift, ch, p.K), hi, lo);
                g_sts128(sb + (c + i) * 4096, hi);
                g_sts128(sb + (unsigned)b_part + (c + i) * 4096, lo);
              }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (c + i < nb) g_sts128(sb + (unsigned)b_part + (c + i) * 4096, lo_of_raw4(v[i]));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        g_mbar_arrive(g_smem_u32(&s_full[st]));
        w_wait += t1 - t0;
        w_work += clock64() - t1;
      }
    }
    if ((p.dbg & 128) && blockIdx.x == 0 && blockIdx.y == 0 && xt == 0) {
      g_gemm_prof[0] = w_wait; g_gemm_prof[1] = w_work; g_gemm_prof[2] = 0; g_gemm_prof[3] = it;
    }
  } else if (warp == T_MMA) {
    // ================================ MMA issuer =============================================
    unsigned it = 0, ccount = 0;
    long long w_acce = 0, w_full = 0, w_issue = 0, t_begin = clock64();
    const unsigned major_bits = (1u << 15) | (1u << 16);  // A and B MN-major
    const unsigned idesc0 = g_idesc(128, kpl <= 256 ? kpl : 256) | major_bits;
    for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x, ++ccount) {
      const int acc = nacc == 2 ? (int)(ccount & 1) : 0;
      const unsigned dbase = tmem + (unsigned)(acc * 256);
      const unsigned use = nacc == 2 ? (ccount >> 1) : ccount;
      long long ta = clock64();
      g_mbar_wait(g_smem_u32(&s_acce[acc]), (use & 1u) ^ 1u);   // epilogue has drained it
      w_acce += clock64() - ta;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int slab_end = min(nslab, (chunk + 1) * p.chunk);
      for (int slab = chunk * p.chunk; slab < slab_end; ++slab, ++it) {
        const int st = it % p.nstages;
        const unsigned ph = (it / p.nstages) & 1u;
        ta = clock64();
        g_mbar_wait(g_smem_u32(&s_full[st]), ph);
        const long long tb = clock64();
        w_full += tb - ta;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned a_hi = smem_base + (unsigned)(st * stage_bytes);
        const unsigned b_hi = a_hi + 2u * (unsigned)a_part;
        const unsigned long long ah0 = g_desc_mn(a_hi, 4096, 512), bh0 = g_desc_mn(b_hi, 4096, 512);
        const unsigned long long al0 = ah0 + (unsigned long long)(a_part >> 4);
        const unsigned long long bl0 = bh0 + (unsigned long long)(b_part >> 4);
        const bool first = slab == chunk * p.chunk;
        if (g_elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 8 reduction rows = two 512-byte K atoms = 64 units
            const unsigned long long o = (unsigned long long)(ks * 64);
            const unsigned accum = (first && ks == 0) ? 0u : 1u;
            if (kpl <= 256) {
              g_mma(dbase, ah0 + o, bh0 + o, idesc0, accum);
              g_mma(dbase, ah0 + o, bl0 + o, idesc0, 1u);
              g_mma(dbase, al0 + o, bh0 + o, idesc0, 1u);
            } else {
              for (int n0 = 0; n0 < kpl; n0 += 256) {
                const int nn = min(256, kpl - n0);
                const unsigned idesc = g_idesc(128, nn) | major_bits;
                const unsigned long long bn = o + (unsigned long long)(n0 * 8);  // n0 * 128 bytes >> 4
                const unsigned d = dbase + (unsigned)n0;
                g_mma(d, ah0 + o, bh0 + bn, idesc, accum);
                g_mma(d, ah0 + o, bl0 + bn, idesc, 1u);
                g_mma(d, al0 + o, bh0 + bn, idesc, 1u);
              }
            }
          }
          g_commit(g_smem_u32(&s_empty[st]));
          if (slab == slab_end - 1) g_commit(g_smem_u32(&s_accf[acc]));
        }
        __syncwarp();
        w_issue += clock64() - tb;
      }
    }
    if ((p.dbg & 128) && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) {
      g_gemm_prof[4] = w_acce; g_gemm_prof[5] = w_full; g_gemm_prof[6] = w_issue;
      g_gemm_prof[7] = clock64() - t_begin;
    }
  } else if (warp < 4) {
    // ================================ epilogue ===============================================
    const int n = m0 + warp * 32 + lane;
    unsigned ccount = 0;
    for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x, ++ccount) {
      const int acc = nacc == 2 ? (int)(ccount & 1) : 0;
      const unsigned use = nacc == 2 ? (ccount >> 1) : ccount;
      const long long te0 = clock64();
      g_mbar_wait(g_smem_u32(&s_accf[acc]), use & 1u);
      const long long te1 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // One partial block per CTA: the chunks a CTA processes are added (round-to-nearest fp32,
      // fixed order) into its own block, which stays in L2 between chunks; the thread that owns
      // an element is the only one that ever touches it.
      float *prow = p.P + ((size_t)blockIdx.x * p.N + n) * p.K + k0;   // this CTA's column range
      const int kleft = p.K - k0;                                        // columns that exist from k0 on
      const bool addto = ccount > 0;
      for (int c0 = 0; c0 < kpl; c0 += 32) {
        unsigned v[32];
        g_tmem_ld32(tmem + (unsigned)(acc * 256 + c0) + ((unsigned)(warp * 32) << 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (n < p.N) {
          if ((p.K & 3) == 0) {
            float4 old[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              old[j] = (addto && c0 + 4 * j < kleft) ? *reinterpret_cast<const float4 *>(prow + c0 + 4 * j)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + 4 * j < kleft)
                *reinterpret_cast<float4 *>(prow + c0 + 4 * j) =
                    make_float4(old[j].x + __uint_as_float(v[4 * j]), old[j].y + __uint_as_float(v[4 * j + 1]),
                                old[j].z + __uint_as_float(v[4 * j + 2]), old[j].w + __uint_as_float(v[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < kleft) prow[c0 + j] = (addto ? prow[c0 + j] : 0.f) + __uint_as_float(v[j]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      g_mbar_arrive(g_smem_u32(&s_acce[acc]));
      if ((p.dbg & 128) && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        g_gemm_prof[8] += te1 - te0; g_gemm_prof[9] += clock64() - te1; g_gemm_prof[10] += 1;
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                 : "memory");
}

// ---- host side: tensor maps through the driver entry point (libcuda is not linked) -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// fp32 matrix (rows x cols, row stride ld floats) as a 2-D tensor; box = box_cols x box_rows
inline bool make_tmap(CUtensorMap *tm, const float *base, long long rows, long long cols, long long ld,
                      int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3)) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shared memory the TMA kernels may use for stages (+ epilogue staging); NESIE_GEMM_SMEM_KB overrides
inline size_t gemm_smem_budget() {
  const char *e = getenv("NESIE_GEMM_SMEM_KB");
  return e ? (size_t)atoi(e) * 1024 : (size_t)G_SMEM_BUDGET;
}

// CTAs of a persistent GEMM launch: all SMs, or NESIE_GEMM_GRID (e.g. 140 leaves eight SMs to the
// single-CTA-per-scene FPS kernels of the input pipeline so that they do not force a second wave)
inline int gemm_grid_sms() {
  static int g = -1;
  if (g < 0) {
    const char *e = getenv("NESIE_GEMM_GRID");
    g = e ? atoi(e) : num_sms();
    if (g < 1 || g > num_sms()) g = num_sms();
  }
  return g;
}

// CTAs of a persistent launch over `ntiles` equal tiles: the smallest grid that still finishes in the same
// number of waves (512 tiles on 148 SMs are 4 waves with 148 CTAs and with 128): the launch is no slower
// and the SMs it leaves alone run the small kernels of the other branches of the step.
// NESIE_GEMM_BALANCE=0 always takes every SM.
inline int gemm_balanced_grid(long long ntiles) {
  static int on = -1;
  if (on < 0) { const char *e = getenv("NESIE_GEMM_BALANCE"); on = (e && e[0] == '0') ? 0 : 1; }
  const int sms = gemm_grid_sms();
  if (ntiles <= sms) return (int)ntiles;
  if (!on) return sms;
  const long long waves = (ntiles + sms - 1) / sms;
  return (int)((ntiles + waves - 1) / waves);
}

// register cap of the build to launch: 64, 88 or 96 (NESIE_GEMM_REGS; default 96)
inline int gemm_tma_regs() {
  const char *e = getenv("NESIE_GEMM_REGS");
  const int r = e ? atoi(e) : 96;
  return r <= 64 ? 64 : (r <= 88 ? 88 : 96);
}

// NESIE_GEMM_PAIR=1 launches the wide NT GEMMs as CTA pairs with multicast weight slabs.  Off by default:
// measured on a B200 it halves the L2 -> shared-memory weight traffic but not the time (65536 x 256 -> 128:
// 38.4 us against 37.2 us; 262144 x 128 -> 256: 104 against 102; pretrain step 14.80 against 14.73 ms) --
// the kernel waits on the LATENCY of its two or three 64-96 KB stages, not on L2 bandwidth, and the pair's
// lockstep adds a little (profiles/r02_ncu_notes.md).
inline bool gemm_pair_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("NESIE_GEMM_PAIR"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

inline void gemm_apply_env_once() {
  static bool done = false;
  if (done) return;
  done = true;
  if (const char *e = getenv("NESIE_GEMM_BACKOFF_NS")) {
    const unsigned v = (unsigned)atoi(e);
    cudaMemcpyToSymbol(g_wait_backoff_ns, &v, sizeof(v));
  }
}

inline bool gemm_tma_enabled() {
  gemm_apply_env_once();
  const char *e = getenv("NESIE_GEMM_PATH");
  return !(e && e[0] == 'r');  // NESIE_GEMM_PATH=reg selects the register-staged loaders
}

}  // namespace
}  // namespace nesie
