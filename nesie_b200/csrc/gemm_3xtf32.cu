// fp32-accurate row GEMM on the 5th-generation tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
//   C[R x N] = A[R x K] * B[N x K]^T        A, B, C fp32;  R ~ 10^5..10^6 rows,  N, K <= 288
//
// This is the shared-MLP GEMM of the SA / FP modules in TRAINING (the reference runs it as cuDNN
// 1x1 Conv2d, ops/pointnet_modules/point_sa_module.py:279-288): forward Y = X W^T and the data
// gradient dX = dY W.  Parity with the fp32 reference has to hold at 1e-5, which a single TF32 MMA
// (10-bit mantissa) cannot give, so every operand is split x = hi + lo with hi = x truncated to TF32
// and the product is accumulated as  hi*hi + hi*lo + lo*hi  in the fp32 TMEM accumulator (the
// dropped lo*lo term is 2^-22 relative).  That is 3 tensor-core MMAs per fp32 MMA -- still an order
// of magnitude above the SIMT sgemm cuBLAS selects for strict fp32.
//
// Structure (one persistent CTA per SM, 672 threads):
//   warps 4-19 loaders : four groups of 128 threads, each filling every fourth pipeline stage; A tile rows (fp32, coalesced float4) -> hi/lo -> UMMA K-major SWIZZLE_128B
//                        slabs of 32 K-elements; the pre-split, pre-swizzled B image arrives by one
//                        cp.async.bulk per part, counted on the same mbarrier
//   warp  20   MMA     : 4 K-steps x 3 tcgen05.mma per slab, tcgen05.commit frees the stage; the
//                        last commit of a tile publishes the accumulator
//   warps 0-3  epilogue: tcgen05.ld -> float4 stores to C; two TMEM accumulators so the epilogue of
//                        tile t overlaps the MMAs of tile t+1
#include "gemm_common.cuh"

namespace nesie {
namespace {

// cycle counters of CTA 0 (NESIE_GEMM_DBG bit 128), read back by nesie_gemm_debug_profile
__device__ long long g_gemm_prof[16];

}  // namespace
}  // namespace nesie

#include "gemm_tma.cuh"

namespace nesie {
namespace {

__global__ void __launch_bounds__(G_THREADS, 1) gemm_nt_3xtf32_kernel(GemmParams p) {
  extern __shared__ unsigned char g_smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(g_smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long s_full[G_MAXSTAGES], s_empty[G_MAXSTAGES], s_accf[2], s_acce[2];
  __shared__ unsigned s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bslab = p.npad * 128;                       // bytes of one B part-slab
  const int stage_bytes = 2 * G_ASLAB + 2 * bslab;      // A hi, A lo, B hi, B lo
  const int ntiles = (p.R + G_TILE - 1) / G_TILE;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     g_smem_u32(&s_tmem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      g_mbar_init(g_smem_u32(&s_full[s]), 128 + 1);  // 128 loader arrivals + the expect_tx arrive
      g_mbar_init(g_smem_u32(&s_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      g_mbar_init(g_smem_u32(&s_accf[a]), 1);
      g_mbar_init(g_smem_u32(&s_acce[a]), 32 * G_EPIW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;

  if (warp >= G_EPIW && warp < G_EPIW + 4 * G_LGROUPS) {
    // ================================ loaders ================================================
    // Bandwidth = bytes in flight / latency: one group only keeps 16 KB of loads in flight per SM
    // (~2.4 TB/s chip-wide); four groups working on four consecutive stages keep 64 KB.
    const int lg = (warp - G_EPIW) >> 2;
    const int lt = (tid - 32 * G_EPIW) & 127;
    const unsigned smem_base = g_smem_u32(smem);
    unsigned it = 0;  // global stage counter
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = (long long)tile * G_TILE;
      for (int ks = 0; ks < p.nslab; ++ks, ++it) {
        // a group owns pipeline stage `lg` outright (groups beyond the stage count stay idle):
        // mbarrier parity waits cannot tell phases two apart, so a stage must never be shared
        if (lg >= p.nstages || (int)(it % p.nstages) != lg) continue;
        const int st = it % p.nstages;
        const unsigned ph = (it / p.nstages) & 1u;
        const unsigned sa_hi = smem_base + (unsigned)(st * stage_bytes);
        const unsigned sa_lo = sa_hi + G_ASLAB;
        const unsigned sb = sa_lo + G_ASLAB;
        const long long t0 = clock64();
        g_mbar_wait(g_smem_u32(&s_empty[st]), ph ^ 1u);
        const long long t1 = clock64();
        if (lt == 0 && (p.dbg & 8)) g_mbar_arrive(g_smem_u32(&s_full[st]));
        if (lt == 0 && !(p.dbg & 8)) {
          g_mbar_expect_tx(g_smem_u32(&s_full[st]), 2u * (unsigned)bslab);
          g_bulk_g2s(sb, p.Bimg + (size_t)ks * bslab, (unsigned)bslab, g_smem_u32(&s_full[st]));
          g_bulk_g2s(sb + bslab, p.Bimg + (size_t)(p.nslab + ks) * bslab, (unsigned)bslab,
                     g_smem_u32(&s_full[st]));
        }
        // A slab: 128 rows x 8 chunks of 4 floats; thread lt owns chunk (lt & 7) of rows lt>>3 + 16j.
        // r & 7 is the same for all eight rows, so the swizzled offsets differ by j * 2048 only.
        const int c = lt & 7, rb = lt >> 3;
        const int k0 = ks * G_SLABK + c * 4;
        const unsigned soff = (unsigned)(rb * 128 + ((c ^ (rb & 7)) << 4));
        float4 v[8];
        if (p.fast) {
          // rows are 16-byte aligned and K is a multiple of 4: a chunk is loaded whole or not at all
          if (k0 < p.K && !(p.dbg & 1)) {
            if (row0 + G_TILE <= p.R) {
              const float *src = p.A + (row0 + rb) * p.lda + k0;
              const long long step = 16 * p.lda;
#pragma unroll
              for (int j = 0; j < 8; ++j, src += step) v[j] = __ldg(reinterpret_cast<const float4 *>(src));
            } else {  // ragged last tile: clamp the row (rows >= R are never stored)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                long long gr = row0 + rb + 16 * j;
                gr = gr < p.R ? gr : (long long)p.R - 1;
                v[j] = __ldg(reinterpret_cast<const float4 *>(p.A + gr * p.lda + k0));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const long long gr = row0 + rb + 16 * j;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < p.R) {
              const float *src = p.A + gr * p.lda + k0;
              if (k0 + 0 < p.K) v[j].x = __ldg(src + 0);
              if (k0 + 1 < p.K) v[j].y = __ldg(src + 1);
              if (k0 + 2 < p.K) v[j].z = __ldg(src + 2);
              if (k0 + 3 < p.K) v[j].w = __ldg(src + 3);
            }
          }
        }
        if (!(p.dbg & 16))
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 hi, lo;
          split_tf32_fast(v[j].x, hi.x, lo.x);
          split_tf32_fast(v[j].y, hi.y, lo.y);
          split_tf32_fast(v[j].z, hi.z, lo.z);
          split_tf32_fast(v[j].w, hi.w, lo.w);
          g_sts128(sa_hi + soff + j * 2048, hi);
          g_sts128(sa_lo + soff + j * 2048, lo);
        }
        const long long t2 = clock64();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        g_mbar_arrive(g_smem_u32(&s_full[st]));
        if ((p.dbg & 128) && blockIdx.x == 0 && lt == 0 && lg == 0) {
          g_gemm_prof[0] += t1 - t0; g_gemm_prof[1] += t2 - t1; g_gemm_prof[2] += clock64() - t2;
          g_gemm_prof[3] += 1;
        }
      }
    }
  } else if (warp == G_EPIW + 4 * G_LGROUPS) {
    // ================================ MMA issuer =============================================
    // The whole warp runs the loop (uniform control flow keeps descriptors in uniform registers:
    // issued from inside an `if (lane == 0)` region every MMA cost an ELECT + 5 R2UR round trip);
    // one elected lane issues the MMAs and commits.
    {
      const unsigned idesc = g_idesc(128, p.npad);
      unsigned it = 0, tcount = 0;
      long long w_acce = 0, w_full = 0, w_issue = 0, t_begin = clock64();
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
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
          const unsigned a_hi = g_smem_u32(smem) + (unsigned)(st * stage_bytes);
          const unsigned long long ah0 = g_desc(a_hi), al0 = ah0 + (G_ASLAB >> 4);
          const unsigned long long bh0 = al0 + (G_ASLAB >> 4), bl0 = bh0 + (unsigned long long)(bslab >> 4);
          if (g_elect_one()) {
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // UMMA_K = 8 tf32 = 32 bytes = 2 descriptor units
                g_mma(d, ah0 + 2 * k, bh0 + 2 * k, idesc, (ks | k) ? 1u : 0u);
                g_mma(d, ah0 + 2 * k, bl0 + 2 * k, idesc, 1u);
                g_mma(d, al0 + 2 * k, bh0 + 2 * k, idesc, 1u);
              }
            }
            if (p.dbg & 64) g_mbar_arrive(g_smem_u32(&s_empty[st]));
            else g_commit(g_smem_u32(&s_empty[st]));  // stage reusable once these MMAs have read it
            if (ks == p.nslab - 1) {
              if (p.dbg & 64) g_mbar_arrive(g_smem_u32(&s_accf[acc]));
              else g_commit(g_smem_u32(&s_accf[acc]));    // accumulator complete
            }
          }
          __syncwarp();
          w_issue += clock64() - tb;
        }
      }
      if ((p.dbg & 128) && blockIdx.x == 0 && lane == 0) {
        g_gemm_prof[4] = w_acce; g_gemm_prof[5] = w_full; g_gemm_prof[6] = w_issue;
        g_gemm_prof[7] = clock64() - t_begin;
      }
    }
  } else if (warp < G_EPIW) {
    // ================================ epilogue ===============================================
    unsigned tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter (fixed by warp % 4), column interleave
      const long long gr = (long long)tile * G_TILE + q * 32 + lane;
      const long long te0 = clock64();
      g_mbar_wait(g_smem_u32(&s_accf[acc]), (tcount >> 1) & 1u);
      const long long te1 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float *crow = p.C + gr * p.ldc;
      const bool vec = ((p.ldc & 3) == 0) && ((p.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
      // A thread owns an accumulator ROW, so storing straight from registers scatters every warp
      // store over 32 rows (32 half-filled sectors per instruction; the LSU time of those stores
      // starved the loaders).  Instead each warp transposes its 32 x 32 block through 4 KB of
      // shared memory and writes 128 contiguous bytes per quarter-warp.
      const unsigned stg = g_smem_u32(smem) + (unsigned)(p.nstages * stage_bytes) + (unsigned)(warp * 4096);
      const int qr = lane >> 3, qc = lane & 7;   // read-back: row qr + 4i, 16-byte chunk qc
      for (int c0 = half * 32; c0 < p.npad; c0 += 32 * (G_EPIW / 4)) {
        unsigned v[32];
        if (p.dbg & 32) break;
        g_tmem_ld32(tmem + (unsigned)(acc * 256 + c0) + ((unsigned)(q * 32) << 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (p.dbg & 2) continue;
        if (vec) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            g_sts128(stg + (unsigned)(lane * 128 + ((j ^ (lane & 7)) << 4)),
                     make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
          __syncwarp();
          const int cc = c0 + qc * 4;
          const long long grow = (long long)tile * G_TILE + q * 32 + qr;
          float *dst = p.C + grow * p.ldc + cc;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = qr + 4 * i;
            const float4 o = g_lds128(stg + (unsigned)(rr * 128 + ((qc ^ (rr & 7)) << 4)));
            if (grow + 4 * i < p.R && cc < p.N) *reinterpret_cast<float4 *>(dst + (long long)(4 * i) * p.ldc) = o;
          }
          __syncwarp();
        } else if (gr < p.R) {
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
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                 : "memory");
}

// =============================================================================================
// Weight gradient:  W'[n x k] = sum_r A[r, n] * B[r, k]    (A = dY (R x N), B = X (R x K))
// The reduction runs over the ROWS, i.e. along the non-contiguous dimension of both operands.
// kind::tf32 with MN-major descriptors (a_major = b_major = 1) returned all-zero accumulators on
// this B200 whatever LBO/SBO were tried, so the loaders TRANSPOSE instead: a lane owns one reduction
// row, reads float4s of it and scatters the four elements into four K-major operand rows (32
// lanes -> 32 distinct words of a 128-byte row: conflict-free), SWIZZLE_128B as in the NT kernel.
// 3xTF32 split as above.  Each CTA accumulates its share of the row slabs in TMEM (split-R) and writes one partial
// [n x k] block; the caller sums the partials (deterministic, no atomics).
// =============================================================================================
struct WgradParams {
  int R, N, K;
  int kp;            // K rounded up to 32 (MN-block granularity)
  int nstages;
  long long lda, ldb;
  const float *A, *B;
  float *P;          // [gridDim.x][N][K] partial sums, one block per CTA
  int nchunks;       // row chunks of `chunk` slabs (one TMEM accumulation each)
  int chunk;         // slabs (of 32 rows) accumulated per TMEM accumulator
  int dbg;           // NESIE_GEMM_DBG & 128: cycle counters of CTA 0
  int mn;            // bit 0: B tile MN-major, bit 1: A tile MN-major (straight float4 copies);
                     // a clear bit selects the transposing loader and a K-major tile
  int vec;           // rows 16-byte aligned, channel counts multiples of 4
};

// The tensor core's fp32 accumulation truncates: the error of a serial in-TMEM reduction grows
// linearly with its length (measured vs float64: 3e-6 after 14 slabs, 1.3e-5 after 64, 5e-5 after
// 220).  So an accumulator only ever sums <= 32 slabs (1024 rows); the partial blocks are then
// added in ordinary round-to-nearest fp32 by the caller.
constexpr int W_CHUNK_MAX = 32, W_CHUNK_MIN = 8;
inline int wgrad_chunk(long long nslab, int mblocks) {
  long long c = nslab / (2LL * (num_sms() / mblocks > 0 ? num_sms() / mblocks : 1));
  if (c > W_CHUNK_MAX) c = W_CHUNK_MAX;
  if (c < W_CHUNK_MIN) c = W_CHUNK_MIN;
  return (int)c;
}

constexpr int W_THREADS = 128 + 128 * G_LGROUPS + 32;

__global__ void __launch_bounds__(W_THREADS, 1) gemm_wgrad_3xtf32_kernel(WgradParams p) {
  extern __shared__ unsigned char g_smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(g_smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long s_full[G_MAXSTAGES], s_empty[G_MAXSTAGES], s_accf[2], s_acce[2];
  __shared__ unsigned s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nacc = p.kp <= 256 ? 2 : 1;     // TMEM accumulators (double-buffered when they fit)
  const int a_part = 128 * 128;             // 128 operand rows (channels) x 32 reduction elements
  const int b_part = p.kp * 128;            // kp operand rows x 32 reduction elements
  const int stage_bytes = 2 * a_part + 2 * b_part;
  const int m0 = blockIdx.y * 128;          // first output row (channel of A) of this CTA
  const int nslab = (p.R + 31) >> 5;        // reduction slabs of 32 rows

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     g_smem_u32(&s_tmem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      g_mbar_init(g_smem_u32(&s_full[s]), 128);
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

  if (warp >= 4 && warp < 4 + 4 * G_LGROUPS) {
    // ================================ loaders ================================================
    const int lg = (warp - 4) >> 2;
    const int lt = (tid - 128) & 127;
    const int kq = p.kp >> 2;  // float4 columns of a B row
    unsigned it = 0;
    for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x)
    for (int slab = chunk * p.chunk; slab < min(nslab, (chunk + 1) * p.chunk); ++slab, ++it) {
      if (lg >= p.nstages || (int)(it % p.nstages) != lg) continue;  // a group owns its stage
      const int st = lg;
      const unsigned ph = (it / p.nstages) & 1u;
      unsigned char *sa_hi = smem + (size_t)st * stage_bytes;
      unsigned char *sa_lo = sa_hi + a_part;
      unsigned char *sb_hi = sa_lo + a_part;
      unsigned char *sb_lo = sb_hi + b_part;
      const long long r0 = (long long)slab * 32;
      const long long t0 = clock64();
      g_mbar_wait(g_smem_u32(&s_empty[st]), ph ^ 1u);
      const long long t1 = clock64();
      const int lw = lt >> 5, ll = lt & 31;       // warp of the group, lane
      const unsigned sa = g_smem_u32(sa_hi), sbb = g_smem_u32(sb_hi);
      // MN-major tiles: the channels of a reduction row are contiguous in memory and in the
      // operand tile, so a warp store covers one 512-byte atom (4 rows x 32 channels) with plain
      // float4 copies: lane = 16-byte chunk (ll & 7) of row (ll >> 3).
      const int c16 = ll & 7, j4 = ll >> 3;
      const unsigned sw = (unsigned)(j4 * 128 + (((c16 >> 1) ^ j4) << 5) + ((c16 & 1) << 4));
      // Transposing (K-major) tiles: lane <-> reduction row (r0 + lane), so the four scalar stores
      // of a float4 go to four operand rows (channels) at 32 distinct words: conflict-free.
      const long long gr = r0 + ll;
      const bool rok = gr < p.R;
      const unsigned tw = (unsigned)(((ll >> 2) << 4) + ((ll & 3) << 2));  // chunk ll>>2, word ll&3
      // MN-major B tile: (32-channel block, 4-row group) pairs, four per warp and batch.  The first
      // batch is requested together with the A tile so that a slab costs one memory round trip.
      const int natoms = (p.kp >> 5) * 8;
      auto load_b = [&](int t0, float4 (&v)[4]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = t0 + 4 * u;
          const int ch = (idx >> 3) * 32 + c16 * 4;
          long long g2 = r0 + (idx & 7) * 4 + j4;
          g2 = g2 < p.R ? g2 : (long long)p.R - 1;  // A is zero there; any finite value will do
          v[u] = (idx < natoms && ch < p.K) ? __ldg(reinterpret_cast<const float4 *>(p.B + g2 * p.ldb + ch))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 vb[4], vb_next[4];
      if (p.mn & 1) load_b(lw, vb);
      // ---- A tile: 128 channels (m0..) x 32 reduction rows
      if (p.mn & 2) {
        float4 v[8];
        const int ch = m0 + lw * 32 + c16 * 4;
        const float *src = p.A + (r0 + j4) * p.lda + ch;
        const long long step = 4 * p.lda;
#pragma unroll
        for (int i = 0; i < 8; ++i, src += step)
          v[i] = (r0 + i * 4 + j4 < p.R && ch < p.N) ? __ldg(reinterpret_cast<const float4 *>(src))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 hi, lo;
          split_tf32_fast(v[i].x, hi.x, lo.x);
          split_tf32_fast(v[i].y, hi.y, lo.y);
          split_tf32_fast(v[i].z, hi.z, lo.z);
          split_tf32_fast(v[i].w, hi.w, lo.w);
          const unsigned off = (unsigned)(lw * 4096 + i * 512) + sw;
          g_sts128(sa + off, hi);
          g_sts128(sa + (unsigned)a_part + off, lo);
        }
      } else {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ch = m0 + lw * 32 + i * 4;
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rok && ch < p.N) {
            const float *src = p.A + gr * p.lda + ch;
            if (p.vec) {
              v[i] = __ldg(reinterpret_cast<const float4 *>(src));
            } else {
              v[i].x = __ldg(src);
              if (ch + 1 < p.N) v[i].y = __ldg(src + 1);
              if (ch + 2 < p.N) v[i].z = __ldg(src + 2);
              if (ch + 3 < p.N) v[i].w = __ldg(src + 3);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = lw * 32 + i * 4 + j;  // operand row; n & 7 == (i * 4 + j) & 7
            float hi, lo;
            split_tf32_fast(e[j], hi, lo);
            const unsigned off = (unsigned)(n * 128) + (tw ^ (unsigned)(((i * 4 + j) & 7) << 4));
            g_sts32(sa + off, hi);
            g_sts32(sa + (unsigned)a_part + off, lo);
          }
        }
      }
      // ---- B tile: kp channels x 32 reduction rows
      if (p.mn & 1) {
        for (int t0 = lw; t0 < natoms; t0 += 16) {
          if (t0 + 16 < natoms) load_b(t0 + 16, vb_next);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx = t0 + 4 * u;
            if (idx < natoms) {
              float4 hi, lo;
              split_tf32_fast(vb[u].x, hi.x, lo.x);
              split_tf32_fast(vb[u].y, hi.y, lo.y);
              split_tf32_fast(vb[u].z, hi.z, lo.z);
              split_tf32_fast(vb[u].w, hi.w, lo.w);
              const unsigned off = (unsigned)((idx >> 3) * 4096 + (idx & 7) * 512) + sw;
              g_sts128(sbb + off, hi);
              g_sts128(sbb + (unsigned)b_part + off, lo);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) vb[u] = vb_next[u];
        }
      } else {
        for (int c4 = lw; c4 < kq; c4 += 4) {
          const int k0 = c4 * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rok && k0 < p.K) {
            const float *src = p.B + gr * p.ldb + k0;
            if (p.vec) {
              v = __ldg(reinterpret_cast<const float4 *>(src));
            } else {
              v.x = __ldg(src);
              if (k0 + 1 < p.K) v.y = __ldg(src + 1);
              if (k0 + 2 < p.K) v.z = __ldg(src + 2);
              if (k0 + 3 < p.K) v.w = __ldg(src + 3);
            }
          }
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = k0 + j;
            float hi, lo;
            split_tf32_fast(e[j], hi, lo);
            const unsigned off = (unsigned)(n * 128) + (tw ^ (unsigned)((n & 7) << 4));
            g_sts32(sbb + off, hi);
            g_sts32(sbb + (unsigned)b_part + off, lo);
          }
        }
      }
      const long long t2 = clock64();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      g_mbar_arrive(g_smem_u32(&s_full[st]));
      if ((p.dbg & 128) && blockIdx.x == 0 && blockIdx.y == 0 && lt == 0 && lg == 0) {
        g_gemm_prof[0] += t1 - t0; g_gemm_prof[1] += t2 - t1; g_gemm_prof[2] += clock64() - t2;
        g_gemm_prof[3] += 1;
      }
    }
  } else if (warp == 4 + 4 * G_LGROUPS) {
    // ================================ MMA issuer =============================================
    {  // whole warp, uniform control flow; one elected lane issues (see the NT kernel)
      unsigned it = 0, ccount = 0;
      long long w_acce = 0, w_full = 0, w_issue = 0, t_begin = clock64();
      // per k-step (8 reduction rows) descriptor increments, in 16-byte units: two 512-byte K atoms
      // of an MN-major tile, or 32 bytes inside the swizzle row of a K-major one
      const int a_kstep = (p.mn & 2) ? 64 : 2, b_kstep = (p.mn & 1) ? 64 : 2;
      const unsigned major_bits = ((p.mn & 2) ? (1u << 15) : 0u) | ((p.mn & 1) ? (1u << 16) : 0u);
      const unsigned idesc0 = g_idesc(128, p.kp <= 256 ? p.kp : 256) | major_bits;
      for (int chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x, ++ccount) {
        const int acc = nacc == 2 ? (int)(ccount & 1) : 0;
        const unsigned dbase = tmem + (unsigned)(acc * 256);
        const unsigned use = nacc == 2 ? (ccount >> 1) : ccount;  // how often this accumulator was used
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
          // The descriptors of a stage differ from a base only in the start-address field (bits
          // 0-13, no carry out for < 256 KB of shared memory): one add per operand and MMA.
          const unsigned a_hi = g_smem_u32(smem) + (unsigned)(st * stage_bytes);
          const unsigned b_hi = a_hi + 2u * (unsigned)a_part;
          const unsigned long long ah0 = (p.mn & 2) ? g_desc_mn(a_hi, 4096, 512) : g_desc(a_hi);
          const unsigned long long bh0 = (p.mn & 1) ? g_desc_mn(b_hi, 4096, 512) : g_desc(b_hi);
          const unsigned long long al0 = ah0 + (unsigned long long)(a_part >> 4);
          const unsigned long long bl0 = bh0 + (unsigned long long)(b_part >> 4);
          const bool first = slab == chunk * p.chunk;
          if (g_elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // UMMA_K = 8 reduction rows
              const unsigned long long ao = (unsigned long long)(ks * a_kstep);
              const unsigned long long bo = (unsigned long long)(ks * b_kstep);
              const unsigned accum = (first && ks == 0) ? 0u : 1u;
              if (p.kp <= 256) {
                g_mma(dbase, ah0 + ao, bh0 + bo, idesc0, accum);
                g_mma(dbase, ah0 + ao, bl0 + bo, idesc0, 1u);
                g_mma(dbase, al0 + ao, bh0 + bo, idesc0, 1u);
              } else {
                for (int n0 = 0; n0 < p.kp; n0 += 256) {
                  const int nn = min(256, p.kp - n0);
                  const unsigned idesc = g_idesc(128, nn) | major_bits;
                  const unsigned long long bn = bo + (unsigned long long)(n0 * 8);  // n0 * 128 bytes >> 4
                  const unsigned d = dbase + (unsigned)n0;
                  g_mma(d, ah0 + ao, bh0 + bn, idesc, accum);
                  g_mma(d, ah0 + ao, bl0 + bn, idesc, 1u);
                  g_mma(d, al0 + ao, bh0 + bn, idesc, 1u);
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
      float *prow = p.P + ((size_t)blockIdx.x * p.N + n) * p.K;  // one partial block per CTA
      const bool addto = ccount > 0;
      for (int c0 = 0; c0 < p.kp; c0 += 32) {
        unsigned v[32];
        g_tmem_ld32(tmem + (unsigned)(acc * 256 + c0) + ((unsigned)(warp * 32) << 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < p.K) prow[c0 + j] = (addto ? prow[c0 + j] : 0.f) + __uint_as_float(v[j]);
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

// B (N x K, element (n,k) at B[n*sn + k*sk]) -> image [hi|lo][nslab][npad][128 B], fp32 split,
// 16-byte chunk c of row n at chunk position c ^ (n & 7); zero padding for n >= N, k >= K.
__global__ void __launch_bounds__(256) pack_b_tf32_kernel(int N, int K, int npad, int nslab,
                                                          long long sn, long long sk,
                                                          const float *__restrict__ B,
                                                          float *__restrict__ img) {
  const int total = nslab * npad * 32;  // elements per part
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 31, n = (i >> 5) % npad, s = (i >> 5) / npad;
    const int k = s * 32 + e;
    const float x = (n < N && k < K) ? B[n * sn + k * sk] : 0.f;
    float hi, lo;
    split_tf32(x, hi, lo);
    const int chunk = e >> 2, w = e & 3;
    const size_t off = ((size_t)s * npad + n) * 32 + (size_t)(((chunk ^ (n & 7)) << 2) + w);
    img[off] = hi;
    img[(size_t)total + off] = lo;
  }
}

// out[i] = sum_p partials[p][i]: blockDim = (32 float4 columns, SP_SLICES slices of the partial
// blocks); every slice adds its blocks in ascending order, the slices are combined in ascending order
// through shared memory -- a fixed association, so the result is deterministic.
constexpr int SP_SLICES = 8;
__global__ void __launch_bounds__(32 * SP_SLICES) sum_partials_kernel(
    int nparts, long long count4, const float4 *__restrict__ partials, float4 *__restrict__ out) {
  __shared__ float4 s_acc[SP_SLICES][32];
  const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < count4) {
    int p = threadIdx.y;
    for (; p + 3 * SP_SLICES < nparts; p += 4 * SP_SLICES) {  // four independent loads in flight
      const float4 a = __ldcs(partials + (size_t)p * count4 + i);
      const float4 b = __ldcs(partials + (size_t)(p + SP_SLICES) * count4 + i);
      const float4 c = __ldcs(partials + (size_t)(p + 2 * SP_SLICES) * count4 + i);
      const float4 d = __ldcs(partials + (size_t)(p + 3 * SP_SLICES) * count4 + i);
      acc.x = (((acc.x + a.x) + b.x) + c.x) + d.x;
      acc.y = (((acc.y + a.y) + b.y) + c.y) + d.y;
      acc.z = (((acc.z + a.z) + b.z) + c.z) + d.z;
      acc.w = (((acc.w + a.w) + b.w) + c.w) + d.w;
    }
    for (; p < nparts; p += SP_SLICES) {
      const float4 a = __ldcs(partials + (size_t)p * count4 + i);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  s_acc[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < count4) {
#pragma unroll
    for (int y = 1; y < SP_SLICES; ++y) {
      const float4 a = s_acc[y][threadIdx.x];
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    out[i] = acc;
  }
}

}  // namespace
}  // namespace nesie

using namespace nesie;

// out[count] = sum over the nparts partial blocks written by nesie_gemm_wgrad_3xtf32 (ascending
// order, round-to-nearest fp32: deterministic).  count must be a multiple of 4, pointers 16-byte aligned.
extern "C" int nesie_gemm_sum_partials(int nparts, long long count, const float *partials, float *out,
                                       void *stream) {
  NESIE_REQUIRE(nparts >= 1 && count >= 0 && (count & 3) == 0, "need nparts >= 1, count % 4 == 0");
  if (count == 0) return NESIE_OK;
  NESIE_REQUIRE(partials && out, "null pointer");
  NESIE_REQUIRE(((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "pointers must be 16-byte aligned");
  const long long count4 = count >> 2;
  sum_partials_kernel<<<(unsigned)((count4 + 31) / 32), dim3(32, SP_SLICES), 0, (cudaStream_t)stream>>>(
      nparts, count4, reinterpret_cast<const float4 *>(partials), reinterpret_cast<float4 *>(out));
  return check_launch("nesie_gemm_sum_partials");
}

extern "C" long long nesie_gemm_b_image_bytes(int n, int k) {
  if (n <= 0 || k <= 0) return 0;
  const long long npad = (n + 15) & ~15, nslab = (k + 31) / 32;
  return 2 * nslab * npad * 128;
}

extern "C" int nesie_gemm_pack_b(int n, int k, long long stride_n, long long stride_k,
                                 const float *b, void *image, void *stream) {
  NESIE_REQUIRE(n >= 1 && n <= 256 && k >= 1, "need 1 <= n <= 256, k >= 1");
  NESIE_REQUIRE(b && image, "null pointer");
  const int npad = (n + 15) & ~15, nslab = (k + 31) / 32;
  const int total = nslab * npad * 32;
  pack_b_tf32_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      n, k, npad, nslab, stride_n, stride_k, b, reinterpret_cast<float *>(image));
  return check_launch("nesie_gemm_pack_b");
}

// Diagnostic: cycle counters of CTA 0 accumulated by NT launches made with NESIE_GEMM_DBG & 128
// (loader: wait-empty, fill, publish, slabs; MMA thread: wait-acc, wait-full, issue, total;
// epilogue: wait-acc, drain, tiles).  Reading resets them.
extern "C" int nesie_gemm_debug_profile(long long *out16) {
  NESIE_REQUIRE(out16, "null pointer");
  NESIE_CUDA(cudaDeviceSynchronize());
  NESIE_CUDA(cudaMemcpyFromSymbol(out16, g_gemm_prof, sizeof(long long) * 16));
  long long zero[16] = {0};
  NESIE_CUDA(cudaMemcpyToSymbol(g_gemm_prof, zero, sizeof(zero)));
  return NESIE_OK;
}

// row-group extras of the TMA kernel's epilogue (GemmTmaParams)
struct GemmRowGroups {
  float *pool_max, *pool_min;
  unsigned char *pool_amax, *pool_amin;
  int pool_u, store_c;
  const float *grp_bias;
  int grp_shift;
};

static int gemm_nt_impl(long long r, int n, int k, const float *a, long long lda,
                        const void *b_image, float *c, long long ldc, const float *pro_scale,
                        const float *pro_shift, float *col_stats, void *stream,
                        const float *bn_y = nullptr, long long ldy = 0, const float *bn_stats = nullptr,
                        const GemmRowGroups *rg = nullptr) {
  NESIE_REQUIRE(r >= 0 && n >= 1 && n <= 256 && k >= 1, "need r >= 0, 1 <= n <= 256, k >= 1");
  NESIE_REQUIRE(r < (1LL << 31) - 256, "too many rows");
  if (r == 0) return NESIE_OK;
  NESIE_REQUIRE(a && b_image && (c || (rg && !rg->store_c)), "null pointer");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(b_image) & 15) == 0, "b_image must be 16-byte aligned");
  GemmParams p;
  p.R = (int)r; p.N = n; p.K = k;
  p.npad = (n + 15) & ~15;
  p.nslab = (k + 31) / 32;
  p.lda = lda; p.ldc = ldc;
  p.A = a; p.Bimg = reinterpret_cast<const unsigned char *>(b_image); p.C = c;
  const int ntiles_all = (p.R + G_TILE - 1) / G_TILE;
  if (gemm_tma_enabled()) {
    CUtensorMap tm;
    if (make_tmap(&tm, a, r, k, lda, G_SLABK, G_TILE, CU_TENSOR_MAP_SWIZZLE_128B)) {
      GemmTmaParams q;
      q.R = p.R; q.N = n; q.K = k; q.npad = p.npad; q.nslab = p.nslab; q.ldc = ldc;
      q.Bimg = p.Bimg; q.C = c;
      q.pro_scale = pro_scale; q.pro_shift = pro_shift; q.col_stats = col_stats;
      q.bn_y = bn_y; q.ldy = ldy; q.bn_stats = bn_stats;
      q.pool_max = q.pool_min = nullptr; q.pool_amax = q.pool_amin = nullptr;
      q.pool_u = 32; q.store_c = 1; q.grp_bias = nullptr; q.grp_shift = 4;
      if (rg) {
        q.pool_max = rg->pool_max; q.pool_amax = rg->pool_amax;
        q.pool_min = rg->pool_min; q.pool_amin = rg->pool_amin;
        q.pool_u = rg->pool_u; q.store_c = rg->store_c;
        q.grp_bias = rg->grp_bias; q.grp_shift = rg->grp_shift;
      }
      { const char *e = getenv("NESIE_GEMM_DBG"); q.dbg = e ? atoi(e) : 0; }
      const size_t stage = 2 * G_ASLAB + 2 * (size_t)p.npad * 128;
      const size_t epi = (size_t)T_EPIW * 4096;
      q.nstages = (int)((gemm_smem_budget() - epi) / stage);
      if (q.nstages < 1) q.nstages = 1;
      if (q.nstages > G_MAXSTAGES) q.nstages = G_MAXSTAGES;
      const size_t smem = (size_t)q.nstages * stage + epi + T_CTRL_BYTES;
      const int regs = gemm_tma_regs();
      int grid = gemm_balanced_grid(ntiles_all);
      // CTA pairs sharing the weight slabs (multicast): where the weight stream outweighs the
      // activations (wide N x K) and there are at least two tiles per CTA
      if (regs == 96 && gemm_pair_enabled() && p.npad * p.nslab >= 128 * 4 && ntiles_all >= 2 * grid &&
          (grid & 1) == 0) {   // (the statistics partials are sized for `grid` CTAs)
        auto kp = gemm_nt_tma_kernel<96, true>;
        NESIE_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, G_MAX_DYN_SMEM));
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3((unsigned)(grid & ~1), 1, 1);
        cfg.blockDim = dim3(T_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cfg.attrs = attr; cfg.numAttrs = 1;
        static int max_pairs = -1;   // co-resident clusters of two (one CTA per SM)
        if (max_pairs < 0) {
          int nc = 0;
          if (cudaOccupancyMaxActiveClusters(&nc, kp, &cfg) != cudaSuccess) { (void)cudaGetLastError(); nc = 0; }
          max_pairs = nc;
        }
        if (max_pairs >= 8) {
          if ((int)cfg.gridDim.x <= 2 * max_pairs) {
          NESIE_CUDA(cudaLaunchKernelEx(&cfg, kp, tm, q));
          return check_launch("nesie_gemm_nt_3xtf32 (pairs)");
          }
        }
      }
      auto kern = regs == 64 ? gemm_nt_tma_kernel<64, false>
                             : (regs == 88 ? gemm_nt_tma_kernel<88, false> : gemm_nt_tma_kernel<96, false>);
      NESIE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G_MAX_DYN_SMEM));
      kern<<<grid, T_THREADS, smem, (cudaStream_t)stream>>>(tm, q);
      return check_launch("nesie_gemm_nt_3xtf32");
    }
  }
  NESIE_REQUIRE(!pro_scale && !col_stats && !bn_y && !rg,
                "the fused prologue / statistics / pooling need the TMA path (16-byte aligned rows)");
  p.fast = ((lda & 3) == 0) && ((k & 3) == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0);
  { const char *e = getenv("NESIE_GEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
  const size_t stage = 2 * G_ASLAB + 2 * (size_t)p.npad * 128;
  p.nstages = (int)((G_SMEM_BUDGET - G_EPI_STAGE) / stage);
  if (p.nstages > G_MAXSTAGES) p.nstages = G_MAXSTAGES;
  const size_t smem = (size_t)p.nstages * stage + G_EPI_STAGE + 1024;
  NESIE_CUDA(cudaFuncSetAttribute(gemm_nt_3xtf32_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, G_MAX_DYN_SMEM));
  const int ntiles = (p.R + G_TILE - 1) / G_TILE;
  int grid = num_sms();
  if (ntiles < grid) grid = ntiles;
  gemm_nt_3xtf32_kernel<<<grid, G_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("nesie_gemm_nt_3xtf32");
}

extern "C" int nesie_gemm_nt_3xtf32(long long r, int n, int k, const float *a, long long lda,
                                    const void *b_image, float *c, long long ldc, void *stream) {
  return gemm_nt_impl(r, n, k, a, lda, b_image, c, ldc, nullptr, nullptr, nullptr, stream);
}

extern "C" int nesie_gemm_fused_supported(long long r, int n, int k, const float *a, long long lda,
                                          long long ldc) {
  return r >= 1 && n >= 4 && n <= 256 && (n & 3) == 0 && (k & 3) == 0 && (lda & 3) == 0 &&
         (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && gemm_tma_enabled() &&
         encode_tiled_fn() != nullptr;
}

extern "C" int nesie_gemm_stats_parts(long long r) {
  if (r <= 0) return 0;
  const long long ntiles = (r + G_TILE - 1) / G_TILE;
  return 4 * gemm_balanced_grid(ntiles);
}

extern "C" int nesie_gemm_nt_3xtf32_fused(long long r, int n, int k, const float *a, long long lda,
                                          const void *b_image, float *c, long long ldc,
                                          const float *pro_scale, const float *pro_shift,
                                          float *col_stats, void *stream) {
  NESIE_REQUIRE((pro_scale == nullptr) == (pro_shift == nullptr), "scale and shift go together");
  NESIE_REQUIRE(r >= 1, "need r >= 1");
  NESIE_REQUIRE(nesie_gemm_fused_supported(r, n, k, a, lda, ldc), "shape / alignment not supported");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(c) & 15) == 0, "c must be 16-byte aligned");
  return gemm_nt_impl(r, n, k, a, lda, b_image, c, ldc, pro_scale, pro_shift, col_stats, stream);
}

// Pooled / group-biased variant (see GemmTmaParams): c may be null (the output itself is not stored)
// when pool_max is given; pool_u is 16 or 32 and divides r; pool_min / pool_amin are optional;
// grp_bias (nullable) is [r / grp_k][n] with grp_k a power of two >= 16 that divides r.
extern "C" int nesie_gemm_nt_3xtf32_pool(long long r, int n, int k, const float *a, long long lda,
                                         const void *b_image, float *c, long long ldc,
                                         const float *pro_scale, const float *pro_shift,
                                         float *col_stats, int pool_u, float *pool_max,
                                         unsigned char *pool_amax, float *pool_min,
                                         unsigned char *pool_amin, const float *grp_bias, int grp_k,
                                         void *stream) {
  NESIE_REQUIRE((pro_scale == nullptr) == (pro_shift == nullptr), "scale and shift go together");
  NESIE_REQUIRE(r >= 1, "need r >= 1");
  NESIE_REQUIRE(nesie_gemm_fused_supported(r, n, k, a, lda, c ? ldc : 4), "shape / alignment not supported");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(c) & 15) == 0, "c must be 16-byte aligned");
  NESIE_REQUIRE(c || pool_max, "nothing to compute: no output and no pooling");
  NESIE_REQUIRE((pool_max == nullptr) == (pool_amax == nullptr) && (pool_min == nullptr) == (pool_amin == nullptr),
                "value and index arrays go together");
  NESIE_REQUIRE(!pool_min || pool_max, "the minimum comes with the maximum");
  GemmRowGroups rg;
  rg.pool_max = pool_max; rg.pool_amax = pool_amax; rg.pool_min = pool_min; rg.pool_amin = pool_amin;
  rg.pool_u = 32; rg.store_c = c ? 1 : 0; rg.grp_bias = grp_bias; rg.grp_shift = 4;
  if (pool_max) {
    NESIE_REQUIRE((pool_u == 16 || pool_u == 32) && r % pool_u == 0, "pool_u must be 16 or 32 and divide r");
    rg.pool_u = pool_u;
  }
  if (grp_bias) {
    NESIE_REQUIRE(grp_k >= 16 && (grp_k & (grp_k - 1)) == 0 && r % grp_k == 0,
                  "grp_k must be a power of two >= 16 that divides r");
    NESIE_REQUIRE((reinterpret_cast<uintptr_t>(grp_bias) & 15) == 0, "grp_bias must be 16-byte aligned");
    int sh = 0;
    while ((1 << sh) < grp_k) ++sh;
    rg.grp_shift = sh;
  }
  return gemm_nt_impl(r, n, k, a, lda, b_image, c, c ? ldc : n, pro_scale, pro_shift, col_stats, stream,
                      nullptr, 0, nullptr, &rg);
}

// B channels per CTA along gridDim.z.  The kernel is bound by shared-memory traffic (TMA fill + the
// hi / lo transform's read and write + the operand reads of 12 MMAs per slab: profiles/r02_ncu_notes.md),
// so splitting K only pays where it buys parallelism or a third pipeline stage without multiplying
// whole CTAs on a sliver of channels: K a multiple of 128, and either two A blocks (N > 128) or few
// rows (measured per shape; NESIE_WGRAD_KSPLIT=0 keeps K whole, =2 always splits beyond 128).
static int wgrad_kpart(long long r, int n, int k) {
  static int ksplit = -1;
  if (ksplit < 0) { const char *e = getenv("NESIE_WGRAD_KSPLIT"); ksplit = e ? atoi(e) : 1; }
  const int kp = (k + 31) & ~31;
  if (!ksplit || kp <= 128) return kp;
  if (ksplit == 2) return 128;
  return (kp % 128 == 0 && (n > 128 || r <= 16384)) ? 128 : kp;
}

extern "C" int nesie_gemm_nt_3xtf32_bnbwd(long long r, int n, int k, const float *a, long long lda,
                                          const void *b_image, float *c, long long ldc,
                                          const float *bn_y, long long ldy, const float *bn_stats,
                                          float *col_stats, void *stream) {
  NESIE_REQUIRE(r >= 1, "need r >= 1");
  NESIE_REQUIRE(bn_y && bn_stats && col_stats, "null pointer");
  NESIE_REQUIRE(nesie_gemm_fused_supported(r, n, k, a, lda, ldc), "shape / alignment not supported");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(c) & 15) == 0, "c must be 16-byte aligned");
  return gemm_nt_impl(r, n, k, a, lda, b_image, c, ldc, nullptr, nullptr, col_stats, stream, bn_y, ldy,
                      bn_stats);
}

// row chunks (one TMEM accumulation each) and CTAs along x (one partial block each)
static void wgrad_plan(long long r, int n, int k, int *chunk, int *nchunks, int *gx) {
  const long long nslab = (r + 31) / 32;
  const int kp = (k + 31) & ~31, kpart = wgrad_kpart(r, n, k);
  const int mblocks = ((n + 127) / 128) * ((kp + kpart - 1) / kpart);   // CTAs per row chunk
  *chunk = wgrad_chunk(nslab, mblocks);
  *nchunks = (int)((nslab + *chunk - 1) / *chunk);
  int g = gemm_grid_sms() / mblocks;
  if (g < 1) g = 1;
  if (g > *nchunks) g = *nchunks;
  *gx = g;
}

extern "C" int nesie_gemm_wgrad_splits(long long r, int n, int k) {
  if (r <= 0 || n <= 0) return 0;
  int chunk, nchunks, gx;
  wgrad_plan(r, n, k, &chunk, &nchunks, &gx);
  return gx;  // one partial block per CTA
}

static int gemm_wgrad_impl(long long r, int n, int k, const float *a, long long lda,
                           const float *b, long long ldb, const float *pro_scale,
                           const float *pro_shift, float *partials, int nsplits, void *stream) {
  NESIE_REQUIRE(r >= 1 && n >= 1 && n <= 256 && k >= 1 && k <= 512, "need r>=1, n<=256, k<=512");
  NESIE_REQUIRE(r < (1LL << 31) - 256, "too many rows");
  NESIE_REQUIRE(a && b && partials, "null pointer");
  NESIE_REQUIRE(nsplits == nesie_gemm_wgrad_splits(r, n, k), "nsplits must come from nesie_gemm_wgrad_splits");
  if (gemm_tma_enabled()) {
    CUtensorMap ta, tb;
    if (make_tmap(&ta, a, r, n, lda, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
        make_tmap(&tb, b, r, k, ldb, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
      WgradTmaParams q;
      q.R = (int)r; q.N = n; q.K = k; q.kp = (k + 31) & ~31; q.P = partials;
      q.kpart = wgrad_kpart(r, n, k);
      q.pro_scale = pro_scale; q.pro_shift = pro_shift;
      { const char *e = getenv("NESIE_GEMM_DBG"); q.dbg = e ? atoi(e) : 0; }
      const size_t stage = 2 * (size_t)(4 * 4096) + 2 * (size_t)(q.kpart >> 5) * 4096;
      q.nstages = (int)((gemm_smem_budget()) / stage);
      if (q.nstages > G_MAXSTAGES) q.nstages = G_MAXSTAGES;
      NESIE_REQUIRE(q.nstages >= 1, "k too large for shared memory");
      const size_t smem = (size_t)q.nstages * stage + 1024;
      const int regs = gemm_tma_regs();
      auto kern = regs == 64 ? gemm_wgrad_tma_kernel<64> : (regs == 88 ? gemm_wgrad_tma_kernel<88> : gemm_wgrad_tma_kernel<96>);
      NESIE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G_MAX_DYN_SMEM));
      int gx;
      wgrad_plan(r, n, k, &q.chunk, &q.nchunks, &gx);
      const int mblocks = (n + 127) / 128, kblocks = (q.kp + q.kpart - 1) / q.kpart;
      kern<<<dim3(gx, mblocks, kblocks), T_THREADS, smem, (cudaStream_t)stream>>>(ta, tb, q);
      return check_launch("nesie_gemm_wgrad_3xtf32");
    }
  }
  NESIE_REQUIRE(!pro_scale, "the fused prologue needs the TMA path (16-byte aligned rows)");
  WgradParams p;
  p.R = (int)r; p.N = n; p.K = k;
  p.kp = (k + 31) & ~31;
  p.lda = lda; p.ldb = ldb;
  p.A = a; p.B = b; p.P = partials;
  const size_t stage = 2 * (size_t)(4 * 4096) + 2 * (size_t)(p.kp >> 5) * 4096;
  p.nstages = (int)((G_SMEM_BUDGET) / stage);
  if (p.nstages > G_MAXSTAGES) p.nstages = G_MAXSTAGES;
  NESIE_REQUIRE(p.nstages >= 1, "k too large for shared memory");
  const size_t smem = (size_t)p.nstages * stage + 1024;
  NESIE_CUDA(cudaFuncSetAttribute(gemm_wgrad_3xtf32_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, G_MAX_DYN_SMEM));
  int gx;
  wgrad_plan(r, n, k, &p.chunk, &p.nchunks, &gx);
  p.vec = ((lda & 3) == 0) && ((ldb & 3) == 0) && ((n & 3) == 0) && ((k & 3) == 0) &&
          (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0);
  p.mn = p.vec ? 3 : 0;
  { const char *e = getenv("NESIE_GEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
  { const char *e = getenv("NESIE_WGRAD_LAYOUT"); if (e && p.vec) p.mn = atoi(e) & 3; }
  const int mblocks = (n + 127) / 128;
  dim3 grid(gx, mblocks);
  gemm_wgrad_3xtf32_kernel<<<grid, W_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("nesie_gemm_wgrad_3xtf32");
}

extern "C" int nesie_gemm_wgrad_3xtf32(long long r, int n, int k, const float *a, long long lda,
                                       const float *b, long long ldb, float *partials,
                                       int nsplits, void *stream) {
  return gemm_wgrad_impl(r, n, k, a, lda, b, ldb, nullptr, nullptr, partials, nsplits, stream);
}

extern "C" int nesie_gemm_wgrad_3xtf32_fused(long long r, int n, int k, const float *a, long long lda,
                                             const float *b, long long ldb, const float *pro_scale,
                                             const float *pro_shift, float *partials, int nsplits,
                                             void *stream) {
  NESIE_REQUIRE(pro_scale && pro_shift, "null pointer");
  NESIE_REQUIRE((k & 3) == 0, "k must be a multiple of 4");
  return gemm_wgrad_impl(r, n, k, a, lda, b, ldb, pro_scale, pro_shift, partials, nsplits, stream);
}
