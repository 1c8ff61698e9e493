// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel / furthest_point_sampling_with_dist_kernel
// (reference: ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:25-141,213-331).
//
// Design (B200-first, not a translation):
//   * The reference runs ONE 1024-thread block per scene and streams xyz (12 B/pt) and the
//     running min-distance array (4 B/pt, read+write) through L2 on every one of the m-1 serial
//     iterations.  Here a scene is owned by a thread-block CLUSTER (1..16 CTAs); every thread
//     keeps its points AND their running min-distances in registers for the whole kernel, so an
//     iteration touches no global memory at all (one 4-byte index store per iteration).
//   * FPS is a chain of m-1 dependent argmax steps, so the kernel is built around the latency
//     of one step: register-resident distance update -> warp argmax with two redux.sync ops
//     (max over the distance bits, then min over a tie-break key) -> EVERY warp sends its
//     20-byte candidate (distance, key, x, y, z) straight into the shared memory of every CTA of
//     the cluster with st.async, whose completion is counted by a transaction mbarrier in the
//     destination CTA -> each warp waits on its own CTA's mbarrier (CTA-scope acquire: no L1
//     invalidate) and reduces the <=128 candidates.  No __syncthreads, no cluster barrier and no
//     second-level "CTA winner" stage sit on the critical path.  (A single-CTA scene uses a
//     shared-memory stage + one __syncthreads instead.)
//   * Bit-exactness.  Distances use the reference's contraction fma(dz,dz,fma(dx,dx,dy*dy)).
//     The reference resolves equal maxima by (a) a strict '>' scan over k = tid, tid+bs, ...
//     inside a thread and (b) a shared-memory tree whose slot t survives ties against slot t+s
//     for s = bs/2 ... 1.  (b) prefers the slot whose index is smallest when its log2(bs) bits
//     are read in REVERSE order (stride 1 is decided last, so bit 0 is the most significant
//     tie-break bit).  Both rules together = "lowest key" with
//         key(k) = bitreverse32(k mod bs) | (k div bs)
//     where bs = the reference's block size opt_n_threads(n).  Points are dealt to threads so
//     that a thread's register slots are already in ascending key order, which lets the inner
//     loop keep the reference's strict '>':
//       T >= bs threads/scene: thread g owns residue r = g mod bs and a contiguous range of
//                              q = k div bs;
//       T <  bs threads/scene: thread g owns the residues g + a*T, visited in bit-reversed
//                              order of a (= ascending bitreverse(r)), each with all its q.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace nesie {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_WARPS_CL = 8;  // warps per CTA in cluster mode (<= 256 threads)

struct __align__(16) Cand {  // one argmax candidate, 32-byte slot (20 bytes used)
  unsigned d;                // distance bits (non-negative float => unsigned order == float order)
  unsigned p;                // tie-break key, lower wins
  float x, y, z;             // the candidate's coordinates (the next centre if it wins)
  unsigned pad[3];
};

struct FpsParams {
  int n, m;
  int bs, log2bs;   // the reference's block size for this n
  int log2a;        // log2(residues per thread) (0 unless T < bs)
  int qn;           // q values per residue per thread
  int nslots;       // (1 << log2a) * qn  <= PPT
  int tscene;       // threads per scene T = CL * blockDim
};

__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned cta_rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes)
               : "memory");
}
// CTA-scope acquire on purpose: a cluster-scope acquire makes ptxas emit CCTL.IVALL (an L1
// invalidate, ~28 % of all stall samples in the first profile).  The payload lives in this
// CTA's own shared memory and is published by the st.async transaction count.
__device__ __forceinline__ bool mbar_try_wait(unsigned mbar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(mbar), "r"(parity)
      : "memory");
  return ok != 0;
}
// remote 16-byte / 4-byte stores that also signal their byte count on the destination mbarrier
__device__ __forceinline__ void st_async_v4(unsigned raddr, unsigned rmbar, unsigned a,
                                            unsigned b, unsigned c, unsigned d) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
      ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rmbar)
      : "memory");
}
__device__ __forceinline__ void st_async_b32(unsigned raddr, unsigned rmbar, unsigned a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(a), "r"(rmbar)
               : "memory");
}

// Argmax over the lanes of a warp: returns the source lane; d/p are replaced by the winner's.
// The maximum is almost always unique, so the tie-break reduction runs only when the ballot of
// "lanes holding the maximum" has more than one bit set (a warp-uniform branch).
__device__ __forceinline__ int warp_argmax(unsigned &d, unsigned &p) {
  const unsigned wd = __reduce_max_sync(FULL, d);
  unsigned win = __ballot_sync(FULL, d == wd);
  if (win & (win - 1)) {
    const unsigned wp = __reduce_min_sync(FULL, d == wd ? p : 0xffffffffu);
    win = __ballot_sync(FULL, d == wd && p == wp);
  }
  const int src = __ffs(win) - 1;
  d = wd;
  p = __shfl_sync(FULL, p, src);
  return src;
}

__device__ __forceinline__ int decode_key(unsigned p, int bs, int log2bs) {
  if (log2bs == 0) return (int)p;
  const unsigned r = __brev(p) & (unsigned)(bs - 1);
  const unsigned q = p & ((1u << (32 - log2bs)) - 1u);
  return (int)(q * (unsigned)bs + r);
}

// residue and q of register slot s of global thread g
// (no integer division: with one residue per thread ap is 0, otherwise it is < 32 / qn)
__device__ __forceinline__ void slot_rq(const FpsParams &pr, int g, int s, int &r, int &q) {
  int ap = 0, sq = s;  // ap = position in bit-reversed residue order, sq = q offset
  int a = 0;
  if (pr.log2a) {
    while (sq >= pr.qn) { sq -= pr.qn; ++ap; }
    a = (int)(__brev((unsigned)ap) >> (32 - pr.log2a));
  }
  r = (g & (pr.bs - 1)) + a * pr.tscene;
  q = (g >> pr.log2bs) * pr.qn + sq;
}

// ------------------------------------------------------------------------------------------
// Register-resident kernel.  grid = b * CL CTAs (cluster = CL consecutive CTAs = one scene).
// Dynamic smem: 3*nslots*blockDim floats (a copy of the CTA's coordinates, read by candidate
// lanes) followed, in cluster mode, by 2 * CL * MAX_WARPS_CL candidate slots.
// ------------------------------------------------------------------------------------------
template <int CL, int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT) fps_reg_kernel(FpsParams pr,
                                                       const float *__restrict__ xyz,
                                                       float *__restrict__ temp,
                                                       int *__restrict__ idx) {
  unsigned rank = 0;
  if constexpr (CL > 1) rank = cg::this_cluster().block_rank();
  const int scene = blockIdx.x / CL;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
  const int g = (int)rank * NT + tid;
  const int n = pr.n, m = pr.m, nslots = pr.nslots;

  xyz += (size_t)scene * n * 3;
  idx += (size_t)scene * m;
  if (temp) temp += (size_t)scene * n;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_pts = reinterpret_cast<float *>(smem_raw);  // [3][nslots][NT]
  Cand *s_cl = reinterpret_cast<Cand *>(smem_raw + (((size_t)3 * nslots * NT * 4 + 15) & ~(size_t)15));
  __shared__ Cand s_wk[2][32];                         // single-CTA mode only
  __shared__ __align__(8) unsigned long long s_mbar[2];

  float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    float x = 0.f, y = 0.f, z = 0.f, d = -2.f;  // -2 never beats the scan's initial best of -1
    if (s < nslots) {
      int r, q;
      slot_rq(pr, g, s, r, q);
      const long long k = (long long)q * pr.bs + r;
      if (r < pr.bs && k < n) {
        x = xyz[k * 3 + 0];
        y = xyz[k * 3 + 1];
        z = xyz[k * 3 + 2];
        d = temp ? temp[k] : 1e10f;  // furthest_point_sample.py:30
      }
      s_pts[(0 * nslots + s) * NT + tid] = x;
      s_pts[(1 * nslots + s) * NT + tid] = y;
      s_pts[(2 * nslots + s) * NT + tid] = z;
    }
    px[s] = x; py[s] = y; pz[s] = z; md[s] = d;
  }

  // old = 0 (furthest_point_sample_cuda.cu:46-47)
  float cx = xyz[0], cy = xyz[1], cz = xyz[2];
  if (rank == 0 && tid == 0) idx[0] = 0;
  if constexpr (CL > 1) {
    if (tid == 0) {
      mbar_init(smem_u32(&s_mbar[0]), 1);
      mbar_init(smem_u32(&s_mbar[1]), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cg::this_cluster().sync();  // every CTA of the cluster is resident, barriers initialised
  }
  const int ncand = CL * NW;
  const int log2nw = 31 - __clz(NW);

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    // two interleaved strict-'>' chains (even / odd slots) halve the dependent-compare latency
    float b0 = -1.f, b1 = -1.f;
    int s0 = 0, s1 = 0;
#pragma unroll
    for (int c4 = 0; c4 < PPT; c4 += 4) {
      if (c4 < nslots) {  // warp-uniform; slots beyond nslots in the chunk hold md = -2
#pragma unroll
        for (int s = c4; s < c4 + 4 && s < PPT; ++s) {
          const float d = sqdist_ref(px[s], py[s], pz[s], cx, cy, cz);
          const float d2 = fminf(d, md[s]);
          md[s] = d2;
          if (s & 1) { if (d2 > b1) { b1 = d2; s1 = s; } }
          else       { if (d2 > b0) { b0 = d2; s0 = s; } }
        }
      }
    }
    const bool odd = (b1 > b0) || (b1 == b0 && s1 < s0);
    const float best = odd ? b1 : b0;
    const int bslot = odd ? s1 : s0;
    // candidate coordinates (issued before the reductions so the LDS latency overlaps them)
    const float bx = s_pts[(0 * nslots + bslot) * NT + tid];
    const float by = s_pts[(1 * nslots + bslot) * NT + tid];
    const float bz = s_pts[(2 * nslots + bslot) * NT + tid];
    unsigned kd = 0u, kp = 0xffffffffu;
    if (best >= 0.f) {
      int r, q;
      slot_rq(pr, g, bslot, r, q);
      kd = __float_as_uint(best);
      kp = (pr.log2bs ? __brev((unsigned)r) : 0u) | (unsigned)q;
    }
    const int src = warp_argmax(kd, kp);
    const float wx = __shfl_sync(FULL, bx, src);
    const float wy = __shfl_sync(FULL, by, src);
    const float wz = __shfl_sync(FULL, bz, src);

    unsigned fd = 0u, fp = 0xffffffffu;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if constexpr (CL == 1) {
      if (lane == 0) {
        Cand *c = &s_wk[par][warp];
        reinterpret_cast<uint4 *>(c)[0] = make_uint4(kd, kp, __float_as_uint(wx), __float_as_uint(wy));
        c->z = wz;
      }
      __syncthreads();
      if (lane < NW) {
        const Cand c = s_wk[par][lane];
        fd = c.d; fp = c.p; fx = c.x; fy = c.y; fz = c.z;
      }
    } else {
      const unsigned mb = smem_u32(&s_mbar[par]);
      if (tid == 0) mbar_arrive_expect_tx(mb, (unsigned)ncand * 20u);
      if (lane < CL) {  // lane L delivers this warp's candidate to CTA L of the cluster
        const unsigned slot = smem_u32(&s_cl[(par * CL + (int)rank) * MAX_WARPS_CL + warp]);
        const unsigned ra = mapa_u32(slot, (unsigned)lane);
        const unsigned rm = mapa_u32(mb, (unsigned)lane);
        st_async_v4(ra, rm, kd, kp, __float_as_uint(wx), __float_as_uint(wy));
        st_async_b32(ra + 16, rm, __float_as_uint(wz));
      }
      const unsigned parity = (unsigned)((j - 1) >> 1) & 1u;
      while (!mbar_try_wait(mb, parity)) {}
      // lane-local argmax over candidates lane, lane+32, ... (all CL*NW slots are valid)
      for (int c = lane; c < ncand; c += 32) {
        const Cand cd = s_cl[(par * CL + (c >> log2nw)) * MAX_WARPS_CL + (c & (NW - 1))];
        if (cd.d > fd || (cd.d == fd && cd.p < fp)) {
          fd = cd.d; fp = cd.p; fx = cd.x; fy = cd.y; fz = cd.z;
        }
      }
    }
    const int s3 = warp_argmax(fd, fp);
    cx = __shfl_sync(FULL, fx, s3);
    cy = __shfl_sync(FULL, fy, s3);
    cz = __shfl_sync(FULL, fz, s3);
    if (rank == 0 && tid == 0) idx[j] = decode_key(fp, pr.bs, pr.log2bs);
  }

  if (temp) {  // hand the final min-distances back, like the reference's in-place buffer
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
      if (s < nslots) {
        int r, q;
        slot_rq(pr, g, s, r, q);
        const long long k = (long long)q * pr.bs + r;
        if (r < pr.bs && k < n) temp[k] = md[s];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Global-memory fallback (n beyond what a 16-CTA cluster holds in registers) and the
// distance-matrix variant.  One CTA per scene, thread t < bs scans k = t, t+bs, ...
// ------------------------------------------------------------------------------------------
template <bool WITH_DIST>
__global__ void __launch_bounds__(1024) fps_generic_kernel(int n, int m, int bs, int log2bs,
                                                           const float *__restrict__ data,
                                                           float *__restrict__ temp,
                                                           int *__restrict__ idx) {
  const int scene = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  data += WITH_DIST ? (size_t)scene * n * n : (size_t)scene * n * 3;
  temp += (size_t)scene * n;
  idx += (size_t)scene * m;
  __shared__ unsigned s_d[2][32], s_p[2][32];
  const unsigned rbits = log2bs ? __brev((unsigned)tid) : 0u;
  int old = 0;
  if (tid == 0) idx[0] = 0;
  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    float best = -1.f;
    int bq = 0;
    if (tid < bs) {
      float x1 = 0.f, y1 = 0.f, z1 = 0.f;
      if (!WITH_DIST) { x1 = data[old * 3 + 0]; y1 = data[old * 3 + 1]; z1 = data[old * 3 + 2]; }
      int q = 0;
      for (int k = tid; k < n; k += bs, ++q) {
        float d;
        if (WITH_DIST) d = data[(size_t)old * n + k];
        else d = sqdist_ref(data[k * 3 + 0], data[k * 3 + 1], data[k * 3 + 2], x1, y1, z1);
        const float d2 = fminf(d, temp[k]);
        temp[k] = d2;
        if (d2 > best) { best = d2; bq = q; }
      }
    }
    unsigned kd = best >= 0.f ? __float_as_uint(best) : 0u;
    unsigned kp = best >= 0.f ? (rbits | (unsigned)bq) : 0xffffffffu;
    const int src = warp_argmax(kd, kp);
    if (lane == src) { s_d[par][warp] = kd; s_p[par][warp] = kp; }
    __syncthreads();
    unsigned d2 = 0u, p2 = 0xffffffffu;
    if (lane < NW) { d2 = s_d[par][lane]; p2 = s_p[par][lane]; }
    warp_argmax(d2, p2);
    old = decode_key(p2, bs, log2bs);
    if (tid == 0) idx[j] = old;
  }
}

// The reference launcher's block size (furthest_point_sample_cuda.cu:11-15), same libm call.
int ref_block_size(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

typedef void (*fps_fn)(FpsParams, const float *, float *, int *);

template <int CL>
fps_fn pick_ppt(int nslots) {
#define NESIE_FPS_CASE(P) if (nslots <= P) return fps_reg_kernel<CL, P, 256>;
  NESIE_FPS_CASE(4) NESIE_FPS_CASE(8) NESIE_FPS_CASE(12) NESIE_FPS_CASE(16)
  NESIE_FPS_CASE(20) NESIE_FPS_CASE(24) NESIE_FPS_CASE(28) NESIE_FPS_CASE(32)
#undef NESIE_FPS_CASE
  return nullptr;
}

// Slot layout for (n, cluster size, threads per CTA); returns false if it needs > 32 slots.
bool plan(int n, int CL, int NT, FpsParams *pr) {
  const int bs = ref_block_size(n), log2bs = ilog2(bs);
  const int T = CL * NT;
  pr->bs = bs; pr->log2bs = log2bs; pr->tscene = T;
  const int Q = ceil_div(n, bs);
  if (T >= bs) {
    pr->log2a = 0;
    pr->qn = ceil_div(Q, T / bs);
  } else {
    pr->log2a = ilog2(bs / T);
    pr->qn = Q;
  }
  pr->nslots = (1 << pr->log2a) * pr->qn;
  return pr->nslots <= 32;
}

int launch_reg(int CL, int NT, int b, FpsParams pr, const float *xyz, float *temp, int *idx,
               cudaStream_t st) {
  fps_fn fn = nullptr;
  switch (CL) {
    case 1: fn = pick_ppt<1>(pr.nslots); break;
    case 2: fn = pick_ppt<2>(pr.nslots); break;
    case 4: fn = pick_ppt<4>(pr.nslots); break;
    case 8: fn = pick_ppt<8>(pr.nslots); break;
    case 16: fn = pick_ppt<16>(pr.nslots); break;
  }
  if (!fn) return NESIE_ERR_UNSUPPORTED;
  size_t smem = (((size_t)3 * pr.nslots * NT * sizeof(float)) + 15) & ~(size_t)15;
  if (CL > 1) smem += (size_t)2 * CL * MAX_WARPS_CL * sizeof(Cand);
  if (smem > 32 * 1024)  // static smem counts against the 48 KB default too
    NESIE_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * CL));
  cfg.blockDim = dim3((unsigned)NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  int nat = 0;
  if (CL > 1) {
    if (CL > 8)
      NESIE_CUDA(cudaFuncSetAttribute((const void *)fn,
                                      cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    nat = 1;
  }
  cfg.attrs = at;
  cfg.numAttrs = nat;
  if (CL > 1) {
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, (const void *)fn, &cfg);
    if (e != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      return NESIE_ERR_UNSUPPORTED;  // caller retries with a smaller cluster
    }
  }
  NESIE_CUDA(cudaLaunchKernelEx(&cfg, fn, pr, xyz, temp, idx));
  return NESIE_OK;
}

constexpr int REG_CAPACITY = 16 * 256 * 32;  // points a 16-CTA cluster holds in registers

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_fps_needs_temp(int b, int n, int m) {
  (void)b; (void)m;
  return n > REG_CAPACITY ? 1 : 0;
}

extern "C" int nesie_fps(int b, int n, int m, const float *xyz, float *temp, int *idx,
                         void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 1 && m >= 0, "need b >= 0, n >= 1, m >= 0");
  NESIE_REQUIRE(xyz && idx, "null pointer");
  if (b == 0 || m == 0) return NESIE_OK;
  cudaStream_t st = (cudaStream_t)stream;

  // tuning knobs (measurement only): cluster size and threads per CTA
  int force_cl = 0, force_nt = 0;
  if (const char *e = getenv("NESIE_FPS_CLUSTER")) force_cl = atoi(e);
  if (const char *e = getenv("NESIE_FPS_THREADS")) force_nt = atoi(e);

  if (n <= REG_CAPACITY) {
    // Preferred shape: as many CTAs per scene as the GPU can co-schedule (<= 16) once a scene
    // is big enough for the cross-CTA exchange to pay off; 128 threads per CTA when that still
    // fits the 32-slot register budget (fewer warps = cheaper reductions), else 256.
    int cl = 1;
    if (n > 4096) {
      // NESIE_FPS_MAX_CLUSTER caps the CTAs per scene: a smaller cluster holds fewer SMs (the
      // kernel is latency-bound, its CTAs mostly wait), which matters when it runs beside kernels
      // that want the whole GPU
      int max_cl = 16;
      if (const char *e = getenv("NESIE_FPS_MAX_CLUSTER")) max_cl = atoi(e) >= 2 ? atoi(e) : 16;
      cl = 2;
      while (cl * 2 <= max_cl && cl * 2 <= 16 && b * cl * 2 <= num_sms()) cl *= 2;
    }
    if (force_cl == 1 || force_cl == 2 || force_cl == 4 || force_cl == 8 || force_cl == 16)
      cl = force_cl;
    const int bs = ref_block_size(n);
    for (int tries = 0; tries < 6; ++tries) {
      FpsParams pr;
      pr.n = n; pr.m = m;
      int NT = cl == 1 ? (bs >= 256 ? 256 : (bs < 32 ? 32 : bs)) : 128;
      if (force_nt == 32 || force_nt == 64 || force_nt == 128 || force_nt == 256) NT = force_nt;
      bool ok = plan(n, cl, NT, &pr);
      if (!ok && NT < 256) { NT = 256; ok = plan(n, cl, NT, &pr); }
      int rc = NESIE_ERR_UNSUPPORTED;
      if (ok) rc = launch_reg(cl, NT, b, pr, xyz, temp, idx, st);
      if (rc != NESIE_ERR_UNSUPPORTED) return rc;
      // does not fit / cluster not schedulable: grow the cluster if slots ran out, else shrink
      if (!ok && cl < 16) cl *= 2;
      else if (ok && cl > 1) cl /= 2;
      else break;
    }
  }
  // global-memory fallback
  if (!temp) {
    set_error("nesie_fps: n=%d needs the global-memory kernel, which requires temp", n);
    return NESIE_ERR_UNSUPPORTED;
  }
  const int bs = ref_block_size(n), log2bs = ilog2(bs);
  const int NT = bs < 32 ? 32 : bs;
  fps_generic_kernel<false><<<b, NT, 0, st>>>(n, m, bs, log2bs, xyz, temp, idx);
  return check_launch("nesie_fps(generic)");
}

extern "C" int nesie_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx,
                                   void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 1 && m >= 0, "need b >= 0, n >= 1, m >= 0");
  NESIE_REQUIRE(dist && temp && idx, "null pointer");
  if (b == 0 || m == 0) return NESIE_OK;
  const int bs = ref_block_size(n), log2bs = ilog2(bs);
  const int NT = bs < 32 ? 32 : bs;
  fps_generic_kernel<true><<<b, NT, 0, (cudaStream_t)stream>>>(n, m, bs, log2bs, dist, temp, idx);
  return check_launch("nesie_fps_with_dist");
}
