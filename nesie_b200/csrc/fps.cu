// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel / furthest_point_sampling_with_dist_kernel
// (reference: ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:25-141,213-331).
//
// Design (B200-first, not a translation):
//   * The reference runs ONE 1024-thread block per scene and streams xyz (12 B/pt) and the
//     running min-distance array (4 B/pt, read+write) through L2 on every one of the m-1 serial
//     iterations.  Here a scene is owned by a thread-block CLUSTER (1, 4, 8 or 16 CTAs); every
//     thread keeps its points AND their running min-distances in registers for the whole kernel,
//     so an iteration touches no global memory at all (one 4-byte index store per iteration).
//   * Per iteration: register-resident distance update -> warp argmax with two redux.sync ops
//     (max over the distance bits, then min over a tie-break key) -> one __syncthreads ->
//     CTA argmax by warp 0 -> the CTA winner (distance, key, x, y, z) is written into every
//     CTA of the cluster through distributed shared memory -> one cluster barrier -> every
//     warp reduces the <=16 CTA winners and already holds the next centre's coordinates.
//   * Bit-exactness.  Distances use the reference's contraction fma(dz,dz,fma(dx,dx,dy*dy)).
//     The reference resolves equal maxima by (a) a strict '>' scan over k = tid, tid+bs, ...
//     inside a thread and (b) a shared-memory tree whose slot t survives ties against slot t+s
//     for s = bs/2 ... 1.  (b) prefers the slot whose index is smallest when its log2(bs) bits
//     are read in REVERSE order (stride 1 is decided last, so bit 0 is the most significant
//     tie-break bit).  Both rules together = "lowest key" with
//         key(k) = bitreverse32(k mod bs) | (k div bs)
//     where bs = the reference's block size opt_n_threads(n).  Points are dealt to threads so
//     that a thread's own slots are already in ascending key order (same residue k mod bs,
//     ascending k div bs), which lets the inner loop keep the reference's strict '>'.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace nesie {
namespace {

constexpr unsigned FULL = 0xffffffffu;

struct __align__(16) Cand {  // one argmax candidate, 32 bytes
  unsigned d;                // distance bits (non-negative float => unsigned order == float order)
  unsigned p;                // tie-break key, lower wins
  float x, y, z;             // the candidate's coordinates (next centre if it wins)
  unsigned pad[3];
};

__device__ __forceinline__ void store_cand(Cand *dst, unsigned d, unsigned p, float x, float y,
                                           float z) {
  uint4 a = make_uint4(d, p, __float_as_uint(x), __float_as_uint(y));
  uint4 b = make_uint4(__float_as_uint(z), 0u, 0u, 0u);
  reinterpret_cast<uint4 *>(dst)[0] = a;
  reinterpret_cast<uint4 *>(dst)[1] = b;
}

// Argmax over the lanes of a warp: returns the source lane; d/p are replaced by the winner's.
__device__ __forceinline__ int warp_argmax(unsigned &d, unsigned &p) {
  const unsigned wd = __reduce_max_sync(FULL, d);
  const unsigned wp = __reduce_min_sync(FULL, d == wd ? p : 0xffffffffu);
  const unsigned win = __ballot_sync(FULL, d == wd && p == wp);
  d = wd;
  p = wp;
  return __ffs(win) - 1;
}

// ---- cluster exchange primitives (sm_90+ PTX) -------------------------------------------------
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
__device__ __forceinline__ bool mbar_try_wait(unsigned mbar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(mbar), "r"(parity)
      : "memory");
  return ok != 0;
}
// remote 16-byte / 4-byte stores that also signal `bytes` on the destination CTA's mbarrier
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

__device__ __forceinline__ int decode_key(unsigned p, int bs, int log2bs) {
  if (log2bs == 0) return (int)p;
  const unsigned r = __brev(p) & (unsigned)(bs - 1);
  const unsigned q = p & ((1u << (32 - log2bs)) - 1u);
  return (int)(q * (unsigned)bs + r);
}

// ------------------------------------------------------------------------------------------
// Register-resident kernel.  grid = b * CL CTAs (cluster = CL consecutive CTAs = one scene).
// Global thread g = rank*blockDim + tid owns residue r = g mod bs and the slot range
// q in [part*qp, part*qp + qp), part = g div bs; point index k = q*bs + r.
// Dynamic smem: 3*qp*blockDim floats (a copy of the CTA's coordinates, read by winner lanes).
// ------------------------------------------------------------------------------------------
template <int CL, int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT) fps_reg_kernel(int n, int m, int bs, int log2bs, int qp,
                                                       int xmode,
                                                       const float *__restrict__ xyz,
                                                       float *__restrict__ temp,
                                                       int *__restrict__ idx) {
  unsigned rank = 0;
  if constexpr (CL > 1) rank = cg::this_cluster().block_rank();
  const int scene = blockIdx.x / CL;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
  const int g = (int)rank * NT + tid;
  const int r = g & (bs - 1);
  const int q0 = (g >> log2bs) * qp;

  xyz += (size_t)scene * n * 3;
  idx += (size_t)scene * m;
  if (temp) temp += (size_t)scene * n;

  extern __shared__ float s_pts[];  // [3][qp][NT]
  __shared__ Cand s_wk[2][32];
  __shared__ Cand s_cl[2][CL];
  __shared__ __align__(8) unsigned long long s_mbar[2];

  float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    const long long k = (long long)(q0 + s) * bs + r;
    const bool valid = (s < qp) && (k < n);
    float x = 0.f, y = 0.f, z = 0.f, d = -2.f;  // -2 never beats the scan's initial best of -1
    if (valid) {
      x = xyz[k * 3 + 0];
      y = xyz[k * 3 + 1];
      z = xyz[k * 3 + 2];
      d = temp ? temp[k] : 1e10f;  // furthest_point_sample.py:30
    }
    px[s] = x; py[s] = y; pz[s] = z; md[s] = d;
    if (s < qp) {
      s_pts[(0 * qp + s) * NT + tid] = x;
      s_pts[(1 * qp + s) * NT + tid] = y;
      s_pts[(2 * qp + s) * NT + tid] = z;
    }
  }
  const unsigned rbits = log2bs ? __brev((unsigned)r) : 0u;

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

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    float best = -1.f;
    int bslot = 0;
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
      if (s < qp) {
        const float d = sqdist_ref(px[s], py[s], pz[s], cx, cy, cz);
        const float d2 = fminf(d, md[s]);
        md[s] = d2;
        if (d2 > best) { best = d2; bslot = s; }
      }
    }
    unsigned kd = best >= 0.f ? __float_as_uint(best) : 0u;
    unsigned kp = best >= 0.f ? (rbits | (unsigned)(q0 + bslot)) : 0xffffffffu;
    const int src = warp_argmax(kd, kp);
    if (lane == src)
      store_cand(&s_wk[par][warp], kd, kp, s_pts[(0 * qp + bslot) * NT + tid],
                 s_pts[(1 * qp + bslot) * NT + tid], s_pts[(2 * qp + bslot) * NT + tid]);
    __syncthreads();

    if (CL == 1 || warp == 0) {
      unsigned d2 = 0u, p2 = 0xffffffffu;
      float x2 = 0.f, y2 = 0.f, z2 = 0.f;
      if (lane < NW) {
        const Cand c = s_wk[par][lane];
        d2 = c.d; p2 = c.p; x2 = c.x; y2 = c.y; z2 = c.z;
      }
      const int s2 = warp_argmax(d2, p2);
      x2 = __shfl_sync(FULL, x2, s2);
      y2 = __shfl_sync(FULL, y2, s2);
      z2 = __shfl_sync(FULL, z2, s2);
      if constexpr (CL == 1) {
        cx = x2; cy = y2; cz = z2;
        if (tid == 0) idx[j] = decode_key(p2, bs, log2bs);
      } else {
        if (xmode == 0) {  // plain DSMEM stores, published by the cluster barrier below
          if (lane < CL) {
            Cand *dst = cg::this_cluster().map_shared_rank(&s_cl[par][rank], lane);
            store_cand(dst, d2, p2, x2, y2, z2);
          }
        } else {  // st.async: the 20 payload bytes arrive together with their mbarrier signal
          if (lane == 0) mbar_arrive_expect_tx(smem_u32(&s_mbar[par]), CL * 20u);
          if (lane < CL) {
            const unsigned ra = mapa_u32(smem_u32(&s_cl[par][rank]), (unsigned)lane);
            const unsigned rm = mapa_u32(smem_u32(&s_mbar[par]), (unsigned)lane);
            st_async_v4(ra, rm, d2, p2, __float_as_uint(x2), __float_as_uint(y2));
            st_async_b32(ra + 16, rm, __float_as_uint(z2));
          }
        }
      }
    }
    if constexpr (CL > 1) {
      if (xmode == 0) cg::this_cluster().sync();
      else while (!mbar_try_wait(smem_u32(&s_mbar[par]), (unsigned)((j - 1) >> 1) & 1u)) {}
      unsigned d3 = 0u, p3 = 0xffffffffu;
      float x3 = 0.f, y3 = 0.f, z3 = 0.f;
      if (lane < CL) {
        const Cand c = s_cl[par][lane];
        d3 = c.d; p3 = c.p; x3 = c.x; y3 = c.y; z3 = c.z;
      }
      const int s3 = warp_argmax(d3, p3);
      cx = __shfl_sync(FULL, x3, s3);
      cy = __shfl_sync(FULL, y3, s3);
      cz = __shfl_sync(FULL, z3, s3);
      if (rank == 0 && tid == 0) idx[j] = decode_key(p3, bs, log2bs);
    }
  }

  if (temp) {  // hand the final min-distances back, like the reference's in-place buffer
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
      const long long k = (long long)(q0 + s) * bs + r;
      if (s < qp && k < n) temp[k] = md[s];
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

typedef void (*fps_fn)(int, int, int, int, int, int, const float *, float *, int *);

template <int CL, int MAXT>
fps_fn pick_ppt(int qp, int *ppt_out) {
#define NESIE_FPS_CASE(P) if (qp <= P) { *ppt_out = P; return fps_reg_kernel<CL, P, MAXT>; }
  NESIE_FPS_CASE(1) NESIE_FPS_CASE(2) NESIE_FPS_CASE(4) NESIE_FPS_CASE(8)
  if constexpr (MAXT <= 256) { NESIE_FPS_CASE(12) NESIE_FPS_CASE(16) NESIE_FPS_CASE(24) NESIE_FPS_CASE(32) }
#undef NESIE_FPS_CASE
  return nullptr;
}

int launch_reg(fps_fn fn, int CL, int NT, int b, int n, int m, int bs, int log2bs, int qp,
               int xmode, const float *xyz, float *temp, int *idx, cudaStream_t st) {
  const size_t smem = (size_t)3 * qp * NT * sizeof(float);
  if (smem > 32 * 1024)  // static smem (candidate slots) counts against the 48 KB default too
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
  NESIE_CUDA(cudaLaunchKernelEx(&cfg, fn, n, m, bs, log2bs, qp, xmode, xyz, temp, idx));
  return NESIE_OK;
}

// Largest n the register kernel covers with a cluster of CL CTAs (256 threads, 32 slots).
int reg_capacity(int CL) { return CL == 1 ? 8192 : (CL * 256 / 1024) * 32 * 1024; }

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_fps_needs_temp(int b, int n, int m) {
  (void)b; (void)m;
  return n > reg_capacity(16) ? 1 : 0;
}

extern "C" int nesie_fps(int b, int n, int m, const float *xyz, float *temp, int *idx,
                         void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 1 && m >= 0, "need b >= 0, n >= 1, m >= 0");
  NESIE_REQUIRE(xyz && idx, "null pointer");
  if (b == 0 || m == 0) return NESIE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int bs = ref_block_size(n), log2bs = ilog2(bs);
  const int Q = ceil_div(n, bs);

  // tuning knobs (measurement only): cluster size, threads per CTA, exchange mechanism
  int force_cl = 0, force_nt = 0, xmode = 1;
  if (const char *e = getenv("NESIE_FPS_CLUSTER")) force_cl = atoi(e);
  if (const char *e = getenv("NESIE_FPS_THREADS")) force_nt = atoi(e);
  if (const char *e = getenv("NESIE_FPS_XMODE")) xmode = atoi(e);

  if (n <= reg_capacity(16)) {
    int cl_min = 1;
    while (reg_capacity(cl_min) < n) cl_min = cl_min == 1 ? 4 : cl_min * 2;
    int cl = cl_min;
    if (n > reg_capacity(1)) {
      int pref = 4;
      while (pref * 2 <= 16 && b * pref * 2 <= num_sms()) pref *= 2;
      if (pref > cl) cl = pref;
    }
    if (force_cl == 1 || force_cl == 4 || force_cl == 8 || force_cl == 16)
      if (force_cl >= cl_min) cl = force_cl;
    for (; cl >= cl_min; cl = (cl == 4 ? 1 : cl / 2)) {
      int NT = cl == 1 ? (bs < 32 ? 32 : bs) : 256;
      if (cl > 1 && (force_nt == 64 || force_nt == 128) && cl * force_nt >= bs) NT = force_nt;
      const int parts = cl * NT / bs;
      const int qp = ceil_div(Q, parts);
      int ppt = 0;
      fps_fn fn = nullptr;
      switch (cl) {
        case 1: fn = pick_ppt<1, 1024>(qp, &ppt); break;
        case 4: fn = pick_ppt<4, 256>(qp, &ppt); break;
        case 8: fn = pick_ppt<8, 256>(qp, &ppt); break;
        case 16: fn = pick_ppt<16, 256>(qp, &ppt); break;
      }
      if (!fn) { if (cl == 1) break; continue; }
      const int rc = launch_reg(fn, cl, NT, b, n, m, bs, log2bs, qp, xmode, xyz, temp, idx, st);
      if (rc != NESIE_ERR_UNSUPPORTED) return rc;
      if (cl == 1) break;
    }
  }
  // global-memory fallback
  if (!temp) {
    set_error("nesie_fps: n=%d needs the global-memory kernel, which requires temp", n);
    return NESIE_ERR_UNSUPPORTED;
  }
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
