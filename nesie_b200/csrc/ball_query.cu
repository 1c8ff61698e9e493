// Ball query for sm_100a.
//
// Replaces ball_query_kernel (reference: ops/ball_query/src/ball_query_cuda.cu:11-54), which
// gives every centre ONE thread that walks all n points serially with 12-byte broadcast loads
// and up to nsample scattered 4-byte stores per hit.
//
// Here a WARP owns C centres and its 32 lanes own 32 consecutive points per step, so the
// "first nsample hits in ascending index order" contract falls out of a ballot + prefix
// popcount with no sorting, the early exit (cnt >= nsample) is warp-uniform, and the point
// stream is staged once per CTA in shared memory (coalesced 16-byte loads) and reused by all
// 8 warps x C centres.  Hits are collected in shared memory and each idx row is written once,
// fully coalesced, including the reference's padding rule (slots >= cnt repeat the first hit,
// rows with no hit are zeros -- the caller's zero fill, ball_query.py:35).
//
// Bit-exactness: d2 = fma(dz,dz,fma(dx,dx,dy*dy)) with d = centre - point, radii squared in
// fp32 (ball_query_cuda.cu:30-31), hit <=> d2 == 0 || (d2 >= min_r2 && d2 < max_r2).
#include "common.cuh"

namespace nesie {
namespace {

constexpr int BQ_THREADS = 256;
constexpr int BQ_WARPS = BQ_THREADS / 32;
constexpr int BQ_TILE = 2048;  // points staged per step (24 KB)

template <int C>
__global__ void __launch_bounds__(BQ_THREADS, 4) ball_query_kernel(
    int n, int m, float min_r2, float max_r2, int nsample, const float *__restrict__ new_xyz,
    const float *__restrict__ xyz, int *__restrict__ idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_xyz = reinterpret_cast<float *>(smem_raw);                  // [BQ_TILE*3]
  int *s_hits = reinterpret_cast<int *>(s_xyz + BQ_TILE * 3);          // [BQ_WARPS][C][nsample]

  const int scene = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = (blockIdx.x * BQ_WARPS + warp) * C;  // first centre of this warp
  xyz += (size_t)scene * n * 3;
  new_xyz += (size_t)scene * m * 3;
  idx += (size_t)scene * m * nsample;
  int *hits = s_hits + (size_t)warp * C * nsample;

  float cx[C], cy[C], cz[C];
  int cnt[C], first[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int ci = c0 + c;
    const bool ok = ci < m;
    cx[c] = ok ? new_xyz[ci * 3 + 0] : 0.f;
    cy[c] = ok ? new_xyz[ci * 3 + 1] : 0.f;
    cz[c] = ok ? new_xyz[ci * 3 + 2] : 0.f;
    cnt[c] = ok ? 0 : nsample;  // out-of-range centres are "already full"
    first[c] = 0;
  }
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
    const int tn = min(BQ_TILE, n - t0);
    // every warp still has work?  (uniform across the CTA via the barrier's reduction)
    bool warp_open = false;
#pragma unroll
    for (int c = 0; c < C; ++c) warp_open |= cnt[c] < nsample;
    if (!__syncthreads_or(warp_open)) break;
    // stage tile: 3*tn floats, contiguous in global memory
    {
      const float *src = xyz + (size_t)t0 * 3;
      const int nf = tn * 3;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nv = nf >> 2;
        for (int i = tid; i < nv; i += BQ_THREADS)
          reinterpret_cast<float4 *>(s_xyz)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
        for (int i = (nv << 2) + tid; i < nf; i += BQ_THREADS) s_xyz[i] = __ldg(src + i);
      } else {
        for (int i = tid; i < nf; i += BQ_THREADS) s_xyz[i] = __ldg(src + i);
      }
    }
    __syncthreads();
    if (warp_open) {
      for (int k0 = 0; k0 < tn; k0 += 32) {
        const int kl = k0 + lane;
        const bool in = kl < tn;
        const float x = in ? s_xyz[kl * 3 + 0] : 0.f;
        const float y = in ? s_xyz[kl * 3 + 1] : 0.f;
        const float z = in ? s_xyz[kl * 3 + 2] : 0.f;
        bool hit[C];
        bool any = false;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float d2 = sqdist_ref(cx[c], cy[c], cz[c], x, y, z);
          hit[c] = in && (d2 == 0.f || (d2 >= min_r2 && d2 < max_r2));
          any |= hit[c];
        }
        // a ball holds a few dozen of the n points: most 32-point steps have no hit at all, so
        // one vote per step (not one per centre) is the common path
        if (__any_sync(0xffffffffu, any)) {
          bool open = false;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const unsigned b = __ballot_sync(0xffffffffu, hit[c]);
            if (b && cnt[c] < nsample) {
              const int pos = cnt[c] + __popc(b & lt_mask);
              if (hit[c] && pos < nsample) hits[c * nsample + pos] = t0 + kl;
              if (cnt[c] == 0) first[c] = t0 + k0 + __ffs(b) - 1;
              cnt[c] += __popc(b);
            }
            open |= cnt[c] < nsample;
          }
          if (!open) break;  // every centre of this warp is full
        }
      }
    }
  }
  __syncwarp();
  // write the rows: [0,cnt) = hits in index order, [cnt,nsample) = first hit (0 if none)
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int ci = c0 + c;
    if (ci < m) {
      const int have = min(cnt[c], nsample);
      for (int l = lane; l < nsample; l += 32)
        idx[(size_t)ci * nsample + l] = l < have ? hits[c * nsample + l] : first[c];
    }
  }
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_ball_query(int b, int n, int m, float min_radius, float max_radius,
                                int nsample, const float *new_xyz, const float *xyz, int *idx,
                                void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, "negative size");
  NESIE_REQUIRE(new_xyz && xyz && idx, "null pointer");
  if (b == 0 || m == 0 || nsample == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  const float max_r2 = max_radius * max_radius;  // fp32, ball_query_cuda.cu:30-31
  const float min_r2 = min_radius * min_radius;
  // centres per warp: as many as keep >= ~2 CTAs per SM in flight
  int C = 4;
  while (C > 1 && (long long)b * ceil_div(m, BQ_WARPS * C) < 2LL * num_sms()) C >>= 1;
  const size_t smem = (size_t)BQ_TILE * 3 * sizeof(float) + (size_t)BQ_WARPS * C * nsample * 4;
  NESIE_REQUIRE(smem <= 200 * 1024, "nsample too large for the shared-memory hit buffer");
  dim3 grid(ceil_div(m, BQ_WARPS * C), b);
#define NESIE_BQ_LAUNCH(CC)                                                                    \
  do {                                                                                         \
    if (smem > 48 * 1024)                                                                      \
      NESIE_CUDA(cudaFuncSetAttribute(ball_query_kernel<CC>,                                   \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ball_query_kernel<CC><<<grid, BQ_THREADS, smem, st>>>(n, m, min_r2, max_r2, nsample,       \
                                                          new_xyz, xyz, idx);                  \
  } while (0)
  if (C == 4) NESIE_BQ_LAUNCH(4);
  else if (C == 2) NESIE_BQ_LAUNCH(2);
  else NESIE_BQ_LAUNCH(1);
#undef NESIE_BQ_LAUNCH
  return check_launch("nesie_ball_query");
}
