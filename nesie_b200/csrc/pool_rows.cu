// Small row-group kernels around the pooled layers of the SidePooling MiniPointNets (reference:
// models/dense_heads/side_pooling_module.py:343-370).  With the maximum over a box's grid points taken
// in the GEMM's epilogue (gemm_tma.cuh, pool_max / pool_amax) the pooled convolution's output never
// exists in HBM; what is left is
//
//   pool_finalize       unit maxima of the epilogue (16 / 32 rows) -> group maxima (k rows) + conv bias
//   pool_wgrad          d_w[c, :] = sum_g d_out[g, c] * a[g * k + arg[g, c], :]  -- the weight gradient
//                       of a max-pooled convolution touches ONE input row per (group, channel): a gather
//                       of groups x n rows instead of a rows x n x K contraction over a dense d_y
//   group_sum_rows      sum of every k consecutive rows (the gradient of a per-group input that was
//                       multiplied through the weights once per group, GemmTmaParams::grp_bias)
//   scatter_rows_add    d_x[g * k + arg[g, c], c] += d[g, c]  (the max's gradient, added in place)
#include "common.cuh"

namespace nesie {
namespace {

constexpr int PR_THREADS = 256;

__global__ void __launch_bounds__(PR_THREADS) pool_finalize_kernel(
    long long groups, int m, int u, int n, const float *__restrict__ pmax,
    const unsigned char *__restrict__ amax, const float *__restrict__ bias, float *__restrict__ out,
    unsigned char *__restrict__ arg) {
  const int tpr = n >> 2;
  const long long total = groups * tpr;
  for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * PR_THREADS) {
    const long long g = i / tpr;
    const int ch = (int)(i - g * tpr) * 4;
    const long long base = (g * m) * n + ch;
    float4 best = *reinterpret_cast<const float4 *>(pmax + base);
    uchar4 a = *reinterpret_cast<const uchar4 *>(amax + base);
    int bx = a.x, by = a.y, bz = a.z, bw = a.w;
    for (int j = 1; j < m; ++j) {   // strict comparisons: the first maximising row wins
      const float4 v = *reinterpret_cast<const float4 *>(pmax + base + (long long)j * n);
      a = *reinterpret_cast<const uchar4 *>(amax + base + (long long)j * n);
      if (v.x > best.x) { best.x = v.x; bx = j * u + a.x; }
      if (v.y > best.y) { best.y = v.y; by = j * u + a.y; }
      if (v.z > best.z) { best.z = v.z; bz = j * u + a.z; }
      if (v.w > best.w) { best.w = v.w; bw = j * u + a.w; }
    }
    if (bias) {
      const float4 b = *reinterpret_cast<const float4 *>(bias + ch);
      best.x += b.x; best.y += b.y; best.z += b.z; best.w += b.w;
    }
    *reinterpret_cast<float4 *>(out + g * n + ch) = best;
    *reinterpret_cast<uchar4 *>(arg + g * n + ch) =
        make_uchar4((unsigned char)bx, (unsigned char)by, (unsigned char)bz, (unsigned char)bw);
  }
}

// One CTA of 64 threads = 32 output channels (blockIdx.y) x all K input columns (4 per thread); it
// walks groups blockIdx.x, blockIdx.x + gridDim.x, ... with the 32 x 4 sums of a thread in registers
// and leaves ONE partial block per blockIdx.x.  The k rows of a group are staged in shared memory
// (with the previous layer's BatchNorm + ReLU applied when scale / shift are given: the operand is
// relu(y_prev * scale + shift), which is never stored), so a (group, channel) pair costs one 16-byte
// shared-memory read and four FMAs per thread.  Several CTAs per SM hide the staging latency.
constexpr int PW_THREADS = 64;
constexpr int PW_CH = 32;

__global__ void __launch_bounds__(PW_THREADS) pool_wgrad_kernel(
    long long groups, int k, int n, int kk, const float *__restrict__ d_out,
    const unsigned char *__restrict__ arg, const float *__restrict__ y_prev,
    const float *__restrict__ scale, const float *__restrict__ shift, float *__restrict__ dw_part) {
  extern __shared__ __align__(16) unsigned char pw_smem[];
  float4 *s_a = reinterpret_cast<float4 *>(pw_smem);          // [k][kk / 4]
  __shared__ float s_d[PW_CH];
  __shared__ int s_j[PW_CH];
  const int t = threadIdx.x, k4 = kk >> 2, c0 = blockIdx.y * PW_CH;
  const bool own = t < k4;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc[PW_CH];
#pragma unroll
  for (int i = 0; i < PW_CH; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int items = k * k4;
  for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
    const float4 *src = reinterpret_cast<const float4 *>(y_prev + g * k * (long long)kk);
    for (int i = t; i < items; i += PW_THREADS) {
      float4 v = __ldg(src + i);
      if (scale) {
        const int col = (i % k4) * 4;
        sc = __ldg(reinterpret_cast<const float4 *>(scale + col));
        sh = __ldg(reinterpret_cast<const float4 *>(shift + col));
        v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
        v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
      }
      s_a[i] = v;
    }
    if (t < PW_CH) {
      const bool ok = c0 + t < n;
      s_d[t] = ok ? d_out[g * n + c0 + t] : 0.f;
      s_j[t] = ok ? (int)arg[g * n + c0 + t] : 0;
    }
    __syncthreads();
    if (own) {
#pragma unroll
      for (int ci = 0; ci < PW_CH; ++ci) {
        const float d = s_d[ci];
        const float4 a = s_a[s_j[ci] * k4 + t];
        acc[ci].x = fmaf(d, a.x, acc[ci].x); acc[ci].y = fmaf(d, a.y, acc[ci].y);
        acc[ci].z = fmaf(d, a.z, acc[ci].z); acc[ci].w = fmaf(d, a.w, acc[ci].w);
      }
    }
    __syncthreads();
  }
  if (own) {
    float *dst = dw_part + ((size_t)blockIdx.x * n + c0) * kk + 4 * t;
#pragma unroll
    for (int ci = 0; ci < PW_CH; ++ci)
      if (c0 + ci < n) *reinterpret_cast<float4 *>(dst + (size_t)ci * kk) = acc[ci];
  }
}

__global__ void __launch_bounds__(PR_THREADS) group_sum_rows_kernel(
    long long groups, int k, int n, const float *__restrict__ x, float *__restrict__ out) {
  const int tpr = n >> 2;
  const long long total = groups * tpr;
  for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * PR_THREADS) {
    const long long g = i / tpr;
    const int ch = (int)(i - g * tpr) * 4;
    const float *row = x + (g * k) * n + ch;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < k; ++r) {
      const float4 v = *reinterpret_cast<const float4 *>(row + (long long)r * n);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4 *>(out + g * n + ch) = s;
  }
}

__global__ void __launch_bounds__(PR_THREADS) scatter_rows_add_kernel(
    long long groups, int k, int n, const float *__restrict__ d, const unsigned char *__restrict__ arg,
    float *__restrict__ d_x) {
  const long long total = groups * n;
  for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * PR_THREADS) {
    const long long g = i / n;
    const int c = (int)(i - g * n);
    d_x[(g * k + arg[i]) * n + c] += d[i];   // (row, channel) is unique per (group, channel): no race
  }
}

int pr_grid(long long items) {
  long long g = (items + PR_THREADS - 1) / PR_THREADS;
  const long long cap = 16LL * num_sms();
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int pw_grid_x(long long groups) {
  const long long cap = num_sms();   // x 4 channel blocks at n = 128: four CTAs per SM
  return (int)(groups < cap ? groups : cap);
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_pool_finalize(long long groups, int k, int u, int n, const float *pmax,
                                   const unsigned char *amax, const float *bias, float *out,
                                   unsigned char *arg, void *stream) {
  NESIE_REQUIRE(groups >= 0 && u >= 1 && k >= u && k % u == 0 && k <= 255 && n >= 4 && (n & 3) == 0,
                "need u | k, k <= 255, n % 4 == 0");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(pmax && amax && out && arg, "null pointer");
  pool_finalize_kernel<<<pr_grid(groups * (n >> 2)), PR_THREADS, 0, (cudaStream_t)stream>>>(
      groups, k / u, u, n, pmax, amax, bias, out, arg);
  return check_launch("nesie_pool_finalize");
}

extern "C" int nesie_pool_wgrad_parts(long long groups) { return groups <= 0 ? 0 : pw_grid_x(groups); }

extern "C" int nesie_pool_wgrad(long long groups, int k, int n, int kk, const float *d_out,
                                const unsigned char *arg, const float *y_prev, const float *scale,
                                const float *shift, float *dw_part, void *stream) {
  NESIE_REQUIRE(groups >= 1 && k >= 1 && k <= 255 && n >= 1, "need groups >= 1, 1 <= k <= 255, n >= 1");
  NESIE_REQUIRE(kk >= 4 && (kk & 3) == 0 && kk <= 4 * PW_THREADS, "need k_in % 4 == 0, k_in <= 256");
  NESIE_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  NESIE_REQUIRE(d_out && arg && y_prev && dw_part, "null pointer");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(y_prev) & 15) == 0 && (reinterpret_cast<uintptr_t>(dw_part) & 15) == 0,
                "y_prev and dw_part must be 16-byte aligned");
  const size_t smem = (size_t)k * kk * sizeof(float);
  NESIE_REQUIRE(smem <= 200 * 1024, "group does not fit shared memory");
  if (smem > 48 * 1024)
    NESIE_CUDA(cudaFuncSetAttribute(pool_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid(pw_grid_x(groups), (n + PW_CH - 1) / PW_CH);
  pool_wgrad_kernel<<<grid, PW_THREADS, smem, (cudaStream_t)stream>>>(groups, k, n, kk, d_out, arg, y_prev,
                                                                      scale, shift, dw_part);
  return check_launch("nesie_pool_wgrad");
}

extern "C" int nesie_group_sum_rows(long long groups, int k, int n, const float *x, float *out,
                                    void *stream) {
  NESIE_REQUIRE(groups >= 0 && k >= 1 && n >= 4 && (n & 3) == 0, "need k >= 1, n % 4 == 0");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(x && out, "null pointer");
  group_sum_rows_kernel<<<pr_grid(groups * (n >> 2)), PR_THREADS, 0, (cudaStream_t)stream>>>(groups, k, n, x, out);
  return check_launch("nesie_group_sum_rows");
}

extern "C" int nesie_scatter_rows_add(long long groups, int k, int n, const float *d,
                                      const unsigned char *arg, float *d_x, void *stream) {
  NESIE_REQUIRE(groups >= 0 && k >= 1 && k <= 255 && n >= 1, "need 1 <= k <= 255, n >= 1");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(d && arg && d_x, "null pointer");
  scatter_rows_add_kernel<<<pr_grid(groups * n), PR_THREADS, 0, (cudaStream_t)stream>>>(groups, k, n, d, arg, d_x);
  return check_launch("nesie_scatter_rows_add");
}
