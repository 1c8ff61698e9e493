// Max over groups of k consecutive rows of a row-major activation, for the MiniPointNets of the
// SidePooling head (reference: models/dense_heads/side_pooling_module.py:360-370):
//   feature        = x + bias
//   feature_global = max over the k grid points of a box           (torch.max(feature, dim=-1))
//   concat == 0 :  out (groups, c)        = feature_global
//   concat == 1 :  out (groups * k, 2c)   = [ feature_global broadcast to the box's rows | feature ]
//                  (torch.cat([feature_global.expand(...), feature], dim=1))
// and the matching backward: the gradient of the max goes to the first maximising row (arg, u8), the
// way torch.max(dim) routes it through its index.  One thread owns 4 channels of one group; the
// reference formulation costs an expand + cat (+ eq / mul / div / sum kernels in the backward).
#include "common.cuh"

namespace nesie {
namespace {

constexpr int GM_THREADS = 256;

template <bool CONCAT>
__global__ void __launch_bounds__(GM_THREADS) group_max_fwd_kernel(
    long long groups, int k, int c, const float *__restrict__ x, const float *__restrict__ bias,
    float *__restrict__ out, unsigned char *__restrict__ arg) {
  const int tpr = c >> 2;
  const long long total = groups * tpr;
  for (long long i = (long long)blockIdx.x * GM_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * GM_THREADS) {
    const long long g = i / tpr;
    const int ch = (int)(i - g * tpr) * 4;
    const float4 b = bias ? *reinterpret_cast<const float4 *>(bias + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float *row = x + (g * k) * c + ch;
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int bx = 0, by = 0, bz = 0, bw = 0;
    for (int r = 0; r < k; ++r) {
      const float4 v = *reinterpret_cast<const float4 *>(row + (long long)r * c);
      const float vx = v.x + b.x, vy = v.y + b.y, vz = v.z + b.z, vw = v.w + b.w;
      if (vx > best.x) { best.x = vx; bx = r; }
      if (vy > best.y) { best.y = vy; by = r; }
      if (vz > best.z) { best.z = vz; bz = r; }
      if (vw > best.w) { best.w = vw; bw = r; }
      if (CONCAT)
        *reinterpret_cast<float4 *>(out + (g * k + r) * (2LL * c) + c + ch) = make_float4(vx, vy, vz, vw);
    }
    if (CONCAT) {
      for (int r = 0; r < k; ++r) *reinterpret_cast<float4 *>(out + (g * k + r) * (2LL * c) + ch) = best;
    } else {
      *reinterpret_cast<float4 *>(out + g * c + ch) = best;
    }
    *reinterpret_cast<uchar4 *>(arg + g * c + ch) =
        make_uchar4((unsigned char)bx, (unsigned char)by, (unsigned char)bz, (unsigned char)bw);
  }
}

// d_bias_part (nullable): (gridDim.x, c) per-CTA column sums of d_x (= the gradient of the folded conv
// bias), summed by the caller; requires (c / 4) | GM_THREADS so that a thread keeps its 4 channels over
// the grid-stride loop (ATen's column sum of the tall d_x matrix cost more than this whole kernel).
template <bool CONCAT>
__global__ void __launch_bounds__(GM_THREADS) group_max_bwd_kernel(
    long long groups, int k, int c, const float *__restrict__ d_out,
    const unsigned char *__restrict__ arg, float *__restrict__ d_x, float *__restrict__ d_bias_part) {
  __shared__ float4 s_red[GM_THREADS];
  const int tpr = c >> 2;
  const long long total = groups * tpr;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * GM_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * GM_THREADS) {
    const long long g = i / tpr;
    const int ch = (int)(i - g * tpr) * 4;
    const uchar4 a = *reinterpret_cast<const uchar4 *>(arg + g * c + ch);
    float4 s;  // gradient of the group's maximum
    if (CONCAT) {
      s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < k; ++r) {
        const float4 v = *reinterpret_cast<const float4 *>(d_out + (g * k + r) * (2LL * c) + ch);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    } else {
      s = *reinterpret_cast<const float4 *>(d_out + g * c + ch);
    }
    for (int r = 0; r < k; ++r) {
      float4 v = CONCAT ? *reinterpret_cast<const float4 *>(d_out + (g * k + r) * (2LL * c) + c + ch)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      if (r == a.x) v.x += s.x;
      if (r == a.y) v.y += s.y;
      if (r == a.z) v.z += s.z;
      if (r == a.w) v.w += s.w;
      *reinterpret_cast<float4 *>(d_x + (g * k + r) * c + ch) = v;
      if (CONCAT) { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    if (!CONCAT) { acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w; }
  }
  if (d_bias_part) {   // fixed-order tree over the threads that own the same channels: deterministic
    s_red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < tpr) {
      float4 t = s_red[threadIdx.x];
      for (int j = threadIdx.x + tpr; j < GM_THREADS; j += tpr) {
        const float4 u = s_red[j];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      *reinterpret_cast<float4 *>(d_bias_part + (size_t)blockIdx.x * c + threadIdx.x * 4) = t;
    }
  }
}

int gm_grid(long long items) {
  long long g = (items + GM_THREADS - 1) / GM_THREADS;
  const long long cap = 16LL * num_sms();
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_group_max_rows_forward(long long groups, int k, int c, const float *x,
                                            const float *bias, float *out, unsigned char *arg,
                                            int concat, void *stream) {
  NESIE_REQUIRE(groups >= 0 && k >= 1 && k <= 255 && c >= 4 && (c & 3) == 0, "need 1 <= k <= 255, c % 4 == 0");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(x && out && arg, "null pointer");
  const int grid = gm_grid(groups * (c >> 2));
  if (concat)
    group_max_fwd_kernel<true><<<grid, GM_THREADS, 0, (cudaStream_t)stream>>>(groups, k, c, x, bias, out, arg);
  else
    group_max_fwd_kernel<false><<<grid, GM_THREADS, 0, (cudaStream_t)stream>>>(groups, k, c, x, bias, out, arg);
  return check_launch("nesie_group_max_rows_forward");
}

extern "C" int nesie_group_max_bias_parts(long long groups, int c) {
  if (c < 4 || (c & 3) || GM_THREADS % (c >> 2)) return 0;   // 0: column sums not available in-kernel
  return gm_grid(groups * (c >> 2));
}

extern "C" int nesie_group_max_rows_backward(long long groups, int k, int c, const float *d_out,
                                             const unsigned char *arg, float *d_x, int concat,
                                             float *d_bias_part, void *stream) {
  NESIE_REQUIRE(groups >= 0 && k >= 1 && k <= 255 && c >= 4 && (c & 3) == 0, "need 1 <= k <= 255, c % 4 == 0");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(d_out && arg && d_x, "null pointer");
  NESIE_REQUIRE(!d_bias_part || GM_THREADS % (c >> 2) == 0, "d_bias_part needs (c / 4) | 256");
  const int grid = gm_grid(groups * (c >> 2));
  if (concat)
    group_max_bwd_kernel<true><<<grid, GM_THREADS, 0, (cudaStream_t)stream>>>(groups, k, c, d_out, arg, d_x, d_bias_part);
  else
    group_max_bwd_kernel<false><<<grid, GM_THREADS, 0, (cudaStream_t)stream>>>(groups, k, c, d_out, arg, d_x, d_bias_part);
  return check_launch("nesie_group_max_rows_backward");
}
