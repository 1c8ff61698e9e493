// A linear layer commuted with the gather in front of it.
//
// The first convolution of a SidePooling MiniPointNet runs on rows [ grid point - box centre | seed
// features interpolated from the 3 nearest seeds ] (reference: models/dense_heads/
// side_pooling_module.py:183-243,343-358), the first convolution of an SA module on rows
// [ (neighbour - centre) / radius | features of the neighbour ] (ops/group_points/group_points.py:
// 104-160 + ops/pointnet_modules/point_sa_module.py:191-211).  Both are  rows @ W^T  with rows that
// are (weighted) copies of a few thousand table rows, so
//
//   y[r, :] = sum_j w[r, j] * (table @ W_f^T)[idx[r, j], :]  +  head[r, :] @ W_x^T
//
// (head given, or computed as (xyz[idx] - center) / radius for the SA grouping)
//
// needs the GEMM only over the TABLE (8192 seeds instead of 65536 .. 262144 grid rows per box set);
// what is left per row is this gather of already-transformed rows plus 3 FMAs per output for the
// coordinate columns.  The kernel also leaves the column sums of y and y^2 (the BatchNorm statistics
// the GEMM epilogue used to deliver).  Backward: d_table[idx[r, j], :] += w[r, j] * d_y[r, :] (vector
// reductions into the L2-resident table, as three_interpolate_grad does), d_W_x from per-CTA partial
// sums; the table's own weight gradient is then an ordinary GEMM over the table rows.
#include "common.cuh"

namespace nesie {
namespace {

constexpr int GL_THREADS = 256;

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// grid (gx, b); thread = 4 columns of one of GL_THREADS / (c / 4) row lanes
template <int J>
__global__ void __launch_bounds__(GL_THREADS) gather_linear_fwd_kernel(
    int c, int m, int n, const float *__restrict__ table, const int *__restrict__ idx,
    const float *__restrict__ weight, const float *__restrict__ head, const float *__restrict__ wx,
    const float *__restrict__ xyz, const float *__restrict__ center, int ns, float inv_radius,
    float *__restrict__ y, float *__restrict__ col_parts) {
  __shared__ float4 s_red[GL_THREADS];
  const int tpr = c >> 2, lanes = GL_THREADS / tpr;
  const int c4 = threadIdx.x % tpr, lane = threadIdx.x / tpr;
  const int b = blockIdx.y;
  const float4 *tab = reinterpret_cast<const float4 *>(table + (size_t)b * m * c) + c4;
  float4 wx0 = make_float4(0.f, 0.f, 0.f, 0.f), wx1 = wx0, wx2 = wx0;   // columns 0..2 of W_x for my 4 outputs
  if (wx) {
    const float *w = wx + (size_t)c4 * 12;
    wx0 = make_float4(__ldg(w), __ldg(w + 3), __ldg(w + 6), __ldg(w + 9));
    wx1 = make_float4(__ldg(w + 1), __ldg(w + 4), __ldg(w + 7), __ldg(w + 10));
    wx2 = make_float4(__ldg(w + 2), __ldg(w + 5), __ldg(w + 8), __ldg(w + 11));
  }
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  for (long long t = (long long)blockIdx.x * lanes + lane; t < n; t += (long long)gridDim.x * lanes) {
    const size_t r = (size_t)b * n + t;
    float4 v;
    int a0;
    if (J == 3) {
      a0 = __ldg(idx + r * 3);
      const int a1 = __ldg(idx + r * 3 + 1), a2 = __ldg(idx + r * 3 + 2);
      const float w0 = __ldg(weight + r * 3), w1 = __ldg(weight + r * 3 + 1), w2 = __ldg(weight + r * 3 + 2);
      const float4 p0 = __ldg(tab + (size_t)a0 * tpr), p1 = __ldg(tab + (size_t)a1 * tpr), p2 = __ldg(tab + (size_t)a2 * tpr);
      v.x = fmaf(w2, p2.x, fmaf(w0, p0.x, w1 * p1.x)); v.y = fmaf(w2, p2.y, fmaf(w0, p0.y, w1 * p1.y));
      v.z = fmaf(w2, p2.z, fmaf(w0, p0.z, w1 * p1.z)); v.w = fmaf(w2, p2.w, fmaf(w0, p0.w, w1 * p1.w));
    } else {
      a0 = __ldg(idx + r);
      v = __ldg(tab + (size_t)a0 * tpr);
      if (weight) { const float w0 = __ldg(weight + r); v.x *= w0; v.y *= w0; v.z *= w0; v.w *= w0; }
    }
    if (wx) {
      float h0, h1, h2;
      if (head) {
        h0 = __ldg(head + r * 3); h1 = __ldg(head + r * 3 + 1); h2 = __ldg(head + r * 3 + 2);
      } else {   // SA grouping: (neighbour - centre) [* 1 / radius], as nesie_group_rows writes it
        const float *p = xyz + ((size_t)b * m + a0) * 3, *q = center + ((size_t)b * (n / ns) + t / ns) * 3;
        h0 = __fsub_rn(__ldg(p), __ldg(q)); h1 = __fsub_rn(__ldg(p + 1), __ldg(q + 1)); h2 = __fsub_rn(__ldg(p + 2), __ldg(q + 2));
        if (inv_radius > 0.f) { h0 = __fmul_rn(h0, inv_radius); h1 = __fmul_rn(h1, inv_radius); h2 = __fmul_rn(h2, inv_radius); }
      }
      v.x = fmaf(h2, wx2.x, fmaf(h1, wx1.x, fmaf(h0, wx0.x, v.x))); v.y = fmaf(h2, wx2.y, fmaf(h1, wx1.y, fmaf(h0, wx0.y, v.y)));
      v.z = fmaf(h2, wx2.z, fmaf(h1, wx1.z, fmaf(h0, wx0.z, v.z))); v.w = fmaf(h2, wx2.w, fmaf(h1, wx1.w, fmaf(h0, wx0.w, v.w)));
    }
    reinterpret_cast<float4 *>(y + r * c)[c4] = v;
    s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
    s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
  }
  if (!col_parts) return;
  float *dst = col_parts + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 2 * c;
#pragma unroll
  for (int h = 0; h < 2; ++h) {          // lanes combined in lane order: deterministic
    __syncthreads();
    s_red[threadIdx.x] = h ? s2 : s1;
    __syncthreads();
    if (threadIdx.x < tpr) {
      float4 a = s_red[threadIdx.x];
      for (int l = 1; l < lanes; ++l) {
        const float4 u = s_red[l * tpr + threadIdx.x];
        a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      }
      reinterpret_cast<float4 *>(dst + h * c)[threadIdx.x] = a;
    }
  }
}

template <int J>
__global__ void __launch_bounds__(GL_THREADS) gather_linear_bwd_kernel(
    int c, int m, int n, const float *__restrict__ d_y, const int *__restrict__ idx,
    const float *__restrict__ weight, const float *__restrict__ head, const float *__restrict__ xyz,
    const float *__restrict__ center, int ns, float inv_radius, float *__restrict__ d_table,
    float *__restrict__ dwx_part) {
  __shared__ float4 s_red[GL_THREADS];
  const int tpr = c >> 2, lanes = GL_THREADS / tpr;
  const int c4 = threadIdx.x % tpr, lane = threadIdx.x / tpr;
  const int b = blockIdx.y;
  float *tab = d_table + (size_t)b * m * c + 4 * c4;
  float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;   // d W_x[:, 0..2] of my 4 outputs
  for (long long t = (long long)blockIdx.x * lanes + lane; t < n; t += (long long)gridDim.x * lanes) {
    const size_t r = (size_t)b * n + t;
    const float4 d = __ldg(reinterpret_cast<const float4 *>(d_y + r * c) + c4);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int a = __ldg(idx + r * J + j);
      const float w = weight ? __ldg(weight + r * J + j) : 1.f;
      red_add_v4(tab + (size_t)a * c, make_float4(w * d.x, w * d.y, w * d.z, w * d.w));
    }
    if (dwx_part) {
      float h0, h1, h2;
      if (head) {
        h0 = __ldg(head + r * 3); h1 = __ldg(head + r * 3 + 1); h2 = __ldg(head + r * 3 + 2);
      } else {
        const float *p = xyz + ((size_t)b * m + __ldg(idx + r * J)) * 3, *q = center + ((size_t)b * (n / ns) + t / ns) * 3;
        h0 = __fsub_rn(__ldg(p), __ldg(q)); h1 = __fsub_rn(__ldg(p + 1), __ldg(q + 1)); h2 = __fsub_rn(__ldg(p + 2), __ldg(q + 2));
        if (inv_radius > 0.f) { h0 = __fmul_rn(h0, inv_radius); h1 = __fmul_rn(h1, inv_radius); h2 = __fmul_rn(h2, inv_radius); }
      }
      g0.x = fmaf(d.x, h0, g0.x); g0.y = fmaf(d.y, h0, g0.y); g0.z = fmaf(d.z, h0, g0.z); g0.w = fmaf(d.w, h0, g0.w);
      g1.x = fmaf(d.x, h1, g1.x); g1.y = fmaf(d.y, h1, g1.y); g1.z = fmaf(d.z, h1, g1.z); g1.w = fmaf(d.w, h1, g1.w);
      g2.x = fmaf(d.x, h2, g2.x); g2.y = fmaf(d.y, h2, g2.y); g2.z = fmaf(d.z, h2, g2.z); g2.w = fmaf(d.w, h2, g2.w);
    }
  }
  if (!dwx_part) return;
  // partial block [c][4] (fourth entry 0) per CTA: rows of W_x for my 4 outputs
  float *dst = dwx_part + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * c * 4;
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    __syncthreads();
    s_red[threadIdx.x] = h == 0 ? g0 : (h == 1 ? g1 : g2);
    __syncthreads();
    if (threadIdx.x < tpr) {
      float4 a = s_red[threadIdx.x];
      for (int l = 1; l < lanes; ++l) {
        const float4 u = s_red[l * tpr + threadIdx.x];
        a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      }
      float *o = dst + (size_t)threadIdx.x * 16 + h;
      o[0] = a.x; o[4] = a.y; o[8] = a.z; o[12] = a.w;
      if (h == 0) { o[3] = 0.f; o[7] = 0.f; o[11] = 0.f; o[15] = 0.f; }
    }
  }
}

int gl_grid_x(int c, int n, int b) {
  const int lanes = GL_THREADS / (c >> 2);
  long long gx = ((long long)n + lanes * 8 - 1) / (lanes * 8);      // >= 8 rows per thread
  const long long cap = (4LL * num_sms() + b - 1) / b;
  if (gx > cap) gx = cap;
  return gx < 1 ? 1 : (int)gx;
}

bool gl_shape_ok(int c) {
  const int tpr = c >> 2;
  return c >= 4 && (c & 3) == 0 && tpr <= GL_THREADS && (tpr & (tpr - 1)) == 0;
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_gather_linear_parts(int b, int c, int n) {
  if (b <= 0 || n <= 0 || !gl_shape_ok(c)) return 0;
  return gl_grid_x(c, n, b) * b;
}

extern "C" int nesie_gather_linear_forward(int b, int c, int m, int n, int j, const float *table,
                                           const int *idx, const float *weight, const float *head,
                                           const float *wx, const float *xyz, const float *center,
                                           int ns, float radius, float *y, float *col_parts,
                                           void *stream) {
  NESIE_REQUIRE(b >= 0 && m >= 1 && n >= 0 && (j == 1 || j == 3), "need b >= 0, m >= 1, n >= 0, j in {1, 3}");
  NESIE_REQUIRE(gl_shape_ok(c), "c / 4 must be a power of two <= 256");
  if (b == 0 || n == 0) return NESIE_OK;
  NESIE_REQUIRE(table && idx && y, "null pointer");
  NESIE_REQUIRE(j == 1 || weight, "three neighbours need their weights");
  NESIE_REQUIRE(!(head && xyz) && (xyz == nullptr) == (center == nullptr), "head, or xyz + center, or neither");
  NESIE_REQUIRE((wx != nullptr) == (head != nullptr || xyz != nullptr), "wx goes with head or xyz + center");
  NESIE_REQUIRE(!xyz || (j == 1 && ns >= 1 && n % ns == 0), "xyz + center: one neighbour per row, ns | n");
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const float inv_radius = radius > 0.f ? 1.0f / radius : 0.f;
  if (!xyz) ns = 1;
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(table) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(col_parts) & 15) == 0, "table, y and col_parts must be 16-byte aligned");
  const dim3 grid(gl_grid_x(c, n, b), b);
  if (j == 3)
    gather_linear_fwd_kernel<3><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(c, m, n, table, idx, weight, head, wx, xyz, center, ns, inv_radius, y, col_parts);
  else
    gather_linear_fwd_kernel<1><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(c, m, n, table, idx, weight, head, wx, xyz, center, ns, inv_radius, y, col_parts);
  return check_launch("nesie_gather_linear_forward");
}

extern "C" int nesie_gather_linear_backward(int b, int c, int m, int n, int j, const float *d_y,
                                            const int *idx, const float *weight, const float *head,
                                            const float *xyz, const float *center, int ns, float radius,
                                            float *d_table, float *dwx_part, void *stream) {
  NESIE_REQUIRE(b >= 0 && m >= 1 && n >= 0 && (j == 1 || j == 3), "need b >= 0, m >= 1, n >= 0, j in {1, 3}");
  NESIE_REQUIRE(gl_shape_ok(c), "c / 4 must be a power of two <= 256");
  if (b == 0 || n == 0) return NESIE_OK;
  NESIE_REQUIRE(d_y && idx && d_table, "null pointer");
  NESIE_REQUIRE(j == 1 || weight, "three neighbours need their weights");
  NESIE_REQUIRE(!(head && xyz) && (xyz == nullptr) == (center == nullptr), "head, or xyz + center, or neither");
  NESIE_REQUIRE((dwx_part != nullptr) == (head != nullptr || xyz != nullptr), "dwx_part goes with head or xyz + center");
  NESIE_REQUIRE(!xyz || (j == 1 && ns >= 1 && n % ns == 0), "xyz + center: one neighbour per row, ns | n");
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const float inv_radius = radius > 0.f ? 1.0f / radius : 0.f;
  if (!xyz) ns = 1;
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(d_table) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_y) & 15) == 0,
                "d_table and d_y must be 16-byte aligned");
  const dim3 grid(gl_grid_x(c, n, b), b);
  if (j == 3)
    gather_linear_bwd_kernel<3><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(c, m, n, d_y, idx, weight, head, xyz, center, ns, inv_radius, d_table, dwx_part);
  else
    gather_linear_bwd_kernel<1><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(c, m, n, d_y, idx, weight, head, xyz, center, ns, inv_radius, d_table, dwx_part);
  return check_launch("nesie_gather_linear_backward");
}
