// three_nn / three_interpolate (+ gradient) for the feature-propagation layers, sm_100a.
//
// Replaces three_nn_kernel (reference: ops/interpolate/src/three_nn_cuda.cu:11-65) and
// three_interpolate_kernel / three_interpolate_grad_kernel
// (reference: ops/interpolate/src/three_interpolate_cuda.cu:11-35,61-84).
//
// three_nn: the source set (<= a few thousand points) is staged in shared memory once per CTA
// with coalesced loads; every thread owns one target and walks the tile with broadcast reads.
// The strict '<' cascade keeps the earliest source on ties, as the reference does.  The
// reference holds its running bests in doubles initialised to 1e40; every compared value is an
// fp32 distance, so fp32 bests initialised to +inf select identically and the stored
// (float)1e40 is +inf.
//
// three_interpolate: the reference puts the channel on gridDim.y and re-reads idx/weight for
// every channel.  Here a thread owns one target point, loads its 3 indices + 3 weights once
// and walks a slab of channels; stores are coalesced along n.  The arithmetic keeps the
// reference's contraction  fma(w2,p2, fma(w0,p0, w1*p1))  (verified in its sm_100a PTX).
#include <math_constants.h>

#include "common.cuh"

namespace nesie {
namespace {

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;  // source points per smem tile (16 KB as float4)

// T targets per thread: one broadcast LDS.128 of a source serves T distance evaluations (with one
// target per thread the three 4-byte shared loads per pair, not the arithmetic, set the pace -- the
// SidePooling grids ask for 650 k targets x 1024 sources per step).  Targets of a thread are
// NN_THREADS apart so loads and stores stay coalesced.
template <int T>
__global__ void __launch_bounds__(NN_THREADS) three_nn_kernel(int n, int m,
                                                              const float *__restrict__ unknown,
                                                              const float *__restrict__ known,
                                                              float *__restrict__ dist2,
                                                              int *__restrict__ idx) {
  __shared__ float4 s_k[NN_TILE];
  const int b = blockIdx.y;
  const int pt0 = blockIdx.x * (NN_THREADS * T) + threadIdx.x;
  unknown += (size_t)b * n * 3;
  known += (size_t)b * m * 3;
  float ux[T], uy[T], uz[T], best1[T], best2[T], best3[T];
  int i1[T], i2[T], i3[T];
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const int pt = pt0 + j * NN_THREADS;
    const bool ok = pt < n;
    ux[j] = ok ? unknown[pt * 3 + 0] : 0.f;
    uy[j] = ok ? unknown[pt * 3 + 1] : 0.f;
    uz[j] = ok ? unknown[pt * 3 + 2] : 0.f;
    best1[j] = best2[j] = best3[j] = CUDART_INF_F;
    i1[j] = i2[j] = i3[j] = 0;
  }
  for (int t0 = 0; t0 < m; t0 += NN_TILE) {
    const int tn = min(NN_TILE, m - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += NN_THREADS) {
      const float *q = known + (size_t)(t0 + i) * 3;
      s_k[i] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f);
    }
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < tn; ++k) {
      const float4 s = s_k[k];
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const float d = sqdist_ref(ux[j], uy[j], uz[j], s.x, s.y, s.z);
        if (d < best3[j]) {            // strict '<': the earliest source keeps its place on ties
          if (d < best1[j]) {
            best3[j] = best2[j]; i3[j] = i2[j];
            best2[j] = best1[j]; i2[j] = i1[j];
            best1[j] = d; i1[j] = t0 + k;
          } else if (d < best2[j]) {
            best3[j] = best2[j]; i3[j] = i2[j];
            best2[j] = d; i2[j] = t0 + k;
          } else {
            best3[j] = d; i3[j] = t0 + k;
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const int pt = pt0 + j * NN_THREADS;
    if (pt < n) {
      float *od = dist2 + ((size_t)b * n + pt) * 3;
      int *oi = idx + ((size_t)b * n + pt) * 3;
      od[0] = best1[j]; od[1] = best2[j]; od[2] = best3[j];
      oi[0] = i1[j]; oi[1] = i2[j]; oi[2] = i3[j];
    }
  }
}

constexpr int TI_THREADS = 128;

// Inverse-distance interpolation straight into the row-major GEMM layout (the grid features of the
// SidePooling quality head, models/dense_heads/side_pooling_module.py:183-243): one warp per target
// row, lanes over channels of the POINT-major source table, so reads and writes are contiguous.
//   rows[t, 0:3]   = head[t, :]              (the grid point relative to its box centre)
//   rows[t, 3+j]   = fma(w2, T[i2, j], fma(w0, T[i0, j], w1 * T[i1, j]))
//   rows[t, 3+c:]  = 0                        (padding up to the row stride)
__global__ void __launch_bounds__(256) interp_rows_kernel(
    int c, int m, int n, int ld, const float *__restrict__ table, const int *__restrict__ idx,
    const float *__restrict__ weight, const float *__restrict__ head, float *__restrict__ rows) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long w0i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 8;
  table += (size_t)b * m * c;
  for (long long t = w0i; t < n; t += nw) {
    const size_t g = (size_t)b * n + t;
    const int a0 = __ldg(idx + g * 3), a1 = __ldg(idx + g * 3 + 1), a2 = __ldg(idx + g * 3 + 2);
    const float w0 = __ldg(weight + g * 3), w1 = __ldg(weight + g * 3 + 1), w2 = __ldg(weight + g * 3 + 2);
    float *o = rows + g * ld;
    if (lane < 3) o[lane] = __ldg(head + g * 3 + lane);
    const float *t0 = table + (size_t)a0 * c, *t1 = table + (size_t)a1 * c, *t2 = table + (size_t)a2 * c;
    for (int j = lane; j < c; j += 32)
      o[3 + j] = __fmaf_rn(w2, __ldg(t2 + j), __fmaf_rn(w0, __ldg(t0 + j), __fmul_rn(w1, __ldg(t1 + j))));
    if (lane < ld - 3 - c) o[3 + c + lane] = 0.f;
  }
}

// grid: (ceil(n/128), channel slabs, b)
__global__ void __launch_bounds__(TI_THREADS) three_interpolate_kernel(
    int c, int m, int n, int slab, const float *__restrict__ points,
    const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out) {
  const int b = blockIdx.z;
  const int pt = blockIdx.x * TI_THREADS + threadIdx.x;
  if (pt >= n) return;
  const int cbeg = blockIdx.y * slab, cend = min(c, cbeg + slab);
  const int *id = idx + ((size_t)b * n + pt) * 3;
  const float *w = weight + ((size_t)b * n + pt) * 3;
  const int a0 = id[0], a1 = id[1], a2 = id[2];
  const float w0 = w[0], w1 = w[1], w2 = w[2];
  points += (size_t)b * c * m;
  out += (size_t)b * c * n + pt;
#pragma unroll 4
  for (int ci = cbeg; ci < cend; ++ci) {
    const float *row = points + (size_t)ci * m;
    const float v = __fmaf_rn(w2, __ldg(row + a2),
                              __fmaf_rn(w0, __ldg(row + a0), __fmul_rn(w1, __ldg(row + a1))));
    out[(size_t)ci * n] = v;
  }
}

__global__ void __launch_bounds__(TI_THREADS) three_interpolate_grad_kernel(
    int c, int n, int m, int slab, const float *__restrict__ grad_out,
    const int *__restrict__ idx, const float *__restrict__ weight,
    float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const int pt = blockIdx.x * TI_THREADS + threadIdx.x;
  if (pt >= n) return;
  const int cbeg = blockIdx.y * slab, cend = min(c, cbeg + slab);
  const int *id = idx + ((size_t)b * n + pt) * 3;
  const float *w = weight + ((size_t)b * n + pt) * 3;
  const int a0 = id[0], a1 = id[1], a2 = id[2];
  const float w0 = w[0], w1 = w[1], w2 = w[2];
  grad_points += (size_t)b * c * m;
  grad_out += (size_t)b * c * n + pt;
  for (int ci = cbeg; ci < cend; ++ci) {
    float *row = grad_points + (size_t)ci * m;
    const float g = grad_out[(size_t)ci * n];
    atomicAdd(row + a0, __fmul_rn(g, w0));
    atomicAdd(row + a1, __fmul_rn(g, w1));
    atomicAdd(row + a2, __fmul_rn(g, w2));
  }
}

int pick_slab_ti(int c, long long ctas) {
  int slab = 32;
  while (slab > 2 && ctas * ceil_div(c, slab) < 4LL * num_sms()) slab >>= 1;
  return slab;
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_three_nn(int b, int n, int m, const float *unknown, const float *known,
                              float *dist2, int *idx, void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 0 && m >= 0, "negative size");
  NESIE_REQUIRE(unknown && known && dist2 && idx, "null pointer");
  if (b == 0 || n == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  // 4 targets per thread once that still leaves two CTAs per SM; one otherwise (the FP layers)
  if ((long long)ceil_div(n, NN_THREADS * 4) * b >= 2LL * num_sms()) {
    dim3 grid(ceil_div(n, NN_THREADS * 4), b);
    three_nn_kernel<4><<<grid, NN_THREADS, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
  } else {
    dim3 grid(ceil_div(n, NN_THREADS), b);
    three_nn_kernel<1><<<grid, NN_THREADS, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
  }
  return check_launch("nesie_three_nn");
}

extern "C" int nesie_three_interpolate(int b, int c, int m, int n, const float *points,
                                       const int *idx, const float *weight, float *out,
                                       void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0, "negative size");
  NESIE_REQUIRE(points && idx && weight && out, "null pointer");
  if (b == 0 || c == 0 || n == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const int gx = ceil_div(n, TI_THREADS);
  const int slab = pick_slab_ti(c, (long long)gx * b);
  dim3 grid(gx, ceil_div(c, slab), b);
  three_interpolate_kernel<<<grid, TI_THREADS, 0, (cudaStream_t)stream>>>(c, m, n, slab, points,
                                                                          idx, weight, out);
  return check_launch("nesie_three_interpolate");
}

extern "C" int nesie_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                            const int *idx, const float *weight,
                                            float *grad_points, void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0, "negative size");
  NESIE_REQUIRE(grad_out && idx && weight && grad_points, "null pointer");
  if (b == 0 || c == 0 || n == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const int gx = ceil_div(n, TI_THREADS);
  const int slab = pick_slab_ti(c, (long long)gx * b);
  dim3 grid(gx, ceil_div(c, slab), b);
  three_interpolate_grad_kernel<<<grid, TI_THREADS, 0, (cudaStream_t)stream>>>(
      c, n, m, slab, grad_out, idx, weight, grad_points);
  return check_launch("nesie_three_interpolate_grad");
}

extern "C" int nesie_interp_rows(int b, int c, int m, int n, const float *table_pm, const int *idx,
                                 const float *weight, const float *head, float *rows, int ld,
                                 void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 1 && m >= 1 && n >= 0, "need b >= 0, c >= 1, m >= 1, n >= 0");
  NESIE_REQUIRE(ld >= 3 + c && ld <= 3 + c + 32, "row stride must be in [3 + c, 3 + c + 32]");
  if (b == 0 || n == 0) return NESIE_OK;
  NESIE_REQUIRE(table_pm && idx && weight && head && rows, "null pointer");
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  int gx = (n + 7) / 8;
  if (gx > 8 * num_sms()) gx = 8 * num_sms();
  interp_rows_kernel<<<dim3(gx, b), 256, 0, (cudaStream_t)stream>>>(c, m, n, ld, table_pm, idx, weight,
                                                                   head, rows);
  return check_launch("nesie_interp_rows");
}
