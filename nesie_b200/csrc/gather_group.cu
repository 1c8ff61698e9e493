// gather_points / grouping_operation (+ gradients) and the fused QueryAndGroup body, sm_100a.
//
// Replaces gather_points_kernel / gather_points_grad_kernel
//   (reference: ops/gather_points/src/gather_points_cuda.cu:8-26,51-70) and
// group_points_kernel / group_points_grad_kernel
//   (reference: ops/group_points/src/group_points_cuda.cu:56-79,10-31).
//
// The reference launches one thread per OUTPUT ELEMENT with the channel on gridDim.y, so every
// channel re-reads the index tensor and issues its own 4-byte scattered load.  These kernels are
// pure HBM traffic (SA2's grouped tensor is 137 MB), so here a thread owns FOUR consecutive
// output positions (one int4 index load, one float4 store per channel) and walks a slab of
// channels, keeping the indices in registers: index bytes are read once per slab instead of
// once per channel and every store is a full 16-byte coalesced write.  The source rows
// (features[b,c,:], <= 160 KB each) are L2-resident; only the output stream touches HBM.
#include "common.cuh"

namespace nesie {
namespace {

constexpr int GG_THREADS = 256;

// out[b,c,j] = src[b,c,idx[b,j]] for j in [0,L); L = npoints (gather) or npoints*nsample (group).
// grid: (ceil(L/4/256), channel slabs, b)
__global__ void __launch_bounds__(GG_THREADS) gather_rows_kernel(
    int c, int n, int L, int slab, const float *__restrict__ src, const int *__restrict__ idx,
    float *__restrict__ out) {
  const int b = blockIdx.z;
  const int cbeg = blockIdx.y * slab, cend = min(c, cbeg + slab);
  const int j0 = (blockIdx.x * GG_THREADS + threadIdx.x) * 4;
  if (j0 >= L) return;
  idx += (size_t)b * L;
  src += (size_t)b * c * n;
  out += (size_t)b * c * L;
  const bool vec = (j0 + 4 <= L) && ((L & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec) {
    const int4 id = __ldg(reinterpret_cast<const int4 *>(idx + j0));
#pragma unroll 4
    for (int ci = cbeg; ci < cend; ++ci) {
      const float *row = src + (size_t)ci * n;
      float4 v;
      v.x = __ldg(row + id.x); v.y = __ldg(row + id.y);
      v.z = __ldg(row + id.z); v.w = __ldg(row + id.w);
      __stcs(reinterpret_cast<float4 *>(out + (size_t)ci * L + j0), v);  // streaming store
    }
  } else {
    const int cntj = min(4, L - j0);
    int id[4];
    for (int t = 0; t < cntj; ++t) id[t] = idx[j0 + t];
    for (int ci = cbeg; ci < cend; ++ci) {
      const float *row = src + (size_t)ci * n;
      for (int t = 0; t < cntj; ++t) out[(size_t)ci * L + j0 + t] = __ldg(row + id[t]);
    }
  }
}

// grad_src[b,c,idx[b,j]] += grad_out[b,c,j].  Same thread mapping; fp32 reduction atomics
// (RED.ADD, no return value) like the reference's atomicAdd -- order is unspecified.
__global__ void __launch_bounds__(GG_THREADS) scatter_rows_kernel(
    int c, int n, int L, int slab, const float *__restrict__ grad_out,
    const int *__restrict__ idx, float *__restrict__ grad_src) {
  const int b = blockIdx.z;
  const int cbeg = blockIdx.y * slab, cend = min(c, cbeg + slab);
  const int j0 = (blockIdx.x * GG_THREADS + threadIdx.x) * 4;
  if (j0 >= L) return;
  idx += (size_t)b * L;
  grad_src += (size_t)b * c * n;
  grad_out += (size_t)b * c * L;
  const int cntj = min(4, L - j0);
  int id[4] = {0, 0, 0, 0};
  for (int t = 0; t < cntj; ++t) id[t] = idx[j0 + t];
  const bool vec = (cntj == 4) && ((L & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(grad_out) & 15) == 0);
  for (int ci = cbeg; ci < cend; ++ci) {
    float *row = grad_src + (size_t)ci * n;
    const float *g = grad_out + (size_t)ci * L + j0;
    if (vec) {
      const float4 v = __ldcs(reinterpret_cast<const float4 *>(g));
      atomicAdd(row + id[0], v.x); atomicAdd(row + id[1], v.y);
      atomicAdd(row + id[2], v.z); atomicAdd(row + id[3], v.w);
    } else {
      for (int t = 0; t < cntj; ++t) atomicAdd(row + id[t], g[t]);
    }
  }
}

// QueryAndGroup body: out[b, 0..2, j, k] = (xyz[b, idx, :] - center[b, j, :]) (/ radius)
//                     out[b, 3+ci, j, k] = features[b, ci, idx]
// The xyz part reads the point-major (b,n,3) array directly: no transposed copy of xyz, no
// separate subtract / divide / concat passes (group_points.py:98-116 launches 5 kernels and
// moves the grouped tensor 4 times).  Arithmetic is one fp32 subtract and one fp32 multiply by
// inv_radius = 1.0f / radius: torch's CUDA `grouped_xyz /= self.max_radius` with a python
// scalar divisor is executed as a multiplication by the fp32 reciprocal
// (ATen BinaryDivTrueKernel.cu, "a * reciprocal(b)"), and the reference only runs on CUDA.
__global__ void __launch_bounds__(GG_THREADS) query_group_concat_kernel(
    int c, int n, int npoints, int nsample, int slab, const float *__restrict__ xyz,
    const float *__restrict__ center, const float *__restrict__ feat,
    const int *__restrict__ idx, float inv_radius, float *__restrict__ out) {
  const int b = blockIdx.z;
  const int L = npoints * nsample;
  const int j0 = (blockIdx.x * GG_THREADS + threadIdx.x) * 4;
  if (j0 >= L) return;
  idx += (size_t)b * L;
  out += (size_t)b * (c + 3) * L;
  const int cntj = min(4, L - j0);
  int id[4] = {0, 0, 0, 0};
  for (int t = 0; t < cntj; ++t) id[t] = idx[j0 + t];
  const bool vec = (cntj == 4) && ((L & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (blockIdx.y == 0) {  // xyz channels
    const float *p = xyz + (size_t)b * n * 3;
    const float *ce = center + (size_t)b * npoints * 3;
    float v[3][4];
    for (int t = 0; t < cntj; ++t) {
      const int grp = (j0 + t) / nsample;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float d = __fsub_rn(__ldg(p + (size_t)id[t] * 3 + a), __ldg(ce + (size_t)grp * 3 + a));
        if (inv_radius > 0.f) d = __fmul_rn(d, inv_radius);
        v[a][t] = d;
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (vec) __stcs(reinterpret_cast<float4 *>(out + (size_t)a * L + j0),
                      make_float4(v[a][0], v[a][1], v[a][2], v[a][3]));
      else for (int t = 0; t < cntj; ++t) out[(size_t)a * L + j0 + t] = v[a][t];
    }
  }
  if (c > 0) {
    const int cbeg = blockIdx.y * slab, cend = min(c, cbeg + slab);
    const float *f = feat + (size_t)b * c * n;
    for (int ci = cbeg; ci < cend; ++ci) {
      const float *row = f + (size_t)ci * n;
      float *o = out + (size_t)(3 + ci) * L + j0;
      if (vec) {
        float4 v;
        v.x = __ldg(row + id[0]); v.y = __ldg(row + id[1]);
        v.z = __ldg(row + id[2]); v.w = __ldg(row + id[3]);
        __stcs(reinterpret_cast<float4 *>(o), v);
      } else {
        for (int t = 0; t < cntj; ++t) o[t] = __ldg(row + id[t]);
      }
    }
  }
}

// channel slab so that the grid has a few waves of CTAs but indices are reused >= 8x
int pick_slab(int c, long long ctas_per_slab_row) {
  int slab = 32;
  while (slab > 4 && ctas_per_slab_row * ceil_div(c, slab) < 4LL * num_sms()) slab >>= 1;
  return slab < 1 ? 1 : slab;
}

int launch_gather(int b, int c, int n, int L, const float *src, const int *idx, float *out,
                  cudaStream_t st, const char *what) {
  if (b == 0 || c == 0 || L == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const int gx = ceil_div(ceil_div(L, 4), GG_THREADS);
  const int slab = pick_slab(c, (long long)gx * b);
  dim3 grid(gx, ceil_div(c, slab), b);
  NESIE_REQUIRE(grid.y <= 65535, "too many channel slabs");
  gather_rows_kernel<<<grid, GG_THREADS, 0, st>>>(c, n, L, slab, src, idx, out);
  return check_launch(what);
}

int launch_scatter(int b, int c, int n, int L, const float *grad_out, const int *idx,
                   float *grad_src, cudaStream_t st, const char *what) {
  if (b == 0 || c == 0 || L == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const int gx = ceil_div(ceil_div(L, 4), GG_THREADS);
  const int slab = pick_slab(c, (long long)gx * b);
  dim3 grid(gx, ceil_div(c, slab), b);
  NESIE_REQUIRE(grid.y <= 65535, "too many channel slabs");
  scatter_rows_kernel<<<grid, GG_THREADS, 0, st>>>(c, n, L, slab, grad_out, idx, grad_src);
  return check_launch(what);
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_gather_points(int b, int c, int n, int npoints, const float *points,
                                   const int *idx, float *out, void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, "negative size");
  NESIE_REQUIRE(points && idx && out, "null pointer");
  return launch_gather(b, c, n, npoints, points, idx, out, (cudaStream_t)stream,
                       "nesie_gather_points");
}

extern "C" int nesie_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                                        const int *idx, float *grad_points, void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, "negative size");
  NESIE_REQUIRE(grad_out && idx && grad_points, "null pointer");
  return launch_scatter(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream,
                        "nesie_gather_points_grad");
}

extern "C" int nesie_group_points(int b, int c, int n, int npoints, int nsample,
                                  const float *points, const int *idx, float *out, void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "negative size");
  NESIE_REQUIRE(points && idx && out, "null pointer");
  NESIE_REQUIRE((long long)npoints * nsample < (1LL << 31), "npoints*nsample overflows int32");
  return launch_gather(b, c, n, npoints * nsample, points, idx, out, (cudaStream_t)stream,
                       "nesie_group_points");
}

extern "C" int nesie_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                       const float *grad_out, const int *idx, float *grad_points,
                                       void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "negative size");
  NESIE_REQUIRE(grad_out && idx && grad_points, "null pointer");
  NESIE_REQUIRE((long long)npoints * nsample < (1LL << 31), "npoints*nsample overflows int32");
  return launch_scatter(b, c, n, npoints * nsample, grad_out, idx, grad_points,
                        (cudaStream_t)stream, "nesie_group_points_grad");
}

extern "C" int nesie_query_group_concat(int b, int c, int n, int npoints, int nsample,
                                        const float *xyz, const float *center_xyz,
                                        const float *features, const int *idx, float radius,
                                        float *out, void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "negative size");
  NESIE_REQUIRE(xyz && center_xyz && idx && out, "null pointer");
  NESIE_REQUIRE(c == 0 || features, "features is NULL with c > 0");
  NESIE_REQUIRE((long long)npoints * nsample < (1LL << 31), "npoints*nsample overflows int32");
  if (b == 0 || npoints == 0 || nsample == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const int L = npoints * nsample;
  const int gx = ceil_div(ceil_div(L, 4), GG_THREADS);
  const int slab = c > 0 ? pick_slab(c, (long long)gx * b) : 1;
  dim3 grid(gx, c > 0 ? ceil_div(c, slab) : 1, b);
  query_group_concat_kernel<<<grid, GG_THREADS, 0, (cudaStream_t)stream>>>(
      c, n, npoints, nsample, slab, xyz, center_xyz, features, idx,
      radius > 0.f ? 1.0f / radius : 0.f, out);
  return check_launch("nesie_query_group_concat");
}

// ---------------------------------------------------------------------------------------------
// Row-major grouped tensor for the GEMM formulation of the shared MLP (training path):
//   rows[(b*npoints + j)*nsample + k, :] = [ (xyz[b, i] - center[b, j]) * inv_radius (3) | table[b, i, :] (c) ]
// with i = idx[b, j, k] and `table` the POINT-major (b, n, c) copy of the features, so a row is one
// contiguous read and one contiguous write.  The channel-major (B, C, M, K) grouped tensor of the
// reference (group_points.py:98-116) and its transposing copy into GEMM layout are never built.
// ---------------------------------------------------------------------------------------------
namespace nesie {
namespace {

__global__ void __launch_bounds__(256) group_rows_kernel(
    int c, int n, int npoints, int nsample, const float *__restrict__ xyz,
    const float *__restrict__ center, const float *__restrict__ table,
    const int *__restrict__ idx, float inv_radius, float *__restrict__ out, long long nrows_scene,
    int c0 /* row stride >= 3 + c; the padding columns are written as zeros */) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  idx += (size_t)b * nrows_scene;
  out += (size_t)b * nrows_scene * c0;
  for (long long row = warp_global; row < nrows_scene; row += nwarps) {
    const int pt = __ldg(idx + row);
    float *o = out + row * c0;
    if (lane < 3) {
      const int grp = (int)(row / nsample);
      float d = __fsub_rn(__ldg(xyz + ((size_t)b * n + pt) * 3 + lane),
                          __ldg(center + ((size_t)b * npoints + grp) * 3 + lane));
      if (inv_radius > 0.f) d = __fmul_rn(d, inv_radius);
      o[lane] = d;
    }
    const float *t = table + ((size_t)b * n + pt) * c;
    for (int j = lane; j < c; j += 32) o[3 + j] = __ldg(t + j);
    if (lane < c0 - 3 - c) o[3 + c + lane] = 0.f;
  }
}

// Rows of four floats (c <= 1: the first SA level, coordinates + height): one thread per row and one
// 16-byte store -- a warp per 16-byte row leaves 29 of its lanes idle over a million rows.
__global__ void __launch_bounds__(256) group_rows4_kernel(
    int c, int n, int npoints, int nsample, const float *__restrict__ xyz,
    const float *__restrict__ center, const float *__restrict__ table,
    const int *__restrict__ idx, float inv_radius, float4 *__restrict__ out, long long nrows_scene) {
  const int b = blockIdx.y;
  idx += (size_t)b * nrows_scene;
  out += (size_t)b * nrows_scene;
  for (long long row = (long long)blockIdx.x * 256 + threadIdx.x; row < nrows_scene;
       row += (long long)gridDim.x * 256) {
    const int pt = __ldg(idx + row);
    const float *p = xyz + ((size_t)b * n + pt) * 3;
    const float *q = center + ((size_t)b * npoints + (int)(row / nsample)) * 3;
    float4 v;
    v.x = __fsub_rn(__ldg(p), __ldg(q));
    v.y = __fsub_rn(__ldg(p + 1), __ldg(q + 1));
    v.z = __fsub_rn(__ldg(p + 2), __ldg(q + 2));
    if (inv_radius > 0.f) { v.x = __fmul_rn(v.x, inv_radius); v.y = __fmul_rn(v.y, inv_radius); v.z = __fmul_rn(v.z, inv_radius); }
    v.w = c ? __ldg(table + (size_t)b * n + pt) : 0.f;
    out[row] = v;
  }
}

// grad_rows (.., 3+c) -> grad_table (b, n, c) [+= by index], grad_xyz (b, n, 3), grad_center (b, m, 3)
__global__ void __launch_bounds__(256) group_rows_grad_kernel(
    int c, int n, int npoints, int nsample, const float *__restrict__ grad_rows,
    const int *__restrict__ idx, float inv_radius, float *__restrict__ grad_table,
    float *__restrict__ grad_xyz, float *__restrict__ grad_center, long long nrows_scene, int c0) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  idx += (size_t)b * nrows_scene;
  grad_rows += (size_t)b * nrows_scene * c0;
  for (long long row = warp_global; row < nrows_scene; row += nwarps) {
    const int pt = __ldg(idx + row);
    const float *g = grad_rows + row * c0;
    if (lane < 3 && (grad_xyz || grad_center)) {
      float v = g[lane];
      if (inv_radius > 0.f) v = __fmul_rn(v, inv_radius);
      if (grad_xyz) atomicAdd(grad_xyz + ((size_t)b * n + pt) * 3 + lane, v);
      if (grad_center) atomicAdd(grad_center + ((size_t)b * npoints + (int)(row / nsample)) * 3 + lane, -v);
    }
    if (grad_table) {
      float *t = grad_table + ((size_t)b * n + pt) * c;
      for (int j = lane; j < c; j += 32) atomicAdd(t + j, g[3 + j]);
    }
  }
}

}  // namespace
}  // namespace nesie

extern "C" int nesie_group_rows(int b, int c, int n, int npoints, int nsample, const float *xyz,
                                const float *center_xyz, const float *table_pm, const int *idx,
                                float radius, float *rows, int ld, void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "negative size");
  NESIE_REQUIRE(ld >= 3 + c && ld <= 3 + c + 32, "row stride must be in [3 + c, 3 + c + 32]");
  NESIE_REQUIRE(xyz && center_xyz && idx && rows && (c == 0 || table_pm), "null pointer");
  if (b == 0 || npoints == 0 || nsample == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const long long nr = (long long)npoints * nsample;
  if (ld == 4 && c <= 1 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0) {
    long long g4 = (nr + 1023) / 1024;   // four rows per thread
    if (g4 > 8LL * nesie::num_sms()) g4 = 8LL * nesie::num_sms();
    nesie::group_rows4_kernel<<<dim3((unsigned)(g4 < 1 ? 1 : g4), b), 256, 0, (cudaStream_t)stream>>>(
        c, n, npoints, nsample, xyz, center_xyz, table_pm, idx, radius > 0.f ? 1.0f / radius : 0.f,
        reinterpret_cast<float4 *>(rows), nr);
    return nesie::check_launch("nesie_group_rows");
  }
  int gx = (int)((nr + 7) / 8);
  if (gx > 8 * nesie::num_sms()) gx = 8 * nesie::num_sms();
  nesie::group_rows_kernel<<<dim3(gx, b), 256, 0, (cudaStream_t)stream>>>(
      c, n, npoints, nsample, xyz, center_xyz, table_pm, idx, radius > 0.f ? 1.0f / radius : 0.f,
      rows, nr, ld);
  return nesie::check_launch("nesie_group_rows");
}

extern "C" int nesie_group_rows_grad(int b, int c, int n, int npoints, int nsample,
                                     const float *grad_rows, const int *idx, float radius,
                                     float *grad_table_pm, float *grad_xyz, float *grad_center,
                                     int ld, void *stream) {
  NESIE_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "negative size");
  NESIE_REQUIRE(ld >= 3 + c, "row stride must be >= 3 + c");
  NESIE_REQUIRE(grad_rows && idx, "null pointer");
  if (b == 0 || npoints == 0 || nsample == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const long long nr = (long long)npoints * nsample;
  int gx = (int)((nr + 7) / 8);
  if (gx > 8 * nesie::num_sms()) gx = 8 * nesie::num_sms();
  nesie::group_rows_grad_kernel<<<dim3(gx, b), 256, 0, (cudaStream_t)stream>>>(
      c, n, npoints, nsample, grad_rows, idx, radius > 0.f ? 1.0f / radius : 0.f,
      c > 0 ? grad_table_pm : nullptr, grad_xyz, grad_center, nr, ld);
  return nesie::check_launch("nesie_group_rows_grad");
}
