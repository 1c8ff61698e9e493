// Target assignment of the train step, three small kernels that replace python loops
// and tensor expands of the reference head:
//
//   vote_targets_kernel     NesieHead.get_targets_single, vote part
//                           (models/dense_heads/nesie_head.py:618-654): DepthInstance3DBoxes
//                           .points_in_boxes (core/bbox/structures/depth_box3d.py:251-277, i.e. the
//                           depth -> LiDAR flip + points_in_boxes_batch, ops/roiaware_pool3d/src/
//                           points_in_boxes_cuda.cu:24-105) followed by the per-box python loop that
//                           fills gt_per_seed = 3 vote slots per point.
//   chamfer_assign_kernel   the two argmins of chamfer_distance (models/losses/chamfer_distance.py:
//                           49-56) without the (B, N, M, 3) expands.
//   sort_vertices_kernel    ops/rotated_iou/cuda_op/sort_vert_kernel.cu:16-134 (polygon vertex order
//                           for the rotated IoU of cal_iou_3d).
//
// All index / mask outputs are bit-identical to the reference formulation: the inclusion test is the
// reference kernel's float / double mix, distances are formed with single-rounding intrinsics in
// torch's operator order, the vertex comparator keeps the reference's double-precision epsilon tests.
#include "common.cuh"

namespace nesie {
namespace {

constexpr int VT_THREADS = 128;
constexpr int VT_MAXBOX = 256;

struct VBox {         // LiDAR-frame box terms that do not depend on the point + the vote centre
  float cx, cy, czc;  // centre in the LiDAR frame (z shifted to the box centre, rounded to float)
  float cosa, sina;
  double hh, hl, hw;
  float gx, gy, gz;   // gravity centre in the depth frame (what the votes point at)
};

__device__ __forceinline__ VBox vbox_pre(const float *d) {
  // depth (x, y, z_bottom, dx, dy, dz, yaw) -> LiDAR: xyz @ [[0,1,0],[-1,0,0],[0,0,1]]^T, sizes
  // (dy, dx, dz), yaw kept (core/bbox/structures/box_3d_mode.py:124-160)
  VBox p;
  const float w = d[4], l = d[3], h = d[5], rz = d[6];
  p.cx = d[1];
  p.cy = -d[0];
  float cz = d[2];
  cz += h / 2.0;
  p.czc = cz;
  const float rot_angle = rz + M_PI / 2;
  p.cosa = cos(rot_angle);
  p.sina = sin(rot_angle);
  p.hh = h / 2.0;
  p.hl = l / 2.0;
  p.hw = w / 2.0;
  p.gx = d[0];
  p.gy = d[1];
  p.gz = __fadd_rn(d[2], __fmul_rn(d[5], 0.5f));   // depth_box3d.py:42-48
  return p;
}

__device__ __forceinline__ int vbox_contains(float x, float y, float z, const VBox &p) {
  if (fabsf(z - p.czc) > p.hh) return 0;
  const float shift_x = x - p.cx, shift_y = y - p.cy;
  const float local_x = shift_x * p.cosa + shift_y * (-p.sina);
  const float local_y = shift_x * p.sina + shift_y * p.cosa;
  return (local_x > -p.hl) & (local_x < p.hl) & (local_y > -p.hw) & (local_y < p.hw);
}

// One thread per (scene, output row).  rows = s when `idx` selects s points of the scene, else n.
__global__ void __launch_bounds__(VT_THREADS) vote_targets_kernel(
    int n, int g, int rows, const float *__restrict__ pts, int pts_stride,
    const float *__restrict__ boxes, const int *__restrict__ nvalid,
    const long long *__restrict__ idx, float *__restrict__ vote_targets,
    long long *__restrict__ vote_mask) {
  __shared__ VBox s_box[VT_MAXBOX];
  const int b = blockIdx.y;
  const int nv = min(nvalid ? nvalid[b] : g, g);
  for (int k = threadIdx.x; k < nv; k += VT_THREADS) s_box[k] = vbox_pre(boxes + ((size_t)b * g + k) * 7);
  __syncthreads();
  const int r = blockIdx.x * VT_THREADS + threadIdx.x;
  if (r >= rows) return;
  long long p = idx ? idx[(size_t)b * rows + r] : r;
  const float *q = pts + ((size_t)b * n + p) * pts_stride;
  const float x = q[0], y = q[1], z = q[2];
  const float lx = y, ly = -x;          // points_lidar = points[..., [1, 0, 2]]; [..., 1] *= -1
  float v[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) v[j] = 0.f;
  int cnt = 0;
  for (int k = 0; k < nv; ++k) {
    if (!vbox_contains(lx, ly, z, s_box[k])) continue;
    const float vx = __fsub_rn(s_box[k].gx, x), vy = __fsub_rn(s_box[k].gy, y),
                vz = __fsub_rn(s_box[k].gz, z);
    if (cnt == 0) {                     // j == 0: the vote fills all three slots (:645-647)
      v[0] = v[3] = v[6] = vx; v[1] = v[4] = v[7] = vy; v[2] = v[5] = v[8] = vz;
    } else if (cnt == 1) {
      v[3] = vx; v[4] = vy; v[5] = vz;
    } else {                            // the counter saturates at 2: later boxes overwrite slot 2
      v[6] = vx; v[7] = vy; v[8] = vz;
    }
    cnt = min(cnt + 1, 2);
  }
  float *o = vote_targets + ((size_t)b * rows + r) * 9;
#pragma unroll
  for (int j = 0; j < 9; ++j) o[j] = v[j];
  vote_mask[(size_t)b * rows + r] = cnt > 0;
}

// ---- chamfer argmins -------------------------------------------------------------------------
// One CTA per scene; dst staged in shared memory.  d = ((dx^2 + dy^2) + dz^2) with every operation
// rounded once (mse_loss(...).sum(-1) of the reference), first minimum wins (torch.min).
constexpr int CH_THREADS = 256;
constexpr int CH_MAXDST = 1024;

__device__ __forceinline__ float sq3(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__global__ void __launch_bounds__(CH_THREADS) chamfer_assign_kernel(
    int n, int m, const float *__restrict__ src, const float *__restrict__ dst,
    const int *__restrict__ nvalid, long long *__restrict__ idx1, long long *__restrict__ idx2) {
  __shared__ float s_dst[CH_MAXDST * 3];
  __shared__ unsigned long long s_best[CH_MAXDST];   // (distance bits << 32 | source index) per dst slot
  const int b = blockIdx.x;
  src += (size_t)b * n * 3;
  dst += (size_t)b * m * 3;
  const int mv = min(nvalid ? nvalid[b] : m, m);
  for (int i = threadIdx.x; i < m * 3; i += CH_THREADS) s_dst[i] = dst[i];
  for (int j = threadIdx.x; j < m; j += CH_THREADS) s_best[j] = ~0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += CH_THREADS) {
    const float x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
    float best = 0.f;
    int arg = 0;
    for (int j = 0; j < m; ++j) {
      const float d = sq3(x, y, z, s_dst[3 * j], s_dst[3 * j + 1], s_dst[3 * j + 2]);
      if (j < mv && (j == 0 || d < best)) { best = d; arg = j; }
      // non-negative floats order like their bit patterns; NaN (0x7fc00000) sorts last
      atomicMin(&s_best[j], ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i);
    }
    if (idx1) idx1[(size_t)b * n + i] = arg;
  }
  __syncthreads();
  if (idx2)
    for (int j = threadIdx.x; j < m; j += CH_THREADS)
      idx2[(size_t)b * m + j] = (long long)(s_best[j] & 0xffffffffull);
}

// ---- sort_vertices ---------------------------------------------------------------------------
constexpr int SV_MAX_IDX = 9, SV_INTER_OFF = 8, SV_M = 24;

// The reference comparator (compare_vertices, sort_vert_kernel.cu:16-40) orders two vertices by
// quadrant and by q = |x| x / n with n = x^2 + y^2 + EPSILON; EPSILON is the double literal 1e-8, so
// its epsilon tests are double comparisons of float values.  q depends on one vertex only: it is
// formed once per vertex (same operations: fma(x, x, y*y) in float, + 1e-8 in double, rounded to
// float, float divide) instead of twice per comparison, and the double comparisons are replaced by
// float comparisons against 1e-8 rounded up / down, which decide identically for every float.
struct SvVertex { float x, y, q; };

__device__ __forceinline__ float sv_q(float x, float y) {
  const float n = (float)((double)__fmaf_rn(x, x, __fmul_rn(y, y)) + 1e-8);
  return __fdiv_rn(__fmul_rn(fabsf(x), x), n);
}

__device__ __forceinline__ bool sv_less(const SvVertex &a, const SvVertex &b) {
  const float eps_up = __double2float_ru(1e-8), eps_dn = __double2float_rd(1e-8);
  if (fabsf(__fsub_rn(a.x, b.x)) < eps_up && fabsf(__fsub_rn(b.y, a.y)) < eps_up) return false;
  if (a.y > 0 && b.y < 0) return true;
  if (a.y < 0 && b.y > 0) return false;
  const float d = __fsub_rn(a.q, b.q);
  if (a.y > 0 && b.y > 0) return d > eps_dn;   // (double)d > 1e-8
  if (a.y < 0 && b.y < 0) return d < eps_up;   // (double)d < 1e-8
  return false;
}

// One thread per polygon, the 24 candidates (x, y, q) in registers (every loop over them is fully
// unrolled), validity as a bit mask; the selection keeps the reference's single ascending scan per
// round, because its comparator is not transitive across the epsilon tests.
__global__ void __launch_bounds__(64) sort_vertices_kernel(
    long long polys, const float *__restrict__ vertices, const unsigned char *__restrict__ mask,
    const int *__restrict__ num_valid, int *__restrict__ idx) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= polys) return;
  SvVertex v[SV_M];
  unsigned valid = 0u;
  const float2 *vp = reinterpret_cast<const float2 *>(vertices) + p * SV_M;
#pragma unroll
  for (int k = 0; k < SV_M; ++k) {
    const float2 c = vp[k];
    v[k].x = c.x; v[k].y = c.y; v[k].q = sv_q(c.x, c.y);
    valid |= (mask[p * SV_M + k] ? 1u : 0u) << k;
  }
  const int nv = num_valid[p];
  int pad = 0;
  {
    const unsigned inv = ~valid & 0x00ffff00u;
    if (inv) pad = __ffs(inv) - 1;
  }
  int out[SV_MAX_IDX];
#pragma unroll
  for (int j = 0; j < SV_MAX_IDX; ++j) out[j] = pad;
  if (nv >= 3) {
    SvVertex first = {1.f, (float)(-1e-8), 0.f};       // the scan's initial "minimum" (:84-85)
    first.q = sv_q(first.x, first.y);
    SvVertex prev = first;
#pragma unroll 1
    for (int j = 0; j < nv && j < SV_MAX_IDX - 1; ++j) {
      SvVertex best = first;
      int take = 0;
#pragma unroll
      for (int k = 0; k < SV_M; ++k) {
        if (((valid >> k) & 1u) && sv_less(v[k], best) && (j == 0 || sv_less(prev, v[k]))) {
          best = v[k];
          take = k;
        }
      }
      // the next round compares against vertices[idx[j]] -- vertex 0 when nothing was selected
      // (the reference's i_take stays 0, :86)
      prev = v[0];
#pragma unroll
      for (int k = 1; k < SV_M; ++k)
        if (k == take) prev = v[k];
#pragma unroll
      for (int jj = 0; jj < SV_MAX_IDX - 1; ++jj)
        if (jj == j) out[jj] = take;
    }
#pragma unroll
    for (int jj = 3; jj < SV_MAX_IDX; ++jj)
      if (jj == nv) out[jj] = out[0];
    if (nv == 8) {                      // two identical boxes: every corner appears twice
      int counter = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 4; k < SV_INTER_OFF; ++k) counter += out[k] == out[a];
      if (counter == 4) {
        out[4] = out[0];
#pragma unroll
        for (int jj = 5; jj < SV_MAX_IDX; ++jj) out[jj] = pad;
      }
    }
  }
  int *o = idx + p * SV_MAX_IDX;
#pragma unroll
  for (int j = 0; j < SV_MAX_IDX; ++j) o[j] = out[j];
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_vote_targets(int b, int n, int g, int rows, const float *pts, int pts_stride,
                                  const float *boxes, const int *nvalid, const long long *idx,
                                  float *vote_targets, long long *vote_mask, void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 0 && g >= 0 && rows >= 0 && pts_stride >= 3, "bad size");
  if (b == 0 || rows == 0) return NESIE_OK;
  NESIE_REQUIRE(pts && vote_targets && vote_mask && (g == 0 || boxes), "null pointer");
  NESIE_REQUIRE(g <= VT_MAXBOX, "more than 256 boxes per scene");
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  NESIE_REQUIRE(idx || rows == n, "rows must equal n without an index list");
  dim3 grid(ceil_div(rows, VT_THREADS), b);
  vote_targets_kernel<<<grid, VT_THREADS, 0, (cudaStream_t)stream>>>(
      n, g, rows, pts, pts_stride, boxes, nvalid, idx, vote_targets, vote_mask);
  return check_launch("nesie_vote_targets");
}

extern "C" int nesie_chamfer_assign(int b, int n, int m, const float *src, const float *dst,
                                    const int *nvalid, long long *idx1, long long *idx2,
                                    void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 0 && m >= 1, "bad size");
  if (b == 0) return NESIE_OK;
  NESIE_REQUIRE(src && dst, "null pointer");
  NESIE_REQUIRE(m <= CH_MAXDST, "more than 1024 destination points");
  chamfer_assign_kernel<<<b, CH_THREADS, 0, (cudaStream_t)stream>>>(n, m, src, dst, nvalid, idx1, idx2);
  return check_launch("nesie_chamfer_assign");
}

extern "C" int nesie_sort_vertices(int b, int n, int m, const float *vertices,
                                   const unsigned char *mask, const int *num_valid, int *idx,
                                   void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 0, "bad size");
  NESIE_REQUIRE(m == SV_M, "m must be 24 (4 + 4 corners, 16 edge intersections)");
  const long long polys = (long long)b * n;
  if (polys == 0) return NESIE_OK;
  NESIE_REQUIRE(vertices && mask && num_valid && idx, "null pointer");
  sort_vertices_kernel<<<(unsigned)((polys + 63) / 64), 64, 0, (cudaStream_t)stream>>>(
      polys, vertices, mask, num_valid, idx);
  return check_launch("nesie_sort_vertices");
}
