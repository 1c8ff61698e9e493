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

// ---------------------------------------------------------------------------------------------
// Fused rotated IoU (cal_iou_3d, ops/rotated_iou/oriented_iou_loss.py:86-109 with cal_iou :38-58 and
// box_intersection_2d.py) forward + gradient with respect to the FIRST box, one thread per box pair.
// The reference evaluates ~135 small tensor ops forward and ~200 backward per call (3 calls per train
// step); here the candidate vertices (4 + 4 corners, 16 edge intersections), the vertex ordering
// (the sort_vertices scan above), the shoelace area, the 3-D IoU and the reverse-mode gradient all
// stay in registers.  Arithmetic follows the tensor formulation operator by operator (single-rounding
// intrinsics, no contraction), so values agree with it to the last bits; the gradient is the exact
// derivative of the same formulas (what autograd computes), including its conventions: no gradient
// through the masks / ordering, `clamp_min` passes the gradient where its input is >= 0, a tie of
// torch.min / torch.max splits it evenly.
namespace nesie {
namespace {

struct F2 { float x, y; };

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fd(float a, float b) { return __fdiv_rn(a, b); }

// corners of (x, y, w, h, alpha): lx = sx * w, ly = sy * h, (lx c - ly s + x, lx s + ly c + y)
__device__ __forceinline__ void rect_corners(float x, float y, float w, float h, float c, float s,
                                             F2 (&out)[4]) {
  const float sx[4] = {0.5f, -0.5f, -0.5f, 0.5f}, sy[4] = {0.5f, 0.5f, -0.5f, -0.5f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lx = fm(sx[i], w), ly = fm(sy[i], h);
    out[i].x = fa(fs(fm(lx, c), fm(ly, s)), x);
    out[i].y = fa(fa(fm(lx, s), fm(ly, c)), y);
  }
}

// corners of r1 inside (or on) rectangle r2 (box_intersection_2d.py:56-83)
__device__ __forceinline__ unsigned corners_inside(const F2 (&r1)[4], const F2 (&r2)[4]) {
  const float abx = fs(r2[1].x, r2[0].x), aby = fs(r2[1].y, r2[0].y);
  const float adx = fs(r2[3].x, r2[0].x), ady = fs(r2[3].y, r2[0].y);
  const float nab = fa(fm(abx, abx), fm(aby, aby)), nad = fa(fm(adx, adx), fm(ady, ady));
  unsigned m = 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float amx = fs(r1[i].x, r2[0].x), amy = fs(r1[i].y, r2[0].y);
    const float pab = fd(fa(fm(abx, amx), fm(aby, amy)), nab);
    const float pad = fd(fa(fm(adx, amx), fm(ady, amy)), nad);
    const bool in = pab > -1e-6f && pab < 1.f + 1e-6f && pad > -1e-6f && pad < 1.f + 1e-6f;
    m |= (in ? 1u : 0u) << i;
  }
  return m;
}

__global__ void __launch_bounds__(64) iou3d_kernel(long long n, const float *__restrict__ box1,
                                                   const float *__restrict__ box2,
                                                   float *__restrict__ iou_out,
                                                   float *__restrict__ jac_out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float *a = box1 + r * 7, *b = box2 + r * 7;
  const float x1 = a[0], y1 = a[1], z1 = a[2], w1 = a[3], h1 = a[4], l1 = a[5], al1 = a[6];
  const float x2 = b[0], y2 = b[1], z2 = b[2], w2 = b[3], h2 = b[4], l2 = b[5], al2 = b[6];
  const float c1 = cosf(al1), s1 = sinf(al1), c2 = cosf(al2), s2 = sinf(al2);
  F2 r1[4], r2[4];
  rect_corners(x1, y1, w1, h1, c1, s1, r1);
  rect_corners(x2, y2, w2, h2, c2, s2, r2);

  // ---- 24 candidate vertices -------------------------------------------------------------------
  SvVertex v[SV_M];
  float tt[16];                         // t of every edge pair (for the gradient)
  unsigned valid = corners_inside(r1, r2) | (corners_inside(r2, r1) << 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i].x = r1[i].x; v[i].y = r1[i].y; v[4 + i].x = r2[i].x; v[4 + i].y = r2[i].y; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const F2 A = r1[i], B = r1[(i + 1) & 3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const F2 C = r2[j], D = r2[(j + 1) & 3];
      const float ya = fs(C.y, D.y), xb = fs(C.x, D.x);
      const float num = fs(fm(fs(A.x, B.x), ya), fm(fs(A.y, B.y), xb));
      const float den_t = fs(fm(fs(A.x, C.x), ya), fm(fs(A.y, C.y), xb));
      const float den_u = fs(fm(fs(A.x, B.x), fs(A.y, C.y)), fm(fs(A.y, B.y), fs(A.x, C.x)));
      const bool par = num == 0.f;
      const float t0 = par ? -1.f : fd(den_t, num);
      const float u0 = par ? -1.f : fd(-den_u, num);
      const bool hit = t0 > 0.f && t0 < 1.f && u0 > 0.f && u0 < 1.f;
      const float t = fd(den_t, fa(num, 1e-8f));
      const int k = 8 + i * 4 + j;
      tt[i * 4 + j] = t;
      v[k].x = hit ? fa(A.x, fm(t, fs(B.x, A.x))) : 0.f;
      v[k].y = hit ? fa(A.y, fm(t, fs(B.y, A.y))) : 0.f;
      valid |= (hit ? 1u : 0u) << k;
    }
  }
  // ---- ordering: mean-normalised copies through the sort_vertices scan ---------------------------
  const int nv = __popc(valid);
  float mx = 0.f, my = 0.f;
#pragma unroll
  for (int k = 0; k < SV_M; ++k) {      // torch.sum over the 24 slots in order, masked
    mx = fa(mx, ((valid >> k) & 1u) ? v[k].x : fm(v[k].x, 0.f));
    my = fa(my, ((valid >> k) & 1u) ? v[k].y : fm(v[k].y, 0.f));
  }
  mx = fd(mx, (float)nv);
  my = fd(my, (float)nv);
  int pad = 0;
  {
    const unsigned inv = ~valid & 0x00ffff00u;
    if (inv) pad = __ffs(inv) - 1;
  }
  int order[SV_MAX_IDX];
#pragma unroll
  for (int j = 0; j < SV_MAX_IDX; ++j) order[j] = pad;
  if (nv >= 3) {
    SvVertex nrm[SV_M];
#pragma unroll
    for (int k = 0; k < SV_M; ++k) {
      nrm[k].x = fs(v[k].x, mx); nrm[k].y = fs(v[k].y, my); nrm[k].q = sv_q(nrm[k].x, nrm[k].y);
    }
    SvVertex first = {1.f, (float)(-1e-8), 0.f};
    first.q = sv_q(first.x, first.y);
    SvVertex prev = first;
#pragma unroll 1
    for (int j = 0; j < nv && j < SV_MAX_IDX - 1; ++j) {
      SvVertex best = first;
      int take = 0;
#pragma unroll
      for (int k = 0; k < SV_M; ++k)
        if (((valid >> k) & 1u) && sv_less(nrm[k], best) && (j == 0 || sv_less(prev, nrm[k]))) {
          best = nrm[k]; take = k;
        }
      prev = nrm[0];
#pragma unroll
      for (int k = 1; k < SV_M; ++k)
        if (k == take) prev = nrm[k];
#pragma unroll
      for (int jj = 0; jj < SV_MAX_IDX - 1; ++jj)
        if (jj == j) order[jj] = take;
    }
#pragma unroll
    for (int jj = 3; jj < SV_MAX_IDX; ++jj)
      if (jj == nv) order[jj] = order[0];
    if (nv == 8) {
      int counter = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 4; k < SV_INTER_OFF; ++k) counter += order[k] == order[q];
      if (counter == 4) {
        order[4] = order[0];
#pragma unroll
        for (int jj = 5; jj < SV_MAX_IDX; ++jj) order[jj] = pad;
      }
    }
  }
  // ---- shoelace area over the 9 selected vertices + its gradient per selected slot ---------------
  float px[SV_MAX_IDX], py[SV_MAX_IDX];
#pragma unroll
  for (int j = 0; j < SV_MAX_IDX; ++j) {
    px[j] = v[0].x; py[j] = v[0].y;
#pragma unroll
    for (int k = 1; k < SV_M; ++k)
      if (k == order[j]) { px[j] = v[k].x; py[j] = v[k].y; }
  }
  float total = 0.f;
#pragma unroll
  for (int j = 0; j + 1 < SV_MAX_IDX; ++j)
    total = fa(total, fs(fm(px[j], py[j + 1]), fm(py[j], px[j + 1])));
  const float inter = fd(fabsf(total), 2.f);
  const float sgn = total > 0.f ? 1.f : (total < 0.f ? -1.f : 0.f);
  // 3-D part
  const float zmax1 = fa(z1, fm(l1, 0.5f)), zmin1 = fs(z1, fm(l1, 0.5f));
  const float zmax2 = fa(z2, fm(l2, 0.5f)), zmin2 = fs(z2, fm(l2, 0.5f));
  const float ztop = fminf(zmax1, zmax2), zbot = fmaxf(zmin1, zmin2);
  const float zraw = fs(ztop, zbot);
  const float zov = fmaxf(zraw, 0.f);
  const float area1 = fm(w1, h1), area2 = fm(w2, h2);
  const float u = fs(fa(area1, area2), inter);
  const float iou2d = fd(inter, u);
  const float i3 = fm(fm(iou2d, u), zov);
  const float vol1 = fm(fm(w1, h1), l1), vol2 = fm(fm(w2, h2), l2);
  const float u3 = fs(fa(vol1, vol2), i3);
  const float iou = fd(i3, u3);
  iou_out[r] = iou;
  if (!jac_out) return;

  // ---- reverse mode: d iou / d (x1, y1, z1, w1, h1, l1, al1) ---------------------------------------
  // iou = i3 / u3, u3 = vol1 + vol2 - i3
  const float g_i3 = 1.f / u3 + i3 / (u3 * u3);            // d iou / d i3 (direct + through u3)
  const float g_vol1 = -i3 / (u3 * u3);
  // i3 = (iou2d * u) * zov
  const float g_zov = g_i3 * (iou2d * u);
  const float g_iou2d = g_i3 * zov * u;
  float g_u = g_i3 * zov * iou2d;
  // iou2d = inter / u ; u = area1 + area2 - inter
  float g_inter = g_iou2d / u;
  g_u += -g_iou2d * inter / (u * u);
  g_inter += -g_u;
  const float g_area1 = g_u;
  float gw = g_area1 * h1 + g_vol1 * (h1 * l1);
  float gh = g_area1 * w1 + g_vol1 * (w1 * l1);
  float gl = g_vol1 * (w1 * h1);
  // z overlap
  float gz = 0.f;
  {
    const float g_zraw = zraw >= 0.f ? g_zov : 0.f;
    const float top1 = zmax1 < zmax2 ? 1.f : (zmax1 == zmax2 ? 0.5f : 0.f);   // d min / d zmax1
    const float bot1 = zmin1 > zmin2 ? 1.f : (zmin1 == zmin2 ? 0.5f : 0.f);   // d max / d zmin1
    const float g_zmax1 = g_zraw * top1, g_zmin1 = -g_zraw * bot1;
    gz = g_zmax1 + g_zmin1;
    gl += 0.5f * g_zmax1 - 0.5f * g_zmin1;
  }
  // area = |total| / 2 -> gradient of every selected slot
  const float g_total = g_inter * 0.5f * sgn;
  float gcx[4] = {0.f, 0.f, 0.f, 0.f}, gcy[4] = {0.f, 0.f, 0.f, 0.f};       // wrt the corners of box 1
#pragma unroll
  for (int j = 0; j < SV_MAX_IDX; ++j) {
    // total = sum_j px[j] py[j+1] - py[j] px[j+1]
    float gx = 0.f, gy = 0.f;
    if (j + 1 < SV_MAX_IDX) { gx += py[j + 1]; gy -= px[j + 1]; }
    if (j > 0) { gx -= py[j - 1]; gy += px[j - 1]; }
    gx *= g_total; gy *= g_total;
    const int k = order[j];
    if (!((valid >> k) & 1u)) continue;                   // padded slot: value 0, gradient 0
    if (k < 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q == k) { gcx[q] += gx; gcy[q] += gy; }
    } else if (k >= 8) {
      // P = A + t (B - A), t = den_t / (num + eps); A = r1[i], B = r1[i+1], C / D of box 2 constant
      const int e = k - 8, i = e >> 2, jj2 = e & 3;
      F2 A = r1[0], B = r1[1], C = r2[0], D = r2[1];
      float t = tt[0];
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (q == e) t = tt[q];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q == i) { A = r1[q]; B = r1[(q + 1) & 3]; }
        if (q == jj2) { C = r2[q]; D = r2[(q + 1) & 3]; }
      }
      const float ya = C.y - D.y, xb = C.x - D.x;
      const float num = (A.x - B.x) * ya - (A.y - B.y) * xb + 1e-8f;
      const float den_t = (A.x - C.x) * ya - (A.y - C.y) * xb;
      // g_t = gx * (B.x - A.x) + gy * (B.y - A.y)
      const float g_t = gx * (B.x - A.x) + gy * (B.y - A.y);
      const float g_den = g_t / num, g_num = -g_t * den_t / (num * num);
      // direct terms: P.x = A.x (1 - t) + t B.x
      float gAx = gx * (1.f - t), gAy = gy * (1.f - t), gBx = gx * t, gBy = gy * t;
      // num = (A.x - B.x) ya - (A.y - B.y) xb ; den_t = (A.x - C.x) ya - (A.y - C.y) xb
      gAx += g_num * ya + g_den * ya;
      gAy += -g_num * xb - g_den * xb;
      gBx += -g_num * ya;
      gBy += g_num * xb;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q == i) { gcx[q] += gAx; gcy[q] += gAy; }
        if (q == ((i + 1) & 3)) { gcx[q] += gBx; gcy[q] += gBy; }
      }
    }
  }
  // corners of box 1 -> (x, y, w, h, alpha)
  float gx1 = 0.f, gy1 = 0.f, gal = 0.f;
  {
    const float sx[4] = {0.5f, -0.5f, -0.5f, 0.5f}, sy[4] = {0.5f, 0.5f, -0.5f, -0.5f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float lx = sx[q] * w1, ly = sy[q] * h1;
      gx1 += gcx[q];
      gy1 += gcy[q];
      gw += gcx[q] * (sx[q] * c1) + gcy[q] * (sx[q] * s1);
      gh += gcx[q] * (-sy[q] * s1) + gcy[q] * (sy[q] * c1);
      gal += gcx[q] * (-lx * s1 - ly * c1) + gcy[q] * (lx * c1 - ly * s1);
    }
  }
  float *jo = jac_out + r * 7;
  jo[0] = gx1; jo[1] = gy1; jo[2] = gz; jo[3] = gw; jo[4] = gh; jo[5] = gl; jo[6] = gal;
}

}  // namespace
}  // namespace nesie

extern "C" int nesie_iou3d(long long n, const float *box1, const float *box2, float *iou,
                           float *jac_box1, void *stream) {
  NESIE_REQUIRE(n >= 0, "negative size");
  if (n == 0) return NESIE_OK;
  NESIE_REQUIRE(box1 && box2 && iou, "null pointer");
  nesie::iou3d_kernel<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(n, box1, box2, iou, jac_box1);
  return nesie::check_launch("nesie_iou3d");
}
