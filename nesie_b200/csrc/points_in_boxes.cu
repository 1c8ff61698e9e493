// points_in_boxes / points_in_boxes_batch (target assignment of the train step).
//
// Replaces points_in_boxes_kernel / points_in_boxes_batch_kernel
// (reference: ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:49-105).  Same arithmetic, including
// the reference's float / double mix (`h / 2.0`, `rz + M_PI / 2` and all comparisons in double, the
// rotated coordinates in float with the contraction nvcc emits for them), so the masks are
// bit-identical to the reference kernel's for every yaw.
//
// The reference gives every thread one point and lets it walk the boxes in global memory, writing
// its (point, box) flags one int at a time: for the batch variant a warp's stores land 4 * nbox bytes
// apart.  Here the boxes of a scene are staged in shared memory with their yaw-dependent terms
// (h / 2, cos, sin) computed once per box instead of once per (point, box), each lane packs the flags
// of 32 boxes into a word, and the warp writes each point's row of 32 flags as one 128-byte store.
#include "common.cuh"

namespace nesie {
namespace {

constexpr int PIB_THREADS = 256;
constexpr int PIB_MAXBOX = 1024;   // boxes staged per pass (36 KB of shared memory)

struct BoxPre {       // per-box terms of check_pt_in_box3d that do not depend on the point
  float cx, cy, czc;  // centre (z shifted from the bottom face to the centre, rounded to float)
  float cosa, sina;
  double hh, hl, hw;  // half extents as the reference forms them (double)
};

__device__ __forceinline__ BoxPre box_pre(const float *b) {
  BoxPre p;
  const float w = b[3], l = b[4], h = b[5], rz = b[6];
  p.cx = b[0];
  p.cy = b[1];
  float cz = b[2];
  cz += h / 2.0;                        // double add, rounded back to float (as in the reference)
  p.czc = cz;
  const float rot_angle = rz + M_PI / 2;
  p.cosa = cos(rot_angle);
  p.sina = sin(rot_angle);
  p.hh = h / 2.0;
  p.hl = l / 2.0;
  p.hw = w / 2.0;
  return p;
}

__device__ __forceinline__ int in_box(float x, float y, float z, const BoxPre &p) {
  if (fabsf(z - p.czc) > p.hh) return 0;
  const float shift_x = x - p.cx, shift_y = y - p.cy;
  const float local_x = shift_x * p.cosa + shift_y * (-p.sina);
  const float local_y = shift_x * p.sina + shift_y * p.cosa;
  return (local_x > -p.hl) & (local_x < p.hl) & (local_y > -p.hw) & (local_y < p.hw);
}

template <bool BATCH>
__global__ void __launch_bounds__(PIB_THREADS) points_in_boxes_kernel(
    int nbox, int npts, const float *__restrict__ boxes, const float *__restrict__ pts,
    int *__restrict__ out) {
  extern __shared__ __align__(8) unsigned char pib_smem[];
  BoxPre *s_box = reinterpret_cast<BoxPre *>(pib_smem);
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * PIB_THREADS + threadIdx.x;
  const int pw = p - lane;                                 // first point of this warp
  boxes += (size_t)b * nbox * 7;
  float x = 0.f, y = 0.f, z = 0.f;
  if (p < npts) {
    const float *q = pts + ((size_t)b * npts + p) * 3;
    x = q[0]; y = q[1]; z = q[2];
  }
  int first = -1;
  for (int k0 = 0; k0 < nbox; k0 += PIB_MAXBOX) {
    const int kn = min(PIB_MAXBOX, nbox - k0);
    __syncthreads();
    for (int k = threadIdx.x; k < kn; k += PIB_THREADS) s_box[k] = box_pre(boxes + (size_t)(k0 + k) * 7);
    __syncthreads();
    for (int c0 = 0; c0 < kn; c0 += 32) {                  // 32 boxes -> one flag word per point
      unsigned mask = 0u;
      const int cn = min(32, kn - c0);
      if (p < npts)
        for (int j = 0; j < cn; ++j) mask |= (unsigned)in_box(x, y, z, s_box[c0 + j]) << j;
      if (BATCH) {
        // row of point pw + i: 32 consecutive ints, written by the whole warp at once
        for (int i = 0; i < 32; ++i) {
          const unsigned m = __shfl_sync(0xffffffffu, mask, i);
          if (pw + i < npts && lane < cn)
            out[((size_t)b * npts + pw + i) * nbox + k0 + c0 + lane] = (m >> lane) & 1u;
        }
      } else if (first < 0 && mask) {
        first = k0 + c0 + __ffs(mask) - 1;
      }
    }
  }
  if (!BATCH && p < npts && first >= 0) out[(size_t)b * npts + p] = first;
}

}  // namespace
}  // namespace nesie

using namespace nesie;

static int pib_launch(bool batch, int b, int nbox, int npts, const float *boxes, const float *pts,
                      int *out, void *stream) {
  NESIE_REQUIRE(b >= 0 && nbox >= 0 && npts >= 0, "negative size");
  if (b == 0 || npts == 0 || nbox == 0) return NESIE_OK;
  NESIE_REQUIRE(boxes && pts && out, "null pointer");
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  const int staged = nbox < PIB_MAXBOX ? nbox : PIB_MAXBOX;
  const size_t smem = (size_t)staged * sizeof(BoxPre);
  dim3 grid((npts + PIB_THREADS - 1) / PIB_THREADS, b);
  if (batch)
    points_in_boxes_kernel<true><<<grid, PIB_THREADS, smem, (cudaStream_t)stream>>>(nbox, npts, boxes, pts, out);
  else
    points_in_boxes_kernel<false><<<grid, PIB_THREADS, smem, (cudaStream_t)stream>>>(nbox, npts, boxes, pts, out);
  return check_launch(batch ? "nesie_points_in_boxes_batch" : "nesie_points_in_boxes");
}

extern "C" int nesie_points_in_boxes(int b, int nbox, int npts, const float *boxes, const float *pts,
                                     int *box_idx_of_points, void *stream) {
  return pib_launch(false, b, nbox, npts, boxes, pts, box_idx_of_points, stream);
}

extern "C" int nesie_points_in_boxes_batch(int b, int nbox, int npts, const float *boxes,
                                           const float *pts, int *box_idx_of_points, void *stream) {
  return pib_launch(true, b, nbox, npts, boxes, pts, box_idx_of_points, stream);
}
