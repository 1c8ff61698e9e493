// Batched class-aware axis-aligned 3D NMS (test-time aligned_3d_nms and the train-time
// lenient pseudo-label NMS), one scene per CTA, everything in shared memory.
//
// Replaces
//   aligned_3d_nms            reference: core/post_processing/box3d_nms.py:129-176
//                             (a python while-loop of ~20 tiny torch launches and one host sync
//                              per picked box, called per scene from nesie_head.py:715-724)
//   lhs_3d_faster_samecls     reference: models/detectors/votenet_nesie.py:733-779
//                             (numpy on the host after >= 7 device->host copies per step)
//
// Parity rules kept bit-for-bit:
//   * every arithmetic step is ONE rounding, in the order the reference evaluates it (torch /
//     numpy launch one elementwise op per python operator, so nothing is fused): the intrinsics
//     __fmul_rn/__fadd_rn/... (fp32) and __dmul_rn/... (fp64) forbid FMA contraction;
//   * max(0, x) and max/min of coordinates propagate NaN like torch.max / np.maximum;
//   * a box survives iff `iou <= thresh` (aligned) resp. is suppressed iff `o > thresh` (lhs),
//     so a NaN IoU (0/0 for two zero-volume boxes) drops the box in the aligned variant and
//     keeps it in the lhs variant, exactly as the reference comparisons do;
//   * the class mask multiplies the IoU (NaN * 0 stays NaN).
// Sorting: the reference uses an unstable argsort, so its result is only defined for distinct
// scores; here equal scores are ordered by ascending index (a stable ascending sort).
#include "common.cuh"

namespace nesie {
namespace {

constexpr int NMS_THREADS = 256;
constexpr int NMS_MAX_N = 1024;

__device__ __forceinline__ float nanmaxf(float a, float b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
__device__ __forceinline__ float nanminf(float a, float b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}
__device__ __forceinline__ double nanmaxd(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
__device__ __forceinline__ double nanmind(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}

// rank of element i in a stable ascending sort of key[0..n)
template <typename T>
__device__ __forceinline__ int stable_rank(const T *key, int n, int i) {
  const T ki = key[i];
  int r = 0;
  for (int j = 0; j < n; ++j) {
    const T kj = key[j];
    r += (kj < ki) || (kj == ki && j < i);
  }
  return r;
}

__global__ void __launch_bounds__(NMS_THREADS) aligned_nms_kernel(
    int max_n, const float *__restrict__ boxes, const float *__restrict__ scores,
    const int *__restrict__ classes, const int *__restrict__ counts, float thresh,
    long long *__restrict__ keep, int *__restrict__ keep_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_box = reinterpret_cast<float *>(smem_raw);   // [6][max_n] SoA
  float *s_area = s_box + 6 * max_n;                    // [max_n]
  float *s_score = s_area + max_n;                      // [max_n]
  int *s_cls = reinterpret_cast<int *>(s_score + max_n);  // [max_n]
  int *s_order = s_cls + max_n;                         // sorted position -> box
  int *s_rank = s_order + max_n;                        // box -> sorted position
  int *s_alive = s_rank + max_n;                        // by box
  __shared__ int s_npick;

  const int scene = blockIdx.x, tid = threadIdx.x;
  const int n = counts ? min(counts[scene], max_n) : max_n;
  boxes += (size_t)scene * max_n * 6;
  scores += (size_t)scene * max_n;
  classes += (size_t)scene * max_n;
  keep += (size_t)scene * max_n;

  for (int i = tid; i < n; i += NMS_THREADS) {
    float v[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) { v[a] = boxes[i * 6 + a]; s_box[a * max_n + i] = v[a]; }
    // area = (x2 - x1) * (y2 - y1) * (z2 - z1), box3d_nms.py:147
    s_area[i] = __fmul_rn(__fmul_rn(__fsub_rn(v[3], v[0]), __fsub_rn(v[4], v[1])),
                          __fsub_rn(v[5], v[2]));
    s_score[i] = scores[i];
    s_cls[i] = classes[i];
    s_alive[i] = 1;
  }
  if (tid == 0) s_npick = 0;
  __syncthreads();
  for (int i = tid; i < n; i += NMS_THREADS) {
    const int r = stable_rank(s_score, n, i);
    s_rank[i] = r;
    s_order[r] = i;
  }
  __syncthreads();

  for (int pos = n - 1; pos >= 0; --pos) {
    const int i = s_order[pos];
    if (!s_alive[i]) continue;  // uniform: no thread has written s_alive since the last barrier
    if (tid == 0) keep[s_npick++] = (long long)i;
    const float ix1 = s_box[0 * max_n + i], iy1 = s_box[1 * max_n + i], iz1 = s_box[2 * max_n + i];
    const float ix2 = s_box[3 * max_n + i], iy2 = s_box[4 * max_n + i], iz2 = s_box[5 * max_n + i];
    const float iarea = s_area[i];
    const int icls = s_cls[i];
    // no barrier needed here: this round only clears flags of boxes j != i
    for (int j = tid; j < n; j += NMS_THREADS) {
      if (!s_alive[j] || s_rank[j] >= pos) continue;
      const float xx1 = nanmaxf(ix1, s_box[0 * max_n + j]);
      const float yy1 = nanmaxf(iy1, s_box[1 * max_n + j]);
      const float zz1 = nanmaxf(iz1, s_box[2 * max_n + j]);
      const float xx2 = nanminf(ix2, s_box[3 * max_n + j]);
      const float yy2 = nanminf(iy2, s_box[4 * max_n + j]);
      const float zz2 = nanminf(iz2, s_box[5 * max_n + j]);
      const float l = nanmaxf(0.f, __fsub_rn(xx2, xx1));
      const float w = nanmaxf(0.f, __fsub_rn(yy2, yy1));
      const float h = nanmaxf(0.f, __fsub_rn(zz2, zz1));
      const float inter = __fmul_rn(__fmul_rn(l, w), h);
      float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, s_area[j]), inter));
      iou = __fmul_rn(iou, icls == s_cls[j] ? 1.f : 0.f);
      if (!(iou <= thresh)) s_alive[j] = 0;
    }
    __syncthreads();
  }
  if (tid == 0) keep_cnt[scene] = s_npick;
}

// lhs_3d_faster_samecls, float64.  Rows of `boxes`: x1,y1,z1,x2,y2,z2,score,cls.
__global__ void __launch_bounds__(NMS_THREADS) lhs_nms_kernel(
    int max_n, const double *__restrict__ boxes, const int *__restrict__ counts, double thresh,
    int old_type, int *__restrict__ pick, int *__restrict__ pick_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_box = reinterpret_cast<double *>(smem_raw);  // [8][max_n] SoA
  double *s_area = s_box + 8 * max_n;
  int *s_order = reinterpret_cast<int *>(s_area + max_n);  // sorted position -> box
  int *s_alive = s_order + max_n;                          // by sorted position
  int *s_supp = s_alive + max_n;                           // by sorted position
  __shared__ int s_npick, s_top;

  const int scene = blockIdx.x, tid = threadIdx.x;
  const int n = counts ? min(counts[scene], max_n) : max_n;
  boxes += (size_t)scene * max_n * 8;
  pick += (size_t)scene * 2 * max_n;

  for (int i = tid; i < n; i += NMS_THREADS) {
    double v[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) { v[a] = boxes[i * 8 + a]; s_box[a * max_n + i] = v[a]; }
    // area = (x2-x1)*(y2-y1)*(z2-z1) + 1e-8, votenet_nesie.py:742
    s_area[i] = __dadd_rn(
        __dmul_rn(__dmul_rn(__dsub_rn(v[3], v[0]), __dsub_rn(v[4], v[1])), __dsub_rn(v[5], v[2])),
        1e-8);
  }
  if (tid == 0) { s_npick = 0; s_top = n - 1; }
  __syncthreads();
  for (int i = tid; i < n; i += NMS_THREADS) {
    const int r = stable_rank(s_box + 6 * max_n, n, i);
    s_order[r] = i;
    s_alive[r] = 1;
  }
  __syncthreads();

  while (true) {
    const int top = s_top;  // highest alive sorted position (uniform)
    if (top < 0) break;
    const int i = s_order[top];
    const double ix1 = s_box[0 * max_n + i], iy1 = s_box[1 * max_n + i], iz1 = s_box[2 * max_n + i];
    const double ix2 = s_box[3 * max_n + i], iy2 = s_box[4 * max_n + i], iz2 = s_box[5 * max_n + i];
    const double iarea = s_area[i], icls = s_box[7 * max_n + i];
    for (int p = tid; p < top; p += NMS_THREADS) {
      int sup = 0;
      if (s_alive[p]) {
        const int j = s_order[p];
        const double xx1 = nanmaxd(ix1, s_box[0 * max_n + j]);
        const double yy1 = nanmaxd(iy1, s_box[1 * max_n + j]);
        const double zz1 = nanmaxd(iz1, s_box[2 * max_n + j]);
        const double xx2 = nanmind(ix2, s_box[3 * max_n + j]);
        const double yy2 = nanmind(iy2, s_box[4 * max_n + j]);
        const double zz2 = nanmind(iz2, s_box[5 * max_n + j]);
        const double l = nanmaxd(0.0, __dsub_rn(xx2, xx1));
        const double w = nanmaxd(0.0, __dsub_rn(yy2, yy1));
        const double h = nanmaxd(0.0, __dsub_rn(zz2, zz1));
        const double inter = __dmul_rn(__dmul_rn(l, w), h);
        double o;
        if (old_type) o = __ddiv_rn(inter, s_area[j]);
        else o = __ddiv_rn(inter, __dsub_rn(__dadd_rn(iarea, s_area[j]), inter));
        o = __dmul_rn(o, icls == s_box[7 * max_n + j] ? 1.0 : 0.0);
        sup = o > thresh;
      }
      s_supp[p] = sup;
    }
    __syncthreads();
    if (tid == 0) {
      int np = s_npick;
      pick[np++] = i;
      int len = 0;
      for (int p = 0; p < top; ++p) len += s_supp[p];
      int want = len / 2;  // the top half of the suppressed set is picked too (:774-775)
      for (int p = top - 1; p >= 0 && want > 0; --p)
        if (s_supp[p]) { pick[np++] = s_order[p]; --want; }
      int nt = -1;
      s_alive[top] = 0;
      for (int p = top - 1; p >= 0; --p) {
        if (s_supp[p]) s_alive[p] = 0;
        else if (s_alive[p] && nt < 0) nt = p;
      }
      s_npick = np;
      s_top = nt;
    }
    __syncthreads();
  }
  if (tid == 0) pick_cnt[scene] = s_npick;
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_aligned_3d_nms_batched(int nscenes, int max_n, const float *boxes,
                                            const float *scores, const int *classes,
                                            const int *counts, float thresh, long long *keep,
                                            int *keep_cnt, void *stream) {
  NESIE_REQUIRE(nscenes >= 0 && max_n >= 0, "negative size");
  NESIE_REQUIRE(max_n <= NMS_MAX_N, "max_n > 1024 boxes per scene is not supported");
  NESIE_REQUIRE(keep_cnt, "null pointer");
  if (nscenes == 0) return NESIE_OK;
  NESIE_REQUIRE(max_n == 0 || (boxes && scores && classes && keep), "null pointer");
  const size_t smem = (size_t)max_n * (8 * sizeof(float) + 4 * sizeof(int));
  aligned_nms_kernel<<<nscenes, NMS_THREADS, smem, (cudaStream_t)stream>>>(
      max_n, boxes, scores, classes, counts, thresh, keep, keep_cnt);
  return check_launch("nesie_aligned_3d_nms_batched");
}

extern "C" int nesie_lhs_nms_batched(int nscenes, int max_n, const double *boxes,
                                     const int *counts, double overlap_threshold, int old_type,
                                     int *pick, int *pick_cnt, void *stream) {
  NESIE_REQUIRE(nscenes >= 0 && max_n >= 0, "negative size");
  NESIE_REQUIRE(max_n <= NMS_MAX_N, "max_n > 1024 boxes per scene is not supported");
  NESIE_REQUIRE(pick_cnt, "null pointer");
  if (nscenes == 0) return NESIE_OK;
  NESIE_REQUIRE(max_n == 0 || (boxes && pick), "null pointer");
  const size_t smem = (size_t)max_n * (9 * sizeof(double) + 3 * sizeof(int));
  if (smem > 48 * 1024)
    NESIE_CUDA(cudaFuncSetAttribute(lhs_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  lhs_nms_kernel<<<nscenes, NMS_THREADS, smem, (cudaStream_t)stream>>>(
      max_n, boxes, counts, overlap_threshold, old_type, pick, pick_cnt);
  return check_launch("nesie_lhs_nms_batched");
}
