/*
 * nesie_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * CPU restatement (plain C + OpenMP) of the arithmetic of the reference's hot-path CUDA
 * kernels, one loop iteration per reference CUDA thread, with the same fp32 contraction
 * nvcc emits for them under its default -fmad=true (verified in PTX for sm_100a):
 *     d = fmaf(dz, dz, fmaf(dx, dx, dy * dy))
 *     o = fmaf(w2, p2, fmaf(w0, p0, w1 * p1))
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  Parity status: pinned on the GPU box against the reference's own
 * kernels compiled unmodified (oracle/_ref, see oracle/build_ref.sh) and, on CPU, against the
 * golden vectors those kernels produced (tests/golden/).
 *
 * Reference files restated (paths under /root/reference/mmdet3d/ops):
 *   furthest_point_sample/src/furthest_point_sample_cuda.cu:11-23,25-141,213-331
 *   ball_query/src/ball_query_cuda.cu:11-54
 *   gather_points/src/gather_points_cuda.cu:8-26,51-70
 *   group_points/src/group_points_cuda.cu:10-31,56-79
 *   interpolate/src/three_nn_cuda.cu:11-65
 *   interpolate/src/three_interpolate_cuda.cu:11-35,61-84
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1; the bench sets the thread count explicitly. */
void nesie_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* furthest_point_sample_cuda.cu:11-15 -- block size the reference launcher picks. */
int nesie_oracle_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}

static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  /* a - b, contraction order as in the compiled reference. */
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* Tree reduction of furthest_point_sample_cuda.cu:17-23,75-136: strides bs/2 .. 1,
 * slot t keeps its own entry unless the partner is strictly larger. */
static int fps_tree_argmax(float *dists, int *dists_i, int bs) {
  for (int s = bs / 2; s >= 1; s >>= 1) {
    for (int t = 0; t < s; ++t) {
      const float v1 = dists[t], v2 = dists[t + s];
      const int i1 = dists_i[t], i2 = dists_i[t + s];
      dists[t] = fmaxf(v1, v2);
      dists_i[t] = v2 > v1 ? i2 : i1;
    }
  }
  return dists_i[0];
}

/* furthest_point_sample_cuda.cu:25-141.  temp must be pre-filled by the caller
 * (1e10, furthest_point_sample.py:30) and is updated in place like the reference's. */
void nesie_oracle_fps(int b, int n, int m, const float *xyz, float *temp, int *idxs) {
  if (m <= 0) return;
  const int bs = nesie_oracle_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float *p = xyz + (size_t)bi * n * 3;
    float *t = temp + (size_t)bi * n;
    int *out = idxs + (size_t)bi * m;
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int tid = 0; tid < bs; ++tid) { dists[tid] = -1.f; dists_i[tid] = 0; }
      /* thread tid walks k = tid, tid+bs, ... ascending; visiting k in plain ascending
       * order gives every tid the same sequence. */
      for (int k0 = 0; k0 < n; k0 += bs) {
        const int lim = (n - k0 < bs) ? n - k0 : bs;
        for (int tid = 0; tid < lim; ++tid) {
          const int k = k0 + tid;
          const float d = sqdist(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
          const float d2 = fminf(d, t[k]);
          t[k] = d2;
          if (d2 > dists[tid]) { dists[tid] = d2; dists_i[tid] = k; }
        }
      }
      old = fps_tree_argmax(dists, dists_i, bs);
      out[j] = old;
    }
    free(dists);
    free(dists_i);
  }
}

/* furthest_point_sample_cuda.cu:213-331: same loop over a (B,N,N) distance matrix. */
void nesie_oracle_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idxs) {
  if (m <= 0) return;
  const int bs = nesie_oracle_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float *dm = dist + (size_t)bi * n * n;
    float *t = temp + (size_t)bi * n;
    int *out = idxs + (size_t)bi * m;
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      for (int tid = 0; tid < bs; ++tid) { dists[tid] = -1.f; dists_i[tid] = 0; }
      for (int k = 0; k < n; ++k) {
        const int tid = k % bs;
        const float d2 = fminf(dm[(size_t)old * n + k], t[k]);
        t[k] = d2;
        if (d2 > dists[tid]) { dists[tid] = d2; dists_i[tid] = k; }
      }
      old = fps_tree_argmax(dists, dists_i, bs);
      out[j] = old;
    }
    free(dists);
    free(dists_i);
  }
}

/* ball_query_cuda.cu:11-54.  idx must be zero-filled by the caller (ball_query.py:35). */
void nesie_oracle_ball_query(int b, int n, int m, float min_radius, float max_radius, int nsample,
                             const float *new_xyz, const float *xyz, int *idx) {
  const float max_r2 = max_radius * max_radius;
  const float min_r2 = min_radius * min_radius;
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi) {
    for (int pt = 0; pt < m; ++pt) {
      const float *c = new_xyz + ((size_t)bi * m + pt) * 3;
      const float *p = xyz + (size_t)bi * n * 3;
      int *o = idx + ((size_t)bi * m + pt) * nsample;
      const float cx = c[0], cy = c[1], cz = c[2];
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        const float d2 = sqdist(cx, cy, cz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 == 0 || (d2 >= min_r2 && d2 < max_r2)) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
          if (cnt >= nsample) break;
        }
      }
    }
  }
}

/* gather_points_cuda.cu:8-26 */
void nesie_oracle_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                                float *out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *src = points + ((size_t)bi * c + ci) * n;
      const int *id = idx + (size_t)bi * m;
      float *dst = out + ((size_t)bi * c + ci) * m;
      for (int j = 0; j < m; ++j) dst[j] = src[id[j]];
    }
}

/* gather_points_cuda.cu:51-70 -- sequential accumulation in ascending m (the reference's
 * atomicAdd order is unspecified; gradient parity is tolerance-based). grad_points zeroed
 * by the caller (gather_points.py:44). */
void nesie_oracle_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                                     const int *idx, float *grad_points) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *g = grad_out + ((size_t)bi * c + ci) * m;
      const int *id = idx + (size_t)bi * m;
      float *dst = grad_points + ((size_t)bi * c + ci) * n;
      for (int j = 0; j < m; ++j) dst[id[j]] += g[j];
    }
}

/* group_points_cuda.cu:56-79 */
void nesie_oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                               const int *idx, float *out) {
  const size_t mk = (size_t)npoints * nsample;
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *src = points + ((size_t)bi * c + ci) * n;
      const int *id = idx + (size_t)bi * mk;
      float *dst = out + ((size_t)bi * c + ci) * mk;
      for (size_t j = 0; j < mk; ++j) dst[j] = src[id[j]];
    }
}

/* group_points_cuda.cu:10-31 */
void nesie_oracle_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                    const float *grad_out, const int *idx, float *grad_points) {
  const size_t mk = (size_t)npoints * nsample;
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *g = grad_out + ((size_t)bi * c + ci) * mk;
      const int *id = idx + (size_t)bi * mk;
      float *dst = grad_points + ((size_t)bi * c + ci) * n;
      for (size_t j = 0; j < mk; ++j) dst[id[j]] += g[j];
    }
}

/* three_nn_cuda.cu:11-65 -- running bests are doubles (line 35); output is d^2
 * (the sqrt is applied in python, three_nn.py:38). */
void nesie_oracle_three_nn(int b, int n, int m, const float *unknown, const float *known,
                           float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int pt = 0; pt < n; ++pt) {
      const float *u = unknown + ((size_t)bi * n + pt) * 3;
      const float *kn = known + (size_t)bi * m * 3;
      const float ux = u[0], uy = u[1], uz = u[2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist(ux, uy, uz, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float *od = dist2 + ((size_t)bi * n + pt) * 3;
      int *oi = idx + ((size_t)bi * n + pt) * 3;
      od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
      oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
}

/* three_interpolate_cuda.cu:11-35 */
void nesie_oracle_three_interpolate(int b, int c, int m, int n, const float *points,
                                    const int *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *src = points + ((size_t)bi * c + ci) * m;
      const int *id = idx + (size_t)bi * n * 3;
      const float *w = weight + (size_t)bi * n * 3;
      float *dst = out + ((size_t)bi * c + ci) * n;
      for (int j = 0; j < n; ++j)
        dst[j] = fmaf(w[j * 3 + 2], src[id[j * 3 + 2]],
                      fmaf(w[j * 3 + 0], src[id[j * 3 + 0]], w[j * 3 + 1] * src[id[j * 3 + 1]]));
    }
}

/* three_interpolate_cuda.cu:61-84, sequential accumulation order. */
void nesie_oracle_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                         const int *idx, const float *weight,
                                         float *grad_points) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci) {
      const float *g = grad_out + ((size_t)bi * c + ci) * n;
      const int *id = idx + (size_t)bi * n * 3;
      const float *w = weight + (size_t)bi * n * 3;
      float *dst = grad_points + ((size_t)bi * c + ci) * m;
      for (int j = 0; j < n; ++j) {
        dst[id[j * 3 + 0]] += g[j] * w[j * 3 + 0];
        dst[id[j * 3 + 1]] += g[j] * w[j * 3 + 1];
        dst[id[j * 3 + 2]] += g[j] * w[j * 3 + 2];
      }
    }
}

/* ---------------------------------------------------------------------------------------------
 * points_in_boxes / points_in_boxes_batch (reference:
 * ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:24-105).  One loop iteration per reference CUDA
 * thread.  The reference mixes float and double exactly as restated here (`h / 2.0`, `M_PI / 2`
 * and the comparisons are double; the rotated coordinates are float with the contraction nvcc
 * emits, checked in its sm_100a SASS:  lx = fma(sx, cos, -(sy * sin)),  ly = fma(sy, cos, sx * sin)).
 * cosf / sinf are libm's here and CUDA's there: they agree for the yaw-0 boxes of this path
 * (ScanNet), for other yaws a point within rounding of a face may differ.
 * batch == 0: out (b, npts) int32 = index of the first containing box, caller-initialised to -1;
 * batch != 0: out (b, npts, nbox) int32 = 1 where the point is inside, caller-initialised to 0. */
static int pib_check(const float *pt, const float *box) {
  const float x = pt[0], y = pt[1], z = pt[2];
  const float cx = box[0], cy = box[1];
  float cz = box[2];
  const float w = box[3], l = box[4], h = box[5], rz = box[6];
  cz = (float)((double)cz + (double)h / 2.0);
  if ((double)fabsf(z - cz) > (double)h / 2.0) return 0;
  const float sx = x - cx, sy = y - cy;
  const float rot = (float)((double)rz + M_PI / 2);
  const float cosa = cosf(rot), sina = sinf(rot);
  const float lx = fmaf(sx, cosa, -(sy * sina));
  const float ly = fmaf(sy, cosa, sx * sina);
  return ((double)lx > -(double)l / 2.0) & ((double)lx < (double)l / 2.0) &
         ((double)ly > -(double)w / 2.0) & ((double)ly < (double)w / 2.0);
}

void nesie_oracle_points_in_boxes(int b, int nbox, int npts, const float *boxes, const float *pts,
                                  int *out, int batch) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int p = 0; p < npts; ++p) {
      const float *pt = pts + ((size_t)bi * npts + p) * 3;
      const float *bx = boxes + (size_t)bi * nbox * 7;
      for (int k = 0; k < nbox; ++k) {
        if (pib_check(pt, bx + (size_t)k * 7)) {
          if (batch) out[((size_t)bi * npts + p) * nbox + k] = 1;
          else { out[(size_t)bi * npts + p] = k; break; }
        }
      }
    }
}

/* ---------------------------------------------------------------------------------------------
 * sort_vertices (reference: ops/rotated_iou/cuda_op/sort_vert_kernel.cu:16-134).
 * One loop iteration per reference CUDA thread (= one polygon).  `EPSILON` is the double literal
 * 1e-8 there, so the comparisons against it are double; n = x*x + y*y contracts to
 * fma(x, x, y*y) in float, then `+ EPSILON` in double, rounded back to float.  The reference's
 * comparator has no return statement when one of the two y's is exactly 0 (or NaN); nvcc's code
 * for that path returns false (checked in the sm_100a SASS of the reference kernel), restated. */
#define SV_MAX_IDX 9
#define SV_INTER_OFF 8
static int sv_less(float x1, float y1, float x2, float y2) {
  const double EPS = 1e-8;
  if ((double)fabsf(x1 - x2) < EPS && (double)fabsf(y2 - y1) < EPS) return 0;
  if (y1 > 0 && y2 < 0) return 1;
  if (y1 < 0 && y2 > 0) return 0;
  const float n1 = (float)((double)fmaf(x1, x1, y1 * y1) + EPS);
  const float n2 = (float)((double)fmaf(x2, x2, y2 * y2) + EPS);
  const float d = fabsf(x1) * x1 / n1 - fabsf(x2) * x2 / n2;
  if (y1 > 0 && y2 > 0) return (double)d > EPS;
  if (y1 < 0 && y2 < 0) return (double)d < EPS;
  return 0;
}

void nesie_oracle_sort_vertices(int b, int n, int m, const float *vertices,
                                const unsigned char *mask, const int *num_valid, int *idx) {
#pragma omp parallel for schedule(static)
  for (long p = 0; p < (long)b * n; ++p) {
    const float *v = vertices + (size_t)p * m * 2;
    const unsigned char *mk = mask + (size_t)p * m;
    int *out = idx + (size_t)p * SV_MAX_IDX;
    const int nv = num_valid[p];
    int pad = 0;
    for (int j = SV_INTER_OFF; j < m; ++j)
      if (!mk[j]) { pad = j; break; }
    if (nv < 3) {
      for (int j = 0; j < SV_MAX_IDX; ++j) out[j] = pad;
      continue;
    }
    for (int j = 0; j < nv; ++j) {
      float x_min = 1.f, y_min = (float)(-1e-8);
      int take = 0;
      for (int k = 0; k < m; ++k) {
        const float x = v[2 * k], y = v[2 * k + 1];
        if (j == 0) {
          if (mk[k] && sv_less(x, y, x_min, y_min)) { x_min = x; y_min = y; take = k; }
        } else {
          const int i2 = out[j - 1];
          const float x2 = v[2 * i2], y2 = v[2 * i2 + 1];
          if (mk[k] && sv_less(x, y, x_min, y_min) && sv_less(x2, y2, x, y)) {
            x_min = x; y_min = y; take = k;
          }
        }
      }
      if (j < SV_MAX_IDX) out[j] = take;
    }
    out[nv] = out[0];
    for (int j = nv + 1; j < SV_MAX_IDX; ++j) out[j] = pad;
    if (nv == 8) {
      int counter = 0;
      for (int j = 0; j < 4; ++j)
        for (int k = 4; k < SV_INTER_OFF; ++k)
          if (out[k] == out[j]) counter++;
      if (counter == 4) {
        out[4] = out[0];
        for (int j = 5; j < SV_MAX_IDX; ++j) out[j] = pad;
      }
    }
  }
}
