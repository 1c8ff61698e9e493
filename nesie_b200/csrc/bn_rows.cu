// Training-mode BatchNorm + ReLU (+ max-pool over the nsample rows of a group) on ROW-major
// activations (R rows x C channels, C % 4 == 0), forward and backward, sm_100a.
//
// In the reference this is the BN2d / ReLU / F.max_pool2d tail of every ConvModule of the SA
// shared MLP (ops/pointnet_modules/point_sa_module.py:149-150,279-288) on channel-major tensors,
// i.e. per layer: statistics pass, normalise pass, in-place ReLU pass, and in backward a ReLU-mask
// pass, a reduction pass and an element pass (plus the pooling passes) -- 46 sweeps over the
// activations of one SA module, all HBM-bound (1.9 GB of activations per step at batch 8).  Here:
//   forward : stats (read Y)  ->  finalize (C threads)  ->  apply: A = relu(Y*scale + shift)
//             last layer: apply+pool: out[g] = relu(max_k(Y*scale+shift)), arg[g] = first argmax
//   backward: reduce (read dA, Y): sum g, sum g*xhat with g = dA * [A > 0] recomputed from Y
//             -> finalize -> apply: dY = scale * (g - mean(g) - xhat * mean(g*xhat))
//             last layer: g exists only at the argmax rows, so the reduction touches G*C values
// 21 sweeps instead of 46.  Statistics are accumulated around a per-channel pivot (row 0) in fp32
// per thread / CTA and combined across CTAs in double, so E[y^2] - E[y]^2 does not cancel.
#include "common.cuh"

namespace nesie {
namespace {

constexpr int BN_THREADS = 256;
constexpr int BN_MAXPART = 1184;  // partial blocks (8 per SM)

// ---- column sums of two quantities per element, 4 channels per thread -------------------------
// mode 0 (forward):  q1 = y - pivot,            q2 = (y - pivot)^2
// mode 1 (backward): q1 = g,                     q2 = g * xhat     (g = dA * [y*scale+shift > 0])
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_colsum_kernel(
    long long R, int C, const float *__restrict__ Y, const float *__restrict__ dA,
    const float *__restrict__ scale, const float *__restrict__ shift,
    const float *__restrict__ mean, const float *__restrict__ invstd,
    float *__restrict__ partial /* [gridDim.x][2][C] */) {
  __shared__ float4 s_acc[BN_THREADS * 2];
  const int tpr = C >> 2;                       // threads per row
  const int rpi = BN_THREADS / tpr;             // rows per block iteration
  const int cx = threadIdx.x % tpr, ry = threadIdx.x / tpr;
  const int c = cx * 4;
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  float4 p0 = a1, sc = a1, sh = a1, mu = a1, is = a1;
  if (ry < rpi) {
    if (MODE == 0) {
      p0 = *reinterpret_cast<const float4 *>(Y + c);  // pivot = row 0
    } else {
      sc = *reinterpret_cast<const float4 *>(scale + c);
      sh = *reinterpret_cast<const float4 *>(shift + c);
      mu = *reinterpret_cast<const float4 *>(mean + c);
      is = *reinterpret_cast<const float4 *>(invstd + c);
    }
    for (long long r = (long long)blockIdx.x * rpi + ry; r < R; r += (long long)gridDim.x * rpi) {
      // backward: plain loads (the apply kernel that follows walks the rows in reverse and finds the
      // tail of this sweep still in the 126 MB L2); forward statistics: streaming loads
      const float4 y = MODE == 0 ? __ldcs(reinterpret_cast<const float4 *>(Y + r * C + c))
                                 : __ldg(reinterpret_cast<const float4 *>(Y + r * C + c));
      if (MODE == 0) {
        const float dx = y.x - p0.x, dy = y.y - p0.y, dz = y.z - p0.z, dw = y.w - p0.w;
        a1.x += dx; a1.y += dy; a1.z += dz; a1.w += dw;
        a2.x = fmaf(dx, dx, a2.x); a2.y = fmaf(dy, dy, a2.y);
        a2.z = fmaf(dz, dz, a2.z); a2.w = fmaf(dw, dw, a2.w);
      } else {
        const float4 d = __ldg(reinterpret_cast<const float4 *>(dA + r * C + c));
        const float gx = fmaf(y.x, sc.x, sh.x) > 0.f ? d.x : 0.f;
        const float gy = fmaf(y.y, sc.y, sh.y) > 0.f ? d.y : 0.f;
        const float gz = fmaf(y.z, sc.z, sh.z) > 0.f ? d.z : 0.f;
        const float gw = fmaf(y.w, sc.w, sh.w) > 0.f ? d.w : 0.f;
        a1.x += gx; a1.y += gy; a1.z += gz; a1.w += gw;
        a2.x = fmaf(gx, (y.x - mu.x) * is.x, a2.x); a2.y = fmaf(gy, (y.y - mu.y) * is.y, a2.y);
        a2.z = fmaf(gz, (y.z - mu.z) * is.z, a2.z); a2.w = fmaf(gw, (y.w - mu.w) * is.w, a2.w);
      }
    }
  }
  s_acc[threadIdx.x] = a1;
  s_acc[BN_THREADS + threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.x < tpr) {  // thread cx sums its channel group over the rpi row slots
    float4 t1 = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t1;
    for (int j = 0; j < rpi; ++j) {
      const float4 u = s_acc[j * tpr + threadIdx.x], v = s_acc[BN_THREADS + j * tpr + threadIdx.x];
      t1.x += u.x; t1.y += u.y; t1.z += u.z; t1.w += u.w;
      t2.x += v.x; t2.y += v.y; t2.z += v.z; t2.w += v.w;
    }
    float *pp = partial + (size_t)blockIdx.x * 2 * C;
    *reinterpret_cast<float4 *>(pp + threadIdx.x * 4) = t1;
    *reinterpret_cast<float4 *>(pp + C + threadIdx.x * 4) = t2;
  }
}

// Sum the per-CTA partials of FIN_CH channels: blockDim = (FIN_CH channels, FIN_SLICES slices); every
// slice walks its share of the partial blocks in double, then the slices are combined through shared
// memory.  Few channels per CTA spread the ~0.6 MB of partials over C / 8 SMs (one SM pulls only
// ~150 GB/s: with 32 channels per CTA the finalize of a 64-channel layer took 9 us on two SMs).
constexpr int FIN_CH = 8, FIN_SLICES = 128;

__device__ __forceinline__ bool reduce_partials(int C, int nparts, const float *__restrict__ partial,
                                                double &s1, double &s2) {
  __shared__ double s_red[2][FIN_SLICES][FIN_CH + 1];
  const int c = blockIdx.x * FIN_CH + threadIdx.x;
  double a1 = 0.0, a2 = 0.0;
  if (c < C) {
    for (int p = threadIdx.y; p < nparts; p += FIN_SLICES) {
      a1 += (double)partial[(size_t)p * 2 * C + c];
      a2 += (double)partial[(size_t)p * 2 * C + C + c];
    }
  }
  s_red[0][threadIdx.y][threadIdx.x] = a1;
  s_red[1][threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  // tree over the slices, fixed order: deterministic
  for (int h = FIN_SLICES / 2; h >= 1; h >>= 1) {
    if (threadIdx.y < h) {
      s_red[0][threadIdx.y][threadIdx.x] += s_red[0][threadIdx.y + h][threadIdx.x];
      s_red[1][threadIdx.y][threadIdx.x] += s_red[1][threadIdx.y + h][threadIdx.x];
    }
    __syncthreads();
  }
  if (threadIdx.y != 0 || c >= C) return false;
  s1 = s_red[0][0][threadIdx.x];
  s2 = s_red[1][0][threadIdx.x];
  return true;
}

// forward finalize: partial sums -> mean / invstd / scale / shift, running stats
__global__ void __launch_bounds__(FIN_CH * FIN_SLICES) bn_fwd_finalize_kernel(
    long long R, int C, int nparts, const float *__restrict__ Y, const float *__restrict__ partial,
    const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float momentum,
    float *__restrict__ running_mean, float *__restrict__ running_var,
    float *__restrict__ stats /* [4][C]: mean, invstd, scale, shift */) {
  double s1, s2;
  if (!reduce_partials(C, nparts, partial, s1, s2)) return;
  const int c = blockIdx.x * FIN_CH + threadIdx.x;
  const double n = (double)R, m1 = s1 / n;
  double var = s2 / n - m1 * m1;  // variance of (y - pivot) == variance of y
  if (var < 0.0) var = 0.0;
  const double mean = (Y ? (double)Y[c] : 0.0) + m1;  // Y == nullptr: pivot-free sums
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  stats[c] = (float)mean;
  stats[C + c] = invstd;
  stats[2 * C + c] = sc;
  stats[3 * C + c] = beta[c] - (float)mean * sc;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    const double unbiased = R > 1 ? var * n / (n - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// backward finalize: sum g -> d beta, sum g*xhat -> d gamma, and the two means used by the apply pass
__global__ void __launch_bounds__(FIN_CH * FIN_SLICES) bn_bwd_finalize_kernel(
    long long R, int C, int nparts, const float *__restrict__ partial, float *__restrict__ dgamma,
    float *__restrict__ dbeta, float *__restrict__ coef /* [2][C]: mean g, mean g*xhat */) {
  double s1, s2;
  if (!reduce_partials(C, nparts, partial, s1, s2)) return;
  const int c = blockIdx.x * FIN_CH + threadIdx.x;
  dbeta[c] = (float)s1;
  dgamma[c] = (float)s2;
  coef[c] = (float)(s1 / (double)R);
  coef[C + c] = (float)(s2 / (double)R);
}

// A = relu(Y * scale + shift)
__global__ void __launch_bounds__(BN_THREADS) bn_relu_apply_kernel(
    long long n4, int C, const float *__restrict__ Y, const float *__restrict__ stats,
    float *__restrict__ A) {
  const float *scale = stats + 2 * C, *shift = stats + 3 * C;
  for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < n4;
       i += (long long)gridDim.x * BN_THREADS) {
    const int c = (int)((i * 4) % C);
    const float4 y = __ldcs(reinterpret_cast<const float4 *>(Y) + i);
    const float4 sc = *reinterpret_cast<const float4 *>(scale + c);
    const float4 sh = *reinterpret_cast<const float4 *>(shift + c);
    float4 a;
    a.x = fmaxf(fmaf(y.x, sc.x, sh.x), 0.f); a.y = fmaxf(fmaf(y.y, sc.y, sh.y), 0.f);
    a.z = fmaxf(fmaf(y.z, sc.z, sh.z), 0.f); a.w = fmaf(y.w, sc.w, sh.w) > 0.f ? fmaf(y.w, sc.w, sh.w) : 0.f;
    reinterpret_cast<float4 *>(A)[i] = a;
  }
}

// last layer: out[g, c] = relu(max_k (Y[g*K + k, c] * scale + shift)); arg = first maximising k
// (255 when the maximum is not positive: ReLU passes no gradient)
__global__ void __launch_bounds__(BN_THREADS) bn_relu_pool_kernel(
    long long G, int K, int C, const float *__restrict__ Y, const float *__restrict__ stats,
    float *__restrict__ out, unsigned char *__restrict__ arg) {
  const float *scale = stats + 2 * C, *shift = stats + 3 * C;
  const int tpr = C >> 2;
  const long long total = G * tpr;
  for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * BN_THREADS) {
    const long long g = i / tpr;
    const int c = (int)(i - g * tpr) * 4;
    const float4 sc = *reinterpret_cast<const float4 *>(scale + c);
    const float4 sh = *reinterpret_cast<const float4 *>(shift + c);
    float4 best = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
    int bx = 0, by = 0, bz = 0, bw = 0;
    const float *row = Y + (g * K) * C + c;
    for (int k = 0; k < K; ++k) {
      const float4 y = __ldcs(reinterpret_cast<const float4 *>(row + (long long)k * C));
      const float vx = fmaf(y.x, sc.x, sh.x), vy = fmaf(y.y, sc.y, sh.y);
      const float vz = fmaf(y.z, sc.z, sh.z), vw = fmaf(y.w, sc.w, sh.w);
      if (vx > best.x) { best.x = vx; bx = k; }
      if (vy > best.y) { best.y = vy; by = k; }
      if (vz > best.z) { best.z = vz; bz = k; }
      if (vw > best.w) { best.w = vw; bw = k; }
    }
    float4 o;
    o.x = fmaxf(best.x, 0.f); o.y = fmaxf(best.y, 0.f); o.z = fmaxf(best.z, 0.f); o.w = fmaxf(best.w, 0.f);
    *reinterpret_cast<float4 *>(out + g * C + c) = o;
    uchar4 a;
    a.x = best.x > 0.f ? (unsigned char)bx : 255; a.y = best.y > 0.f ? (unsigned char)by : 255;
    a.z = best.z > 0.f ? (unsigned char)bz : 255; a.w = best.w > 0.f ? (unsigned char)bw : 255;
    *reinterpret_cast<uchar4 *>(arg + g * C + c) = a;
  }
}

// pooled backward reduction: g lives only at (group, argmax row): G*C terms
__global__ void __launch_bounds__(BN_THREADS) bn_pool_bwd_colsum_kernel(
    long long G, int K, int C, const float *__restrict__ Y, const float *__restrict__ dOut,
    const unsigned char *__restrict__ arg, const float *__restrict__ stats,
    float *__restrict__ partial) {
  __shared__ float s1[BN_THREADS], s2[BN_THREADS];
  const float *mean = stats, *invstd = stats + C;
  // thread <-> channel (threadIdx.x % C-chunk); rows of groups strided over y and blocks
  const int lanes = C < BN_THREADS ? C : BN_THREADS;
  const int gpi = BN_THREADS / lanes;  // groups per block iteration
  for (int cb = 0; cb < C; cb += lanes) {
    const int c = cb + threadIdx.x % lanes, gy = threadIdx.x / lanes;
    float a1 = 0.f, a2 = 0.f;
    if (gy < gpi) {
      const float mu = mean[c], is = invstd[c];
      for (long long g = (long long)blockIdx.x * gpi + gy; g < G; g += (long long)gridDim.x * gpi) {
        const unsigned k = arg[g * C + c];
        if (k != 255) {
          const float d = dOut[g * C + c];
          const float y = Y[(g * K + k) * C + c];
          a1 += d;
          a2 = fmaf(d, (y - mu) * is, a2);
        }
      }
    }
    s1[threadIdx.x] = a1; s2[threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.x < lanes) {
      float t1 = 0.f, t2 = 0.f;
      for (int j = 0; j < gpi; ++j) { t1 += s1[j * lanes + threadIdx.x]; t2 += s2[j * lanes + threadIdx.x]; }
      partial[(size_t)blockIdx.x * 2 * C + cb + threadIdx.x] = t1;
      partial[(size_t)blockIdx.x * 2 * C + C + cb + threadIdx.x] = t2;
    }
    __syncthreads();
  }
}

// dY = scale * (g - mean_g - xhat * mean_gx);  g = dA*[A>0] (dense) or dOut at the argmax row (pooled)
template <bool POOLED>
__global__ void __launch_bounds__(BN_THREADS) bn_relu_bwd_apply_kernel(
    long long R, int K, int C, const float *__restrict__ Y, const float *__restrict__ dA,
    const unsigned char *__restrict__ arg, const float *__restrict__ stats,
    const float *__restrict__ coef, float *__restrict__ dY) {
  const float *mean = stats, *invstd = stats + C, *scale = stats + 2 * C, *shift = stats + 3 * C;
  const int tpr = C >> 2;
  const long long n4 = R * tpr;
  // rows from the last to the first: the statistics sweep that ran just before read them in
  // ascending order, so its most recent ~100 MB are served from L2 here
  for (long long j = (long long)blockIdx.x * BN_THREADS + threadIdx.x; j < n4;
       j += (long long)gridDim.x * BN_THREADS) {
    const long long i = n4 - 1 - j;
    const long long r = i / tpr;
    const int c = (int)(i - r * tpr) * 4;
    const float4 y = __ldcs(reinterpret_cast<const float4 *>(Y) + i);
    const float4 mu = *reinterpret_cast<const float4 *>(mean + c);
    const float4 is = *reinterpret_cast<const float4 *>(invstd + c);
    const float4 sc = *reinterpret_cast<const float4 *>(scale + c);
    const float4 m1 = *reinterpret_cast<const float4 *>(coef + c);
    const float4 m2 = *reinterpret_cast<const float4 *>(coef + C + c);
    float4 g;
    if (POOLED) {
      const long long grp = r / K;
      const unsigned k = (unsigned)(r - grp * K);
      const uchar4 a = *reinterpret_cast<const uchar4 *>(arg + grp * C + c);
      const float4 d = *reinterpret_cast<const float4 *>(dA + grp * C + c);
      g.x = a.x == k ? d.x : 0.f; g.y = a.y == k ? d.y : 0.f;
      g.z = a.z == k ? d.z : 0.f; g.w = a.w == k ? d.w : 0.f;
    } else {
      const float4 sh = *reinterpret_cast<const float4 *>(shift + c);
      const float4 d = __ldcs(reinterpret_cast<const float4 *>(dA) + i);
      g.x = fmaf(y.x, sc.x, sh.x) > 0.f ? d.x : 0.f; g.y = fmaf(y.y, sc.y, sh.y) > 0.f ? d.y : 0.f;
      g.z = fmaf(y.z, sc.z, sh.z) > 0.f ? d.z : 0.f; g.w = fmaf(y.w, sc.w, sh.w) > 0.f ? d.w : 0.f;
    }
    float4 o;
    o.x = sc.x * (g.x - m1.x - (y.x - mu.x) * is.x * m2.x);
    o.y = sc.y * (g.y - m1.y - (y.y - mu.y) * is.y * m2.y);
    o.z = sc.z * (g.z - m1.z - (y.z - mu.z) * is.z * m2.z);
    o.w = sc.w * (g.w - m1.w - (y.w - mu.w) * is.w * m2.w);
    reinterpret_cast<float4 *>(dY)[i] = o;
  }
}

int parts_for(long long rows, int C) {
  const int rpi = BN_THREADS / (C >> 2);
  long long g = (rows + rpi - 1) / rpi;
  const int cap = 8 * num_sms() < BN_MAXPART ? 8 * num_sms() : BN_MAXPART;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}
int grid_for(long long items) {
  long long g = (items + BN_THREADS - 1) / BN_THREADS;
  const long long cap = 16LL * num_sms();
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}
bool shape_ok(long long R, int C) { return R >= 1 && C >= 4 && C <= 1024 && (C & 3) == 0; }

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" long long nesie_bn_rows_workspace_bytes(int c) {
  return (long long)BN_MAXPART * 2 * c * sizeof(float);
}

// Forward statistics + apply.  stats (4*C) receives mean, invstd, scale, shift (saved for backward).
// K == 0: A (R, C) = relu(bn(Y)).   K > 0: pooled (R/K, C) + arg (R/K, C) uint8, A is not written.
// col_partials != nullptr: [nparts][2][c] pivot-free column sums (sum y, sum y^2) produced by the
// GEMM epilogue (nesie_gemm_nt_3xtf32_fused) replace the statistics sweep over y.
// a_or_pooled == nullptr: statistics only (the consumer applies scale / shift + ReLU itself).
static int bn_forward_impl(long long r, int c, int k, const float *y, const float *gamma,
                           const float *beta, float eps, float momentum, float *running_mean,
                           float *running_var, const float *col_partials, int nparts_in,
                           float *stats, float *a_or_pooled, unsigned char *arg, void *workspace,
                           void *stream) {
  NESIE_REQUIRE(shape_ok(r, c), "need R >= 1 and C a multiple of 4 in [4, 1024]");
  NESIE_REQUIRE(k >= 0 && k <= 254 && (k == 0 || r % k == 0), "bad pooling window");
  NESIE_REQUIRE(y && gamma && beta && stats && (!a_or_pooled || k == 0 || arg), "null pointer");
  NESIE_REQUIRE(col_partials || workspace, "null pointer");
  NESIE_REQUIRE(!col_partials || nparts_in >= 1, "need nparts >= 1 with col_partials");
  cudaStream_t st = (cudaStream_t)stream;
  if (col_partials) {
    bn_fwd_finalize_kernel<<<ceil_div(c, FIN_CH), dim3(FIN_CH, FIN_SLICES), 0, st>>>(
        r, c, nparts_in, nullptr, col_partials, gamma, beta, eps, momentum, running_mean, running_var,
        stats);
  } else {
    float *partial = reinterpret_cast<float *>(workspace);
    const int nparts = parts_for(r, c);
    bn_colsum_kernel<0><<<nparts, BN_THREADS, 0, st>>>(r, c, y, nullptr, nullptr, nullptr, nullptr,
                                                      nullptr, partial);
    bn_fwd_finalize_kernel<<<ceil_div(c, FIN_CH), dim3(FIN_CH, FIN_SLICES), 0, st>>>(
        r, c, nparts, y, partial, gamma, beta, eps, momentum, running_mean, running_var, stats);
  }
  if (!a_or_pooled) return check_launch("nesie_bn_rows_forward");
  if (k == 0) {
    const long long n4 = r * (c >> 2);
    bn_relu_apply_kernel<<<grid_for(n4), BN_THREADS, 0, st>>>(n4, c, y, stats, a_or_pooled);
  } else {
    const long long G = r / k;
    bn_relu_pool_kernel<<<grid_for(G * (c >> 2)), BN_THREADS, 0, st>>>(G, k, c, y, stats,
                                                                      a_or_pooled, arg);
  }
  return check_launch("nesie_bn_relu_rows_forward");
}

extern "C" int nesie_bn_relu_rows_forward(long long r, int c, int k, const float *y,
                                          const float *gamma, const float *beta, float eps,
                                          float momentum, float *running_mean, float *running_var,
                                          float *stats, float *a_or_pooled, unsigned char *arg,
                                          void *workspace, void *stream) {
  NESIE_REQUIRE(a_or_pooled && workspace, "null pointer");
  return bn_forward_impl(r, c, k, y, gamma, beta, eps, momentum, running_mean, running_var, nullptr, 0,
                         stats, a_or_pooled, arg, workspace, stream);
}

extern "C" int nesie_bn_rows_forward_fused(long long r, int c, int k, const float *y,
                                           const float *gamma, const float *beta, float eps,
                                           float momentum, float *running_mean, float *running_var,
                                           const float *col_partials, int nparts, float *stats,
                                           float *a_or_pooled, unsigned char *arg, void *workspace,
                                           void *stream) {
  return bn_forward_impl(r, c, k, y, gamma, beta, eps, momentum, running_mean, running_var,
                         col_partials, nparts, stats, a_or_pooled, arg, workspace, stream);
}

// Backward of a dense (un-pooled) layer whose statistics sweep already happened in the epilogue of the
// data-gradient GEMM that produced d_a (nesie_gemm_nt_3xtf32_bnbwd): finalize + apply only.
extern "C" int nesie_bn_relu_rows_backward_fused(long long r, int c, const float *y, const float *d_a,
                                                 const float *stats, const float *col_partials,
                                                 int nparts, float *d_y, float *d_gamma, float *d_beta,
                                                 void *workspace, void *stream) {
  NESIE_REQUIRE(shape_ok(r, c), "need R >= 1 and C a multiple of 4 in [4, 1024]");
  NESIE_REQUIRE(y && d_a && stats && col_partials && d_y && d_gamma && d_beta && workspace, "null pointer");
  NESIE_REQUIRE(nparts >= 1, "need nparts >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  float *coef = reinterpret_cast<float *>(workspace) + (size_t)(BN_MAXPART - 1) * 2 * c;
  bn_bwd_finalize_kernel<<<ceil_div(c, FIN_CH), dim3(FIN_CH, FIN_SLICES), 0, st>>>(r, c, nparts, col_partials,
                                                                                 d_gamma, d_beta, coef);
  const long long n4 = r * (c >> 2);
  bn_relu_bwd_apply_kernel<false><<<grid_for(n4), BN_THREADS, 0, st>>>(r, 1, c, y, d_a, nullptr, stats, coef, d_y);
  return check_launch("nesie_bn_relu_rows_backward_fused");
}

// Backward.  K == 0: d_a is (R, C).  K > 0: d_a is the pooled gradient (R/K, C) and arg the forward's.
extern "C" int nesie_bn_relu_rows_backward(long long r, int c, int k, const float *y,
                                           const float *d_a, const unsigned char *arg,
                                           const float *stats, float *d_y, float *d_gamma,
                                           float *d_beta, void *workspace, void *stream) {
  NESIE_REQUIRE(shape_ok(r, c), "need R >= 1 and C a multiple of 4 in [4, 1024]");
  NESIE_REQUIRE(k >= 0 && k <= 254 && (k == 0 || r % k == 0), "bad pooling window");
  NESIE_REQUIRE(y && d_a && stats && d_y && d_gamma && d_beta && workspace && (k == 0 || arg), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  float *partial = reinterpret_cast<float *>(workspace);
  float *coef = partial + (size_t)(BN_MAXPART - 1) * 2 * c;  // last partial slot doubles as coef
  int nparts;
  if (k == 0) {
    nparts = parts_for(r, c);
    if (nparts > BN_MAXPART - 1) nparts = BN_MAXPART - 1;
    bn_colsum_kernel<1><<<nparts, BN_THREADS, 0, st>>>(r, c, y, d_a, stats + 2 * c, stats + 3 * c,
                                                      stats, stats + c, partial);
  } else {
    const long long G = r / k;
    const int lanes = c < BN_THREADS ? c : BN_THREADS;
    const int gpi = BN_THREADS / lanes;
    long long g = (G + gpi - 1) / gpi;
    if (g > 4 * num_sms()) g = 4 * num_sms();
    nparts = (int)g;
    bn_pool_bwd_colsum_kernel<<<nparts, BN_THREADS, 0, st>>>(G, k, c, y, d_a, arg, stats, partial);
  }
  bn_bwd_finalize_kernel<<<ceil_div(c, FIN_CH), dim3(FIN_CH, FIN_SLICES), 0, st>>>(r, c, nparts, partial, d_gamma, d_beta, coef);
  const long long n4 = r * (c >> 2);
  if (k == 0)
    bn_relu_bwd_apply_kernel<false><<<grid_for(n4), BN_THREADS, 0, st>>>(r, 1, c, y, d_a, nullptr,
                                                                        stats, coef, d_y);
  else
    bn_relu_bwd_apply_kernel<true><<<grid_for(n4), BN_THREADS, 0, st>>>(r, k, c, y, d_a, arg, stats,
                                                                       coef, d_y);
  return check_launch("nesie_bn_relu_rows_backward");
}
