// Small row-group kernels around the pooled layers of the SidePooling MiniPointNets (reference:
// models/dense_heads/side_pooling_module.py:343-370).  With the maximum over a box's grid points taken
// in the GEMM's epilogue (gemm_tma.cuh, pool_max / pool_amax) the pooled convolution's output never
// exists in HBM; what is left is
//
//   pool_finalize       unit maxima of the epilogue (16 / 32 rows) -> group maxima (k rows) + conv bias
//   pool_wgrad          d_w[c, :] = sum_g d_out[g, c] * a[g * k + arg[g, c], :]  -- the weight gradient
//                       of a max-pooled convolution touches ONE input row per (group, channel): a gather
//                       of groups x n rows instead of a rows x n x K contraction over a dense d_y
//   pool_dgrad          d_a[g k + j, :] = sum_{c: arg[g, c] = j} d_out[g, c] * w[c, :]  -- likewise n rows of
//                       w per group instead of a product with a dense d_y
//   group_sum_rows      sum of every k consecutive rows (the gradient of a per-group input that was
//                       multiplied through the weights once per group, GemmTmaParams::grp_bias)
//   scatter_rows_add    d_x[g * k + arg[g, c], c] += d[g, c]  (the max's gradient, added in place)
#include "common.cuh"

namespace nesie {
namespace {

constexpr int PR_THREADS = 256;

__global__ void __launch_bounds__(PR_THREADS) pool_finalize_kernel(
    long long groups, int m, int u, int n, const float *__restrict__ pmax,
    const unsigned char *__restrict__ amax, const float *__restrict__ bias, float *__restrict__ out,
    unsigned char *__restrict__ arg) {
  const int tpr = n >> 2;
  const long long total = groups * tpr;
  for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * PR_THREADS) {
    const long long g = i / tpr;
    const int ch = (int)(i - g * tpr) * 4;
    const long long base = (g * m) * n + ch;
    float4 best = *reinterpret_cast<const float4 *>(pmax + base);
    uchar4 a = *reinterpret_cast<const uchar4 *>(amax + base);
    int bx = a.x, by = a.y, bz = a.z, bw = a.w;
    for (int j = 1; j < m; ++j) {   // strict comparisons: the first maximising row wins
      const float4 v = *reinterpret_cast<const float4 *>(pmax + base + (long long)j * n);
      a = *reinterpret_cast<const uchar4 *>(amax + base + (long long)j * n);
      if (v.x > best.x) { best.x = v.x; bx = j * u + a.x; }
      if (v.y > best.y) { best.y = v.y; by = j * u + a.y; }
      if (v.z > best.z) { best.z = v.z; bz = j * u + a.z; }
      if (v.w > best.w) { best.w = v.w; bw = j * u + a.w; }
    }
    if (bias) {
      const float4 b = *reinterpret_cast<const float4 *>(bias + ch);
      best.x += b.x; best.y += b.y; best.z += b.z; best.w += b.w;
    }
    *reinterpret_cast<float4 *>(out + g * n + ch) = best;
    *reinterpret_cast<uchar4 *>(arg + g * n + ch) =
        make_uchar4((unsigned char)bx, (unsigned char)by, (unsigned char)bz, (unsigned char)bw);
  }
}

// BatchNorm + ReLU + max-pool of an SA level's last layer from the unit extrema of the GEMM epilogue:
// relu(y * scale + shift) is monotone in y, increasing for scale >= 0 and decreasing otherwise, so the
// pooled value is relu(scale * (max or min over the group) + shift) and the pre-activation is not read
// again.  arg = first row of the group that attains it (255 when the maximum is not positive: the ReLU
// passes no gradient), as bn_relu_pool_kernel reports it.
__global__ void __launch_bounds__(PR_THREADS) bn_pool_finalize_kernel(
    long long groups, int m, int u, int n, const float *__restrict__ pmax,
    const unsigned char *__restrict__ amax, const float *__restrict__ pmin,
    const unsigned char *__restrict__ amin, const float *__restrict__ stats, float *__restrict__ out,
    unsigned char *__restrict__ arg) {
  const float *scale = stats + 2 * n, *shift = stats + 3 * n;
  const long long total = groups * n;
  for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * PR_THREADS) {
    const long long g = i / n;
    const int c = (int)(i - g * n);
    const float sc = scale[c], sh = shift[c];
    const long long base = (g * m) * n + c;
    float best;
    int bi;
    if (sc >= 0.f) {
      best = pmax[base]; bi = amax[base];
      for (int j = 1; j < m; ++j) {
        const float v = pmax[base + (long long)j * n];
        if (v > best) { best = v; bi = j * u + amax[base + (long long)j * n]; }
      }
      if (sc == 0.f) bi = 0;            // every row ties
    } else {
      best = pmin[base]; bi = amin[base];
      for (int j = 1; j < m; ++j) {
        const float v = pmin[base + (long long)j * n];
        if (v < best) { best = v; bi = j * u + amin[base + (long long)j * n]; }
      }
    }
    const float z = fmaf(best, sc, sh);
    out[i] = fmaxf(z, 0.f);
    arg[i] = z > 0.f ? (unsigned char)bi : (unsigned char)255;
  }
}

// One CTA = all n output channels x a slice of 64 input columns (blockIdx.y); thread t owns the 32
// channels of block t >> 4 and the four columns (t & 15) * 4 of the slice, 32 x 4 sums in registers.  It
// walks groups blockIdx.x, blockIdx.x + gridDim.x, ... and leaves ONE partial block per blockIdx.x.
// The slice of a group's k rows is staged in shared memory (with the previous layer's BatchNorm + ReLU
// applied when scale / shift are given: the operand is relu(y_prev * scale + shift), never stored) --
// every input element is read once -- and the next group's slice is already in flight in registers
// while the current one is used: a (group, channel) pair costs one 16-byte shared-memory read and four
// FMAs per thread.
constexpr int PW_CH = 32;     // channels per thread
constexpr int PW_COLS = 64;   // input columns per CTA
constexpr int PW_MAXT = 128;  // threads: 16 per 32-channel block (n <= 256)

template <int IT, int D>      // staged 16-byte items per thread (k * 16 / blockDim.x; 0: any k, no prefetch), ring depth
__global__ void __launch_bounds__(PW_MAXT) pool_wgrad_kernel(
    long long groups, int k, int n, int kk, const float *__restrict__ d_out,
    const unsigned char *__restrict__ arg, const float *__restrict__ y_prev,
    const float *__restrict__ scale, const float *__restrict__ shift, float *__restrict__ dw_part) {
  extern __shared__ __align__(16) unsigned char pw_smem[];
  float4 *s_a = reinterpret_cast<float4 *>(pw_smem);          // [k][16]
  __shared__ float s_d[8 * PW_CH];
  __shared__ int s_j[8 * PW_CH];
  const int t = threadIdx.x, nt = blockDim.x;
  const int col0 = blockIdx.y * PW_COLS;
  const int cb = t >> 4, q = t & 15;
  const int c0 = cb * PW_CH;
  const int mycol = col0 + 4 * q;
  const bool own = mycol < kk;
  float4 acc[PW_CH];
#pragma unroll
  for (int i = 0; i < PW_CH; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int items = k * 16;
  // staged item i: row i >> 4, 16-byte column chunk i & 15 of the slice; blockDim.x is a multiple of 16,
  // so a thread always stages chunk q and keeps that chunk's scale / shift in registers.  Raw values
  // are fetched (nothing depends on them until they are stored: the loads of a whole group overlap)
  // and the BatchNorm + ReLU is applied on the way into shared memory.
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale && own) {
    sc = __ldg(reinterpret_cast<const float4 *>(scale + mycol));
    sh = __ldg(reinterpret_cast<const float4 *>(shift + mycol));
  }
  const float *ybase = y_prev + mycol;
  auto fetch = [&](long long g, int i) -> float4 {
    return own ? __ldg(reinterpret_cast<const float4 *>(ybase + (g * k + (i >> 4)) * (long long)kk))
               : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto act = [&](float4 v) -> float4 {
    if (scale) {
      v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
      v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
    }
    return v;
  };
  // ring of D groups in flight in registers (slot s holds group blockIdx.x + (s + D * round) * gridDim.x)
  float4 pre[D][IT > 0 ? IT : 1];
  float pd[D][2];
  int pj[D][2];
  auto prefetch = [&](int s, long long g) {
    if (g >= groups) return;
    if (IT > 0) {
#pragma unroll
      for (int i = 0; i < IT; ++i) pre[s][i] = fetch(g, t + i * nt);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = t + h * nt;
      const bool ok = c < n;
      pd[s][h] = ok ? __ldg(d_out + g * n + c) : 0.f;
      pj[s][h] = ok ? (int)__ldg(arg + g * n + c) : 0;
    }
  };
#pragma unroll
  for (int s = 0; s < D; ++s) prefetch(s, (long long)blockIdx.x + (long long)s * gridDim.x);
  for (long long gb = blockIdx.x; gb < groups; gb += (long long)D * gridDim.x) {
#pragma unroll
    for (int s = 0; s < D; ++s) {
      const long long g = gb + (long long)s * gridDim.x;
      if (g < groups) {           // uniform over the CTA
        if (IT > 0) {
#pragma unroll
          for (int i = 0; i < IT; ++i) s_a[t + i * nt] = act(pre[s][i]);
        } else {
          for (int i = t; i < items; i += nt) s_a[i] = act(fetch(g, i));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (t + h * nt < 8 * PW_CH) { s_d[t + h * nt] = pd[s][h]; s_j[t + h * nt] = pj[s][h]; }
        __syncthreads();
        prefetch(s, g + (long long)D * gridDim.x);   // travels while the staged groups are used
        if (own) {
#pragma unroll
          for (int ci = 0; ci < PW_CH; ++ci) {
            const float d = s_d[c0 + ci];
            const float4 a = s_a[s_j[c0 + ci] * 16 + q];
            acc[ci].x = fmaf(d, a.x, acc[ci].x); acc[ci].y = fmaf(d, a.y, acc[ci].y);
            acc[ci].z = fmaf(d, a.z, acc[ci].z); acc[ci].w = fmaf(d, a.w, acc[ci].w);
          }
        }
        __syncthreads();
      }
    }
  }
  if (own) {
    float *dst = dw_part + ((size_t)blockIdx.x * n + c0) * kk + mycol;
#pragma unroll
    for (int ci = 0; ci < PW_CH; ++ci)
      if (c0 + ci < n) *reinterpret_cast<float4 *>(dst + (size_t)ci * kk) = acc[ci];
  }
}

// Data gradient of the same layer: d_a[g k + j, :] = sum over the channels c whose maximum sits in row
// j of d_out[g, c] * w[c, :] -- n rows of w per group instead of a (k x n) x (n x K) product.  w lives
// in shared memory (its rows are read groups x n times: from L2 that alone is 0.5 GB per launch at 4096
// boxes).  Per group the channels are counting-sorted by row (rank within a warp from match.any, warps
// in order: ascending channel within a row, so the sums are deterministic) and every row walks its own
// short list.  One persistent CTA per SM; thread = 4 columns x one of 8 row lanes; the next group's
// d_out / arg are in flight while the current group is processed.
constexpr int PD_THREADS = 512;

__global__ void __launch_bounds__(PD_THREADS) pool_dgrad_kernel(
    long long groups, int k, int n, int kk, const float *__restrict__ d_out,
    const unsigned char *__restrict__ arg, const float *__restrict__ w, float *__restrict__ d_a) {
  extern __shared__ __align__(16) unsigned char pd_smem[];
  float4 *s_w = reinterpret_cast<float4 *>(pd_smem);            // [n][kk / 4]
  __shared__ int s_cnt[8][256];     // [warp of channels][row]: channels of that warp in that row
  __shared__ int s_start[257];      // first list position of a row
  __shared__ int s_c[256];          // channels sorted by (row, channel)
  __shared__ float s_dv[256];       // their d_out
  const int t = threadIdx.x, q = t & 63, rl = t >> 6, lane = t & 31, wp = t >> 5;
  const int nwarp = (n + 31) >> 5, k4 = kk >> 2;
  for (int i = t; i < n * k4; i += PD_THREADS) s_w[i] = __ldg(reinterpret_cast<const float4 *>(w) + i);
  for (int i = t; i < 8 * 256; i += PD_THREADS) (&s_cnt[0][0])[i] = 0;
  long long g = blockIdx.x;
  float nd = 0.f;
  int na = 0;
  if (g < groups && t < n) { nd = __ldg(d_out + g * n + t); na = (int)__ldg(arg + g * n + t); }
  __syncthreads();
  for (; g < groups; g += gridDim.x) {
    const float dcur = nd;
    const int jcur = na;
    int rank = 0;
    if (wp < nwarp) {   // whole warps; lanes beyond n take part with a row no channel has
      const int key = t < n ? jcur : 256 + lane;
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      rank = __popc(peers & ((1u << lane) - 1u));
      if (t < n && rank == 0) s_cnt[wp][jcur] = __popc(peers);
    }
    __syncthreads();
    if (t < 32) {   // one warp, rows 2 lane and 2 lane + 1 (k <= 64): exclusive prefix over rows, then warps
      int tot[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * lane + h;
        tot[h] = 0;
        if (j < k)
          for (int ww = 0; ww < nwarp; ++ww) tot[h] += s_cnt[ww][j];
      }
      int incl = tot[0] + tot[1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int before = incl - tot[0] - tot[1];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * lane + h;
        if (j < k) {
          s_start[j] = before;
          for (int ww = 0; ww < nwarp; ++ww) {
            const int c = s_cnt[ww][j];
            s_cnt[ww][j] = before;       // becomes the base position of (warp, row)
            before += c;
          }
        }
      }
      if (lane == 0) s_start[k] = n;
    }
    __syncthreads();
    if (t < n) {
      const int pos = s_cnt[wp][jcur] + rank;
      s_c[pos] = t;
      s_dv[pos] = dcur;
    }
    const long long gn = g + gridDim.x;
    if (gn < groups && t < n) { nd = __ldg(d_out + gn * n + t); na = (int)__ldg(arg + gn * n + t); }
    __syncthreads();
    if (t < k)
      for (int ww = 0; ww < nwarp; ++ww) s_cnt[ww][t] = 0;   // for the next group (read again after 2 barriers)
    for (int j = rl; j < k; j += 8) {
      const int i0 = s_start[j], i1 = s_start[j + 1];
      for (int c4 = q; c4 < k4; c4 += 64) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = i0; i < i1; i += 4) {   // four list entries in flight (a single one is a chain of
          float d[4];                        // dependent shared-memory round trips); order kept
          float4 wv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ii = min(i + u, i1 - 1);
            d[u] = i + u < i1 ? s_dv[ii] : 0.f;
            wv[u] = s_w[s_c[ii] * k4 + c4];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc.x = fmaf(d[u], wv[u].x, acc.x); acc.y = fmaf(d[u], wv[u].y, acc.y);
            acc.z = fmaf(d[u], wv[u].z, acc.z); acc.w = fmaf(d[u], wv[u].w, acc.w);
          }
        }
        reinterpret_cast<float4 *>(d_a + (g * k + j) * (long long)kk)[c4] = acc;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PR_THREADS) group_sum_rows_kernel(
    long long groups, int k, int n, const float *__restrict__ x, float *__restrict__ out) {
  const int tpr = n >> 2;
  const long long total = groups * tpr;
  for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * PR_THREADS) {
    const long long g = i / tpr;
    const int ch = (int)(i - g * tpr) * 4;
    const float *row = x + (g * k) * n + ch;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < k; ++r) {
      const float4 v = *reinterpret_cast<const float4 *>(row + (long long)r * n);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4 *>(out + g * n + ch) = s;
  }
}

// Column sums of a (rows, n) matrix in ONE launch and a fixed summation order: CTA b sums rows
// b, b + gridDim.x, ... (thread = 4 columns x one of 256 / (n / 4) row lanes, lanes combined through
// shared memory in lane order), writes its partial row, and the CTA that finishes last adds the
// partial rows in a fixed order.  (ATen's reduction of a 4096 x 256 matrix over dim 0 takes 14 us.)
__global__ void __launch_bounds__(PR_THREADS) colsum_rows_kernel(
    long long rows, int n, const float *__restrict__ x, float *__restrict__ part,
    unsigned *__restrict__ counter, float *__restrict__ out) {
  __shared__ float4 s_red[PR_THREADS];
  __shared__ bool s_last;
  const int tpr = n >> 2;                 // threads per row
  const int lanes = PR_THREADS / tpr;     // row lanes per CTA
  const int c4 = threadIdx.x % tpr, lane = threadIdx.x / tpr;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
    const long long step = (long long)gridDim.x * lanes;
    long long r = (long long)blockIdx.x * lanes + lane;
#pragma unroll 8
    for (; r < rows; r += step) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * n) + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  s_red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < tpr) {
    float4 t = s_red[threadIdx.x];
    for (int l = 1; l < lanes; ++l) {
      const float4 u = s_red[l * tpr + threadIdx.x];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    reinterpret_cast<float4 *>(part + (size_t)blockIdx.x * n)[threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // partial rows lane, lane + lanes, ... per row lane (independent loads), then the lanes in order
  s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
#pragma unroll 8
    for (unsigned b = lane; b < gridDim.x; b += lanes) {
      const float4 u = __ldcg(reinterpret_cast<const float4 *>(part + (size_t)b * n) + c4);
      s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w;
    }
  }
  __syncthreads();
  s_red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < tpr) {
    float4 t = s_red[threadIdx.x];
    for (int l = 1; l < lanes; ++l) {
      const float4 u = s_red[l * tpr + threadIdx.x];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    reinterpret_cast<float4 *>(out)[threadIdx.x] = t;
  }
  if (threadIdx.x == 0) *counter = 0;     // ready for the next launch (stream order)
}

__global__ void __launch_bounds__(PR_THREADS) scatter_rows_add_kernel(
    long long groups, int k, int n, const float *__restrict__ d, const unsigned char *__restrict__ arg,
    float *__restrict__ d_x) {
  const long long total = groups * n;
  for (long long i = (long long)blockIdx.x * PR_THREADS + threadIdx.x; i < total;
       i += (long long)gridDim.x * PR_THREADS) {
    const long long g = i / n;
    const int c = (int)(i - g * n);
    d_x[(g * k + arg[i]) * n + c] += d[i];   // (row, channel) is unique per (group, channel): no race
  }
}

int pr_grid(long long items) {
  long long g = (items + PR_THREADS - 1) / PR_THREADS;
  const long long cap = 16LL * num_sms();
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int pw_grid_x(long long groups) {
  const long long cap = num_sms();   // x 4 column slices at k_in = 256: four CTAs per SM
  return (int)(groups < cap ? groups : cap);
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_pool_finalize(long long groups, int k, int u, int n, const float *pmax,
                                   const unsigned char *amax, const float *bias, float *out,
                                   unsigned char *arg, void *stream) {
  NESIE_REQUIRE(groups >= 0 && u >= 1 && k >= u && k % u == 0 && k <= 255 && n >= 4 && (n & 3) == 0,
                "need u | k, k <= 255, n % 4 == 0");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(pmax && amax && out && arg, "null pointer");
  pool_finalize_kernel<<<pr_grid(groups * (n >> 2)), PR_THREADS, 0, (cudaStream_t)stream>>>(
      groups, k / u, u, n, pmax, amax, bias, out, arg);
  return check_launch("nesie_pool_finalize");
}

extern "C" int nesie_bn_pool_finalize(long long groups, int k, int u, int n, const float *pmax,
                                      const unsigned char *amax, const float *pmin,
                                      const unsigned char *amin, const float *stats, float *out,
                                      unsigned char *arg, void *stream) {
  NESIE_REQUIRE(groups >= 0 && u >= 1 && k >= u && k % u == 0 && k <= 254 && n >= 1, "need u | k, k <= 254, n >= 1");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(pmax && amax && pmin && amin && stats && out && arg, "null pointer");
  bn_pool_finalize_kernel<<<pr_grid(groups * n), PR_THREADS, 0, (cudaStream_t)stream>>>(
      groups, k / u, u, n, pmax, amax, pmin, amin, stats, out, arg);
  return check_launch("nesie_bn_pool_finalize");
}

extern "C" int nesie_pool_wgrad_parts(long long groups) { return groups <= 0 ? 0 : pw_grid_x(groups); }

extern "C" int nesie_pool_wgrad(long long groups, int k, int n, int kk, const float *d_out,
                                const unsigned char *arg, const float *y_prev, const float *scale,
                                const float *shift, float *dw_part, void *stream) {
  NESIE_REQUIRE(groups >= 1 && k >= 1 && k <= 255 && n >= 1 && n <= 8 * PW_CH, "need groups >= 1, 1 <= k <= 255, 1 <= n <= 256");
  NESIE_REQUIRE(kk >= 4 && (kk & 3) == 0, "need k_in % 4 == 0");
  NESIE_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  NESIE_REQUIRE(d_out && arg && y_prev && dw_part, "null pointer");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(y_prev) & 15) == 0 && (reinterpret_cast<uintptr_t>(dw_part) & 15) == 0,
                "y_prev and dw_part must be 16-byte aligned");
  const int threads = 16 * ((n + PW_CH - 1) / PW_CH);
  const size_t smem = (size_t)k * PW_COLS * sizeof(float);
  const dim3 grid(pw_grid_x(groups), (kk + PW_COLS - 1) / PW_COLS);
  const int items = k * 16;
  const int it = items % threads == 0 ? items / threads : 0;
#define NESIE_PW_LAUNCH(IT, D)                                                                                 \
  pool_wgrad_kernel<IT, D><<<grid, threads, smem, (cudaStream_t)stream>>>(groups, k, n, kk, d_out, arg, y_prev, \
                                                                           scale, shift, dw_part)
  if (it == 2) NESIE_PW_LAUNCH(2, 4);
  else if (it == 4) NESIE_PW_LAUNCH(4, 3);
  else if (it == 8) NESIE_PW_LAUNCH(8, 2);
  else if (it == 16) NESIE_PW_LAUNCH(16, 1);
  else NESIE_PW_LAUNCH(0, 1);
#undef NESIE_PW_LAUNCH
  return check_launch("nesie_pool_wgrad");
}

extern "C" int nesie_pool_dgrad(long long groups, int k, int n, int kk, const float *d_out,
                                const unsigned char *arg, const float *w, float *d_a, void *stream) {
  NESIE_REQUIRE(groups >= 0 && k >= 1 && k <= 64 && n >= 1 && n <= 256, "need 1 <= k <= 64, 1 <= n <= 256");
  NESIE_REQUIRE(kk >= 4 && (kk & 3) == 0, "need k_in % 4 == 0");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(d_out && arg && w && d_a, "null pointer");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_a) & 15) == 0,
                "w and d_a must be 16-byte aligned");
  const size_t smem = (size_t)n * kk * sizeof(float);
  NESIE_REQUIRE(smem <= 216 * 1024, "the weight matrix does not fit shared memory (n * k_in <= 55296)");
  NESIE_CUDA(cudaFuncSetAttribute(pool_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long cap = num_sms();
  pool_dgrad_kernel<<<(int)(groups < cap ? groups : cap), PD_THREADS, smem, (cudaStream_t)stream>>>(
      groups, k, n, kk, d_out, arg, w, d_a);
  return check_launch("nesie_pool_dgrad");
}

extern "C" int nesie_group_sum_rows(long long groups, int k, int n, const float *x, float *out,
                                    void *stream) {
  NESIE_REQUIRE(groups >= 0 && k >= 1 && n >= 4 && (n & 3) == 0, "need k >= 1, n % 4 == 0");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(x && out, "null pointer");
  group_sum_rows_kernel<<<pr_grid(groups * (n >> 2)), PR_THREADS, 0, (cudaStream_t)stream>>>(groups, k, n, x, out);
  return check_launch("nesie_group_sum_rows");
}

extern "C" int nesie_scatter_rows_add(long long groups, int k, int n, const float *d,
                                      const unsigned char *arg, float *d_x, void *stream) {
  NESIE_REQUIRE(groups >= 0 && k >= 1 && k <= 255 && n >= 1, "need 1 <= k <= 255, n >= 1");
  if (groups == 0) return NESIE_OK;
  NESIE_REQUIRE(d && arg && d_x, "null pointer");
  scatter_rows_add_kernel<<<pr_grid(groups * n), PR_THREADS, 0, (cudaStream_t)stream>>>(groups, k, n, d, arg, d_x);
  return check_launch("nesie_scatter_rows_add");
}

/* work: nesie_colsum_rows_workspace() bytes, zero-initialised ONCE by the caller (the kernel leaves its
 * counter at zero); launches that share a workspace must be stream-ordered. */
extern "C" long long nesie_colsum_rows_workspace(int n) { return 256 + (long long)64 * n * sizeof(float); }

extern "C" int nesie_colsum_rows(long long rows, int n, const float *x, float *out, void *work,
                                 void *stream) {
  NESIE_REQUIRE(rows >= 1 && n >= 4 && (n & 3) == 0 && n <= 4 * PR_THREADS, "need rows >= 1, n % 4 == 0, n <= 1024");
  NESIE_REQUIRE(x && out && work, "null pointer");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(work) & 255) == 0, "alignment");
  const int lanes = PR_THREADS / (n >> 2);
  long long grid = (rows + (long long)lanes * 8 - 1) / ((long long)lanes * 8);   // >= 8 rows per thread
  if (grid > 64) grid = 64;
  if (grid < 1) grid = 1;
  colsum_rows_kernel<<<(int)grid, PR_THREADS, 0, (cudaStream_t)stream>>>(
      rows, n, x, reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(work) + 256),
      reinterpret_cast<unsigned *>(work), out);
  return check_launch("nesie_colsum_rows");
}
