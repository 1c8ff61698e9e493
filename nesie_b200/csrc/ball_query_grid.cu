// Ball query through a uniform grid, sm_100a -- the HBM-bound formulation for large scenes.
//
// Same contract as nesie_ball_query / the reference kernel (ops/ball_query/src/
// ball_query_cuda.cu:11-54): per centre the first `nsample` points IN ASCENDING INDEX ORDER with
// d2 == 0 || (min_r^2 <= d2 < max_r^2), d2 = fma(dz,dz,fma(dx,dx,dy*dy)) (centre - point), short
// rows padded with the first hit, empty rows zero.  The brute-force kernel tests every
// (centre, point) pair: 655 M distance tests at SA1 (batch 8), fp32-issue-bound.  Here each scene
// is binned once into cells of edge >= max_radius and a centre only tests the points of its 27
// neighbouring cells (~100-200 candidates instead of 40000); because the grid visits points out
// of index order, the hits are then ranked by index inside the warp.  The hit SET is identical to
// brute force by construction (same distance expression on the same coordinates; cells are
// 1e-4 wider than the radius and points/centres use the same monotone fp32 cell function), and the
// rank step restores the reference's order, so the output is bit-identical.
//
//   bq_grid_build : one CTA per scene; bounding box -> cell size -> histogram -> exclusive scan ->
//                   fill, all in shared memory (<= 32768 cells); writes the cell offsets and the
//                   cell-sorted points as float4 (x, y, z, index) to the workspace
//   bq_grid_query : one warp per centre; 9 contiguous runs of sorted points (x-neighbours are
//                   adjacent cells), coalesced float4 loads, ballot-compacted hit list in shared
//                   memory, rank-by-counting, one coalesced row store.  A centre with more than
//                   HMAX hits (pathological duplicates) falls back to the ordered scan, which in
//                   that regime exits after a few hundred points.
#include <math_constants.h>

#include "common.cuh"

namespace nesie {
namespace {

constexpr int NCMAX = 32768;      // cells per scene (128 KB histogram in shared memory)
constexpr int MAXDIM = 256;       // cells per axis
constexpr int BUILD_THREADS = 1024;
constexpr int Q_THREADS = 256;
constexpr int Q_WARPS = Q_THREADS / 32;
constexpr int HMAX = 512;         // hits kept per centre before the ordered-scan fallback

struct GridInfo {
  float ox, oy, oz, inv_cs;
  int nx, ny, nz, ncell;
};

__device__ __forceinline__ int cell_coord(float x, float o, float inv_cs, int dim) {
  const float u = (x - o) * inv_cs;
  // fmaxf/fminf drop NaN; the clamp keeps far-away centres on the border cell
  return (int)fminf(fmaxf(floorf(u), 0.f), (float)(dim - 1));
}

// workspace per scene: [GridInfo (32 B)] [cell_start int32 x (NCMAX + 1), padded] [sorted float4 x n]
__host__ __device__ inline size_t ws_cells_off() { return 32; }
__host__ __device__ inline size_t ws_sorted_off() { return 32 + (size_t)(NCMAX + 4) * 4; }
__host__ __device__ inline size_t ws_scene_bytes(int n) { return ws_sorted_off() + (size_t)n * 16; }

__global__ void __launch_bounds__(BUILD_THREADS) bq_grid_build(int n, float cell_min,
                                                               const float *__restrict__ xyz,
                                                               unsigned char *__restrict__ ws) {
  extern __shared__ int s_hist[];  // [NCMAX]
  __shared__ float s_red[6][32];
  __shared__ GridInfo s_g;
  __shared__ int s_part[BUILD_THREADS / 32];
  const int scene = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)scene * n * 3;
  ws += (size_t)scene * ws_scene_bytes(n);
  GridInfo *ginfo = reinterpret_cast<GridInfo *>(ws);
  int *cell_start = reinterpret_cast<int *>(ws + ws_cells_off());
  float4 *sorted = reinterpret_cast<float4 *>(ws + ws_sorted_off());

  // ---- bounding box ---------------------------------------------------------------------------
  float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
  float hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int k = tid; k < n; k += BUILD_THREADS) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = xyz[k * 3 + a];
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { s_red[a][warp] = lo[a]; s_red[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    float l[3], h[3];
    for (int a = 0; a < 3; ++a) {
      l[a] = CUDART_INF_F; h[a] = -CUDART_INF_F;
      for (int w = 0; w < BUILD_THREADS / 32; ++w) {
        l[a] = fminf(l[a], s_red[a][w]);
        h[a] = fmaxf(h[a], s_red[3 + a][w]);
      }
      if (!(h[a] >= l[a]) || !isfinite(l[a]) || !isfinite(h[a])) { l[a] = 0.f; h[a] = 0.f; }
    }
    float cs = cell_min;
    if (!(cs > 0.f)) {   // nearest-neighbour use: about two points per cell of the bounding box
      const float ex = h[0] - l[0], ey = h[1] - l[1], ez = h[2] - l[2];
      const float emax = fmaxf(ex, fmaxf(ey, ez));
      cs = 1.25f * cbrtf(fmaxf(ex, emax * 0.05f) * fmaxf(ey, emax * 0.05f) * fmaxf(ez, emax * 0.05f) /
                         (float)max(n, 1));
      cs = fmaxf(cs, fmaxf(emax * (1.0f / 200.f), 1e-6f));
    }
    int nx, ny, nz;
    while (true) {
      nx = min(MAXDIM, (int)((h[0] - l[0]) / cs) + 1);
      ny = min(MAXDIM, (int)((h[1] - l[1]) / cs) + 1);
      nz = min(MAXDIM, (int)((h[2] - l[2]) / cs) + 1);
      // an axis clipped at MAXDIM needs a larger cell as well (the clamp would fold cells)
      const bool clipped = (h[0] - l[0]) / cs >= MAXDIM || (h[1] - l[1]) / cs >= MAXDIM ||
                           (h[2] - l[2]) / cs >= MAXDIM;
      if (!clipped && (long long)nx * ny * nz <= NCMAX) break;
      cs *= 1.25f;
    }
    s_g.ox = l[0]; s_g.oy = l[1]; s_g.oz = l[2];
    s_g.inv_cs = 1.0f / cs;
    s_g.nx = nx; s_g.ny = ny; s_g.nz = nz; s_g.ncell = nx * ny * nz;
    *ginfo = s_g;
  }
  __syncthreads();
  const GridInfo g = s_g;
  // ---- histogram ------------------------------------------------------------------------------
  for (int c = tid; c < g.ncell; c += BUILD_THREADS) s_hist[c] = 0;
  __syncthreads();
  for (int k = tid; k < n; k += BUILD_THREADS) {
    const int c = (cell_coord(xyz[k * 3 + 2], g.oz, g.inv_cs, g.nz) * g.ny +
                   cell_coord(xyz[k * 3 + 1], g.oy, g.inv_cs, g.ny)) * g.nx +
                  cell_coord(xyz[k * 3 + 0], g.ox, g.inv_cs, g.nx);
    atomicAdd(&s_hist[c], 1);
  }
  __syncthreads();
  // ---- exclusive scan (contiguous chunk per thread + block scan of the chunk sums) -------------
  const int per = (g.ncell + BUILD_THREADS - 1) / BUILD_THREADS;
  const int beg = min(g.ncell, tid * per), end = min(g.ncell, beg + per);
  int sum = 0;
  for (int c = beg; c < end; ++c) sum += s_hist[c];
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_part[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    s_part[lane] = v;  // inclusive over warps
  }
  __syncthreads();
  int run = incl - sum + (warp ? s_part[warp - 1] : 0);  // exclusive prefix of this chunk
  for (int c = beg; c < end; ++c) {
    const int cnt = s_hist[c];
    s_hist[c] = run;
    cell_start[c] = run;
    run += cnt;
  }
  if (tid == 0) cell_start[g.ncell] = n;
  __syncthreads();
  // ---- fill -----------------------------------------------------------------------------------
  for (int k = tid; k < n; k += BUILD_THREADS) {
    const float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
    const int c = (cell_coord(z, g.oz, g.inv_cs, g.nz) * g.ny +
                   cell_coord(y, g.oy, g.inv_cs, g.ny)) * g.nx +
                  cell_coord(x, g.ox, g.inv_cs, g.nx);
    const int pos = atomicAdd(&s_hist[c], 1);
    sorted[pos] = make_float4(x, y, z, __int_as_float(k));
  }
}

// ---------------------------------------------------------------------------------------------
// three_nn over the same grid (the SidePooling grids ask for 650 k targets x 1024 seeds per step;
// brute force is 0.7 G distance evaluations).  One thread per target walks the shells of cells around
// its own cell (x-neighbours are contiguous runs of the sorted points) and keeps the three smallest
// (d2, index) pairs in lexicographic order -- exactly the brute-force kernel's result, whose strict '<'
// cascade over ascending indices keeps the earliest index among equal distances.  After shell rho every
// unvisited point lies beyond one of the block's interior faces, i.e. at least `gap` away along that
// axis; the walk stops once the third-best d2 is below gap^2 (with a relative margin for the rounding of
// the cell function and of d2), or when the block covers the grid.  Same d2 expression (sqdist_ref) as
// three_nn_kernel, so distances and indices are bit-identical.
constexpr int NNG_THREADS = 128;

__device__ __forceinline__ void nn3_insert(float d, int id, float &b1, float &b2, float &b3, int &i1,
                                           int &i2, int &i3) {
  if (d < b3 || (d == b3 && id < i3)) {
    if (d < b1 || (d == b1 && id < i1)) {
      b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = id;
    } else if (d < b2 || (d == b2 && id < i2)) {
      b3 = b2; i3 = i2; b2 = d; i2 = id;
    } else {
      b3 = d; i3 = id;
    }
  }
}

__global__ void __launch_bounds__(NNG_THREADS) three_nn_grid_kernel(
    int n, int m, const float *__restrict__ unknown, const unsigned char *__restrict__ ws,
    float *__restrict__ dist2, int *__restrict__ idx) {
  const int scene = blockIdx.y;
  const int pt = blockIdx.x * NNG_THREADS + threadIdx.x;
  if (pt >= n) return;
  ws += (size_t)scene * ws_scene_bytes(m);
  const GridInfo g = *reinterpret_cast<const GridInfo *>(ws);
  const int *cell_start = reinterpret_cast<const int *>(ws + ws_cells_off());
  const float4 *sorted = reinterpret_cast<const float4 *>(ws + ws_sorted_off());
  const float *u = unknown + ((size_t)scene * n + pt) * 3;
  const float ux = u[0], uy = u[1], uz = u[2];
  const int cx = cell_coord(ux, g.ox, g.inv_cs, g.nx), cy = cell_coord(uy, g.oy, g.inv_cs, g.ny),
            cz = cell_coord(uz, g.oz, g.inv_cs, g.nz);
  const float cs = 1.0f / g.inv_cs;
  float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int rho = 0;; ++rho) {
    const int z0 = max(cz - rho, 0), z1 = min(cz + rho, g.nz - 1);
    const int y0 = max(cy - rho, 0), y1 = min(cy + rho, g.ny - 1);
    const int x0 = max(cx - rho, 0), x1 = min(cx + rho, g.nx - 1);
    for (int z = z0; z <= z1; ++z) {
      for (int y = y0; y <= y1; ++y) {
        const int rowc = (z * g.ny + y) * g.nx;
        const bool shell = (z - cz == rho) || (cz - z == rho) || (y - cy == rho) || (cy - y == rho);
        // shell rows: the whole x range; interior rows: only the two end cells (when they exist)
        for (int part = 0; part < 2; ++part) {
          int xa, xb;
          if (shell) {
            if (part) break;
            xa = x0; xb = x1;
          } else {
            if (rho == 0) break;
            xa = xb = part ? cx + rho : cx - rho;
            if (xa < 0 || xa >= g.nx) continue;
          }
          const int beg = __ldg(cell_start + rowc + xa), end = __ldg(cell_start + rowc + xb + 1);
          for (int k = beg; k < end; ++k) {
            const float4 sp = __ldg(sorted + k);
            nn3_insert(sqdist_ref(ux, uy, uz, sp.x, sp.y, sp.z), __float_as_int(sp.w), b1, b2, b3, i1, i2, i3);
          }
        }
      }
    }
    // distance to the nearest interior face of the visited block
    float gap = CUDART_INF_F;
    if (cx - rho > 0) gap = fminf(gap, ux - (g.ox + (float)(cx - rho) * cs));
    if (cx + rho < g.nx - 1) gap = fminf(gap, (g.ox + (float)(cx + rho + 1) * cs) - ux);
    if (cy - rho > 0) gap = fminf(gap, uy - (g.oy + (float)(cy - rho) * cs));
    if (cy + rho < g.ny - 1) gap = fminf(gap, (g.oy + (float)(cy + rho + 1) * cs) - uy);
    if (cz - rho > 0) gap = fminf(gap, uz - (g.oz + (float)(cz - rho) * cs));
    if (cz + rho < g.nz - 1) gap = fminf(gap, (g.oz + (float)(cz + rho + 1) * cs) - uz);
    if (gap == CUDART_INF_F) break;                       // the block covers the whole grid
    gap -= 1e-4f * cs;                                    // rounding of the cell function
    if (gap > 0.f && b3 < gap * gap * 0.9999f) break;     // nothing unvisited can enter the top three
  }
  float *od = dist2 + ((size_t)scene * n + pt) * 3;
  int *oi = idx + ((size_t)scene * n + pt) * 3;
  od[0] = b1; od[1] = b2; od[2] = b3;
  oi[0] = i1; oi[1] = i2; oi[2] = i3;
}

__global__ void __launch_bounds__(Q_THREADS) bq_grid_query(
    int n, int m, float min_r2, float max_r2, int nsample, const float *__restrict__ new_xyz,
    const float *__restrict__ xyz, const unsigned char *__restrict__ ws, int *__restrict__ idx) {
  extern __shared__ int s_dyn[];  // [Q_WARPS][HMAX] hits, then [Q_WARPS][nsample] rows
  const int scene = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci = blockIdx.x * Q_WARPS + warp;
  if (ci >= m) return;  // whole warp
  int *hits = s_dyn + warp * HMAX;
  int *rowbuf = s_dyn + Q_WARPS * HMAX + warp * nsample;
  ws += (size_t)scene * ws_scene_bytes(n);
  const GridInfo g = *reinterpret_cast<const GridInfo *>(ws);
  const int *cell_start = reinterpret_cast<const int *>(ws + ws_cells_off());
  const float4 *sorted = reinterpret_cast<const float4 *>(ws + ws_sorted_off());
  xyz += (size_t)scene * n * 3;
  const float *c = new_xyz + ((size_t)scene * m + ci) * 3;
  const float cx = c[0], cy = c[1], cz = c[2];
  int *out = idx + ((size_t)scene * m + ci) * nsample;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int gx = cell_coord(cx, g.ox, g.inv_cs, g.nx);
  const int gy = cell_coord(cy, g.oy, g.inv_cs, g.ny);
  const int gz = cell_coord(cz, g.oz, g.inv_cs, g.nz);
  const int x0 = max(gx - 1, 0), x1 = min(gx + 1, g.nx - 1);
  int cnt = 0;
  for (int z = max(gz - 1, 0); z <= min(gz + 1, g.nz - 1); ++z) {
    for (int y = max(gy - 1, 0); y <= min(gy + 1, g.ny - 1); ++y) {
      const int row = (z * g.ny + y) * g.nx;
      const int s = cell_start[row + x0], e = cell_start[row + x1 + 1];
      for (int j0 = s; j0 < e; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        int k = 0;
        if (j < e) {
          const float4 p = __ldg(sorted + j);
          const float d2 = sqdist_ref(cx, cy, cz, p.x, p.y, p.z);
          hit = d2 == 0.f || (d2 >= min_r2 && d2 < max_r2);
          k = __float_as_int(p.w);
        }
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (b) {
          const int pos = cnt + __popc(b & lt_mask);
          if (hit && pos < HMAX) hits[pos] = k;
          cnt += __popc(b);
        }
      }
    }
  }
  __syncwarp();
  if (cnt > HMAX) {
    // pathological density: ordered scan over the original points (exits after ~nsample hits)
    int have = 0, first = 0;
    for (int k0 = 0; k0 < n && have < nsample; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < n) {
        const float d2 = sqdist_ref(cx, cy, cz, xyz[k * 3 + 0], xyz[k * 3 + 1], xyz[k * 3 + 2]);
        hit = d2 == 0.f || (d2 >= min_r2 && d2 < max_r2);
      }
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (b) {
        const int pos = have + __popc(b & lt_mask);
        if (hit && pos < nsample) rowbuf[pos] = k;
        if (have == 0) first = k0 + __ffs(b) - 1;
        have += __popc(b);
      }
    }
    __syncwarp();
    have = min(have, nsample);
    for (int l = lane; l < nsample; l += 32) out[l] = l < have ? rowbuf[l] : first;
    return;
  }
  // rank by counting: hit indices are distinct, rank = number of smaller indices
  int mn = 0x7fffffff;
  for (int i = lane; i < cnt; i += 32) {
    const int h = hits[i];
    int rank = 0;
    for (int j = 0; j < cnt; ++j) rank += hits[j] < h;
    if (rank < nsample) rowbuf[rank] = h;
    mn = min(mn, h);
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  __syncwarp();
  const int have = min(cnt, nsample);
  const int pad = cnt ? mn : 0;
  for (int l = lane; l < nsample; l += 32) out[l] = l < have ? rowbuf[l] : pad;
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" long long nesie_ball_query_grid_workspace(int b, int n, int m) {
  (void)m;
  if (b <= 0 || n <= 0) return 0;
  return (long long)b * (long long)ws_scene_bytes(n);
}

extern "C" int nesie_ball_query_grid(int b, int n, int m, float min_radius, float max_radius,
                                     int nsample, const float *new_xyz, const float *xyz, int *idx,
                                     void *workspace, long long workspace_bytes, void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, "negative size");
  NESIE_REQUIRE(new_xyz && xyz && idx, "null pointer");
  if (b == 0 || m == 0 || nsample == 0) return NESIE_OK;
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  NESIE_REQUIRE(n >= 1 && max_radius > 0.f && max_radius < 1e18f,
                "the grid variant needs n >= 1 and a finite positive max_radius");
  NESIE_REQUIRE(nsample <= 1024, "nsample > 1024");
  NESIE_REQUIRE(workspace && workspace_bytes >= nesie_ball_query_grid_workspace(b, n, m),
                "workspace too small (see nesie_ball_query_grid_workspace)");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const float max_r2 = max_radius * max_radius;  // fp32, ball_query_cuda.cu:30-31
  const float min_r2 = min_radius * min_radius;
  const float cell_min = max_radius * 1.0001f + 1e-30f;
  static bool attr_set = false;
  if (!attr_set) {
    NESIE_CUDA(cudaFuncSetAttribute(bq_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    NCMAX * 4));
    attr_set = true;
  }
  bq_grid_build<<<b, BUILD_THREADS, NCMAX * 4, st>>>(n, cell_min, xyz,
                                                     reinterpret_cast<unsigned char *>(workspace));
  int rc = check_launch("nesie_ball_query_grid(build)");
  if (rc) return rc;
  const size_t smem = (size_t)Q_WARPS * (HMAX + nsample) * sizeof(int);
  dim3 grid(ceil_div(m, Q_WARPS), b);
  bq_grid_query<<<grid, Q_THREADS, smem, st>>>(n, m, min_r2, max_r2, nsample, new_xyz, xyz,
                                               reinterpret_cast<const unsigned char *>(workspace),
                                               idx);
  return check_launch("nesie_ball_query_grid(query)");
}

// three_nn through the grid: `workspace` as for nesie_ball_query_grid with n := m (the SOURCE count;
// nesie_ball_query_grid_workspace(b, m, 0) bytes).  build != 0 bins the sources first; pass 0 to reuse
// the grid of a previous call over the same sources (stream-ordered).  Results are bit-identical to
// nesie_three_nn.
extern "C" int nesie_three_nn_grid(int b, int n, int m, const float *unknown, const float *known,
                                   float *dist2, int *idx, void *workspace, long long workspace_bytes,
                                   int build, void *stream) {
  NESIE_REQUIRE(b >= 0 && n >= 0 && m >= 1, "need b >= 0, n >= 0, m >= 1");
  if (b == 0 || n == 0) return NESIE_OK;
  NESIE_REQUIRE(unknown && known && dist2 && idx, "null pointer");
  NESIE_REQUIRE(b <= 65535, "b > 65535");
  NESIE_REQUIRE(workspace && workspace_bytes >= nesie_ball_query_grid_workspace(b, m, 0),
                "workspace too small (see nesie_ball_query_grid_workspace)");
  NESIE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (build) {
    static bool attr_set = false;
    if (!attr_set) {
      NESIE_CUDA(cudaFuncSetAttribute(bq_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, NCMAX * 4));
      attr_set = true;
    }
    bq_grid_build<<<b, BUILD_THREADS, NCMAX * 4, st>>>(m, 0.f, known, reinterpret_cast<unsigned char *>(workspace));
    const int rc = check_launch("nesie_three_nn_grid(build)");
    if (rc) return rc;
  }
  three_nn_grid_kernel<<<dim3(ceil_div(n, NNG_THREADS), b), NNG_THREADS, 0, st>>>(
      n, m, unknown, reinterpret_cast<const unsigned char *>(workspace), dist2, idx);
  return check_launch("nesie_three_nn_grid");
}
