// Per-side uncertainty box-regression loss (forward + backward) and the teacher EMA update.
//
// Loss: replaces the ~15 elementwise / reduction launches plus a rows-long python gather loop of
//   NesieHead.loss            reference: models/dense_heads/nesie_head.py:332-349
//   SurfaceLoss (MSE branch)  reference: models/losses/surface_loss.py:57-61
//   Bbox2Surface              reference: models/losses/surface_loss.py:90-100
// with ONE kernel each way.  Per (row, side) element, in the reference's evaluation order
// (one rounding per python operator, no FMA):
//   tgt   = centre -/+ 0.5 * size                       (Bbox2Surface)
//   L     = loss_weight * (((pred - tgt)^2) * w)        (mmdet MSELoss, reduction 'none')
//   s     = side_scores[row, side, argmax_c sem_scores[row, c]]   (first maximum)
//   sigma = ((0.8 * s) * s - 1.8 * s) + 1
//   out   = exp(-sigma) * L + (alpha * sigma) * w       -> summed over all elements
//
// EMA: replaces the per-tensor python loop of SimiTeacherHook.hooks_after_train_iter
//   (reference: core/utils/simi_teacher_hook.py:54-64): ema.mul_(1-m).add_(param, alpha=m),
//   i.e. fma(m, param, ema * (1-m)) as ATen's CUDA add-with-alpha kernel computes it, over one
//   flat buffer with 16-byte accesses.
#include "common.cuh"

namespace nesie {
namespace {

constexpr int SL_THREADS = 256;

__device__ __forceinline__ int argmax_first(const float *v, int ncls) {
  int best = 0;
  float bv = v[0];
  for (int c = 1; c < ncls; ++c) {
    const float x = v[c];
    if (x > bv || (x != x && bv == bv)) { bv = x; best = c; }  // torch.max treats NaN as max
  }
  return best;
}

struct SideTerms { float diff, L, s, sigma, e, w; int cls; };

__device__ __forceinline__ SideTerms side_terms(int row, int side, int ncls,
                                                const float *surface_pred,
                                                const float *box_targets,
                                                const float *side_scores,
                                                const float *sem_scores, const float *weight,
                                                float loss_weight) {
  SideTerms t;
  const float *bx = box_targets + (size_t)row * 7;
  const int a = side < 3 ? side : side - 3;
  const float half = __fmul_rn(0.5f, bx[3 + a]);
  const float tgt = side < 3 ? __fsub_rn(bx[a], half) : __fadd_rn(bx[a], half);
  t.w = weight[(size_t)row * 6 + side];
  t.diff = __fsub_rn(surface_pred[(size_t)row * 6 + side], tgt);
  t.L = __fmul_rn(loss_weight, __fmul_rn(__fmul_rn(t.diff, t.diff), t.w));
  t.cls = argmax_first(sem_scores + (size_t)row * ncls, ncls);
  t.s = side_scores[((size_t)row * 6 + side) * ncls + t.cls];
  t.sigma = __fadd_rn(__fsub_rn(__fmul_rn(__fmul_rn(0.8f, t.s), t.s), __fmul_rn(1.8f, t.s)), 1.f);
  t.e = expf(-t.sigma);
  return t;
}

__global__ void __launch_bounds__(SL_THREADS) side_loss_fwd_kernel(
    int rows, int ncls, const float *__restrict__ surface_pred,
    const float *__restrict__ box_targets, const float *__restrict__ side_scores,
    const float *__restrict__ sem_scores, const float *__restrict__ weight, float loss_weight,
    float alpha, float *__restrict__ loss_out, float *__restrict__ sigma_out) {
  __shared__ float s_part[SL_THREADS / 32];
  const int total = rows * 6;
  float acc = 0.f;
  for (int e = blockIdx.x * SL_THREADS + threadIdx.x; e < total; e += gridDim.x * SL_THREADS) {
    const int row = e / 6, side = e - row * 6;
    const SideTerms t = side_terms(row, side, ncls, surface_pred, box_targets, side_scores,
                                   sem_scores, weight, loss_weight);
    if (sigma_out) sigma_out[e] = t.sigma;
    acc += __fadd_rn(__fmul_rn(t.e, t.L), __fmul_rn(__fmul_rn(alpha, t.sigma), t.w));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < SL_THREADS / 32 ? s_part[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss_out, v);
  }
}

__global__ void __launch_bounds__(SL_THREADS) side_loss_bwd_kernel(
    int rows, int ncls, const float *__restrict__ surface_pred,
    const float *__restrict__ box_targets, const float *__restrict__ side_scores,
    const float *__restrict__ sem_scores, const float *__restrict__ weight, float loss_weight,
    float alpha, const float *__restrict__ grad_loss, const float *__restrict__ grad_sigma,
    float *__restrict__ grad_surface_pred, float *__restrict__ grad_side_scores) {
  const int e = blockIdx.x * SL_THREADS + threadIdx.x;
  if (e >= rows * 6) return;
  const int row = e / 6, side = e - row * 6;
  const SideTerms t = side_terms(row, side, ncls, surface_pred, box_targets, side_scores,
                                 sem_scores, weight, loss_weight);
  const float g = grad_loss ? grad_loss[0] : 0.f;
  // d out / d pred = e * loss_weight * w * 2 * diff
  if (grad_surface_pred)
    grad_surface_pred[e] = g * t.e * loss_weight * t.w * 2.f * t.diff;
  // d out / d sigma = -e * L + alpha * w ; d sigma / d s = 1.6 s - 1.8
  if (grad_side_scores) {
    float gs = g * (-t.e * t.L + alpha * t.w);
    if (grad_sigma) gs += grad_sigma[e];
    grad_side_scores[((size_t)row * 6 + side) * ncls + t.cls] = gs * (1.6f * t.s - 1.8f);
  }
}

__global__ void __launch_bounds__(256) ema_kernel(long long count, float *__restrict__ ema,
                                                  const float *__restrict__ param, float decay,
                                                  float momentum) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(ema) | reinterpret_cast<uintptr_t>(param)) & 15) == 0;
  long long done = 0;
  if (vec) {
    const long long nv = count >> 2;
    for (long long i = tid; i < nv; i += stride) {
      float4 a = reinterpret_cast<float4 *>(ema)[i];
      const float4 p = __ldg(reinterpret_cast<const float4 *>(param) + i);
      a.x = __fmaf_rn(momentum, p.x, __fmul_rn(a.x, decay));
      a.y = __fmaf_rn(momentum, p.y, __fmul_rn(a.y, decay));
      a.z = __fmaf_rn(momentum, p.z, __fmul_rn(a.z, decay));
      a.w = __fmaf_rn(momentum, p.w, __fmul_rn(a.w, decay));
      reinterpret_cast<float4 *>(ema)[i] = a;
    }
    done = nv << 2;
  }
  for (long long i = done + tid; i < count; i += stride)
    ema[i] = __fmaf_rn(momentum, param[i], __fmul_rn(ema[i], decay));
}

}  // namespace
}  // namespace nesie

using namespace nesie;

extern "C" int nesie_side_uncertainty_loss(int rows, int ncls, const float *surface_pred,
                                           const float *box_targets, const float *side_scores,
                                           const float *sem_scores, const float *weight,
                                           float loss_weight, float alpha, float *loss_out,
                                           float *sigma_out, void *stream) {
  NESIE_REQUIRE(rows >= 0 && ncls >= 1, "need rows >= 0, ncls >= 1");
  NESIE_REQUIRE(loss_out, "null pointer");
  if (rows == 0) return NESIE_OK;
  NESIE_REQUIRE(surface_pred && box_targets && side_scores && sem_scores && weight, "null pointer");
  int grid = ceil_div(rows * 6, SL_THREADS);
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  side_loss_fwd_kernel<<<grid, SL_THREADS, 0, (cudaStream_t)stream>>>(
      rows, ncls, surface_pred, box_targets, side_scores, sem_scores, weight, loss_weight, alpha,
      loss_out, sigma_out);
  return check_launch("nesie_side_uncertainty_loss");
}

extern "C" int nesie_side_uncertainty_loss_grad(int rows, int ncls, const float *surface_pred,
                                                const float *box_targets,
                                                const float *side_scores, const float *sem_scores,
                                                const float *weight, float loss_weight,
                                                float alpha, const float *grad_loss,
                                                const float *grad_sigma,
                                                float *grad_surface_pred, float *grad_side_scores,
                                                void *stream) {
  NESIE_REQUIRE(rows >= 0 && ncls >= 1, "need rows >= 0, ncls >= 1");
  if (rows == 0) return NESIE_OK;
  NESIE_REQUIRE(surface_pred && box_targets && side_scores && sem_scores && weight, "null pointer");
  side_loss_bwd_kernel<<<ceil_div(rows * 6, SL_THREADS), SL_THREADS, 0, (cudaStream_t)stream>>>(
      rows, ncls, surface_pred, box_targets, side_scores, sem_scores, weight, loss_weight, alpha,
      grad_loss, grad_sigma, grad_surface_pred, grad_side_scores);
  return check_launch("nesie_side_uncertainty_loss_grad");
}

extern "C" int nesie_ema_update(long long count, float *ema, const float *param, float decay,
                                float momentum, void *stream) {
  NESIE_REQUIRE(count >= 0, "negative count");
  if (count == 0) return NESIE_OK;
  NESIE_REQUIRE(ema && param, "null pointer");
  long long blocks = (count / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
  ema_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(count, ema, param, decay, momentum);
  return check_launch("nesie_ema_update");
}
