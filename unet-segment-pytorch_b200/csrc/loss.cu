// Segmentation-loss statistics and their gradient (DiceLoss / BalancedCELoss / DiceBCELoss,
// unet/utils/loss.py:45-85, :110-150, :184-191).
//
// Both reference losses are functions of four per-(image, class) sums over the pixels:
//   cnt[n,c]  = #{i : t_i = c}                      (loss.py:139-140, F.one_hot sum :71)
//   ce[n,c]   = sum_{i : t_i = c} -log softmax(z_i)[c]   (F.cross_entropy, :129, split by class)
//   I[n,c]    = sum_{i : t_i = c} softmax(z_i)[c]        (intersection, :70)
//   P[n,c]    = sum_i softmax(z_i)[c]                    (:71)
// seg_stats computes them in one pass over the fp32 NCHW logits + int64 targets with
// warp-shuffle reductions (16 B/pixel for 2 classes); the O(N*C) combination into the scalar
// loss stays on the host side of the ABI.  seg_stats_bwd is the second pass:
//   dz[i,k] = dce[n,t_i] (p_ik - y_ik) + p_ik (g_ik - sum_c g_ic p_ic),
//   g_ic = dI[n,c] [t_i = c] + dP[n,c]
#include "launch.cuh"
#include "../../include/unetb200.h"
#include "conv.h"
#include "vec.cuh"

namespace ub2 {

static constexpr int kLossThreads = 256;
static constexpr int kLossMaxC = 8;

template <int C>
__device__ __forceinline__ void softmax_px(const float* __restrict__ z, long long HW, long long r,
                                           float (&p)[C], float& lse) {
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    p[c] = __ldg(z + c * HW + r);
    m = fmaxf(m, p[c]);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    p[c] = expf(p[c] - m);
    s += p[c];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] *= inv;
  lse = m + logf(s);
}

// grid = (blocks_per_image, N); partial rows: [n][block][C][4] doubles
template <int C>
__global__ void __launch_bounds__(kLossThreads)
seg_stats_kernel(const float* __restrict__ logits, const long long* __restrict__ targets, long long HW,
                 double* partials) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_red[kLossThreads / 32][C * 4];
  const int n = blockIdx.y;
  const float* z = logits + static_cast<size_t>(n) * C * HW;
  const long long* t = targets + static_cast<size_t>(n) * HW;
  float acc[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[c][k] = 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < HW;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    float p[C], lse;
    softmax_px<C>(z, HW, r, p, lse);
    const long long tv = __ldg(t + r);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const bool hit = (tv == c);
      acc[c][0] += hit ? 1.f : 0.f;
      acc[c][1] += hit ? (lse - __ldg(z + c * HW + r)) : 0.f;
      acc[c][2] += hit ? p[c] : 0.f;
      acc[c][3] += p[c];
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float w = warp_sum(acc[c][k]);
      if (lane == 0) s_red[warp][c * 4 + k] = w;
    }
  __syncthreads();
  if (threadIdx.x < C * 4) {
    double s = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) s += static_cast<double>(s_red[w][threadIdx.x]);
    partials[(static_cast<size_t>(n) * gridDim.x + blockIdx.x) * C * 4 + threadIdx.x] = s;
  }
}

// stats[n][k][c] (k = cnt, ce, I, P) as fp32;  grid = (ceil(C*4/32), N), blockDim = (32, 32)
__global__ void seg_stats_finalize_kernel(const double* __restrict__ partials, int blocks, int N, int C,
                                          float* stats) {
  pdl_trigger();
  pdl_wait();
  __shared__ double smem[32 * 33];
  const int n = blockIdx.y;
  const int r = blockIdx.x * 32 + threadIdx.x;
  double s[1];
  rows_sum<1>(partials + static_cast<size_t>(n) * blocks * C * 4, blocks, C * 4, r, s, smem);
  if (threadIdx.y != 0 || r >= C * 4) return;
  const int c = r / 4, k = r % 4;
  stats[(static_cast<size_t>(n) * 4 + k) * C + c] = static_cast<float>(s[0]);
}

// DiceBCELoss (unet/utils/loss.py:153-191) from the statistics table, value and gradient table in
// one tiny launch instead of ~40 O(N*C) tensor expressions and their autograd:
//   bce  = (1/N) sum_n [ ce[n,0] (1-cw)/(cnt[n,0]+s_ce) + ce[n,1] cw/(cnt[n,1]+s_ce) ]   (loss.py:134-148)
//   dice = (2 I + s_d) / (P + cnt + s_d),  loss_dice = 1 - mean over n and c >= c0 of dice   (loss.py:70-85)
//   loss = ce_weight * bce + dice_weight * loss_dice                                         (loss.py:184-191)
// stats (N,4,C) = {cnt, ce, I, P}; coef (N,3,C) = {dL/dce, dL/dI, dL/dP}.  One block, fixed order.
__global__ void __launch_bounds__(256)
dice_bce_head_kernel(const float* __restrict__ stats, int N, int C, float ce_weight, float dice_weight,
                     float class_weight, float ce_smooth, float dice_smooth, int ignore_background,
                     float* __restrict__ loss, float* __restrict__ coef) {
  pdl_trigger();
  pdl_wait();
  __shared__ double s_bce[256], s_dice[256];
  const int c0 = (ignore_background && C > 1) ? 1 : 0;
  const float inv_n = 1.f / static_cast<float>(N);
  const float inv_d = 1.f / static_cast<float>(N * (C - c0));
  double bce = 0.0, dice = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* st = stats + static_cast<size_t>(n) * 4 * C;
    float* cf = coef + static_cast<size_t>(n) * 3 * C;
    for (int c = 0; c < C; ++c) {
      const float cnt = st[c], ce = st[C + c], inter = st[2 * C + c], psum = st[3 * C + c];
      float dce = 0.f, dI = 0.f, dP = 0.f;
      if (c < 2) {   // the balanced CE weights classes 0 and 1 only (binary masks, loss.py:139-145)
        const float wgt = (c == 0 ? 1.f - class_weight : class_weight) / (cnt + ce_smooth);
        bce += static_cast<double>(ce * wgt);
        dce = ce_weight * wgt * inv_n;
      }
      if (c >= c0) {
        const float den = psum + cnt + dice_smooth;
        const float d = (2.f * inter + dice_smooth) / den;
        dice += static_cast<double>(d);
        dI = -dice_weight * inv_d * 2.f / den;
        dP = dice_weight * inv_d * d / den;
      }
      cf[c] = dce;
      cf[C + c] = dI;
      cf[2 * C + c] = dP;
    }
  }
  s_bce[threadIdx.x] = bce;
  s_dice[threadIdx.x] = dice;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0, d = 0.0;
    for (int i = 0; i < 256; ++i) { b += s_bce[i]; d += s_dice[i]; }
    loss[0] = static_cast<float>(ce_weight * (b * inv_n) + dice_weight * (1.0 - d * inv_d));
  }
}

// coef[n][3][C] = {dce, dI, dP}; gscale (optional, one float) = the upstream gradient of the loss
template <int C>
__global__ void __launch_bounds__(kLossThreads)
seg_stats_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ targets,
                     const float* __restrict__ coef, const float* __restrict__ gscale, long long HW,
                     float* __restrict__ dlogits) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  const float* z = logits + static_cast<size_t>(n) * C * HW;
  const long long* t = targets + static_cast<size_t>(n) * HW;
  float* dz = dlogits + static_cast<size_t>(n) * C * HW;
  float dce[C], dI[C], dP[C];
  const float gs = gscale != nullptr ? __ldg(gscale) : 1.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    dce[c] = gs * __ldg(coef + (static_cast<size_t>(n) * 3 + 0) * C + c);
    dI[c] = gs * __ldg(coef + (static_cast<size_t>(n) * 3 + 1) * C + c);
    dP[c] = gs * __ldg(coef + (static_cast<size_t>(n) * 3 + 2) * C + c);
  }
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < HW;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    float p[C], lse;
    softmax_px<C>(z, HW, r, p, lse);
    const long long tv = __ldg(t + r);
    float wce = 0.f, gp = 0.f;
    float g[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const bool hit = (tv == c);
      if (hit) wce = dce[c];
      g[c] = dP[c] + (hit ? dI[c] : 0.f);
      gp = fmaf(g[c], p[c], gp);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float y = (tv == c) ? 1.f : 0.f;
      dz[c * HW + r] = wce * (p[c] - y) + p[c] * (g[c] - gp);
    }
  }
}

template <int C>
static int launch_stats(const float* logits, const long long* targets, int N, long long HW,
                        double* partials, int blocks, float* stats, cudaStream_t s) {
  launch(seg_stats_kernel<C>, dim3(blocks, N), kLossThreads, 0, s, logits, targets, HW, partials);
  launch(seg_stats_finalize_kernel, dim3((C * 4 + 31) / 32, N), dim3(32, 32), 0, s, partials, blocks, N, C, stats);
  return static_cast<int>(cudaGetLastError());
}
template <int C>
static int launch_bwd(const float* logits, const long long* targets, const float* coef, const float* gscale,
                      int N, long long HW, float* dlogits, int blocks, cudaStream_t s) {
  launch(seg_stats_bwd_kernel<C>, dim3(blocks, N), kLossThreads, 0, s, logits, targets, coef, gscale, HW, dlogits);
  return static_cast<int>(cudaGetLastError());
}

static int loss_blocks(int N, long long HW) {
  long long per = (HW + kLossThreads * 4 - 1) / (kLossThreads * 4);
  long long cap = (static_cast<long long>(num_sms()) * 8 + N - 1) / N;
  if (per > cap) per = cap;
  return static_cast<int>(per < 1 ? 1 : per);
}

}  // namespace ub2

using namespace ub2;

extern "C" {

int ub2_seg_stats_blocks(int N, long long HW) {
  if (N <= 0 || HW <= 0) return UB2_ERR_SHAPE;
  return loss_blocks(N, HW);
}

int ub2_seg_stats(const float* logits, const long long* targets, int N, int C, long long HW,
                  double* partials, int blocks, float* stats, void* stream) {
  if (N <= 0 || HW <= 0 || C < 1 || C > kLossMaxC) return UB2_ERR_SHAPE;
  if (blocks != loss_blocks(N, HW)) return UB2_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 1: return launch_stats<1>(logits, targets, N, HW, partials, blocks, stats, s);
    case 2: return launch_stats<2>(logits, targets, N, HW, partials, blocks, stats, s);
    case 3: return launch_stats<3>(logits, targets, N, HW, partials, blocks, stats, s);
    case 4: return launch_stats<4>(logits, targets, N, HW, partials, blocks, stats, s);
    case 5: return launch_stats<5>(logits, targets, N, HW, partials, blocks, stats, s);
    case 6: return launch_stats<6>(logits, targets, N, HW, partials, blocks, stats, s);
    case 7: return launch_stats<7>(logits, targets, N, HW, partials, blocks, stats, s);
    default: return launch_stats<8>(logits, targets, N, HW, partials, blocks, stats, s);
  }
}

int ub2_dice_bce_head(const float* stats, int N, int C, float ce_weight, float dice_weight, float class_weight,
                      float ce_smooth, float dice_smooth, int ignore_background, float* loss, float* coef,
                      void* stream) {
  if (N <= 0 || C < 1 || C > kLossMaxC) return UB2_ERR_SHAPE;
  launch(dice_bce_head_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), stats, N, C, ce_weight, dice_weight, class_weight, ce_smooth, dice_smooth, ignore_background, loss, coef);
  return static_cast<int>(cudaGetLastError());
}

int ub2_seg_stats_bwd(const float* logits, const long long* targets, const float* coef, const float* gscale,
                      int N, int C, long long HW, float* dlogits, void* stream) {
  if (N <= 0 || HW <= 0 || C < 1 || C > kLossMaxC) return UB2_ERR_SHAPE;
  const int blocks = loss_blocks(N, HW);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 1: return launch_bwd<1>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
    case 2: return launch_bwd<2>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
    case 3: return launch_bwd<3>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
    case 4: return launch_bwd<4>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
    case 5: return launch_bwd<5>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
    case 6: return launch_bwd<6>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
    case 7: return launch_bwd<7>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
    default: return launch_bwd<8>(logits, targets, coef, gscale, N, HW, dlogits, blocks, s);
  }
}

}  // extern "C"
