// Bandwidth-bound BatchNorm / ReLU / MaxPool passes over NHWC bf16 activations.
//
// Train-mode forward of one conv->BN->ReLU stage (layers.py:32-34) is
//   conv kernel (raw bf16 output + per-CTA sum / sum-sq rows)
//   -> bn_finalize (batch mean / biased var, running-stat update, scale/shift)
//   -> bn_act     (a = relu(scale*y + shift); optionally the 2x2 max-pooled copy
//                  that Down (layers.py:56) feeds to the next stage).
// Backward: bn_bwd<false> reduces sum(dz), sum(dz*xhat); bn_bwd_finalize turns them
// into dgamma/dbeta (+ the three per-channel coefficients); bn_bwd<true> writes
// dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)).  The gradient arriving through
// the max-pool is routed to the window's first maximum, recomputed from y.
//
// Threads walk 2x2 pixel windows x 8-channel (128-bit) vectors; a thread's channel
// group is fixed so per-channel coefficients live in registers.
#include "../../include/unetb200.h"
#include "conv.h"
#include "vec.cuh"

namespace ub2 {

static constexpr int kBnThreads = 256;

// ------------------------------------------------------------------------------ finalize
__global__ void bn_finalize_kernel(const double* __restrict__ partials, int rows, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_var, long long* nbt,
                                   float momentum, float eps, float* scale, float* shift, float* mean,
                                   float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int r = 0; r < rows; ++r) {
    s1 += partials[(static_cast<size_t>(r) * 2 + 0) * C + c];
    s2 += partials[(static_cast<size_t>(r) * 2 + 1) * C + c];
  }
  const double m = s1 / count;
  double var = s2 / count - m * m;
  if (var < 0.0) var = 0.0;
  const double inv = 1.0 / sqrt(var + static_cast<double>(eps));
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  scale[c] = static_cast<float>(g * inv);
  shift[c] = static_cast<float>(b - m * g * inv);
  mean[c] = static_cast<float>(m);
  invstd[c] = static_cast<float>(inv);
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(m);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv,
                                      float eps, int C, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv = 1.f / sqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f;
  scale[c] = g * inv;
  shift[c] = (beta ? beta[c] : 0.f) - rm[c] * g * inv;
}

// ------------------------------------------------------------------------------ window walk
struct WinGeom {
  int N, H, W, C, cgs, Hc, Wc;
  long long windows;
};
static WinGeom make_geom(int N, int H, int W, int C) {
  WinGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.cgs = C / 8;
  g.Hc = (H + 1) / 2; g.Wc = (W + 1) / 2;
  g.windows = static_cast<long long>(N) * g.Hc * g.Wc;
  return g;
}

// ------------------------------------------------------------------------------ forward apply
__global__ void __launch_bounds__(kBnThreads)
bn_act_kernel(const __nv_bfloat16* __restrict__ y, int ld_y, const float* __restrict__ scale,
              const float* __restrict__ shift, __nv_bfloat16* a, int ld_a, __nv_bfloat16* pooled,
              int ld_p, int relu, WinGeom g) {
  const int lanes = blockDim.x / g.cgs;
  const int lane = threadIdx.x / g.cgs;
  const int cg = threadIdx.x % g.cgs;
  if (lane >= lanes) return;
  F8 sc, sh;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc.v[i] = scale ? scale[cg * 8 + i] : 1.f;
    sh.v[i] = shift ? shift[cg * 8 + i] : 0.f;
  }
  const int Hp = g.H / 2, Wp = g.W / 2;
  for (long long wi = static_cast<long long>(blockIdx.x) * lanes + lane; wi < g.windows;
       wi += static_cast<long long>(gridDim.x) * lanes) {
    const int wc = static_cast<int>(wi % g.Wc);
    const int hc = static_cast<int>((wi / g.Wc) % g.Hc);
    const int n = static_cast<int>(wi / (static_cast<long long>(g.Wc) * g.Hc));
    F8 mx;
#pragma unroll
    for (int i = 0; i < 8; ++i) mx.v[i] = -INFINITY;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int h = hc * 2 + (d >> 1), w = wc * 2 + (d & 1);
      if (h < g.H && w < g.W) {
        const size_t pix = (static_cast<size_t>(n) * g.H + h) * g.W + w;
        F8 v = load8_stream(y + pix * ld_y + cg * 8);
        uint4 packed;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v.v[i] = fmaf(v.v[i], sc.v[i], sh.v[i]);
          if (relu) v.v[i] = fmaxf(v.v[i], 0.f);
        }
        packed = pack8(v);
        if (a != nullptr) *reinterpret_cast<uint4*>(a + pix * ld_a + cg * 8) = packed;
        const F8 r = unpack8(packed);
#pragma unroll
        for (int i = 0; i < 8; ++i) mx.v[i] = fmaxf(mx.v[i], r.v[i]);
      }
    }
    if (pooled != nullptr && hc < Hp && wc < Wp) {
      const size_t pp = (static_cast<size_t>(n) * Hp + hc) * Wp + wc;
      store8(pooled + pp * ld_p + cg * 8, mx);
    }
  }
}

// ------------------------------------------------------------------------------ backward
template <bool APPLY>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_kernel(const __nv_bfloat16* __restrict__ dA, int ld_da, const __nv_bfloat16* __restrict__ dP,
              int ld_dp, const __nv_bfloat16* __restrict__ y, int ld_y,
              const float* __restrict__ scale, const float* __restrict__ shift,
              const float* __restrict__ mean, const float* __restrict__ invstd,
              const float* __restrict__ coef, __nv_bfloat16* dY, int ld_dy, double* partials,
              int relu, WinGeom g) {
  extern __shared__ float s_red[];
  const int lanes = blockDim.x / g.cgs;
  const int lane = threadIdx.x / g.cgs;
  const int cg = threadIdx.x % g.cgs;
  const bool active = lane < lanes;
  F8 sc, sh, mu, is, c1, c2, c3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cg * 8 + i;
    sc.v[i] = scale[c]; sh.v[i] = shift[c]; mu.v[i] = mean[c]; is.v[i] = invstd[c];
    if (APPLY) {
      c1.v[i] = coef[c]; c2.v[i] = coef[g.C + c]; c3.v[i] = coef[2 * g.C + c];
    }
  }
  F8 s1, s2;
#pragma unroll
  for (int i = 0; i < 8; ++i) s1.v[i] = s2.v[i] = 0.f;
  const int Hp = g.H / 2, Wp = g.W / 2;
  if (active) {
    for (long long wi = static_cast<long long>(blockIdx.x) * lanes + lane; wi < g.windows;
         wi += static_cast<long long>(gridDim.x) * lanes) {
      const int wc = static_cast<int>(wi % g.Wc);
      const int hc = static_cast<int>((wi / g.Wc) % g.Hc);
      const int n = static_cast<int>(wi / (static_cast<long long>(g.Wc) * g.Hc));
      F8 yv[4], zv[4];
      bool inb[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int h = hc * 2 + (d >> 1), w = wc * 2 + (d & 1);
        inb[d] = (h < g.H && w < g.W);
        if (inb[d]) {
          const size_t pix = (static_cast<size_t>(n) * g.H + h) * g.W + w;
          yv[d] = load8_stream(y + pix * ld_y + cg * 8);
#pragma unroll
          for (int i = 0; i < 8; ++i) zv[d].v[i] = fmaf(yv[d].v[i], sc.v[i], sh.v[i]);
        }
      }
      // gradient through the 2x2 max-pool goes to the first maximum of the stored activation
      int amax[8];
      F8 gp;
      const bool pooled_win = (dP != nullptr) && hc < Hp && wc < Wp;
      if (pooled_win) {
        const size_t pp = (static_cast<size_t>(n) * Hp + hc) * Wp + wc;
        gp = load8_stream(dP + pp * ld_dp + cg * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float best = -INFINITY;
          int bi = 0;
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            float av = relu ? fmaxf(zv[d].v[i], 0.f) : zv[d].v[i];
            av = __bfloat162float(__float2bfloat16_rn(av));
            if (av > best) { best = av; bi = d; }
          }
          amax[i] = bi;
        }
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (!inb[d]) continue;
        const int h = hc * 2 + (d >> 1), w = wc * 2 + (d & 1);
        const size_t pix = (static_cast<size_t>(n) * g.H + h) * g.W + w;
        F8 dz;
        if (dA != nullptr) {
          dz = load8_stream(dA + pix * ld_da + cg * 8);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) dz.v[i] = 0.f;
        }
        F8 out;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float gsum = dz.v[i];
          if (pooled_win && amax[i] == d) gsum += gp.v[i];
          if (relu && !(zv[d].v[i] > 0.f)) gsum = 0.f;
          const float xh = (yv[d].v[i] - mu.v[i]) * is.v[i];
          if (APPLY) {
            out.v[i] = c1.v[i] * (gsum - c2.v[i] - xh * c3.v[i]);
          } else {
            s1.v[i] += gsum;
            s2.v[i] += gsum * xh;
          }
        }
        if (APPLY) store8(dY + pix * ld_dy + cg * 8, out);
      }
    }
  }
  if (!APPLY) {
    // block reduction over pixel lanes: s_red[lane][cg][16]
    float* mine = s_red + (static_cast<size_t>(lane) * g.cgs + cg) * 16;
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { mine[i] = s1.v[i]; mine[8 + i] = s2.v[i]; }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < g.cgs * 16; idx += blockDim.x) {
      double acc = 0.0;
      for (int l = 0; l < lanes; ++l) acc += static_cast<double>(s_red[static_cast<size_t>(l) * g.cgs * 16 + idx]);
      const int cgi = idx / 16, k = idx % 16;
      const int c = cgi * 8 + (k & 7);
      partials[(static_cast<size_t>(blockIdx.x) * 2 + (k >> 3)) * g.C + c] = acc;
    }
  }
}

// dgamma/dbeta accumulate into the fp32 .grad tensors; coef = {gamma*invstd, dbeta/M, dgamma/M}
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ partials, int rows, int C,
                                       double count, const float* __restrict__ gamma,
                                       const float* __restrict__ invstd, float* dgamma, float* dbeta,
                                       float* coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int r = 0; r < rows; ++r) {
    s1 += partials[(static_cast<size_t>(r) * 2 + 0) * C + c];
    s2 += partials[(static_cast<size_t>(r) * 2 + 1) * C + c];
  }
  if (dbeta) dbeta[c] += static_cast<float>(s1);
  if (dgamma) dgamma[c] += static_cast<float>(s2);
  const float g = gamma ? gamma[c] : 1.f;
  coef[c] = g * invstd[c];
  coef[C + c] = static_cast<float>(s1 / count);
  coef[2 * C + c] = static_cast<float>(s2 / count);
}

static int bn_block(int cgs) { return cgs * (kBnThreads / cgs); }

}  // namespace ub2

using namespace ub2;

extern "C" {

int ub2_bn_finalize(const double* partials, int rows, int C, double count, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, long long* nbt,
                    float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                    void* stream) {
  if (C <= 0 || rows <= 0) return UB2_ERR_SHAPE;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      partials, rows, C, count, gamma, beta, running_mean, running_var, nbt, momentum, eps, scale,
      shift, mean, invstd);
  return static_cast<int>(cudaGetLastError());
}

int ub2_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, int C, float* scale, float* shift,
                       void* stream) {
  if (C <= 0) return UB2_ERR_SHAPE;
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      gamma, beta, running_mean, running_var, eps, C, scale, shift);
  return static_cast<int>(cudaGetLastError());
}

int ub2_bn_act(const void* y, int ld_y, const float* scale, const float* shift, void* a, int ld_a,
               void* pooled, int ld_p, int N, int H, int W, int C, int relu, void* stream) {
  if (C % 8 != 0 || C / 8 > kBnThreads || N <= 0) return UB2_ERR_SHAPE;
  if (ld_y % 8 || (a && ld_a % 8) || (pooled && ld_p % 8)) return UB2_ERR_ALIGN;
  WinGeom g = make_geom(N, H, W, C);
  const int block = bn_block(g.cgs);
  const int lanes = block / g.cgs;
  const int grid = stream_grid(g.windows, lanes, num_sms(), 8);
  bn_act_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), ld_y, scale, shift, static_cast<__nv_bfloat16*>(a), ld_a,
      static_cast<__nv_bfloat16*>(pooled), ld_p, relu, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_bn_bwd_rows(int N, int H, int W, int C) {
  if (C % 8 != 0 || C / 8 > kBnThreads) return UB2_ERR_SHAPE;
  WinGeom g = make_geom(N, H, W, C);
  const int lanes = bn_block(g.cgs) / g.cgs;
  return stream_grid(g.windows, lanes, num_sms(), 4);
}

int ub2_bn_bwd_reduce(const void* dA, int ld_da, const void* dP, int ld_dp, const void* y, int ld_y,
                      const float* scale, const float* shift, const float* mean, const float* invstd,
                      double* partials, int rows, int N, int H, int W, int C, int relu, void* stream) {
  if (C % 8 != 0 || C / 8 > kBnThreads || N <= 0) return UB2_ERR_SHAPE;
  WinGeom g = make_geom(N, H, W, C);
  const int block = bn_block(g.cgs);
  const int lanes = block / g.cgs;
  const int grid = stream_grid(g.windows, lanes, num_sms(), 4);
  if (grid != rows) return UB2_ERR_WORKSPACE;
  const size_t smem = static_cast<size_t>(lanes) * g.cgs * 16 * sizeof(float);
  bn_bwd_kernel<false><<<grid, block, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dA), ld_da, static_cast<const __nv_bfloat16*>(dP), ld_dp,
      static_cast<const __nv_bfloat16*>(y), ld_y, scale, shift, mean, invstd, nullptr, nullptr, 0,
      partials, relu, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_bn_bwd_finalize(const double* partials, int rows, int C, double count, const float* gamma,
                        const float* invstd, float* dgamma, float* dbeta, float* coef, void* stream) {
  if (C <= 0 || rows <= 0) return UB2_ERR_SHAPE;
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      partials, rows, C, count, gamma, invstd, dgamma, dbeta, coef);
  return static_cast<int>(cudaGetLastError());
}

int ub2_bn_bwd_apply(const void* dA, int ld_da, const void* dP, int ld_dp, const void* y, int ld_y,
                     const float* scale, const float* shift, const float* mean, const float* invstd,
                     const float* coef, void* dY, int ld_dy, int N, int H, int W, int C, int relu,
                     void* stream) {
  if (C % 8 != 0 || C / 8 > kBnThreads || N <= 0) return UB2_ERR_SHAPE;
  WinGeom g = make_geom(N, H, W, C);
  const int block = bn_block(g.cgs);
  const int lanes = block / g.cgs;
  const int grid = stream_grid(g.windows, lanes, num_sms(), 8);
  bn_bwd_kernel<true><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dA), ld_da, static_cast<const __nv_bfloat16*>(dP), ld_dp,
      static_cast<const __nv_bfloat16*>(y), ld_y, scale, shift, mean, invstd, coef,
      static_cast<__nv_bfloat16*>(dY), ld_dy, nullptr, relu, g);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
