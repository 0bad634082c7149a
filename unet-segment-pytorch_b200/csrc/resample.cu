// Bilinear (align_corners=True) up-sampling of NHWC bf16 tensors and its transpose.
//
// Forward restates nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
// followed by F.pad to the skip tensor's size (layers.py:78, :98-102, :212, :247-251)
// and F.interpolate(g, size=x.shape[2:]) of the attention gate (layers.py:183):
// the (hin,win) source is resampled to (hu,wu) and placed at offset (pt,pl) inside an
// (Ho,Wo) output whose remaining border is zero.  Source index arithmetic follows
// ATen: scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = int(src), lambda = src-i0.
// Backward is the exact transpose in gather form (deterministic, no atomics).
#include "../../include/unetb200.h"
#include "conv.h"
#include "resample.cuh"
#include "vec.cuh"

namespace ub2 {

struct UpGeom {
  int N, hin, win, hu, wu, Ho, Wo, pt, pl, C, cgs;
  float rh, rw;
};

// Launch geometry for both kernels: blockDim = (cgs' , 256/cgs'), grid = (ceil(width/blockDim.y),
// height, N): no integer division in the kernels, the row taps are block-uniform.
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, __nv_bfloat16* __restrict__ out,
                    int ld_out, UpGeom g) {
  const int wo = blockIdx.x * blockDim.y + threadIdx.y;
  const int ho = blockIdx.y;
  const int n = blockIdx.z;
  if (wo >= g.Wo) return;
  const int uh = ho - g.pt, uw = wo - g.pl;
  const bool inside = !(uh < 0 || uh >= g.hu || uw < 0 || uw >= g.wu);
  int h0 = 0, h1 = 0, w0 = 0, w1 = 0;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  if (inside) {
    src_index(g.rh, uh, g.hin, h0, h1, a0, a1);
    src_index(g.rw, uw, g.win, w0, w1, b0, b1);
  }
  const __nv_bfloat16* base = in + static_cast<size_t>(n) * g.hin * g.win * ld_in;
  __nv_bfloat16* dst = out + ((static_cast<size_t>(n) * g.Ho + ho) * g.Wo + wo) * ld_out;
  for (int cg = threadIdx.x; cg < g.cgs; cg += blockDim.x) {
    F8 o;
    if (!inside) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
    } else {
      const F8 v00 = load8(base + (static_cast<size_t>(h0) * g.win + w0) * ld_in + cg * 8);
      const F8 v01 = load8(base + (static_cast<size_t>(h0) * g.win + w1) * ld_in + cg * 8);
      const F8 v10 = load8(base + (static_cast<size_t>(h1) * g.win + w0) * ld_in + cg * 8);
      const F8 v11 = load8(base + (static_cast<size_t>(h1) * g.win + w1) * ld_in + cg * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        o.v[k] = a0 * (b0 * v00.v[k] + b1 * v01.v[k]) + a1 * (b0 * v10.v[k] + b1 * v11.v[k]);
    }
    store8(dst + cg * 8, o);
  }
}

// weight with which up-sampled index `u` reads source index `i`
__device__ __forceinline__ float tap_weight(float r, int u, int in, int i) {
  int i0, i1;
  float l0, l1;
  src_index(r, u, in, i0, i1, l0, l1);
  float w = 0.f;
  if (i0 == i) w += l0;
  if (i1 == i) w += l1;  // at the clamped edge both taps read the same source
  return w;
}

__device__ __forceinline__ void dst_range(float r, int i, int out, int& lo, int& hi) {
  if (r <= 0.f) {
    lo = 0;
    hi = out - 1;
    return;
  }
  lo = static_cast<int>(floorf((static_cast<float>(i) - 1.f) / r)) - 1;
  hi = static_cast<int>(ceilf((static_cast<float>(i) + 1.f) / r)) + 1;
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
}

static constexpr int kMaxTaps = 8;

__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ld_dout,
                    __nv_bfloat16* __restrict__ din, int ld_din, int accumulate, UpGeom g) {
  const int wi = blockIdx.x * blockDim.y + threadIdx.y;
  const int hi = blockIdx.y;
  const int n = blockIdx.z;
  if (wi >= g.win) return;
  int ulo, uhi, vlo, vhi;
  dst_range(g.rh, hi, g.hu, ulo, uhi);
  dst_range(g.rw, wi, g.wu, vlo, vhi);
  // the (<= 6 for a 2x up-sampling) column taps that touch this source column
  int vs[kMaxTaps];
  float wv[kMaxTaps];
  int nv = 0;
  bool overflow = false;
  for (int v = vlo; v <= vhi; ++v) {
    const float ww = tap_weight(g.rw, v, g.win, wi);
    const int wo = v + g.pl;
    if (ww == 0.f || wo < 0 || wo >= g.Wo) continue;
    if (nv < kMaxTaps) {
#pragma unroll
      for (int k = 0; k < kMaxTaps; ++k)
        if (k == nv) { vs[k] = wo; wv[k] = ww; }
      ++nv;
    } else {
      overflow = true;
    }
  }
  __nv_bfloat16* dst = din + ((static_cast<size_t>(n) * g.hin + hi) * g.win + wi) * ld_din;
  for (int cg = threadIdx.x; cg < g.cgs; cg += blockDim.x) {
    F8 acc;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
    for (int u = ulo; u <= uhi; ++u) {
      const float wh = tap_weight(g.rh, u, g.hin, hi);
      const int ho = u + g.pt;
      if (wh == 0.f || ho < 0 || ho >= g.Ho) continue;
      const __nv_bfloat16* row = dout + (static_cast<size_t>(n) * g.Ho + ho) * g.Wo * ld_dout + cg * 8;
      if (!overflow) {
#pragma unroll
        for (int k = 0; k < kMaxTaps; ++k) {
          if (k < nv) {
            const F8 d = load8(row + static_cast<size_t>(vs[k]) * ld_dout);
            const float wt = wh * wv[k];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc.v[c] = fmaf(wt, d.v[c], acc.v[c]);
          }
        }
      } else {
        for (int v = vlo; v <= vhi; ++v) {
          const float ww = tap_weight(g.rw, v, g.win, wi);
          const int wo = v + g.pl;
          if (ww == 0.f || wo < 0 || wo >= g.Wo) continue;
          const F8 d = load8(row + static_cast<size_t>(wo) * ld_dout);
          const float wt = wh * ww;
#pragma unroll
          for (int c = 0; c < 8; ++c) acc.v[c] = fmaf(wt, d.v[c], acc.v[c]);
        }
      }
    }
    if (accumulate) {
      const F8 old = load8(dst + cg * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc.v[k] += old.v[k];
    }
    store8(dst + cg * 8, acc);
  }
}

static dim3 up_block(int cgs) {
  int bx = 1;
  while (bx < cgs && bx < 256) bx *= 2;
  return dim3(bx, 256 / bx);
}

static UpGeom up_geom(int N, int hin, int win, int hu, int wu, int Ho, int Wo, int C) {
  UpGeom g;
  g.N = N; g.hin = hin; g.win = win; g.hu = hu; g.wu = wu; g.Ho = Ho; g.Wo = Wo; g.C = C;
  g.cgs = C / 8;
  g.pt = (Ho - hu) / 2;
  g.pl = (Wo - wu) / 2;
  g.rh = hu > 1 ? static_cast<float>(hin - 1) / static_cast<float>(hu - 1) : 0.f;
  g.rw = wu > 1 ? static_cast<float>(win - 1) / static_cast<float>(wu - 1) : 0.f;
  return g;
}

}  // namespace ub2

using namespace ub2;

extern "C" {

int ub2_upsample_fwd(const void* in, int ld_in, void* out, int ld_out, int N, int hin, int win, int hu,
                     int wu, int Ho, int Wo, int C, void* stream) {
  if (C % 8 != 0 || Ho < hu || Wo < wu || N <= 0) return UB2_ERR_SHAPE;
  if (ld_in % 8 || ld_out % 8) return UB2_ERR_ALIGN;
  UpGeom g = up_geom(N, hin, win, hu, wu, Ho, Wo, C);
  const dim3 block = up_block(g.cgs);
  const dim3 grid((Wo + block.y - 1) / block.y, Ho, N);
  upsample_fwd_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_upsample_bwd(const void* dout, int ld_dout, void* din, int ld_din, int accumulate, int N,
                     int hin, int win, int hu, int wu, int Ho, int Wo, int C, void* stream) {
  if (C % 8 != 0 || Ho < hu || Wo < wu || N <= 0) return UB2_ERR_SHAPE;
  if (ld_dout % 8 || ld_din % 8) return UB2_ERR_ALIGN;
  UpGeom g = up_geom(N, hin, win, hu, wu, Ho, Wo, C);
  const dim3 block = up_block(g.cgs);
  const dim3 grid((win + block.y - 1) / block.y, hin, N);
  upsample_bwd_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), ld_dout, static_cast<__nv_bfloat16*>(din), ld_din,
      accumulate, g);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
