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

__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, __nv_bfloat16* __restrict__ out,
                    int ld_out, UpGeom g) {
  const long long total = static_cast<long long>(g.N) * g.Ho * g.Wo * g.cgs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % g.cgs);
    long long pix = i / g.cgs;
    const int wo = static_cast<int>(pix % g.Wo);
    const int ho = static_cast<int>((pix / g.Wo) % g.Ho);
    const int n = static_cast<int>(pix / (static_cast<long long>(g.Wo) * g.Ho));
    const int uh = ho - g.pt, uw = wo - g.pl;
    F8 o;
    if (uh < 0 || uh >= g.hu || uw < 0 || uw >= g.wu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
    } else {
      int h0, h1, w0, w1;
      float a0, a1, b0, b1;
      src_index(g.rh, uh, g.hin, h0, h1, a0, a1);
      src_index(g.rw, uw, g.win, w0, w1, b0, b1);
      const __nv_bfloat16* base = in + static_cast<size_t>(n) * g.hin * g.win * ld_in + cg * 8;
      const F8 v00 = load8(base + (static_cast<size_t>(h0) * g.win + w0) * ld_in);
      const F8 v01 = load8(base + (static_cast<size_t>(h0) * g.win + w1) * ld_in);
      const F8 v10 = load8(base + (static_cast<size_t>(h1) * g.win + w0) * ld_in);
      const F8 v11 = load8(base + (static_cast<size_t>(h1) * g.win + w1) * ld_in);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        o.v[k] = a0 * (b0 * v00.v[k] + b1 * v01.v[k]) + a1 * (b0 * v10.v[k] + b1 * v11.v[k]);
    }
    store8(out + static_cast<size_t>(pix) * ld_out + cg * 8, o);
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

__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int ld_dout,
                    __nv_bfloat16* __restrict__ din, int ld_din, int accumulate, UpGeom g) {
  const long long total = static_cast<long long>(g.N) * g.hin * g.win * g.cgs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % g.cgs);
    long long pix = i / g.cgs;
    const int wi = static_cast<int>(pix % g.win);
    const int hi = static_cast<int>((pix / g.win) % g.hin);
    const int n = static_cast<int>(pix / (static_cast<long long>(g.win) * g.hin));
    int ulo, uhi, vlo, vhi;
    dst_range(g.rh, hi, g.hu, ulo, uhi);
    dst_range(g.rw, wi, g.wu, vlo, vhi);
    F8 acc;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
    for (int u = ulo; u <= uhi; ++u) {
      const float wh = tap_weight(g.rh, u, g.hin, hi);
      if (wh == 0.f) continue;
      const int ho = u + g.pt;
      if (ho < 0 || ho >= g.Ho) continue;
      for (int v = vlo; v <= vhi; ++v) {
        const float ww = tap_weight(g.rw, v, g.win, wi);
        if (ww == 0.f) continue;
        const int wo = v + g.pl;
        if (wo < 0 || wo >= g.Wo) continue;
        const F8 d = load8(dout + ((static_cast<size_t>(n) * g.Ho + ho) * g.Wo + wo) * ld_dout + cg * 8);
        const float wt = wh * ww;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.v[k] = fmaf(wt, d.v[k], acc.v[k]);
      }
    }
    __nv_bfloat16* dst = din + static_cast<size_t>(pix) * ld_din + cg * 8;
    if (accumulate) {
      const F8 old = load8(dst);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc.v[k] += old.v[k];
    }
    store8(dst, acc);
  }
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
  const long long total = static_cast<long long>(N) * Ho * Wo * g.cgs;
  upsample_fwd_kernel<<<stream_grid(total, 256, num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_upsample_bwd(const void* dout, int ld_dout, void* din, int ld_din, int accumulate, int N,
                     int hin, int win, int hu, int wu, int Ho, int Wo, int C, void* stream) {
  if (C % 8 != 0 || Ho < hu || Wo < wu || N <= 0) return UB2_ERR_SHAPE;
  if (ld_dout % 8 || ld_din % 8) return UB2_ERR_ALIGN;
  UpGeom g = up_geom(N, hin, win, hu, wu, Ho, Wo, C);
  const long long total = static_cast<long long>(N) * hin * win * g.cgs;
  upsample_bwd_kernel<<<stream_grid(total, 256, num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), ld_dout, static_cast<__nv_bfloat16*>(din), ld_din,
      accumulate, g);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
