// Bilinear align_corners=True source-index arithmetic shared by the resampling and
// attention-gate kernels (ATen: scale = (in-1)/(out-1) in fp32, src = scale*dst).
#pragma once
#include "vec.cuh"

namespace ub2 {

__device__ __forceinline__ void src_index(float r, int dst, int in, int& i0, int& i1, float& l0,
                                          float& l1) {
  const float s = r * static_cast<float>(dst);
  i0 = static_cast<int>(s);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = s - static_cast<float>(i0);
  l0 = 1.f - l1;
}

// Low-resolution (N,hin,win,*) tensor sampled at full-resolution pixel (ho,wo) of an (Ho,Wo) grid.
struct LowRes {
  int hin, win, Ho, Wo;
  float rh, rw;
};
inline LowRes make_lowres(int hin, int win, int Ho, int Wo) {
  LowRes g;
  g.hin = hin; g.win = win; g.Ho = Ho; g.Wo = Wo;
  g.rh = Ho > 1 ? static_cast<float>(hin - 1) / static_cast<float>(Ho - 1) : 0.f;
  g.rw = Wo > 1 ? static_cast<float>(win - 1) / static_cast<float>(Wo - 1) : 0.f;
  return g;
}
__device__ __forceinline__ F8 interp8(const __nv_bfloat16* __restrict__ q, int ld, const LowRes& g,
                                      int n, int ho, int wo, int cg) {
  int h0, h1, w0, w1;
  float a0, a1, b0, b1;
  src_index(g.rh, ho, g.hin, h0, h1, a0, a1);
  src_index(g.rw, wo, g.win, w0, w1, b0, b1);
  const __nv_bfloat16* base = q + static_cast<size_t>(n) * g.hin * g.win * ld + cg * 8;
  const F8 v00 = load8(base + (static_cast<size_t>(h0) * g.win + w0) * ld);
  const F8 v01 = load8(base + (static_cast<size_t>(h0) * g.win + w1) * ld);
  const F8 v10 = load8(base + (static_cast<size_t>(h1) * g.win + w0) * ld);
  const F8 v11 = load8(base + (static_cast<size_t>(h1) * g.win + w1) * ld);
  F8 o;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    o.v[k] = a0 * (b0 * v00.v[k] + b1 * v01.v[k]) + a1 * (b0 * v10.v[k] + b1 * v11.v[k]);
  return o;
}

}  // namespace ub2
