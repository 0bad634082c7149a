// Bilinear (align_corners=True) up-sampling of NHWC bf16 tensors and its transpose.
//
// Forward restates nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
// followed by F.pad to the skip tensor's size (layers.py:78, :98-102, :212, :247-251)
// and F.interpolate(g, size=x.shape[2:]) of the attention gate (layers.py:183):
// the (hin,win) source is resampled to (hu,wu) and placed at offset (pt,pl) inside an
// (Ho,Wo) output whose remaining border is zero.  Source index arithmetic follows
// ATen: scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = int(src), lambda = src-i0.
// Backward is the exact transpose in gather form (deterministic, no atomics).
#include "launch.cuh"
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
  pdl_trigger();
  pdl_wait();
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

// Column-strip forward: a thread owns one destination column x 8 channels and walks kStripRows
// destination rows, keeping the two horizontally interpolated source rows it is between in
// registers (they change every other row for a 2x up-sampling).  ~50 instructions per 16-byte
// output instead of ~230 (the per-pixel kernel above is issue bound: ncu issue-active 75 %).
static constexpr int kStripRows = 16;

__global__ void __launch_bounds__(256)
upsample_fwd_strip_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, __nv_bfloat16* __restrict__ out,
                          int ld_out, UpGeom g) {
  pdl_trigger();
  pdl_wait();
  const int wo = blockIdx.x * blockDim.y + threadIdx.y;
  const int n = blockIdx.z;
  if (wo >= g.Wo) return;
  const int uw = wo - g.pl;
  const bool col_inside = uw >= 0 && uw < g.wu;
  int w0 = 0, w1 = 0;
  float b0 = 0.f, b1 = 0.f;
  if (col_inside) src_index(g.rw, uw, g.win, w0, w1, b0, b1);
  const __nv_bfloat16* base = in + static_cast<size_t>(n) * g.hin * g.win * ld_in;
  const int ho_begin = blockIdx.y * kStripRows;
  const int ho_end = min(ho_begin + kStripRows, g.Ho);
  for (int cg = threadIdx.x; cg < g.cgs; cg += blockDim.x) {
    auto hrow = [&](int h) {   // source row h interpolated at this column
      const F8 a = load8(base + (static_cast<size_t>(h) * g.win + w0) * ld_in + cg * 8);
      const F8 b = load8(base + (static_cast<size_t>(h) * g.win + w1) * ld_in + cg * 8);
      F8 r;
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[k] = b0 * a.v[k] + b1 * b.v[k];
      return r;
    };
    int cur0 = -1, cur1 = -1;
    F8 top, bot;
#pragma unroll
    for (int k = 0; k < 8; ++k) top.v[k] = bot.v[k] = 0.f;
    for (int ho = ho_begin; ho < ho_end; ++ho) {
      const int uh = ho - g.pt;
      F8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
      if (col_inside && uh >= 0 && uh < g.hu) {
        int h0, h1;
        float a0, a1;
        src_index(g.rh, uh, g.hin, h0, h1, a0, a1);   // block-uniform
        if (h0 != cur0) {
          top = (h0 == cur1) ? bot : hrow(h0);
          cur0 = h0;
        }
        if (h1 != cur1) {
          bot = (h1 == cur0) ? top : hrow(h1);
          cur1 = h1;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = a0 * top.v[k] + a1 * bot.v[k];
      }
      store8(out + ((static_cast<size_t>(n) * g.Ho + ho) * g.Wo + wo) * ld_out + cg * 8, o);
    }
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
  pdl_trigger();
  pdl_wait();
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

// Tiled transpose for the common case (scale <= ~2/3, i.e. real up-sampling): a block owns an
// 8 x 16 tile of source pixels for a slab of 16 channels, stages the part of `dout` that touches
// the tile (every element once, 32-byte segments) in shared memory and gathers from there.
// The direct kernel above reads each dout element ~16 times through L1/L2.
static constexpr int kBT_H = 8, kBT_W = 16;
static constexpr int kBT_RH = 24, kBT_RW = 40;   // staged region capacity (rows x columns of dout)

__global__ void __launch_bounds__(256)
upsample_bwd_tiled_kernel(const __nv_bfloat16* __restrict__ dout, int ld_dout,
                          __nv_bfloat16* __restrict__ din, int ld_din, int accumulate, int slabs, UpGeom g) {
  pdl_trigger();
  pdl_wait();
  __shared__ uint4 region[kBT_RH * kBT_RW * 2];
  const int slab = blockIdx.z % slabs;
  const int n = blockIdx.z / slabs;
  const int h0 = blockIdx.y * kBT_H, w0 = blockIdx.x * kBT_W;
  const int h_last = min(h0 + kBT_H, g.hin) - 1, w_last = min(w0 + kBT_W, g.win) - 1;
  int ulo, uhi, vlo, vhi, t0, t1;
  dst_range(g.rh, h0, g.hu, ulo, t0);
  dst_range(g.rh, h_last, g.hu, t1, uhi);
  dst_range(g.rw, w0, g.wu, vlo, t0);
  dst_range(g.rw, w_last, g.wu, t1, vhi);
  const int RW = vhi - vlo + 1, RH = uhi - ulo + 1;   // host guarantees RH <= kBT_RH, RW <= kBT_RW
  const int tid = threadIdx.x;
  const __nv_bfloat16* src = dout + static_cast<size_t>(n) * g.Ho * g.Wo * ld_dout + slab * 16;
  // all of this thread's loads are issued before the first shared-memory store (a load -> store
  // loop keeps one 16-byte request in flight per thread and runs at a quarter of HBM speed)
  constexpr int kLoads = (kBT_RH * kBT_RW * 2 + 255) / 256;
  uint4 stage[kLoads];
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {   // the staged region has a fixed pitch: constant divisors
    const int idx = tid + k * 256;
    const int col = (idx >> 1) % kBT_RW;
    const int row = (idx >> 1) / kBT_RW;
    const int ho = ulo + row + g.pt, wo = vlo + col + g.pl;
    stage[k] = make_uint4(0u, 0u, 0u, 0u);
    if (row < RH && col < RW && ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo)
      stage[k] = __ldg(reinterpret_cast<const uint4*>(src + (static_cast<size_t>(ho) * g.Wo + wo) * ld_dout) + (idx & 1));
  }
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {
    const int idx = tid + k * 256;
    if (idx < kBT_RH * kBT_RW * 2) region[idx] = stage[k];
  }
  __syncthreads();
  const int half = tid & 1;
  const int wi = w0 + ((tid >> 1) & (kBT_W - 1));
  const int hi = h0 + (tid >> 5);
  if (wi >= g.win || hi >= g.hin) return;
  // Transposed bilinear weights are the hat function max(0, 1 - |r*u - i|) of the same fp32
  // product ATen forms; with r >= 0.4 (host check) at most five destination rows / columns touch
  // one source row / column, starting at the first u with r*u > i - 1.
  const float fi = static_cast<float>(hi), fj = static_cast<float>(wi);
  int u0 = static_cast<int>(floorf((fi - 1.f) / g.rh)) + 1;
  int v0 = static_cast<int>(floorf((fj - 1.f) / g.rw)) + 1;
  if (u0 < 0) u0 = 0;
  if (v0 < 0) v0 = 0;
  float wv[5];
  int cv[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int v = v0 + k;
    wv[k] = (v < g.wu) ? fmaxf(0.f, 1.f - fabsf(g.rw * static_cast<float>(v) - fj)) : 0.f;
    cv[k] = (min(v, vhi) - vlo) * 2;   // clamped: a zero weight never multiplies stale shared memory
  }
  F8 acc;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
#pragma unroll
  for (int kr = 0; kr < 5; ++kr) {
    const int u = u0 + kr;
    const float wh = (u < g.hu) ? fmaxf(0.f, 1.f - fabsf(g.rh * static_cast<float>(u) - fi)) : 0.f;
    if (wh == 0.f) continue;   // warp-uniform: a warp is one source row of the tile
    const uint4* rowp = region + (min(u, uhi) - ulo) * (kBT_RW * 2) + half;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const F8 d = unpack8(rowp[cv[k]]);
      const float wt = wh * wv[k];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc.v[c] = fmaf(wt, d.v[c], acc.v[c]);
    }
  }
  __nv_bfloat16* dst = din + ((static_cast<size_t>(n) * g.hin + hi) * g.win + wi) * ld_din + slab * 16 + half * 8;
  if (accumulate) {
    const F8 old = load8(dst);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.v[k] += old.v[k];
  }
  store8(dst, acc);
}

// host mirror of dst_range's span: can the tiled kernel stage every tile's region?
static bool bwd_tiled_ok(const UpGeom& g) {
  if (g.C % 16 != 0 || g.rh <= 0.f || g.rw <= 0.f) return false;
  const float span_h = (static_cast<float>(kBT_H) + 1.f) / g.rh + 5.f;
  const float span_w = (static_cast<float>(kBT_W) + 1.f) / g.rw + 5.f;
  // and every source column at most kMaxTaps candidate taps (2/r + 3 <= 8)
  return span_h <= kBT_RH && span_w <= kBT_RW && g.rw >= 0.4f && g.rh >= 0.4f;
}

// Column-strip transpose for real up-sampling (0.4 <= scale <= 0.6 on both axes): a thread owns one
// SOURCE column x 8 channels and a strip of kBwdStrip source rows; it walks the destination rows
// that touch the strip once, reduces each horizontally (<= 5 hat-weighted taps) and adds the result
// to the two source rows it lies between, which are kept in registers and written when the walk has
// passed them.  ~220 instructions per 16-byte output instead of ~1000 for the gather kernels (both
// were issue bound), deterministic, no shared memory.
// Strip length: 16 rows when that still fills the machine; shorter strips (more, shorter walks: a
// thread's walk is a chain of dependent load latencies, ~1.8 us per destination row) when the batch
// is small — at batch 4 every level has fewer threads than one wave even with 16-row strips.
static constexpr int kBwdStrip = 16;

__global__ void __launch_bounds__(256)
upsample_bwd_strip_kernel(const __nv_bfloat16* __restrict__ dout, int ld_dout, __nv_bfloat16* __restrict__ din,
                          int ld_din, int accumulate, int strip, UpGeom g) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.x * blockDim.y + threadIdx.y;   // source column
  const int n = blockIdx.z;
  if (j >= g.win) return;
  const int ia = blockIdx.y * strip, ib = min(ia + strip, g.hin);
  // column taps: destination columns v0 .. v0+4 (see upsample_bwd_tiled_kernel)
  const float fj = static_cast<float>(j);
  int v0 = static_cast<int>(floorf((fj - 1.f) / g.rw)) + 1;
  if (v0 < 0) v0 = 0;
  float wv[5];
  int cv[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int v = v0 + k;
    wv[k] = (v < g.wu) ? fmaxf(0.f, 1.f - fabsf(g.rw * static_cast<float>(v) - fj)) : 0.f;
    cv[k] = min(v, g.wu - 1) + g.pl;
  }
  // destination rows whose upper tap is one of the rows ia-1 .. ib-1 (block-uniform, with margins;
  // the exact classification below uses ATen's own index arithmetic)
  int u_lo = static_cast<int>(floorf(static_cast<float>(ia - 1) / g.rh)) - 1;
  int u_hi = static_cast<int>(ceilf(static_cast<float>(ib) / g.rh)) + 1;
  if (u_lo < 0) u_lo = 0;
  if (u_hi > g.hu - 1) u_hi = g.hu - 1;
  const __nv_bfloat16* src = dout + static_cast<size_t>(n) * g.Ho * g.Wo * ld_dout;
  __nv_bfloat16* dst = din + (static_cast<size_t>(n) * g.hin * g.win + j) * ld_din;
  for (int cg = threadIdx.x; cg < g.cgs; cg += blockDim.x) {
    auto flush = [&](int row, const F8& acc) {
      if (row < ia || row >= ib) return;
      __nv_bfloat16* q = dst + static_cast<size_t>(row) * g.win * ld_din + cg * 8;
      F8 o = acc;
      if (accumulate) {
        const F8 old = load8(q);
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] += old.v[k];
      }
      store8(q, o);
    };
    F8 acc0, acc1;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc0.v[k] = acc1.v[k] = 0.f;
    int r_cur = -2;
    for (int u = u_lo; u <= u_hi; ++u) {
      int i0, i1;
      float l0, l1;
      src_index(g.rh, u, g.hin, i0, i1, l0, l1);
      if (i0 < ia - 1 || i0 > ib - 1) continue;   // touches no row of this strip
      if (r_cur == -2) r_cur = i0;
      while (i0 > r_cur) {                          // the walk has passed row r_cur
        flush(r_cur, acc0);
        acc0 = acc1;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc1.v[k] = 0.f;
        ++r_cur;
      }
      const __nv_bfloat16* rowp = src + static_cast<size_t>(u + g.pt) * g.Wo * ld_dout + cg * 8;
      F8 t;
#pragma unroll
      for (int k = 0; k < 8; ++k) t.v[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        if (wv[k] != 0.f) {
          const F8 d = load8(rowp + static_cast<size_t>(cv[k]) * ld_dout);
#pragma unroll
          for (int c = 0; c < 8; ++c) t.v[c] = fmaf(wv[k], d.v[c], t.v[c]);
        }
      }
      if (i1 == i0) {   // clamped last row: both taps are the same source row
#pragma unroll
        for (int k = 0; k < 8; ++k) acc0.v[k] = fmaf(l0 + l1, t.v[k], acc0.v[k]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc0.v[k] = fmaf(l0, t.v[k], acc0.v[k]);
          acc1.v[k] = fmaf(l1, t.v[k], acc1.v[k]);
        }
      }
    }
    if (r_cur != -2) {
      flush(r_cur, acc0);
      flush(r_cur + 1, acc1);
    }
  }
}

static bool bwd_strip_ok(const UpGeom& g) {
  return g.rh >= 0.4f && g.rh <= 0.6f && g.rw >= 0.4f && g.rw <= 0.6f && g.hin >= 2 * kBwdStrip;
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

// ------------------------------------------------------------------------------ fp32 planes
// F.interpolate(logits, size, mode='bilinear', align_corners=True) of the deep-supervision heads
// (unet/models/unet.py:206-208): fp32 NCHW planes (N * n_classes of them), any scale.  Same index
// arithmetic and summation form as ATen's upsample_bilinear2d.  A few MB per step: one thread
// per pixel.
__global__ void __launch_bounds__(256)
resize_planes_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, LowRes g) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(planes) * g.Ho * g.Wo;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(idx % g.Wo);
    const long long row = idx / g.Wo;
    const int ho = static_cast<int>(row % g.Ho);
    const long long p = row / g.Ho;
    int h0, h1, w0, w1;
    float a0, a1, b0, b1;
    src_index(g.rh, ho, g.hin, h0, h1, a0, a1);
    src_index(g.rw, wo, g.win, w0, w1, b0, b1);
    const float* base = in + p * g.hin * g.win;
    const float v00 = __ldg(base + h0 * g.win + w0), v01 = __ldg(base + h0 * g.win + w1);
    const float v10 = __ldg(base + h1 * g.win + w0), v11 = __ldg(base + h1 * g.win + w1);
    out[idx] = a0 * (b0 * v00 + b1 * v01) + a1 * (b0 * v10 + b1 * v11);
  }
}

// Its transpose in gather form: din[p][i][j] = sum over the destination pixels that read (i,j).
// Deterministic (ATen scatters with atomics).
__global__ void __launch_bounds__(256)
resize_planes_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int planes, LowRes g) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(planes) * g.hin * g.win;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(idx % g.win);
    const long long row = idx / g.win;
    const int i = static_cast<int>(row % g.hin);
    const long long p = row / g.hin;
    int ylo, yhi, xlo, xhi;
    dst_range(g.rh, i, g.Ho, ylo, yhi);
    dst_range(g.rw, j, g.Wo, xlo, xhi);
    const float* base = dout + p * g.Ho * g.Wo;
    float acc = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      const float wy = tap_weight(g.rh, y, g.hin, i);
      if (wy == 0.f) continue;
      float rowacc = 0.f;
      for (int x = xlo; x <= xhi; ++x) {
        const float wx = tap_weight(g.rw, x, g.win, j);
        if (wx != 0.f) rowacc = fmaf(wx, __ldg(base + static_cast<long long>(y) * g.Wo + x), rowacc);
      }
      acc = fmaf(wy, rowacc, acc);
    }
    din[idx] = acc;
  }
}

}  // namespace ub2

using namespace ub2;

extern "C" {

int ub2_upsample_fwd(const void* in, int ld_in, void* out, int ld_out, int N, int hin, int win, int hu,
                     int wu, int Ho, int Wo, int C, void* stream) {
  if (C <= 0 || C % 8 != 0 || Ho < hu || Wo < wu || N <= 0 || hin <= 0 || win <= 0 || hu <= 0 || wu <= 0)
    return UB2_ERR_SHAPE;
  if (ld_in % 8 || ld_out % 8) return UB2_ERR_ALIGN;
  UpGeom g = up_geom(N, hin, win, hu, wu, Ho, Wo, C);
  if (Ho >= 2 * kStripRows) {
    const dim3 sblock = up_block(g.cgs);
    const dim3 sgrid((Wo + sblock.y - 1) / sblock.y, (Ho + kStripRows - 1) / kStripRows, N);
    launch(upsample_fwd_strip_kernel, sgrid, sblock, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, g);
    return static_cast<int>(cudaGetLastError());
  }
  const dim3 block = up_block(g.cgs);
  const dim3 grid((Wo + block.y - 1) / block.y, Ho, N);
  launch(upsample_fwd_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_upsample_bwd(const void* dout, int ld_dout, void* din, int ld_din, int accumulate, int N,
                     int hin, int win, int hu, int wu, int Ho, int Wo, int C, void* stream) {
  if (C <= 0 || C % 8 != 0 || Ho < hu || Wo < wu || N <= 0 || hin <= 0 || win <= 0 || hu <= 0 || wu <= 0)
    return UB2_ERR_SHAPE;
  if (ld_dout % 8 || ld_din % 8) return UB2_ERR_ALIGN;
  UpGeom g = up_geom(N, hin, win, hu, wu, Ho, Wo, C);
  if (bwd_strip_ok(g)) {
    const dim3 sblock = up_block(g.cgs);
    const long long wave = static_cast<long long>(num_sms()) * 2048;
    int strip = kBwdStrip;
    while (strip > 2 && static_cast<long long>(N) * ((hin + strip - 1) / strip) * win * g.cgs < wave) strip >>= 1;
    const dim3 sgrid((win + sblock.y - 1) / sblock.y, (hin + strip - 1) / strip, N);
    launch(upsample_bwd_strip_kernel, sgrid, sblock, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(dout), ld_dout, static_cast<__nv_bfloat16*>(din), ld_din, accumulate, strip, g);
    return static_cast<int>(cudaGetLastError());
  }
  if (bwd_tiled_ok(g)) {
    const int slabs = C / 16;
    const dim3 tgrid((win + kBT_W - 1) / kBT_W, (hin + kBT_H - 1) / kBT_H, N * slabs);
    launch(upsample_bwd_tiled_kernel, tgrid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(dout), ld_dout, static_cast<__nv_bfloat16*>(din), ld_din, accumulate, slabs, g);
    return static_cast<int>(cudaGetLastError());
  }
  const dim3 block = up_block(g.cgs);
  const dim3 grid((win + block.y - 1) / block.y, hin, N);
  launch(upsample_bwd_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(dout), ld_dout, static_cast<__nv_bfloat16*>(din), ld_din, accumulate, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_resize_planes_fwd(const float* in, float* out, int planes, int hin, int win, int Ho, int Wo, void* stream) {
  if (planes <= 0 || hin <= 0 || win <= 0 || Ho <= 0 || Wo <= 0) return UB2_ERR_SHAPE;
  const long long total = static_cast<long long>(planes) * Ho * Wo;
  launch(resize_planes_fwd_kernel, stream_grid(total, 256, num_sms(), 8), 256, 0, static_cast<cudaStream_t>(stream), in, out, planes, make_lowres(hin, win, Ho, Wo));
  return static_cast<int>(cudaGetLastError());
}

int ub2_resize_planes_bwd(const float* dout, float* din, int planes, int hin, int win, int Ho, int Wo,
                          void* stream) {
  if (planes <= 0 || hin <= 0 || win <= 0 || Ho <= 0 || Wo <= 0) return UB2_ERR_SHAPE;
  const long long total = static_cast<long long>(planes) * hin * win;
  launch(resize_planes_bwd_kernel, stream_grid(total, 256, num_sms(), 8), 256, 0, static_cast<cudaStream_t>(stream), dout, din, planes, make_lowres(hin, win, Ho, Wo));
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
