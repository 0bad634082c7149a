// Attention gate (AttentionGate.forward, unet/models/layers.py:171-192) as bandwidth-bound
// passes around the two tcgen05 1x1 projections:
//
//   q  = W_g . g            (low resolution: the 1x1 conv commutes with bilinear resampling)
//   xp = W_x . x            (full resolution, statistics in the conv epilogue)
//   gate_upstats : batch statistics of up(q) for BN_g          (layers.py:153,183,186)
//   gate_psi     : psi_raw = w_psi . relu(BN_g(up q) + BN_x(xp)) + its statistics (:188, :164)
//   gate_apply   : a = sigmoid(BN_psi(psi_raw)); out = x * a   (:165-166, :192)
// and the matching backward passes.  The three train-mode BatchNorms force three
// grid-wide reductions, hence three phases; in eval mode the same kernels run with
// folded running statistics and no reductions.
//
// Thread mapping: `tpp` (power of two <= 32) consecutive threads share a pixel and split its
// 8-channel vectors; per-pixel channel reductions are shuffle reductions inside that group,
// per-channel pixel reductions stay in registers (a thread's channels are fixed) and are
// combined per block in shared memory, then across blocks by a finalize kernel in double.
#include "launch.cuh"
#include "../../include/unetb200.h"
#include "conv.h"
#include "resample.cuh"
#include "vec.cuh"

namespace ub2 {

static constexpr int kGateThreads = 256;
static constexpr int kMaxG = 2;  // 8-channel groups per thread (channels <= 512)

struct GateGeom {
  int N, H, W, C, cgs, tpp, slots;
  int pixels;
  int wt, ht;   // strip kernels: column tiles of `slots` pixels x row strips of kGateStrip rows
  LowRes lr;
};
static constexpr int kGateStrip = 16;

static int gate_tpp(int cgs) {
  int t = 1;
  while (t < cgs && t < 32) t *= 2;
  return t;
}
static int make_gate_geom(GateGeom* g, int N, int H, int W, int C, int hin, int win) {
  if (C <= 0 || C % 8 != 0 || N <= 0 || H <= 0 || W <= 0) return UB2_ERR_SHAPE;
  if (static_cast<double>(N) * H * W * (C / 8) >= 2.0e9) return UB2_ERR_SHAPE;  // 32-bit pixel indices
  g->N = N; g->H = H; g->W = W; g->C = C; g->cgs = C / 8;
  g->tpp = gate_tpp(g->cgs);
  if ((g->cgs + g->tpp - 1) / g->tpp > kMaxG) return UB2_ERR_SHAPE;
  g->slots = kGateThreads / g->tpp;
  g->pixels = static_cast<int>(N) * H * W;
  g->wt = (W + g->slots - 1) / g->slots;
  g->ht = (H + kGateStrip - 1) / kGateStrip;
  g->lr = make_lowres(hin > 0 ? hin : 1, win > 0 ? win : 1, H, W);
  return 0;
}
static int gate_strip_grid(const GateGeom& g) { return g.wt * g.ht * g.N; }
static int gate_grid(const GateGeom& g, int per_sm) {
  return stream_grid(g.pixels, g.slots, num_sms(), per_sm);
}

__device__ __forceinline__ float group_sum(float v, int tpp) {
  for (int o = tpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduce per-thread channel accumulators over the pixel slots of a block and write one
// row of doubles: out_row[ns*C + channel].
template <int G, int NS>
__device__ __forceinline__ void block_reduce_channels(float (&acc)[G][NS][8], const GateGeom& g,
                                                      int slot, int j, double* out_row, float* smem) {
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
#pragma unroll
    for (int ns = 0; ns < NS; ++ns) {
#pragma unroll
      for (int k = 0; k < 8; ++k) smem[(slot * g.tpp + j) * 8 + k] = acc[gi][ns][k];
      __syncthreads();
      for (int idx = threadIdx.x; idx < g.tpp * 8; idx += blockDim.x) {
        const int jj = idx >> 3, k = idx & 7;
        const int cg = jj + gi * g.tpp;
        if (cg < g.cgs) {
          double s = 0.0;
          for (int sl = 0; sl < g.slots; ++sl) s += static_cast<double>(smem[(sl * g.tpp + jj) * 8 + k]);
          out_row[static_cast<size_t>(ns) * g.C + cg * 8 + k] = s;
        }
      }
      __syncthreads();
    }
  }
}

// Same for per-thread scalars (NS values), written as out_row[ns].
template <int NS>
__device__ __forceinline__ void block_reduce_scalars(float (&acc)[NS], double* out_row, float* smem) {
#pragma unroll
  for (int ns = 0; ns < NS; ++ns) {
    const float w = warp_sum(acc[ns]);
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < (blockDim.x >> 5); ++i) s += static_cast<double>(smem[i]);
      out_row[ns] = s;
    }
    __syncthreads();
  }
}

// Warp-uniform trip count (the shuffle reductions need every lane): `pv` masks the tail.
#define GATE_PIXEL_LOOP(g)                                                                  \
  const int slot = threadIdx.x / (g).tpp;                                                   \
  const int j = threadIdx.x % (g).tpp;                                                      \
  for (int base = static_cast<int>(blockIdx.x) * (g).slots; base < (g).pixels;  \
       base += static_cast<int>(gridDim.x) * (g).slots)

#define GATE_PIX(g)                     \
  const int pix = base + slot;    \
  const bool pv = pix < (g).pixels;

#define GATE_DECODE(g)                                                        \
  const int wo = static_cast<int>(pix % (g).W);                               \
  const int ho = static_cast<int>((pix / (g).W) % (g).H);                     \
  const int n = static_cast<int>(pix / (static_cast<int>((g).W) * (g).H));

// Strip walk for the passes that need up(q): a block owns `slots` columns x kGateStrip rows of one
// image, a thread one column (x `tpp` lanes over the channels) and walks the rows.  No per-pixel
// integer division, and the two horizontally interpolated low-resolution rows a destination row lies
// between stay in registers (they change every other row for a 2x up-sampling).  ncu on the
// per-pixel form: ~320 instructions per pixel-vector, half of them integer address arithmetic.
template <int G>
struct QRoll {
  int cur0, cur1, w0, w1;
  float a0, a1, b0, b1;
  F8 top[G], bot[G];
};
template <int G>
__device__ __forceinline__ void qroll_init(QRoll<G>& r, const GateGeom& g, int wo) {
  r.cur0 = r.cur1 = -1;
  src_index(g.lr.rw, wo < g.W ? wo : g.W - 1, g.lr.win, r.w0, r.w1, r.b0, r.b1);
#pragma unroll
  for (int gi = 0; gi < G; ++gi)
#pragma unroll
    for (int k = 0; k < 8; ++k) r.top[gi].v[k] = r.bot[gi].v[k] = 0.f;
}
template <int G>
__device__ __forceinline__ void qroll_row(QRoll<G>& r, const __nv_bfloat16* __restrict__ q, int ld_q,
                                          const GateGeom& g, int n, int ho, int j, bool colv) {
  int h0, h1;
  src_index(g.lr.rh, ho, g.lr.hin, h0, h1, r.a0, r.a1);   // block-uniform
  const __nv_bfloat16* base = q + static_cast<size_t>(n) * g.lr.hin * g.lr.win * ld_q;
  auto hrow = [&](int h, int gi) {
    F8 o;
    const int cg = j + gi * g.tpp;
    if (colv && cg < g.cgs) {
      const F8 a = load8(base + (static_cast<size_t>(h) * g.lr.win + r.w0) * ld_q + cg * 8);
      const F8 b = load8(base + (static_cast<size_t>(h) * g.lr.win + r.w1) * ld_q + cg * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = r.b0 * a.v[k] + r.b1 * b.v[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
    }
    return o;
  };
  if (h0 != r.cur0) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) r.top[gi] = (h0 == r.cur1) ? r.bot[gi] : hrow(h0, gi);
    r.cur0 = h0;
  }
  if (h1 != r.cur1) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) r.bot[gi] = (h1 == r.cur0) ? r.top[gi] : hrow(h1, gi);
    r.cur1 = h1;
  }
}
template <int G>
__device__ __forceinline__ F8 qroll_get(const QRoll<G>& r, int gi) {
  F8 u;
#pragma unroll
  for (int k = 0; k < 8; ++k) u.v[k] = r.a0 * r.top[gi].v[k] + r.a1 * r.bot[gi].v[k];
  return u;
}

// for (row of the strip) { pix, pv, u = up(q) at this pixel via qroll_get(roll, gi) }
#define GATE_STRIP_LOOP(g, G_)                                                               \
  const int slot = threadIdx.x / (g).tpp;                                                    \
  const int j = threadIdx.x % (g).tpp;                                                       \
  const int bw_ = static_cast<int>(blockIdx.x) % (g).wt;                                     \
  const int bh_ = (static_cast<int>(blockIdx.x) / (g).wt) % (g).ht;                          \
  const int n = static_cast<int>(blockIdx.x) / ((g).wt * (g).ht);                            \
  const int wo = bw_ * (g).slots + slot;                                                     \
  const bool pv = wo < (g).W;                                                                \
  QRoll<G_> roll;                                                                            \
  qroll_init<G_>(roll, g, wo);                                                               \
  const int ho_end_ = min((bh_ + 1) * kGateStrip, (g).H);                                    \
  for (int ho = bh_ * kGateStrip; ho < ho_end_; ++ho)

#define GATE_STRIP_ROW(g, G_)                                                                \
  const int pix = (n * (g).H + ho) * (g).W + (pv ? wo : 0);                                  \
  qroll_row<G_>(roll, q, ld_q, g, n, ho, j, pv);

// ------------------------------------------------------------------------------ forward
template <int G>
__global__ void __launch_bounds__(kGateThreads)
gate_upstats_kernel(const __nv_bfloat16* __restrict__ q, int ld_q, double* partials, GateGeom g) {
  pdl_trigger();
  pdl_wait();
  __shared__ float smem[kGateThreads * 8];
  float acc[G][2][8];
#pragma unroll
  for (int a = 0; a < G; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][b][k] = 0.f;
  GATE_STRIP_LOOP(g, G) {
    GATE_STRIP_ROW(g, G)
    (void)pix;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int cg = j + gi * g.tpp;
      if (pv && cg < g.cgs) {
        const F8 u = qroll_get<G>(roll, gi);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[gi][0][k] += u.v[k];
          acc[gi][1][k] = fmaf(u.v[k], u.v[k], acc[gi][1][k]);
        }
      }
    }
  }
  const int slot2 = threadIdx.x / g.tpp, j2 = threadIdx.x % g.tpp;
  block_reduce_channels<G, 2>(acc, g, slot2, j2, partials + static_cast<size_t>(blockIdx.x) * 2 * g.C, smem);
}

// Per-channel coefficient vectors of the thread's first channel group, kept in registers
// (the thread -> channel-group mapping is fixed); further groups (Ci > 256) read through L1.
struct GateVec {
  F8 sg, sx, h, w;
};
__device__ __forceinline__ GateVec gate_vec(const float* sg, const float* hg, const float* sx,
                                            const float* hx, const float* wpsi, int cg, int cgs) {
  GateVec v;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = (cg < cgs ? cg : 0) * 8 + k;
    v.sg.v[k] = __ldg(sg + c);
    v.sx.v[k] = __ldg(sx + c);
    v.h.v[k] = __ldg(hg + c) + __ldg(hx + c);
    v.w.v[k] = __ldg(wpsi + c);
  }
  return v;
}

template <int G>
__global__ void __launch_bounds__(kGateThreads)
gate_psi_kernel(const __nv_bfloat16* __restrict__ q, int ld_q, const __nv_bfloat16* __restrict__ xp,
                int ld_xp, const float* __restrict__ sg, const float* __restrict__ hg,
                const float* __restrict__ sx, const float* __restrict__ hx,
                const float* __restrict__ wpsi, float* __restrict__ psi_raw, double* partials,
                GateGeom g) {
  pdl_trigger();
  pdl_wait();
  __shared__ float smem[kGateThreads / 32];
  float st[2] = {0.f, 0.f};
  const GateVec v0 = gate_vec(sg, hg, sx, hx, wpsi, threadIdx.x % g.tpp, g.cgs);
  GATE_STRIP_LOOP(g, G) {
    GATE_STRIP_ROW(g, G)
    float dot = 0.f;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int cg = j + gi * g.tpp;
      if (pv && cg < g.cgs) {
        const GateVec v = (gi == 0) ? v0 : gate_vec(sg, hg, sx, hx, wpsi, cg, g.cgs);
        const F8 u = qroll_get<G>(roll, gi);
        const F8 xv = load8_stream(xp + static_cast<size_t>(pix) * ld_xp + cg * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float t = fmaf(u.v[k], v.sg.v[k], fmaf(xv.v[k], v.sx.v[k], v.h.v[k]));
          dot = fmaf(v.w.v[k], fmaxf(t, 0.f), dot);
        }
      }
    }
    dot = group_sum(dot, g.tpp);
    if (pv && j == 0) {
      psi_raw[pix] = dot;
      st[0] += dot;
      st[1] = fmaf(dot, dot, st[1]);
    }
  }
  if (partials != nullptr) block_reduce_scalars<2>(st, partials + static_cast<size_t>(blockIdx.x) * 2, smem);
}

__global__ void __launch_bounds__(256)
gate_apply_kernel(const float* __restrict__ psi_raw, const float* __restrict__ spsi,
                  const float* __restrict__ hpsi, const __nv_bfloat16* __restrict__ x, int ld_x,
                  __nv_bfloat16* __restrict__ out, int ld_out, float* __restrict__ a_out,
                  int pixels, int cgs) {
  pdl_trigger();
  pdl_wait();
  const float s = __ldg(spsi), h = __ldg(hpsi);
  const int total = pixels * cgs;
  const int stride = static_cast<int>(gridDim.x) * blockDim.x;
  // four vectors per trip, loads first: the copy runs at HBM speed only with several 128-bit
  // requests in flight per thread
  for (int i0 = static_cast<int>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    uint4 raw[4];
    float pr[4];
    int pixv[4], cgv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * stride;
      const bool ok = i < total;
      cgv[u] = ok ? i % cgs : 0;
      pixv[u] = ok ? i / cgs : -1;
      raw[u] = make_uint4(0u, 0u, 0u, 0u);
      pr[u] = 0.f;
      if (ok) {
        raw[u] = ld_stream16(x + static_cast<size_t>(pixv[u]) * ld_x + cgv[u] * 8);
        pr[u] = __ldg(psi_raw + pixv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pixv[u] < 0) continue;
      const float z = fmaf(pr[u], s, h);
      const float a = 1.f / (1.f + __expf(-z));
      F8 v = unpack8(raw[u]);
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] *= a;
      store8(out + static_cast<size_t>(pixv[u]) * ld_out + cgv[u] * 8, v);
      if (cgv[u] == 0 && a_out != nullptr) a_out[pixv[u]] = a;
    }
  }
}

// ------------------------------------------------------------------------------ backward
// da = sum_c dOut_c x_c ; d(BN_psi out) = da * a (1-a) ; dx_direct = dOut * a
template <int G>
__global__ void __launch_bounds__(kGateThreads)
gate_bwd_a_kernel(const __nv_bfloat16* __restrict__ dout, int ld_do, const __nv_bfloat16* __restrict__ x,
                  int ld_x, const float* __restrict__ a, const float* __restrict__ psi_raw,
                  __nv_bfloat16* __restrict__ dx, int ld_dx, float* __restrict__ dpsin,
                  double* partials, GateGeom g) {
  pdl_trigger();
  pdl_wait();
  __shared__ float smem[kGateThreads / 32];
  float st[2] = {0.f, 0.f};  // sum dn, sum dn*psi_raw (bn_bwd_finalize convention)
  GATE_PIXEL_LOOP(g) {
    GATE_PIX(g)
    const float av = pv ? __ldg(a + pix) : 0.f;
    float dot = 0.f;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int cg = j + gi * g.tpp;
      if (pv && cg < g.cgs) {
        F8 d = load8_stream(dout + static_cast<size_t>(pix) * ld_do + cg * 8);
        const F8 xv = load8_stream(x + static_cast<size_t>(pix) * ld_x + cg * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          dot = fmaf(d.v[k], xv.v[k], dot);
          d.v[k] *= av;
        }
        store8(dx + static_cast<size_t>(pix) * ld_dx + cg * 8, d);
      }
    }
    dot = group_sum(dot, g.tpp);
    if (pv && j == 0) {
      const float dn = dot * av * (1.f - av);
      dpsin[pix] = dn;
      st[0] += dn;
      st[1] = fmaf(dn, __ldg(psi_raw + pix), st[1]);
    }
  }
  block_reduce_scalars<2>(st, partials + static_cast<size_t>(blockIdx.x) * 2, smem);
}

// ds_c = dpsi_raw * w_psi_c * [t_c > 0]; channel sums (raw, the finalize converts them):
// sum ds, sum ds*xp, sum ds*up(q), sum dpsi_raw*relu(t)
template <int G>
__global__ void __launch_bounds__(kGateThreads)
gate_bwd_s_kernel(const float* __restrict__ dpsin, const float* __restrict__ psi_raw,
                  const float* __restrict__ coef_psi, const __nv_bfloat16* __restrict__ q, int ld_q,
                  const __nv_bfloat16* __restrict__ xp, int ld_xp, const float* __restrict__ sg,
                  const float* __restrict__ hg, const float* __restrict__ sx,
                  const float* __restrict__ hx, const float* __restrict__ wpsi,
                  __nv_bfloat16* __restrict__ ds, int ld_ds, double* partials, GateGeom g) {
  pdl_trigger();
  pdl_wait();
  __shared__ float smem[kGateThreads * 8];
  const float cA = __ldg(coef_psi), cB = __ldg(coef_psi + 1), cC = __ldg(coef_psi + 2);
  const GateVec v0 = gate_vec(sg, hg, sx, hx, wpsi, threadIdx.x % g.tpp, g.cgs);
  float acc[G][4][8];
#pragma unroll
  for (int a = 0; a < G; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][b][k] = 0.f;
  GATE_STRIP_LOOP(g, G) {
    GATE_STRIP_ROW(g, G)
    // BN_psi backward: d psi_raw = A*dn + B*psi_raw + C
    const float dpr = pv ? fmaf(cA, __ldg(dpsin + pix), fmaf(cB, __ldg(psi_raw + pix), cC)) : 0.f;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int cg = j + gi * g.tpp;
      if (pv && cg < g.cgs) {
        const GateVec v = (gi == 0) ? v0 : gate_vec(sg, hg, sx, hx, wpsi, cg, g.cgs);
        const F8 u = qroll_get<G>(roll, gi);
        const F8 xv = load8_stream(xp + static_cast<size_t>(pix) * ld_xp + cg * 8);
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float t = fmaf(u.v[k], v.sg.v[k], fmaf(xv.v[k], v.sx.v[k], v.h.v[k]));
          const float d = (t > 0.f) ? dpr * v.w.v[k] : 0.f;
          o.v[k] = d;
          acc[gi][0][k] += d;
          acc[gi][1][k] = fmaf(d, xv.v[k], acc[gi][1][k]);
          acc[gi][2][k] = fmaf(d, u.v[k], acc[gi][2][k]);
          acc[gi][3][k] = fmaf(dpr, fmaxf(t, 0.f), acc[gi][3][k]);
        }
        store8(ds + static_cast<size_t>(pix) * ld_ds + cg * 8, o);
      }
    }
  }
  const int slot2 = threadIdx.x / g.tpp, j2 = threadIdx.x % g.tpp;
  block_reduce_channels<G, 4>(acc, g, slot2, j2, partials + static_cast<size_t>(blockIdx.x) * 4 * g.C, smem);
}

// Raw sums -> parameter gradients and the backward coefficients of the two BatchNorms:
//   dxp = coef0*ds + coef1*xp + coef2 ;  d(up q) = coef3*ds + coef4*up(q) + coef5
// blockDim = (8, 128): see rows_sum_wide in vec.cuh
__global__ void gate_bwd_finalize_kernel(const double* __restrict__ partials, int rows, int C,
                                         double count, const float* __restrict__ gamma_x,
                                         const float* __restrict__ mean_x,
                                         const float* __restrict__ invstd_x,
                                         const float* __restrict__ gamma_g,
                                         const float* __restrict__ mean_g,
                                         const float* __restrict__ invstd_g, int frozen, float* dgamma_x,
                                         float* dbeta_x, float* dgamma_g, float* dbeta_g, float* dwpsi,
                                         float* coef) {
  pdl_trigger();
  pdl_wait();
  __shared__ double smem[4 * 128 * 9];
  const int c = blockIdx.x * 8 + threadIdx.x;
  double s[4];
  rows_sum_wide<4>(partials, rows, C, c, s, smem);
  if (threadIdx.y != 0 || c >= C) return;
  const double mx = mean_x[c], ix = invstd_x[c], mg = mean_g[c], ig = invstd_g[c];
  const double db = s[0];
  const double dgx = ix * (s[1] - mx * s[0]);
  const double dgg = ig * (s[2] - mg * s[0]);
  if (dbeta_x) dbeta_x[c] += static_cast<float>(db);
  if (dgamma_x) dgamma_x[c] += static_cast<float>(dgx);
  if (dbeta_g) dbeta_g[c] += static_cast<float>(db);
  if (dgamma_g) dgamma_g[c] += static_cast<float>(dgg);
  if (dwpsi) dwpsi[c] += static_cast<float>(s[3]);
  const double Ax = gamma_x[c] * ix, Ag = gamma_g[c] * ig;
  const double Bx = frozen ? 0.0 : -Ax * ix * dgx / count;
  const double Bg = frozen ? 0.0 : -Ag * ig * dgg / count;
  coef[0 * C + c] = static_cast<float>(Ax);
  coef[1 * C + c] = static_cast<float>(Bx);
  coef[2 * C + c] = static_cast<float>(frozen ? 0.0 : -Ax * db / count - Bx * mx);
  coef[3 * C + c] = static_cast<float>(Ag);
  coef[4 * C + c] = static_cast<float>(Bg);
  coef[5 * C + c] = static_cast<float>(frozen ? 0.0 : -Ag * db / count - Bg * mg);
}

// dxp = BN_x backward of ds; dgup = BN_g backward of ds (full resolution, later up-sample^T)
template <int G>
__global__ void __launch_bounds__(kGateThreads)
gate_bwd_xg_kernel(const __nv_bfloat16* __restrict__ ds, int ld_ds, const __nv_bfloat16* __restrict__ xp,
                   int ld_xp, const __nv_bfloat16* __restrict__ q, int ld_q,
                   const float* __restrict__ coef, __nv_bfloat16* __restrict__ dxp, int ld_dxp,
                   __nv_bfloat16* __restrict__ dgup, int ld_dg, GateGeom g) {
  pdl_trigger();
  pdl_wait();
  F8 cf[6];
  {
    const int cg0 = threadIdx.x % g.tpp;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k) cf[r].v[k] = __ldg(coef + r * g.C + (cg0 < g.cgs ? cg0 : 0) * 8 + k);
  }
  GATE_STRIP_LOOP(g, G) {
    GATE_STRIP_ROW(g, G)
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const int cg = j + gi * g.tpp;
      if (pv && cg < g.cgs) {
        const F8 d = load8_stream(ds + static_cast<size_t>(pix) * ld_ds + cg * 8);
        const F8 xv = load8_stream(xp + static_cast<size_t>(pix) * ld_xp + cg * 8);
        const F8 u = qroll_get<G>(roll, gi);
        F8 ox, og;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (gi == 0) {
            ox.v[k] = fmaf(cf[0].v[k], d.v[k], fmaf(cf[1].v[k], xv.v[k], cf[2].v[k]));
            og.v[k] = fmaf(cf[3].v[k], d.v[k], fmaf(cf[4].v[k], u.v[k], cf[5].v[k]));
          } else {
            const int c = cg * 8 + k;
            ox.v[k] = fmaf(__ldg(coef + c), d.v[k], fmaf(__ldg(coef + g.C + c), xv.v[k], __ldg(coef + 2 * g.C + c)));
            og.v[k] = fmaf(__ldg(coef + 3 * g.C + c), d.v[k],
                           fmaf(__ldg(coef + 4 * g.C + c), u.v[k], __ldg(coef + 5 * g.C + c)));
          }
        }
        store8(dxp + static_cast<size_t>(pix) * ld_dxp + cg * 8, ox);
        store8(dgup + static_cast<size_t>(pix) * ld_dg + cg * 8, og);
      }
    }
  }
}

}  // namespace ub2

using namespace ub2;
typedef const __nv_bfloat16* cbf;
typedef __nv_bfloat16* bf;

extern "C" {

int ub2_gate_rows(int N, int H, int W, int C) {
  GateGeom g;
  int rc = make_gate_geom(&g, N, H, W, C, 1, 1);
  if (rc) return rc;
  return gate_grid(g, 8);
}

int ub2_gate_strip_rows(int N, int H, int W, int C) {
  GateGeom g;
  int rc = make_gate_geom(&g, N, H, W, C, 1, 1);
  if (rc) return rc;
  return gate_strip_grid(g);
}

int ub2_gate_upstats(const void* q, int ld_q, int N, int hin, int win, int H, int W, int Ci,
                     double* partials, int rows, void* stream) {
  GateGeom g;
  int rc = make_gate_geom(&g, N, H, W, Ci, hin, win);
  if (rc) return rc;
  const int grid = gate_strip_grid(g);
  if (grid != rows) return UB2_ERR_WORKSPACE;
  if (g.cgs > g.tpp)
    launch(gate_upstats_kernel<2>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), static_cast<cbf>(q), ld_q, partials, g);
  else
    launch(gate_upstats_kernel<1>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), static_cast<cbf>(q), ld_q, partials, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_gate_psi(const void* q, int ld_q, const void* xp, int ld_xp, const float* scale_g,
                 const float* shift_g, const float* scale_x, const float* shift_x, const float* wpsi,
                 float* psi_raw, double* partials, int rows, int N, int hin, int win, int H, int W,
                 int Ci, void* stream) {
  GateGeom g;
  int rc = make_gate_geom(&g, N, H, W, Ci, hin, win);
  if (rc) return rc;
  const int grid = gate_strip_grid(g);
  if (partials != nullptr && grid != rows) return UB2_ERR_WORKSPACE;
  if (g.cgs > g.tpp)
    launch(gate_psi_kernel<2>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), static_cast<cbf>(q), ld_q, static_cast<cbf>(xp), ld_xp, scale_g, shift_g, scale_x, shift_x, wpsi, psi_raw, partials, g);
  else
    launch(gate_psi_kernel<1>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), static_cast<cbf>(q), ld_q, static_cast<cbf>(xp), ld_xp, scale_g, shift_g, scale_x, shift_x, wpsi, psi_raw, partials, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_gate_apply(const float* psi_raw, const float* scale_psi, const float* shift_psi, const void* x,
                   int ld_x, void* out, int ld_out, float* a_out, int N, int H, int W, int Cx,
                   void* stream) {
  if (Cx % 8 != 0 || N <= 0 || static_cast<double>(N) * H * W * (Cx / 8) >= 2.0e9) return UB2_ERR_SHAPE;
  const int pixels = static_cast<int>(N) * H * W;
  launch(gate_apply_kernel, stream_grid(pixels * (Cx / 8), 256, num_sms()), 256, 0, static_cast<cudaStream_t>(stream), psi_raw, scale_psi, shift_psi, static_cast<cbf>(x), ld_x, static_cast<bf>(out), ld_out, a_out, pixels, Cx / 8);
  return static_cast<int>(cudaGetLastError());
}

int ub2_gate_bwd_a(const void* dout, int ld_do, const void* x, int ld_x, const float* a,
                   const float* psi_raw, void* dx, int ld_dx, float* dpsin, double* partials, int rows,
                   int N, int H, int W, int Cx, void* stream) {
  GateGeom g;
  int rc = make_gate_geom(&g, N, H, W, Cx, 1, 1);
  if (rc) return rc;
  const int grid = gate_grid(g, 8);
  if (grid != rows) return UB2_ERR_WORKSPACE;
  if (g.cgs > g.tpp)
    launch(gate_bwd_a_kernel<2>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), static_cast<cbf>(dout), ld_do, static_cast<cbf>(x), ld_x, a, psi_raw, static_cast<bf>(dx), ld_dx, dpsin, partials, g);
  else
    launch(gate_bwd_a_kernel<1>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), static_cast<cbf>(dout), ld_do, static_cast<cbf>(x), ld_x, a, psi_raw, static_cast<bf>(dx), ld_dx, dpsin, partials, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_gate_bwd_s(const float* dpsin, const float* psi_raw, const float* coef_psi, const void* q,
                   int ld_q, const void* xp, int ld_xp, const float* scale_g, const float* shift_g,
                   const float* scale_x, const float* shift_x, const float* wpsi, void* ds, int ld_ds,
                   double* partials, int rows, int N, int hin, int win, int H, int W, int Ci,
                   void* stream) {
  GateGeom g;
  int rc = make_gate_geom(&g, N, H, W, Ci, hin, win);
  if (rc) return rc;
  const int grid = gate_strip_grid(g);
  if (grid != rows) return UB2_ERR_WORKSPACE;
  if (g.cgs > g.tpp)
    launch(gate_bwd_s_kernel<2>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), dpsin, psi_raw, coef_psi, static_cast<cbf>(q), ld_q, static_cast<cbf>(xp), ld_xp, scale_g, shift_g, scale_x, shift_x, wpsi, static_cast<bf>(ds), ld_ds, partials, g);
  else
    launch(gate_bwd_s_kernel<1>, grid, kGateThreads, 0, static_cast<cudaStream_t>(stream), dpsin, psi_raw, coef_psi, static_cast<cbf>(q), ld_q, static_cast<cbf>(xp), ld_xp, scale_g, shift_g, scale_x, shift_x, wpsi, static_cast<bf>(ds), ld_ds, partials, g);
  return static_cast<int>(cudaGetLastError());
}

int ub2_gate_bwd_finalize(const double* partials, int rows, int Ci, double count, const float* gamma_x,
                          const float* mean_x, const float* invstd_x, const float* gamma_g,
                          const float* mean_g, const float* invstd_g, int frozen, float* dgamma_x,
                          float* dbeta_x, float* dgamma_g, float* dbeta_g, float* dwpsi, float* coef,
                          void* stream) {
  if (Ci <= 0 || rows <= 0) return UB2_ERR_SHAPE;
  launch(gate_bwd_finalize_kernel, (Ci + 7) / 8, dim3(8, 128), 0, static_cast<cudaStream_t>(stream), partials, rows, Ci, count, gamma_x, mean_x, invstd_x, gamma_g, mean_g, invstd_g, frozen, dgamma_x, dbeta_x, dgamma_g, dbeta_g, dwpsi, coef);
  return static_cast<int>(cudaGetLastError());
}

int ub2_gate_bwd_xg(const void* ds, int ld_ds, const void* xp, int ld_xp, const void* q, int ld_q,
                    const float* coef, void* dxp, int ld_dxp, void* dgup, int ld_dg, int N, int hin,
                    int win, int H, int W, int Ci, void* stream) {
  GateGeom g;
  int rc = make_gate_geom(&g, N, H, W, Ci, hin, win);
  if (rc) return rc;
  const int grid = gate_strip_grid(g);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g.cgs > g.tpp)
    launch(gate_bwd_xg_kernel<2>, grid, kGateThreads, 0, s, static_cast<cbf>(ds), ld_ds, static_cast<cbf>(xp), ld_xp, static_cast<cbf>(q), ld_q, coef, static_cast<bf>(dxp), ld_dxp, static_cast<bf>(dgup), ld_dg, g);
  else
    launch(gate_bwd_xg_kernel<1>, grid, kGateThreads, 0, s, static_cast<cbf>(ds), ld_ds, static_cast<cbf>(xp), ld_xp, static_cast<cbf>(q), ld_q, coef, static_cast<bf>(dxp), ld_dxp, static_cast<bf>(dgup), ld_dg, g);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
