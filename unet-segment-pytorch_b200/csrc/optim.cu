// Fused optimizer tail of the training step (scripts/train.py:139-143 with the optimizer built at
// train.py:346-350): torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) followed by
// torch.optim.AdamW.step(), as two multi-tensor launches over a pointer table instead of ~25
// foreach passes.  HBM bound: 16 B read + 12 B written per parameter (+4 B for the norm pass).
//
//   total = sqrt(sum g^2);  c = min(1, max_norm / (total + 1e-6))            (clip_grad_norm_)
//   g' = c*g;  p *= 1 - lr*wd;  m += (g' - m)(1-b1);  v = b2*v + (1-b2) g'^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)                  (AdamW, amsgrad=False)
//
// Everything the host would have to know per step (step count, learning rate) lives in device
// memory, so the step can sit inside a captured CUDA graph while a scheduler changes the rate.
#include "launch.cuh"
#include "../../include/unetb200.h"
#include "conv.h"
#include "vec.cuh"

namespace ub2 {

static constexpr int kOptThreads = 256;

__device__ __forceinline__ double block_sum_double(double v, double* smem /* [8] */) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) t += smem[w];
    smem[0] = t;
  }
  __syncthreads();
  t = smem[0];
  __syncthreads();
  return t;
}

// chunks[b] = {tensor index, first element}; one block per chunk.
__global__ void __launch_bounds__(kOptThreads)
grad_sumsq_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ numel,
                  const int2* __restrict__ chunks, int chunk_elems, int T, double* __restrict__ partial,
                  float* __restrict__ step) {
  pdl_trigger();
  pdl_wait();
  __shared__ double s_red[kOptThreads / 32];
  const int2 ch = chunks[blockIdx.x];
  const float* g = reinterpret_cast<const float*>(ptrs[1 * T + ch.x]);
  const long long n = numel[ch.x];
  const long long begin = ch.y;
  const long long end = (begin + chunk_elems < n) ? begin + chunk_elems : n;
  double acc = 0.0;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long vend = begin + ((end - begin) & ~3LL);
    for (long long i = begin + threadIdx.x * 4LL; i < vend; i += kOptThreads * 4LL) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += static_cast<double>(v.x * v.x + v.y * v.y) + static_cast<double>(v.z * v.z + v.w * v.w);
    }
    for (long long i = vend + threadIdx.x; i < end; i += kOptThreads) acc += static_cast<double>(g[i] * g[i]);
  } else {
    for (long long i = begin + threadIdx.x; i < end; i += kOptThreads) acc += static_cast<double>(g[i] * g[i]);
  }
  const double t = block_sum_double(acc, s_red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = t;
    if (blockIdx.x == 0) step[0] += 1.f;  // the optimizer step counter (read by adamw_kernel)
  }
}

struct AdamCoef {
  float clip, decay, b1, b2, one_m_b1, one_m_b2, step_size, inv_sqrt_bc2, eps;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamCoef& c) {
  g *= c.clip;
  p *= c.decay;
  m += (g - m) * c.one_m_b1;
  v = v * c.b2 + c.one_m_b2 * g * g;
  const float denom = sqrtf(v) * c.inv_sqrt_bc2 + c.eps;
  p -= c.step_size * (m / denom);
}

// hyper rows (one per parameter group): {lr, beta1, beta2, eps, weight_decay, max_norm, -, -};
// max_norm is taken from row 0 (the clip is global, <= 0 disables it).
__global__ void __launch_bounds__(kOptThreads)
adamw_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ numel,
             const int* __restrict__ group, const int2* __restrict__ chunks, int nchunks, int chunk_elems,
             int T, const double* __restrict__ partial, const float* __restrict__ hyper,
             const float* __restrict__ step, float* __restrict__ total_norm, int write_grads) {
  pdl_trigger();
  pdl_wait();
  __shared__ double s_red[kOptThreads / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nchunks; i += kOptThreads) acc += partial[i];
  const float norm = static_cast<float>(sqrt(block_sum_double(acc, s_red)));
  if (blockIdx.x == 0 && threadIdx.x == 0 && total_norm != nullptr) total_norm[0] = norm;

  const int2 ch = chunks[blockIdx.x];
  const int t = ch.x;
  const float* h = hyper + 8 * group[t];
  const float max_norm = hyper[5];
  AdamCoef c;
  c.clip = 1.f;
  if (max_norm > 0.f) c.clip = fminf(max_norm / (norm + 1e-6f), 1.f);
  const float lr = h[0];
  c.b1 = h[1];
  c.b2 = h[2];
  c.eps = h[3];
  c.decay = 1.f - lr * h[4];
  c.one_m_b1 = 1.f - c.b1;
  c.one_m_b2 = 1.f - c.b2;
  const float st = step[0];
  c.step_size = lr / (1.f - powf(c.b1, st));
  c.inv_sqrt_bc2 = rsqrtf(1.f - powf(c.b2, st));

  float* p = reinterpret_cast<float*>(ptrs[0 * T + t]);
  float* g = reinterpret_cast<float*>(ptrs[1 * T + t]);
  float* m = reinterpret_cast<float*>(ptrs[2 * T + t]);
  float* v = reinterpret_cast<float*>(ptrs[3 * T + t]);
  const long long n = numel[t];
  const long long begin = ch.y;
  const long long end = (begin + chunk_elems < n) ? begin + chunk_elems : n;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  long long vend = begin;
  if (aligned) {
    vend = begin + ((end - begin) & ~3LL);
    for (long long i = begin + threadIdx.x * 4LL; i < vend; i += kOptThreads * 4LL) {
      float4 pp = *reinterpret_cast<float4*>(p + i);
      float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      adam_update(pp.x, gg.x, mm.x, vv.x, c);
      adam_update(pp.y, gg.y, mm.y, vv.y, c);
      adam_update(pp.z, gg.z, mm.z, vv.z, c);
      adam_update(pp.w, gg.w, mm.w, vv.w, c);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (write_grads) {
        gg.x *= c.clip; gg.y *= c.clip; gg.z *= c.clip; gg.w *= c.clip;
        *reinterpret_cast<float4*>(g + i) = gg;
      }
    }
  }
  for (long long i = vend + threadIdx.x; i < end; i += kOptThreads) {
    float pp = p[i], mm = m[i], vv = v[i];
    const float gg = g[i];
    adam_update(pp, gg, mm, vv, c);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
    if (write_grads) g[i] = gg * c.clip;
  }
}

// ModelEMA.update (unet/utils/general.py:155-184) as one multi-tensor launch: parameters
// ema = decay*ema + (1-decay)*p, buffers copied (fp32 running statistics, int64 counters).
// desc (T,4) int64 = {dst, src, numel, kind: 0 lerp fp32, 1 copy fp32, 2 copy int64}; decay from
// device memory (graph-capturable while the host ramps it up during warm-up).
__global__ void __launch_bounds__(kOptThreads)
ema_update_kernel(const long long* __restrict__ desc, const int2* __restrict__ chunks, int chunk_elems,
                  const float* __restrict__ decay_ptr) {
  pdl_trigger();
  pdl_wait();
  const int2 ch = chunks[blockIdx.x];
  const long long* d = desc + 4 * ch.x;
  const long long n = d[2];
  const long long begin = ch.y;
  const long long end = (begin + chunk_elems < n) ? begin + chunk_elems : n;
  const int kind = static_cast<int>(d[3]);
  if (kind == 2) {
    long long* dst = reinterpret_cast<long long*>(d[0]);
    const long long* src = reinterpret_cast<const long long*>(d[1]);
    for (long long i = begin + threadIdx.x; i < end; i += kOptThreads) dst[i] = src[i];
    return;
  }
  float* dst = reinterpret_cast<float*>(d[0]);
  const float* src = reinterpret_cast<const float*>(d[1]);
  if (kind == 1) {
    for (long long i = begin + threadIdx.x; i < end; i += kOptThreads) dst[i] = src[i];
    return;
  }
  const float decay = __ldg(decay_ptr);
  const float alpha = 1.f - decay;
  long long vend = begin;
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    vend = begin + ((end - begin) & ~3LL);
    for (long long i = begin + threadIdx.x * 4LL; i < vend; i += kOptThreads * 4LL) {
      float4 e = *reinterpret_cast<float4*>(dst + i);
      const float4 p = *reinterpret_cast<const float4*>(src + i);
      // mul_(decay).add_(p, alpha = 1 - decay): two roundings, as ATen
      e.x = e.x * decay + alpha * p.x;
      e.y = e.y * decay + alpha * p.y;
      e.z = e.z * decay + alpha * p.z;
      e.w = e.w * decay + alpha * p.w;
      *reinterpret_cast<float4*>(dst + i) = e;
    }
  }
  for (long long i = vend + threadIdx.x; i < end; i += kOptThreads) dst[i] = dst[i] * decay + alpha * src[i];
}

}  // namespace ub2

using namespace ub2;

extern "C" {

int ub2_adamw_chunk_elems(void) { return 16384; }

int ub2_grad_sumsq(const long long* ptrs, const long long* numel, const int* chunks, int nchunks, int T,
                   double* partial, float* step, void* stream) {
  if (nchunks <= 0 || T <= 0) return UB2_ERR_SHAPE;
  launch(grad_sumsq_kernel, nchunks, kOptThreads, 0, static_cast<cudaStream_t>(stream), ptrs, numel, reinterpret_cast<const int2*>(chunks), ub2_adamw_chunk_elems(), T, partial, step);
  return static_cast<int>(cudaGetLastError());
}

int ub2_adamw_step(const long long* ptrs, const long long* numel, const int* group, const int* chunks,
                   int nchunks, int T, const double* partial, const float* hyper, const float* step,
                   float* total_norm, int write_grads, void* stream) {
  if (nchunks <= 0 || T <= 0) return UB2_ERR_SHAPE;
  launch(adamw_kernel, nchunks, kOptThreads, 0, static_cast<cudaStream_t>(stream), ptrs, numel, group, reinterpret_cast<const int2*>(chunks), nchunks, ub2_adamw_chunk_elems(), T, partial, hyper, step, total_norm, write_grads);
  return static_cast<int>(cudaGetLastError());
}

int ub2_ema_update(const long long* desc, const int* chunks, int nchunks, const float* decay, void* stream) {
  if (nchunks <= 0) return UB2_ERR_SHAPE;
  launch(ema_update_kernel, nchunks, kOptThreads, 0, static_cast<cudaStream_t>(stream), desc, reinterpret_cast<const int2*>(chunks), ub2_adamw_chunk_elems(), decay);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
