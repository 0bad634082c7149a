// Confusion-matrix accumulation for SegmentationMetrics.update (unet/utils/metrics.py:55-84).
//
// The reference moves predictions and targets to the host and increments cm[t, p] in a
// per-pixel Python loop.  Here one pass over the device tensors builds an int64
// (C+1) x (C+1) histogram: row/column C collects everything the reference skips (labels or
// predictions outside [0, C), or target == ignore_index), so compute_iou / compute_dice
// (metrics.py:160-227), which count such pixels on the prediction side, can be derived from
// the same matrix.  Integer arithmetic: results are bit-exact and order independent.
// argmax follows torch: the first maximum wins (strict >).
#include "launch.cuh"
#include "../../include/unetb200.h"
#include "conv.h"
#include "vec.cuh"

namespace ub2 {

static constexpr int kCmThreads = 256;
static constexpr int kCmMaxC = 32;

// MODE 0: fp32 logits (N,C,H,W); MODE 1: int64 class indices (N,H,W);
// MODE 2: fp32 logits with a probability threshold on class 1 of a 2-class softmax
//         (predict.py:155-159: softmax(z)[1] > thr  <=>  sigmoid(z1 - z0) > thr)
template <int MODE>
__global__ void __launch_bounds__(kCmThreads)
confusion_kernel(const void* __restrict__ pred, const long long* __restrict__ target, int N, int C,
                 long long HW, long long ignore_index, int has_ignore, float threshold,
                 unsigned long long* cm, unsigned char* mask_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ unsigned int s_cm[(kCmMaxC + 1) * (kCmMaxC + 1)];
  const int B = C + 1;
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) s_cm[i] = 0u;
  __syncthreads();
  const long long total = static_cast<long long>(N) * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long p;
    if (MODE == 1) {
      p = __ldg(static_cast<const long long*>(pred) + i);
    } else {
      const long long n = i / HW, r = i % HW;
      const float* z = static_cast<const float*>(pred) + n * C * HW + r;
      if (MODE == 2) {
        // softmax(dim=1)[1] exactly as F.softmax forms it: exp(z - max) / sum
        const float z0 = __ldg(z), z1 = __ldg(z + HW);
        const float m = fmaxf(z0, z1);
        const float e0 = expf(z0 - m), e1 = expf(z1 - m);
        const float prob = e1 / (e0 + e1);
        p = (prob > threshold) ? 1 : 0;
        if (mask_out != nullptr) mask_out[i] = p ? 255 : 0;
      } else {
        float best = __ldg(z);
        p = 0;
        for (int c = 1; c < C; ++c) {
          const float v = __ldg(z + c * HW);
          if (v > best) {
            best = v;
            p = c;
          }
        }
      }
    }
    long long t = __ldg(target + i);
    if (has_ignore && t == ignore_index) t = C;
    if (t < 0 || t >= C) t = C;
    if (p < 0 || p >= C) p = C;
    // warp-aggregated increment: one shared atomic per distinct bin per warp
    const int bin = static_cast<int>(t) * B + static_cast<int>(p);
    const unsigned peers = __match_any_sync(__activemask(), bin);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cm[bin], static_cast<unsigned>(__popc(peers)));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * B; i += blockDim.x)
    if (s_cm[i] != 0u) atomicAdd(&cm[i], static_cast<unsigned long long>(s_cm[i]));
}

}  // namespace ub2

using namespace ub2;

extern "C" {

// mode: 0 logits argmax, 1 class indices, 2 thresholded 2-class softmax (mask_out optional uint8)
int ub2_confusion(const void* pred, const long long* target, int mode, int N, int C, long long HW,
                  long long ignore_index, int has_ignore, float threshold, long long* cm,
                  unsigned char* mask_out, void* stream) {
  if (C < 1 || C > kCmMaxC || N <= 0 || HW <= 0) return UB2_ERR_SHAPE;
  if (mode == 2 && C != 2) return UB2_ERR_SHAPE;
  const long long total = static_cast<long long>(N) * HW;
  // a block's shared counters are 32-bit: keep its share of the pixels far below 2^32
  int grid = stream_grid(total, kCmThreads, num_sms(), 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long* out = reinterpret_cast<unsigned long long*>(cm);
  if (mode == 0)
    launch(confusion_kernel<0>, grid, kCmThreads, 0, s, pred, target, N, C, HW, ignore_index, has_ignore, threshold, out, mask_out);
  else if (mode == 1)
    launch(confusion_kernel<1>, grid, kCmThreads, 0, s, pred, target, N, C, HW, ignore_index, has_ignore, threshold, out, mask_out);
  else
    launch(confusion_kernel<2>, grid, kCmThreads, 0, s, pred, target, N, C, HW, ignore_index, has_ignore, threshold, out, mask_out);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
