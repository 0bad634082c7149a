// Parameter layout conversion at the boundary.  Parameters stay torch-owned fp32 tensors in
// the reference's OIHW layout (the checkpoint wire format); the tensor-core
// kernels read bf16 packs that are rebuilt from them:
//   fwd pack   (Cout, taps, Cin)         : K-major B operand of the forward implicit GEMM
//   dgrad pack (Cin, taps flipped, Cout) : the same kernel computes the data gradient
// and the split-K partial weight gradients are folded back into the OIHW fp32 .grad.
#include "../../include/unetb200.h"
#include "conv.h"
#include "vec.cuh"

namespace ub2 {

__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                   __nv_bfloat16* __restrict__ dgrad, int Cout, int Cin, int taps,
                                   const float* __restrict__ out_scale) {
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // i indexes OIHW: ((co*Cin + ci)*taps + t)
    const int t = static_cast<int>(i % taps);
    const int ci = static_cast<int>((i / taps) % Cin);
    const int co = static_cast<int>(i / (static_cast<long long>(taps) * Cin));
    float v = __ldg(w + i);
    if (out_scale != nullptr) v *= __ldg(out_scale + co);  // BatchNorm folded into the weights (eval)
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    if (fwd != nullptr) fwd[(static_cast<size_t>(co) * taps + t) * Cin + ci] = b;
    if (dgrad != nullptr) dgrad[(static_cast<size_t>(ci) * taps + (taps - 1 - t)) * Cout + co] = b;
  }
}

// grad[co][ci][t] += sum_s partial[s][t*Cin + ci][co];  blockDim = (32, 8): x walks the partial
// layout (coalesced), y strides over the splits; fixed summation order (deterministic).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cout, int Cin,
                                    int taps, float* __restrict__ grad) {
  __shared__ float s_red[8][33];
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  const long long i = static_cast<long long>(blockIdx.x) * 32 + threadIdx.x;
  float a = 0.f;
  if (i < total) {
    for (int s = threadIdx.y; s < splits; s += 8) a += __ldg(partial + s * total + i);
  }
  s_red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && i < total) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += s_red[y][threadIdx.x];
    const int co = static_cast<int>(i % Cout);
    const long long m = i / Cout;
    const int ci = static_cast<int>(m % Cin);
    const int tp = static_cast<int>(m / Cin);
    grad[(static_cast<size_t>(co) * Cin + ci) * taps + tp] += t;
  }
}

}  // namespace ub2

using namespace ub2;

extern "C" {

int ub2_version(void) { return 100; }

int ub2_pack_conv_weight(const float* w, void* fwd, void* dgrad, int Cout, int Cin, int taps,
                         const float* out_scale, void* stream) {
  if (Cout <= 0 || Cin <= 0 || (taps != 1 && taps != 9)) return UB2_ERR_SHAPE;
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  pack_weight_kernel<<<stream_grid(total, 256, num_sms(), 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(fwd), static_cast<__nv_bfloat16*>(dgrad), Cout, Cin, taps, out_scale);
  return static_cast<int>(cudaGetLastError());
}

int ub2_wgrad_reduce(const float* partial, int splits, int Cout, int Cin, int taps, float* grad,
                     void* stream) {
  if (Cout <= 0 || Cin <= 0 || splits <= 0) return UB2_ERR_SHAPE;
  const long long total = static_cast<long long>(Cout) * Cin * taps;
  wgrad_reduce_kernel<<<static_cast<unsigned>((total + 31) / 32), dim3(32, 8), 0,
                        static_cast<cudaStream_t>(stream)>>>(partial, splits, Cout, Cin, taps, grad);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
