// Probe: tcgen05.mma kind::tf32 with MN-major (channel-contiguous) operands, as the TF32 weight gradient
// needs: D[128 x 64] = sum_k A[k][m] * B[k][n], A = (64 pixels x 128 channels) fp32, B = (64 x 64) fp32, both
// written by TMA as 32-channel (128-byte) sub-boxes with the 128-byte swizzle.  Variants of the descriptor
// fields are tried; the one that reproduces the host result is what conv_wgrad.cu uses.
// Measured on B200 (sm_100a, CUDA 12.9, driver 580): with the MN-major (transpose) bits set, kind::tf32 returns ZEROS for
// both LBO/SBO assignments (no fault); without them it runs as K-major (a mismatch here, by construction).  kind::tf32 has
// no MN-major operand mode on this part — the fp32 mode's weight gradient therefore uses bf16 hi/lo splits (fp32_train.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/umma_tf32_mn tools/probe/umma_tf32_mn.cu -lcuda
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "../../unet-segment-pytorch_b200/csrc/ptx.cuh"
using namespace ub2;

constexpr int kK = 64, kM = 128, kN = 64;

__global__ void probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
                      int variant) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;                    // 4 sub-boxes x (64 pixels x 128 B) = 32 KB
  uint8_t* sb = base + 32 * 1024;        // 2 sub-boxes x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 48 * 1024);
  uint32_t* tm = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tm, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tm;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 48 * 1024);
    for (int j = 0; j < 4; ++j) tma_load_2d(sa + j * 8192, &tmA, &bars[0], j * 32, 0);
    for (int j = 0; j < 2; ++j) tma_load_2d(sb + j * 8192, &tmB, &bars[0], j * 32, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    uint32_t idesc = make_idesc_tf32(kM, kN);
    if (variant != 2) idesc |= (1u << 15) | (1u << 16);
    uint32_t lbo = 8192, sbo = 1024;
    if (variant == 1) { lbo = 1024; sbo = 8192; }
    for (int k = 0; k < kK / 8; ++k) {
      const uint64_t da = make_smem_desc(smem_u32(sa) + k * 1024, lbo, sbo, 2);
      const uint64_t db = make_smem_desc(smem_u32(sb) + k * 1024, lbo, sbo, 2);
      umma_tf32(tmem, da, db, idesc, k != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int j = 0; j < 2; ++j) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + j * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * kN + j * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

static float tf32r(float f) { uint32_t u; memcpy(&u, &f, 4); u = (u + 0x1000) & ~0x1FFFu; memcpy(&f, &u, 4); return f; }

int main() {
  std::vector<float> A(kK * kM), B(kK * kN);
  srand(1);
  for (auto& v : A) v = tf32r((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = tf32r((rand() % 2001 - 1000) / 1000.f);
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, kM * kN * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap tA, tB;
  cuuint64_t dimsA[2] = {kM, kK}, dimsB[2] = {kN, kK}, strA[1] = {kM * 4}, strB[1] = {kN * 4};
  cuuint32_t box[2] = {32, kK}, es[2] = {1, 1};
  CUresult r1 = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dimsA, strA, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, strB, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d %d\n", (int)r1, (int)r2);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> O(kM * kN);
  for (int variant = 0; variant < 3; ++variant) {
    cudaMemset(dO, 0, kM * kN * 4);
    probe<<<1, 128, 64 * 1024>>>(tA, tB, dO, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxabs = 0, refabs = 0;
    for (int m = 0; m < kM; ++m)
      for (int n = 0; n < kN; ++n) {
        double ref = 0;
        for (int k = 0; k < kK; ++k) ref += double(A[k * kM + m]) * B[k * kN + n];
        maxerr = std::max(maxerr, std::fabs(ref - O[m * kN + n]));
        maxabs = std::max(maxabs, (double)std::fabs(O[m * kN + n]));
        refabs = std::max(refabs, std::fabs(ref));
      }
    printf("variant %d (%s): max err %.4g, max |out| %.4g, max |ref| %.4g  %s\n", variant,
           variant == 0 ? "LBO=sub-box stride, SBO=1024, MN-major bits" : variant == 1 ? "LBO/SBO swapped" : "no MN-major bits",
           maxerr, maxabs, refabs, maxerr < 1e-3 ? "OK" : "MISMATCH");
    printf("   out[0][0..3] = %.4f %.4f %.4f %.4f\n", O[0], O[1], O[2], O[3]);
  }
  return 0;
}
