// Probe: can a tcgen05.mma A descriptor start at an arbitrary 128-byte row of a TMA-written
// 128B-swizzled tile (i.e. is the swizzle a function of absolute smem address bits)?
// D[128 x 64] = A[shift : shift+128, 0:64] * B[64 x 64]^T for several shifts and base_offset modes.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "../../unet-segment-pytorch_b200/csrc/ptx.cuh"
using namespace ub2;

constexpr int kRows = 160;

__global__ void probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      float* out, int shift, int mode) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;                    // kRows x 128 B
  uint8_t* sb = base + 24 * 1024;        // 64 x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 40 * 1024);
  uint32_t* tm = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tm, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tm;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], kRows * 128 + 64 * 128);
    tma_load_2d(sa, &tmA, &bars[0], 0, 0);
    tma_load_2d(sb, &tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t a_addr = smem_u32(sa) + shift * 128;
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024, 2);
      if (mode == 1) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      const uint64_t db = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024, 2);
      umma_bf16(tmem, da, db, idesc, k != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int j = 0; j < 2; ++j) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + j * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + j * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

static float bf(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t tobf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return uint16_t(u >> 16); }

int main() {
  std::vector<uint16_t> A(kRows * 64), B(64 * 64);
  srand(1);
  for (auto& v : A) v = tobf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = tobf((rand() % 2001 - 1000) / 1000.f);
  uint16_t *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap tA, tB;
  cuuint64_t dimsA[2] = {64, kRows}, dimsB[2] = {64, 64}, str[1] = {128};
  cuuint32_t boxA[2] = {64, kRows}, boxB[2] = {64, 64}, es[2] = {1, 1};
  enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, str, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> O(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 16, 31};
  for (int mode = 0; mode < 2; ++mode)
    for (int s : shifts) {
      probe<<<1, 128, 64 * 1024>>>(tA, tB, dO, s, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, s, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += bf(A[(m + s) * 64 + k]) * bf(B[n * 64 + k]);
          maxerr = std::max(maxerr, std::fabs(ref - O[m * 64 + n]));
        }
      printf("mode %d (base_offset %s) shift %2d: max err %.4g %s\n", mode, mode ? "=(addr>>7)&7" : "=0", s, maxerr,
             maxerr < 1e-2 ? "OK" : "MISMATCH");
    }
  return 0;
}
