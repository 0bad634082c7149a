// Host-side TMA tensor-map construction.  The driver's cuTensorMapEncodeTiled is
// resolved at run time through the CUDA runtime, so the library does not link
// libcuda and still loads on a box without a driver (symbol checks on CPU).
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "conv.h"

namespace ub2 {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_once;

static void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  (void)cudaGetLastError();
}

static CUtensorMapSwizzle swizzle_mode(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int make_tmap_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int ld,
                   const uint32_t box[4], int swizzle_bytes, int elem_bytes) {
  std::call_once(g_once, resolve_encode);
  if (!g_encode) return UB2_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * elem_bytes) % 16 != 0) return UB2_ERR_ALIGN;
  const cuuint64_t eb = static_cast<cuuint64_t>(elem_bytes);
  const CUtensorMapDataType dtype = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W),
                        static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * eb, static_cast<cuuint64_t>(W) * ld * eb,
                           static_cast<cuuint64_t>(H) * W * ld * eb};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, dtype, 4, const_cast<void*>(base), dims,
                        strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_mode(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : UB2_ERR_DRIVER;
}

int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld,
                 uint32_t bc, uint32_t br, int swizzle_bytes, int elem_bytes) {
  std::call_once(g_once, resolve_encode);
  if (!g_encode) return UB2_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * elem_bytes) % 16 != 0) return UB2_ERR_ALIGN;
  const CUtensorMapDataType dtype = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t bx[2] = {bc, br};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, dtype, 2, const_cast<void*>(base), dims,
                        strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_mode(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : UB2_ERR_DRIVER;
}

static thread_local int g_variant = 0;
void note_variant(int code) { g_variant = code; }
int last_variant() { return g_variant; }

int device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

int num_sms() {
  static PerDevice<int> cache;
  int& cached = cache.ref();
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
    // UB2_RESERVE_SMS=k: size every grid for k SMs fewer (persistent kernels then leave room for the NCCL kernels
    // of an overlapped all-reduce instead of being split into two waves by them)
    static const int reserve = [] { const char* e = getenv("UB2_RESERVE_SMS"); return e ? atoi(e) : 0; }();
    if (reserve > 0 && cached - reserve >= 8) cached -= reserve;
  }
  return cached;
}

}  // namespace ub2
