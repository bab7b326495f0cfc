// Host-side helpers: error codes, TMA descriptor encoding (driver entry point, no -lcuda), launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/ssv_b200.h"

namespace ssvb {

// profiling hooks (defined in core.cu; off by default, used by bench.py only)
enum ProfKind { PROF_SIM_FWD = 0, PROF_SIM_BWD = 1, PROF_GEMM = 2, PROF_NKINDS = 3 };
void count_launch();
int prof_begin(int kind, cudaStream_t s);  // returns a record slot or -1
void prof_end(int slot, cudaStream_t s);

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// carve a workspace: returns aligned pointer and advances the cursor
struct Carver {
  uint8_t* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  size_t used() const { return (off + 255) & ~static_cast<size_t>(255); }
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// Small per-thread cache of encoded tensor maps keyed on everything that goes into the encoding.  A descriptor is a
// pure function of (pointer, shape, stride, box, type): the losses are called with the same buffers step after step
// (torch's caching allocator hands the same blocks back), and each cuTensorMapEncodeTiled costs 1-2 us of the
// launch-bound small-shape path (four encodes per NT-Xent fwd+bwd).
struct TmapKey {
  const void* ptr;
  int64_t rows, cols, ld;
  int box_rows, kind;  // kind: 0 bf16 operand, 1 fp16 operand, 2 fp32 output, 3 bf16 output
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && kind == o.kind;
  }
};
struct TmapCache {
  static constexpr int N = 32;
  TmapKey keys[N];
  CUtensorMap maps[N];
  int used = 0, next = 0;
  const CUtensorMap* find(const TmapKey& k) const {
    for (int i = 0; i < used; ++i)
      if (keys[i] == k) return &maps[i];
    return nullptr;
  }
  void put(const TmapKey& k, const CUtensorMap& m) {
    const int i = used < N ? used++ : (next++ % N);
    keys[i] = k;
    maps[i] = m;
  }
};
inline TmapCache& tmap_cache() {
  static thread_local TmapCache c;
  return c;
}

// bf16 row-major [rows x cols], leading dimension ld (elements, ld*2 % 16 == 0).
// Box = 64 columns (128 B, SWIZZLE_128B) x box_rows rows; out-of-bounds reads are zero-filled.
inline int make_tmap_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                          bool fp16 = false) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return SSVB_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return SSVB_ERR_ALIGNMENT;
  const TmapKey key{ptr, rows, cols, ld, box_rows, fp16 ? 1 : 0};
  if (const CUtensorMap* hit = tmap_cache().find(key)) {
    *out = *hit;
    return SSVB_OK;
  }
  // the driver call needs a context bound to THIS thread; autograd's backward thread may not have touched the runtime
  // yet.  cudaFree(0) binds the primary context of the current device (once per thread; the retry below stays as a net).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapDataType dt = fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    // the driver call needs a context bound to THIS thread; autograd's backward thread may not have touched the
    // runtime yet.  cudaFree(0) binds the primary context of the current device, then retry once.
    cudaFree(nullptr);
    r = fn(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS && getenv("SSVB_DEBUG"))
    fprintf(stderr, "[ssv_b200] cuTensorMapEncodeTiled -> %d (ptr=%p rows=%lld cols=%lld ld=%lld box_rows=%d)\n",
            static_cast<int>(r), ptr, static_cast<long long>(rows), static_cast<long long>(cols),
            static_cast<long long>(ld), box_rows);
  if (r == CUDA_SUCCESS) tmap_cache().put(key, *out);
  return r == CUDA_SUCCESS ? SSVB_OK : SSVB_ERR_DRIVER;
}

// OUTPUT map of the GEMM epilogue's TMA stores: row-major [rows x cols] of 4-byte (fp32) or 2-byte (bf16) elements,
// box = 128 bytes of one row x 32 rows, SWIZZLE_128B (the staging layout of gemm_kernels.cuh); stores beyond
// rows / cols are clipped by the TMA unit.
inline int make_tmap_out(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int elem_bytes) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return SSVB_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elem_bytes) & 15)) return SSVB_ERR_ALIGNMENT;
  const TmapKey key{ptr, rows, cols, ld, 32, elem_bytes == 4 ? 2 : 3};
  if (const CUtensorMap* hit = tmap_cache().find(key)) {
    *out = *hit;
    return SSVB_OK;
  }
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), 32u};
  cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("SSVB_DEBUG"))
    fprintf(stderr, "[ssv_b200] cuTensorMapEncodeTiled(out) -> %d (ptr=%p rows=%lld cols=%lld ld=%lld eb=%d)\n",
            static_cast<int>(r), ptr, static_cast<long long>(rows), static_cast<long long>(cols),
            static_cast<long long>(ld), elem_bytes);
  if (r == CUDA_SUCCESS) tmap_cache().put(key, *out);
  return r == CUDA_SUCCESS ? SSVB_OK : SSVB_ERR_DRIVER;
}

inline int check_device_sm100() {
  static int cached = -1;  // per-process; all devices of one box are identical
  if (cached >= 0) return cached;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SSVB_ERR_ARCH;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cached = (major == 10) ? SSVB_OK : SSVB_ERR_ARCH;
  return cached;
}

inline int num_sms() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  return n;
}

#define SSVB_CUDA(expr)                                  \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return static_cast<int>(_e);  \
  } while (0)
#define SSVB_TRY(expr)          \
  do {                          \
    int _rc = (expr);           \
    if (_rc != SSVB_OK) return _rc; \
  } while (0)
#define SSVB_LAUNCH_CHECK()          \
  do {                               \
    ::ssvb::count_launch();          \
    SSVB_CUDA(cudaGetLastError());   \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device) instead of on every launch:
// the attribute is per device and sticky, and the call costs ~1 us of the launch-bound small-shape path.
template <auto Kern>   // the kernel itself is the template argument: one static per kernel instantiation
int set_smem_once(int smem) {
  static unsigned long long done_mask = 0;  // bit = device ordinal
  int dev = 0;
  SSVB_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(__atomic_load_n(&done_mask, __ATOMIC_ACQUIRE) & bit)) {
    SSVB_CUDA(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    __atomic_fetch_or(&done_mask, bit, __ATOMIC_RELEASE);
  }
  return SSVB_OK;
}

}  // namespace ssvb
