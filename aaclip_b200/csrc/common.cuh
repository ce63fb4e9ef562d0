// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

namespace host {

// Error codes of the C ABI (include/aaclip_b200.h).
enum : int {
  OK = 0,
  ERR_INVALID = -1,   // bad argument / unsupported shape
  ERR_CUDA = -2,      // CUDA runtime or driver error
  ERR_NO_DEVICE = -3, // no sm_100 device
  ERR_STATE = -4      // call order / missing weights
};

inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

#define AACLIP_CUDA_CHECK(expr)                                                                      \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return host::fail(host::ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// bf16 tensor of rank `rank` (dims[0] innermost, contiguous), 128B-swizzled boxes whose inner extent is
// 64 elements (128 B).  strides_bytes[i] is the byte stride of dims[i+1].  OOB elements read as zero.
inline int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return fail(ERR_INVALID, "TMA base must be 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) return fail(ERR_INVALID, "TMA stride %d = %llu B is not a multiple of 16", i, (unsigned long long)gstr[i]);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return OK;
}

// [rows, cols] row-major bf16 (row pitch `ld` elements), box = {64 cols, box_rows}
inline int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {64, box_rows};
  return make_tmap_bf16(tm, base, 2, dims, str, box);
}

// output tensor [rows, cols] (row pitch ld elements) of bf16 or fp32; box = 32 rows x 128 B, 128B swizzle
inline int make_tmap_out(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, bool is_bf16) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  const uint64_t es = is_bf16 ? 2 : 4;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * es};
  cuuint32_t box[2] = {is_bf16 ? 64u : 32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || gstr[0] % 16 != 0)
    return fail(ERR_INVALID, "TMA output base/pitch must be 16-byte aligned");
  CUresult r = enc(tm, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERR_CUDA, "cuTensorMapEncodeTiled (output) failed with CUresult %d", int(r));
  return OK;
}

// Kernel launch, optionally with the programmatic-dependent-launch attribute (AACLIP_PDL=1).  Only for kernels
// that call ptx::grid_dep_sync() before their first global-memory access.  Measured on B200 at the bench shape
// (30-step A/B, twice): 2310 img/s with the attribute vs 2349 without - the prologues it overlaps are already short
// next to 100-250 us kernels - so it is OFF by default.
inline bool pdl_enabled() {
  static const int v = getenv("AACLIP_PDL") ? atoi(getenv("AACLIP_PDL")) : 0;
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// The context-free entry points (building blocks, head, loader transform) run on the device that owns their first
// device pointer and give the caller's current device back on return: a stream handle and a kernel's shared-memory
// opt-in belong to one device, and the thread's current device belongs to the caller (PyTorch).
struct PointerDeviceGuard {
  int prev = -1, dev = -1;
  explicit PointerDeviceGuard(const void* p) {
    cudaPointerAttributes at;
    if (p != nullptr && cudaPointerGetAttributes(&at, p) == cudaSuccess &&
        (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged))
      dev = at.device;
    else
      cudaGetLastError();   // host / unregistered pointer: stay on the current device (argument checks report it)
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~PointerDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  PointerDeviceGuard(const PointerDeviceGuard&) = delete;
  PointerDeviceGuard& operator=(const PointerDeviceGuard&) = delete;
};

inline int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}

}  // namespace host
