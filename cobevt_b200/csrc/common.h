// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>          // CUtensorMap (the encoder is fetched through the runtime, libcuda is not linked)
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <utility>
#include "../../include/cobevt_b200.h"

namespace cobevt {

char* last_error_buf();  // thread-local, 512 bytes

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(COBEVT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return COBEVT_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Function attributes (max dynamic smem) and the SM count are PER DEVICE: caches are indexed by the current device.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int num_sms() {
  static int sms[kMaxDevices] = {0};
  const int dev = current_device();
  if (sms[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = v > 0 ? v : 148;
  }
  return sms[dev];
}

// Programmatic dependent launch: every kernel of the library is launched with the stream-serialisation attribute, so in a
// stream (or a captured graph) kernel N+1's CTAs may be scheduled, run their prologue (barrier init, TMEM allocation,
// constant weight / bias staging) and park in `griddepcontrol.wait` while kernel N drains.  Each kernel executes
// pdl_wait() (sm100.cuh) before its first access to memory another kernel may have written or may still read.
// COBEVT_PDL=0 in the environment launches everything fully serialised (A/B timing, debugging).
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("COBEVT_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Timing-experiment knobs (COBEVT_CONV_DEBUG / COBEVT_ATTN_DEBUG produce WRONG results by design) exist only in
// builds made with -DCOBEVT_DEBUG_KNOBS; the release library compiles them out.
#ifdef COBEVT_DEBUG_KNOBS
#define COBEVT_DBG(args, bit) ((args).dbg & (bit))
#else
#define COBEVT_DBG(args, bit) false
#endif

// cuTensorMapEncodeTiled through the runtime's driver entry point table (nullptr if unavailable)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace cobevt
