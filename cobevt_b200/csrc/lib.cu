// Library-wide C-ABI entry points: version and the thread-local error string.
#include "common.h"

namespace cobevt {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace cobevt

extern "C" int cobevt_version(void) { return 110; }   // 110: CobevtAttnParams.bias_layout, cobevt_struct_bytes

// sizeof of the parameter structs as THIS build sees them: a binding (ctypes, cgo, JNI ...) compares them with its own
// declarations before the first call instead of discovering a layout drift as wrong results
extern "C" int32_t cobevt_struct_bytes(int32_t which) {
  switch (which) {
    case 0: return (int32_t)sizeof(CobevtConvParams);
    case 1: return (int32_t)sizeof(CobevtAttnParams);
    case 2: return (int32_t)sizeof(CobevtMlpParams);
    case 3: return (int32_t)sizeof(CobevtTokenMap);
    default: return -1;
  }
}
extern "C" const char* cobevt_last_error(void) { return cobevt::last_error_buf(); }
