// Library-wide C-ABI entry points: version and the thread-local error string.
#include "common.h"

namespace cobevt {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace cobevt

extern "C" int cobevt_version(void) { return 100; }
extern "C" const char* cobevt_last_error(void) { return cobevt::last_error_buf(); }
