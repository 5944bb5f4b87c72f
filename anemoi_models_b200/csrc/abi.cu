// C-ABI plumbing shared by every entry point of libanemoi_b200 (see include/anemoi_b200.h).
#include "common.cuh"

namespace ab2 {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace ab2

extern "C" int ab2_version(void) { return 100; /* 0.1.0 */ }
extern "C" const char* ab2_last_error(void) { return ab2::last_error_buf(); }
