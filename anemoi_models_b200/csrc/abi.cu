// C-ABI plumbing shared by every entry point of libanemoi_b200 (see include/anemoi_b200.h).
#include "common.cuh"

namespace ab2 {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
long long* launch_counter() {
  static long long n = 0;
  return &n;
}
}  // namespace ab2

extern "C" int ab2_version(void) { return 101; /* 0.1.1: ab2_gemm gained a_seg / a_seg_len */ }
extern "C" const char* ab2_last_error(void) { return ab2::last_error_buf(); }
extern "C" long long ab2_launch_count(void) { return __atomic_load_n(ab2::launch_counter(), __ATOMIC_RELAXED); }
