// Library-level entry points: version, thread-local error text, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gloria {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

long long& launch_counter() {
  static thread_local long long n = 0;
  return n;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(err_buf(), 512, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return GLORIA_ERR_CUDA_BASE + (int)e;
}

}  // namespace gloria

extern "C" int gloria_b200_version(void) { return 100; }

extern "C" const char* gloria_b200_last_error(void) { return gloria::err_buf(); }

extern "C" long long gloria_b200_launch_count(int reset) {
  long long v = gloria::launch_counter();
  if (reset) gloria::launch_counter() = 0;
  return v;
}
