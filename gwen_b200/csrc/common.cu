// Error convention and device queries of libgwen_b200 (see include/gwen_b200.h).
#include "common.cuh"

namespace gwen {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

static int g_sm_reserve = 0;
int sm_reserve() { return g_sm_reserve; }

}  // namespace gwen

extern "C" int gwen_set_sm_reserve(int n) {
  int old = gwen::g_sm_reserve;
  gwen::g_sm_reserve = n < 0 ? 0 : n;
  return old;
}
extern "C" int gwen_version(void) { return GWEN_ABI_VERSION; }
extern "C" const char* gwen_last_error(void) { return gwen::err_buf(); }
