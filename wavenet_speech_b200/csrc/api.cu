// Error reporting / device check for libwnb200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace wnb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace wnb

extern "C" const char* wnb200_last_error(void) { return wnb::g_err; }
extern "C" int wnb200_version(void) { return 100; }

extern "C" int wnb200_check_device(void) {
  int dev = 0;
  WNB_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  WNB_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    wnb::set_error("libwnb200 is built for sm_100a (B200) only; device %d is sm_%d%d (%s)", dev, prop.major,
                   prop.minor, prop.name);
    return 4;
  }
  return 0;
}
