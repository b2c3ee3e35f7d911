// C-ABI plumbing shared by all entry points: thread-local last-error string,
// launch checking, device query.  No C++ exception crosses the ABI, the library
// never allocates or frees device memory and never synchronises the device.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void oq_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int oq_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    oq_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int oq_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

extern "C" const char* oq_last_error(void) { return g_err; }
extern "C" int oq_version(void) { return 100; }
extern "C" int oq_sm_count(void) { return oq_num_sms(); }
