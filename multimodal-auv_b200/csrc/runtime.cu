// C-ABI runtime glue: version, error string, device capability check.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

char* mauv_err_buf() { return g_err; }

int mauv_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static thread_local const unsigned int* g_sample_base = nullptr;
const unsigned int* mauv_sample_base() { return g_sample_base; }

int mauv_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

extern "C" {

int mauv_version(void) { return 100; }  // 0.1.0

const char* mauv_last_error(void) { return g_err; }

// 0 when the current device can run this library (compute capability 10.x); the kernels are
// compiled for sm_100a only and there is no fallback path.
int mauv_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  MAUV_CUDA(cudaGetDevice(&dev));
  MAUV_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MAUV_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return mauv_set_error(MAUV_ERR_UNSUPPORTED_ARCH,
                          "mauv_b200 requires an sm_100a (Blackwell B200) device, found sm_%d%d", major, minor);
  return MAUV_OK;
}

int mauv_num_sms_c(void) { return mauv_num_sms(); }

// Device-resident Philox sample-id base for the calling thread's subsequent sampling launches (NULL = none): every
// kernel that draws eps adds *sample_base to its sample ids when it RUNS. A CUDA graph captured with a base pointer set
// therefore replays with whatever value the word holds at replay time - fresh Monte-Carlo draws for every batch from one
// recorded graph (the reference draws fresh eps on every pass, inference/predictors.py:54-66).
int mauv_set_sample_base(const unsigned int* sample_base) {
  g_sample_base = sample_base;
  return MAUV_OK;
}

}  // extern "C"
