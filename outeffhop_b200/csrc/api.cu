// Version / error plumbing of the C ABI (include/oeh.h).
#include "oeh_common.cuh"

namespace oeh {

char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

int set_error(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return status;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace oeh

extern "C" int oeh_version(void) { return OEH_VERSION; }

extern "C" const char* oeh_last_error(void) { return oeh::last_error_buf(); }

extern "C" int oeh_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

// Measurement helper (not part of include/oeh.h; bench.py binds it like oeh_set_override): the SM clock the GPU is
// running at right now, from clock64 against %globaltimer over ~20 us on one SM.  Queued on `stream` behind the kernel
// under test it reads the clock that kernel ran at (power management moves in milliseconds), which nvidia-smi sampling
// cannot resolve for a launch of a few milliseconds.
namespace {
__global__ void sm_clock_probe_kernel(float* out_mhz) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < 20000ull);
  const long long c1 = clock64();
  *out_mhz = (float)(c1 - c0) * 1000.f / (float)(t1 - t0);
}
}  // namespace

extern "C" int oeh_sm_clock_probe(float* out_mhz, void* stream) {
  if (!out_mhz) return oeh::set_error(OEH_ERR_BAD_ARG, "oeh_sm_clock_probe: null output");
  sm_clock_probe_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out_mhz);
  return cudaGetLastError() == cudaSuccess ? OEH_OK : oeh::set_error(OEH_ERR_CUDA, "oeh_sm_clock_probe: launch failed");
}
