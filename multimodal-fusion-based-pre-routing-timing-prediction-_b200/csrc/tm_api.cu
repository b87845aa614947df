// Library-level entry points: version, error string, launch counter, device probe.
#include "tm_common.cuh"

namespace tmk {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace tmk

extern "C" int tm_version(void) { return 100; }  // 0.1.0
extern "C" const char* tm_last_error(void) { return tmk::g_err; }
extern "C" long long tm_launch_count(void) { return tmk::g_launches.load(); }
extern "C" int tm_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  TM_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  TM_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}
