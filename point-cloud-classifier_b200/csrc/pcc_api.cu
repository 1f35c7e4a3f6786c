// Library-level entry points: error string, version, device check.
#include "pcc_common.cuh"

namespace pcc {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace pcc

extern "C" const char* pcc_last_error(void) { return pcc::g_last_error.c_str(); }

extern "C" int pcc_version(void) { return 100; }

extern "C" int pcc_check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return pcc::fail(__func__, "no CUDA device visible; libpcc has no CPU fallback");
  }
  if (device < 0 || device >= count) return pcc::fail(__func__, "device index out of range");
  cudaDeviceProp prop;
  PCC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    return pcc::fail(__func__, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                   "; libpcc is built for sm_100a (B200) only");
  }
  return 0;
}
