// Error plumbing and library identification for the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

unsigned long long g_launches = 0;

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  return PF_OK;
}

}  // namespace pf

extern "C" const char* pf_last_error(void) { return pf::g_err; }
extern "C" int pf_abi_version(void) { return PF_ABI_VERSION; }

// Number of kernels this library has launched since load (bench.py's gpu_launches).
extern "C" unsigned long long pf_launch_count(void) { return pf::g_launches; }
