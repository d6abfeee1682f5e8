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

// Strided 2-D copy between host and device (cudaMemcpy2DAsync, direction inferred): `height` rows of
// `width_bytes`, row pitches in bytes.  The model parameters are user-visible host matrices
// (TW: K x N float64); a rank of a sharded run owns a COLUMN range of them, which this moves
// with one DMA descriptor instead of a host-side gather into a temporary.  Copies that touch
// pageable host memory have completed when the call returns (the stream is synchronised).
extern "C" int pf_copy_2d(void* dst, int64_t dpitch, const void* src, int64_t spitch,
                          int64_t width_bytes, int64_t height, void* stream) {
  if (width_bytes <= 0 || height <= 0) return PF_OK;
  PF_REQUIRE(dpitch >= width_bytes && spitch >= width_bytes, "pf_copy_2d: pitch < width");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes,
                                    (size_t)height, cudaMemcpyDefault, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    pf::set_error("pf_copy_2d: %s", cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  return PF_OK;
}
