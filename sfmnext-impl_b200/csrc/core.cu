// Error reporting and device checks for libsqlx.
#include "common.cuh"

namespace sqlx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();  // clear the non-sticky launch error so later calls are not poisoned
    return SQLX_ECUDA;
  }
  return SQLX_OK;
}

}  // namespace sqlx

extern "C" const char* sqlx_last_error(void) { return sqlx::g_err; }

extern "C" int sqlx_version(void) { return 100; }

extern "C" int sqlx_device_ok(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    cudaGetLastError();
    sqlx::set_error("cudaGetDeviceProperties(%d) failed", device);
    return 0;
  }
  return prop.major == 10 ? 1 : 0;
}
