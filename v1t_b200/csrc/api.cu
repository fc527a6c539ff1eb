// error plumbing + version for the C-ABI
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"

namespace v1t {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace v1t

extern "C" const char* v1t_last_error(void) { return v1t::g_err; }
extern "C" int v1t_version(void) { return 1; }

extern "C" int v1t_dropout_mask(float* out, int64_t n, uint64_t seed, uint32_t site, float p, void* stream) {
  V1T_CHECK_ARG(out && n >= 0 && p >= 0.f && p < 1.f, "dropout_mask: bad argument");
  if (n == 0) return V1T_OK;
  return v1t::dropout_mask(out, n, v1t::DropSpec{seed, site, p}, (cudaStream_t)stream);
}
