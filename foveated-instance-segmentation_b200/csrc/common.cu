// Error channel + ABI version of libfovea_b200.so.
#include <stdarg.h>

#include "common.cuh"

namespace fovea {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace fovea

extern "C" const char* fovea_last_error(void) { return fovea::g_err; }
extern "C" int fovea_abi_version(void) { return FOVEA_ABI_VERSION; }
