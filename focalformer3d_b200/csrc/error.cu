#include "common.cuh"
#include <string.h>

namespace ff3d {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ff3d

extern "C" const char* ff3d_last_error(void) { return ff3d::g_err; }
extern "C" int ff3d_version(void) { return 100; }
