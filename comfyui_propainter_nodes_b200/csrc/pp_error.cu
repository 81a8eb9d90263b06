#include "pp_common.cuh"
#include <stdarg.h>
#include <string.h>
static thread_local char g_err[1024] = "";
void pp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" __attribute__((visibility("default"))) const char* pp_last_error(void) { return g_err; }
