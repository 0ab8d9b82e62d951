// ABI bookkeeping: version, per-thread error string.
#include <stdarg.h>
#include "common.cuh"
#include "fira_b200.h"

namespace {
thread_local char g_err[512] = "";
}

void fira_set_error(int code, const char* fmt, ...) {
  int n = snprintf(g_err, sizeof(g_err), "[fira error %d] ", code);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err + n, sizeof(g_err) - n, fmt, ap);
  va_end(ap);
}

extern "C" {
int fira_version(void) { return 2; }
const char* fira_last_error_string(void) { return g_err; }
int fira_built_arch(void) { return 100; }
}
