#include "common.hpp"

#include <cstdarg>
#include <cstdio>

namespace pb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

}  // namespace pb

extern "C" {
const char* pb_last_error(void) { return pb::get_error(); }
const char* pb_version(void) { return "probly_b200 0.1 (sm_100a)"; }
}
