// runtime.cu -- error string, launch counter, ABI version.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace rmnet {
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
// Read on every launch (a getenv of a short name: negligible next to a kernel launch) so a test can flip it.
bool pdl_enabled() {
  const char *e = getenv("RMNET_DISABLE_PDL");
  return !(e && e[0] == '1');
}
}  // namespace rmnet

extern "C" {
int rmnet_abi_version(void) { return RMNET_ABI_VERSION; }
const char *rmnet_last_error(void) { return rmnet::g_err; }
long long rmnet_launch_count(void) { return rmnet::g_launches; }
void rmnet_launch_count_reset(void) { rmnet::g_launches = 0; }
}
