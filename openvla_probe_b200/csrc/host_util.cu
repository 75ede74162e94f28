#include "host_util.h"

#include <stdarg.h>
#include <stdio.h>

namespace ovla {

static thread_local char g_err[1024] = "";
static long long g_launches = 0;

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
const char* last_error() { return g_err; }
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }
void reset_launch_count() { g_launches = 0; }

}  // namespace ovla
