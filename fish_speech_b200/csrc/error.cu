// Last-error storage for the C-ABI (one message per host thread).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace fsb {

static thread_local char g_err[1024] = {0};
thread_local int g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FSB_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace fsb
