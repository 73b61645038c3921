// vfkio_util.cpp -- error reporting shared by the host companion library (libvfkio.so).
#include <cstdio>
#include "vfk_io.h"

static thread_local char g_err[512] = "";
extern "C" const char *vfkio_last_error(void) { return g_err; }
int vfkio_fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
