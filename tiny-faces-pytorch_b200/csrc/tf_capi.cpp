// Error string + version entry points of the C-ABI.
#include "tf_common.cuh"
#include <mutex>

static thread_local char g_err[1024] = "";

void tf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

TF_API const char* tf_last_error_string(void) { return g_err; }
TF_API int tf_version(void) { return 100; }
