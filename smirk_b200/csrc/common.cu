// Error channel + version for the smirk_b200 C ABI.
#include "common.cuh"
#include <stdarg.h>

namespace smk {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace smk

extern "C" int smk_version(void) { return SMK_VERSION; }
extern "C" const char* smk_last_error(void) { return smk::g_err; }
