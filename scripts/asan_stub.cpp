#include "../dennou-ccm_b200/csrc/dccm_common.h"
#include <cstdarg>
namespace dccm {
static thread_local std::string g_err;
void set_error(const char *fmt, ...) { char b[1024]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap); g_err = b; }
int fail(int code, const char *fmt, ...) { char b[1024]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap); g_err = b; return code; }
}
extern "C" const char *dccm_last_error(void) { return dccm::g_err.c_str(); }
