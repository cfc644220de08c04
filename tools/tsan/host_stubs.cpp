#include <string>
#include "scasa_cuda.h"
static thread_local std::string g_err;
namespace scasa { void set_global_error(const std::string &m) { g_err = m; } }
extern "C" const char *scasa_last_error(const scasa_ctx *) { return g_err.c_str(); }
