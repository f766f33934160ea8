// tg_api.cu — version / error reporting of the C-ABI (include/topicgcn.h).
#include <stdarg.h>

#include "tg_common.cuh"

namespace tg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return TG_ERR_CUDA;
}

}  // namespace tg

extern "C" {

int tg_version(void) { return TG_VERSION; }

const char* tg_last_error(void) { return tg::g_err; }

const char* tg_status_string(int status) {
    switch (status) {
        case TG_OK: return "ok";
        case TG_ERR_INVALID_ARG: return "invalid argument";
        case TG_ERR_CUDA: return "CUDA runtime error";
        case TG_ERR_UNSUPPORTED: return "unsupported configuration";
        case TG_ERR_WORKSPACE: return "workspace too small";
        case TG_ERR_OVERFLOW: return "index overflow";
        default: return "unknown status";
    }
}

}  // extern "C"
