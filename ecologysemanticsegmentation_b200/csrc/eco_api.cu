// Error plumbing, version and device queries of the C ABI (include/ecoloss.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "eco_common.cuh"

namespace eco {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

int sm_count_cached(int device) {
    static int cache[64];
    if (device < 0 || device >= 64) return -1;
    int v = cache[device];
    if (v > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) {
        set_error("cannot query SM count of device %d", device);
        return -1;
    }
    cache[device] = v;
    return v;
}

}  // namespace eco

extern "C" const char* eco_version(void) { return "ecoloss 0.1.0 (sm_100a)"; }
extern "C" const char* eco_last_error(void) { return eco::g_err; }
extern "C" int eco_sm_count(int device) { return eco::sm_count_cached(device); }
