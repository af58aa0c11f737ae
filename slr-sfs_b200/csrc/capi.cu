// Library-wide pieces of the C ABI: version, error text, device properties.
#include <stdlib.h>
#include <string.h>

#include "slr_host.h"

namespace {
thread_local char g_last_error[512] = "";
}

namespace slr_host {

int fail(int code, const char* msg)
{
    strncpy(g_last_error, msg ? msg : "unknown error", sizeof(g_last_error) - 1);
    g_last_error[sizeof(g_last_error) - 1] = '\0';
    return code;
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

bool gather_staged()
{
    const char* e = getenv("SLR_GATHER_MODE");
    return e && strcmp(e, "staged") == 0;
}

bool index_direct()
{
    const char* e = getenv("SLR_GATHER_MODE");
    return !(e && (strcmp(e, "staged") == 0 || strcmp(e, "bins") == 0));
}

}  // namespace slr_host

extern "C" int slr_version(void) { return 100; }  // 0.1.0

extern "C" const char* slr_last_error_string(void) { return g_last_error; }
