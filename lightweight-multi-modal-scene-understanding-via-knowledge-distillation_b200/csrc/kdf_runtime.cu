// Library plumbing: error string, ABI version, device info.
#include <stdarg.h>

#include "kdf_common.cuh"

namespace kdf {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
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

}  // namespace kdf

extern "C" {

int kdf_abi_version(void) { return KDF_ABI_VERSION; }

const char *kdf_last_error(void) { return kdf::g_err; }

int kdf_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    KDF_CUDA(cudaGetDevice(&dev));
    int sms = 0, maj = 0, min = 0;
    KDF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    KDF_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    KDF_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return KDF_OK;
}

}  // extern "C"
