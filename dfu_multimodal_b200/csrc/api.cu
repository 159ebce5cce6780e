// Library-level entry points: error text, device checks.
#include "common.cuh"
#include "../../include/dfu_b200.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";

void dfu_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int physical_sms() {
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

// SM budget of the launches that follow (0 = whole device; process-wide: the host serialises launches): every kernel sizes its grid from dfu_num_sms(), so two
// streams given complementary budgets share the device side by side (the tensor-bound ViT branch next to the
// HBM-bound ResNet branch) instead of time-slicing it with full-width launches.
static volatile int g_sm_budget = 0;

int dfu_num_sms() {
    const int n = physical_sms();
    return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}
extern "C" int dfu_set_sm_budget(int32_t n_sms) {
    g_sm_budget = n_sms > 0 ? n_sms : 0;
    return DFU_OK;
}

bool dfu_pdl_enabled() {
    // off by default: measured on B200 (profiles/r02c_*): 31.76 ms per train step with, 31.51 ms without -- CUDA-graph replay
    // already hides the launch latency this was meant to save; DFU_PDL=1 turns it on
    static const bool on = []() { const char* e = getenv("DFU_PDL"); return e && e[0] == '1'; }();
    return on;
}

extern "C" const char* dfu_last_error(void) { return g_err; }
extern "C" int dfu_version(void) { return 100; }
extern "C" int dfu_sm_count(void) { return physical_sms(); }

extern "C" int dfu_device_ok(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        dfu_set_error("no CUDA device: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return DFU_ERR_CUDA;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        dfu_set_error("device %d has compute capability %d.x; libdfu_b200 is sm_100a only", dev, major);
        return DFU_ERR_UNSUPPORTED;
    }
    return DFU_OK;
}
