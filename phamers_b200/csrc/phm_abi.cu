// Library-wide pieces of the C-ABI: error state, device capabilities, tuning options.
#include <stdarg.h>

#include "score_common.cuh"

namespace phm {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int time_kernels = 0;

static int g_sm_count = 0, g_smem_optin = 0;

static void query_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
}

int sm_count() {
    if (g_sm_count == 0) query_device();
    return g_sm_count;
}
int max_smem_optin() {
    if (g_smem_optin == 0) query_device();
    return g_smem_optin;
}

extern int hist_stride_for_k4;
extern int hist_contigs_per_item;
extern int hist_stride_for_k5;
extern int hist_warps_k6;
extern int hist_tma;
extern int hist_canonical_swizzle;
extern int hist_plan;
extern int score_path;
extern int score_collect_stats;
extern int score_list_pass;
extern int score_debug;
int kmer_hist_last_ms(float *ms);

}  // namespace phm

using namespace phm;

extern "C" int phm_version(void) { return PHM_VERSION; }

extern "C" const char *phm_last_error(void) { return g_error; }

extern "C" int phm_device_caps(phm_caps *out) {
    PHM_REQUIRE(out != nullptr, "out is null");
    int dev = 0;
    PHM_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    PHM_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    out->device = dev;
    out->sm_major = prop.major;
    out->sm_minor = prop.minor;
    out->sm_count = prop.multiProcessorCount;
    out->max_smem_optin = (int32_t)prop.sharedMemPerBlockOptin;
    out->hbm_bytes = (int64_t)prop.totalGlobalMem;
    if (prop.major != 10) {
        set_error("phamers_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
        return PHM_E_UNSUPPORTED;
    }
    return PHM_OK;
}

// Tuning knobs for experiments (bench.py --opt name=value).  Unknown names are an error.
extern "C" int phm_set_option(const char *name, int64_t value) {
    PHM_REQUIRE(name != nullptr, "name is null");
    if (!strcmp(name, "hist_stride_k4")) { PHM_REQUIRE(value == 1 || value == 2, "1 or 2"); hist_stride_for_k4 = (int)value; return PHM_OK; }
    if (!strcmp(name, "hist_stride_k5")) { PHM_REQUIRE(value >= 0 && value <= 2, "0 (automatic), 1 or 2"); hist_stride_for_k5 = (int)value; return PHM_OK; }
    if (!strcmp(name, "hist_canonical_swizzle")) { hist_canonical_swizzle = value != 0; return PHM_OK; }
    if (!strcmp(name, "hist_plan")) { hist_plan = value != 0; return PHM_OK; }
    if (!strcmp(name, "hist_tma")) { hist_tma = value != 0; return PHM_OK; }
    if (!strcmp(name, "hist_warps_k6")) { PHM_REQUIRE(value == 4 || value == 13, "4 or 13"); hist_warps_k6 = (int)value; return PHM_OK; }
    if (!strcmp(name, "hist_contigs_per_item")) { PHM_REQUIRE(value >= 0 && value <= 4096, "0 (automatic) .. 4096"); hist_contigs_per_item = (int)value; return PHM_OK; }
    if (!strcmp(name, "score_path")) { PHM_REQUIRE(value >= 0 && value <= 2, "0 auto, 1 exact, 2 tensor cores"); score_path = (int)value; return PHM_OK; }
    if (!strcmp(name, "time_kernels")) { time_kernels = value != 0; return PHM_OK; }
    if (!strcmp(name, "score_debug")) { score_debug = (int)value; return PHM_OK; }
    if (!strcmp(name, "score_stats")) { score_collect_stats = value != 0; return PHM_OK; }
    if (!strcmp(name, "score_list_pass")) { score_list_pass = value != 0; return PHM_OK; }
    set_error("unknown option '%s'", name);
    return PHM_E_ARG;
}

// Device time of the most recent launch of a named hot kernel (bench.py's roofline numerator); synchronises on that launch.
extern "C" int phm_last_kernel_ms(const char *kernel, float *ms) {
    PHM_REQUIRE(kernel != nullptr && ms != nullptr, "null pointer");
    if (!strcmp(kernel, "score_tc_kernel")) return tc::score_tc_last_ms(ms);
    if (!strcmp(kernel, "kmer_hist_kernel")) return kmer_hist_last_ms(ms);
    set_error("no timing hook for kernel '%s'", kernel);
    return PHM_E_ARG;
}
