// Non-kernel parts of the C ABI (include/ojdf.h): version, error strings, launch counter.
#include <atomic>

#include "ojdf_internal.h"

namespace ojdf {

static std::atomic<uint64_t> g_launches{0};

void make_pose(Pose &P, const float *Kinv, const float *E, const double *origin, double res)
{
    for (int i = 0; i < 9; ++i) P.kinv[i] = Kinv[i];
    for (int i = 0; i < 12; ++i) P.e[i] = E[i];
    P.res = res;
    for (int a = 0; a < 3; ++a) {
        P.origin[a] = origin[a];
        volatile double num = (double)E[4 * a + 3] - origin[a];     // eye = E[:3,3] (modules/extractor.py:57)
        volatile double q = num / res;
        P.ev[a] = q;
    }
}

int launched(int n)
{
    g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace ojdf

extern "C" int ojdf_version(void) { return OJDF_VERSION; }

extern "C" uint64_t ojdf_launch_count(void) { return ojdf::g_launches.load(std::memory_order_relaxed); }

extern "C" const char *ojdf_error_string(int code)
{
    switch (code) {
        case 0: return "success";
        case OJDF_ERR_BADARG: return "ojdf: bad argument (null pointer, non-positive size, or P not odd / > 33)";
        case OJDF_ERR_WORKSPACE: return "ojdf: workspace missing or too small";
        case OJDF_ERR_TOOLARGE: return "ojdf: grid or entry count does not fit 32-bit keys";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "ojdf: unknown error";
    }
}
