// Extractor hot path (modules/extractor.py:24-79) on sm_100a: ONE kernel per frame.
//
//   extract_kernel : a block owns R consecutive rays.  Phase 1: R threads unproject their pixel (f32 FMA chain,
//                    modules/extractor.py:82-120) and build the per-ray record (f64 voxel-space centre + unit direction,
//                    :309-318) in shared memory -- or read a record computed earlier (ojdf_rays: the pipeline needs the
//                    records before the volumes are final, to plan the integration on a side stream).  Phase 2: one thread
//                    per (ray, sample): 3 axis set-ups in f64, 8 corner indices / weights (:533-593), 16 fp16 gathers of
//                    the TSDF and weight volumes (:640-681), the 8-term f64 sums in the reference's order, and coalesced
//                    f32 stores -- optionally straight into FusionNet's pixel-major input buffers
//                    ([values | weights | depth or label], modules/pipeline.py:74-102), which replaces a separate
//                    packing pass.
// The optional materialisation of points / indices / weights (the 177 MB the reference always builds) is a template
// flag so the product path never pays for it.
#include <cstring>

#include "ojdf_internal.h"

namespace ojdf {

constexpr int kExtractThreads = 288;        // R rays x P samples, R = kExtractThreads / P (P = 9: 32 rays)

struct Pack {                                // FusionNet input buffers (NULL out_a: no packing)
    float *out_a, *out_b;                    // (N, stride) each; head B optional
    const float *last_a, *last_b;            // (N) last channel of each head: depth frame / normalised label frame
    int stride;
};

template <bool FULL>
__global__ void __launch_bounds__(kExtractThreads)
extract_kernel(const float *__restrict__ depth, const float *__restrict__ world_in, const double *__restrict__ ray_in, int h, int w,
               Pose pose, const __half *__restrict__ tsdf, const __half *__restrict__ wvol, int X, int Y, int Z, int P, int R,
               float *__restrict__ out_vals, float *__restrict__ out_wts, float *__restrict__ out_world, double *__restrict__ out_ray,
               double *__restrict__ out_points, long long *__restrict__ out_idx, double *__restrict__ out_w, Pack pack)
{
    __shared__ double s_ray[kExtractThreads][6 + 1];          // +1: odd pitch in doubles, no bank conflicts in phase 2
    const int N = h * w;
    const int n0 = blockIdx.x * R;
    if ((int)threadIdx.x < R) {
        const int n = n0 + threadIdx.x;
        if (n < N) {
            double rec[6];
            if (ray_in) {
                const double2 *rp = reinterpret_cast<const double2 *>(ray_in + 6 * (size_t)n);
                const double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
                rec[0] = r0.x; rec[1] = r0.y; rec[2] = r1.x; rec[3] = r1.y; rec[4] = r2.x; rec[5] = r2.y;
            } else {
                float wp[3];
                if (world_in) {
                    wp[0] = world_in[3 * n]; wp[1] = world_in[3 * n + 1]; wp[2] = world_in[3 * n + 2];
                } else {
                    const int r = n / w, c = n - r * w;
                    unproject_pixel(pose, r, c, depth[n], wp);
                }
                if (out_world) { out_world[3 * n] = wp[0]; out_world[3 * n + 1] = wp[1]; out_world[3 * n + 2] = wp[2]; }
                ray_record(pose, wp, rec);
                if (out_ray) {
                    double2 *o = reinterpret_cast<double2 *>(out_ray + 6 * (size_t)n);
                    o[0] = make_double2(rec[0], rec[1]);
                    o[1] = make_double2(rec[2], rec[3]);
                    o[2] = make_double2(rec[4], rec[5]);
                }
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) s_ray[threadIdx.x][a] = rec[a];
        }
    }
    __syncthreads();
    if (!tsdf) return;                                          // rays only (ojdf_rays)
    const int lr = threadIdx.x / P, k = threadIdx.x - lr * P;   // local ray, sample
    const int n = n0 + lr;
    if (lr >= R || n >= N) return;
    const long long t = (long long)n * P + k;
    const int i = k - P / 2;
    const double *rc = s_ray[lr];
    const double px = ray_sample(rc[0], rc[3], i), py = ray_sample(rc[1], rc[4], i), pz = ray_sample(rc[2], rc[5], i);
    const Axis ax = axis_setup(px), ay = axis_setup(py), az = axis_setup(pz);

    float v[8], g[8];
    double wc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        long long ix, iy, iz;
        const bool ok = corner_index(ax, ay, az, c, X, Y, Z, ix, iy, iz);
        wc[c] = corner_weight(ax, ay, az, c);
        v[c] = -0.1f;                                  // modules/extractor.py:663
        g[c] = 0.0f;                                   // modules/extractor.py:664
        if (ok) {
            const long long lin = (ix * Y + iy) * (long long)Z + iz;
            v[c] = __half2float(__ldg(tsdf + lin));
            g[c] = __half2float(__ldg(wvol + lin));
        }
        if (FULL) {
            long long *oi = out_idx + (t * 8 + c) * 3;
            oi[0] = ix; oi[1] = iy; oi[2] = iz;
            out_w[t * 8 + c] = wc[c];
        }
    }
    double tv[8], tw[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { tv[c] = __dmul_rn((double)v[c], wc[c]); tw[c] = __dmul_rn((double)g[c], wc[c]); }
    // ATen's row-sum order over 8 contiguous f64
    const double sv = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(tv[0], tv[4]), __dadd_rn(tv[1], tv[5])), __dadd_rn(tv[2], tv[6])), __dadd_rn(tv[3], tv[7]));
    const double sw = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(tw[0], tw[4]), __dadd_rn(tw[1], tw[5])), __dadd_rn(tw[2], tw[6])), __dadd_rn(tw[3], tw[7]));
    const float fv = (float)sv, fw = (float)sw;
    out_vals[t] = fv;
    out_wts[t] = fw;
    if (pack.out_a) {                                           // FusionNet input: [values(P) | weights(P) | last]
        float *a = pack.out_a + (size_t)n * pack.stride;
        a[k] = fv; a[P + k] = fw;
        if (k == 0) a[2 * P] = pack.last_a[n];
        if (pack.out_b) {
            float *b = pack.out_b + (size_t)n * pack.stride;
            b[k] = fv; b[P + k] = fw;
            if (k == 0) b[2 * P] = pack.last_b[n];
        }
    }
    if (FULL) { out_points[3 * t] = px; out_points[3 * t + 1] = py; out_points[3 * t + 2] = pz; }
}

static int launch_extract(const float *depth, const float *world_in, const double *ray_in, int h, int w, const Pose &pose,
                          const __half *tsdf, const __half *wvol, int X, int Y, int Z, int P, float *out_vals, float *out_wts,
                          float *out_world, double *out_ray, double *out_points, long long *out_idx, double *out_w, const Pack &pack,
                          cudaStream_t s)
{
    const int N = h * w;
    const int R = kExtractThreads / P;
    const unsigned blocks = (unsigned)((N + R - 1) / R);
    if (out_points)
        extract_kernel<true><<<blocks, kExtractThreads, 0, s>>>(depth, world_in, ray_in, h, w, pose, tsdf, wvol, X, Y, Z, P, R, out_vals,
                                                                out_wts, out_world, out_ray, out_points, out_idx, out_w, pack);
    else
        extract_kernel<false><<<blocks, kExtractThreads, 0, s>>>(depth, world_in, ray_in, h, w, pose, tsdf, wvol, X, Y, Z, P, R, out_vals,
                                                                 out_wts, out_world, out_ray, nullptr, nullptr, nullptr, pack);
    return launched(1);
}

}  // namespace ojdf

using namespace ojdf;

static const float kIdent[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

extern "C" int ojdf_unproject(const float *depth_dev, int h, int w, const float *Kinv_host, const float *E_host,
                              float *world_dev, void *stream)
{
    if (!depth_dev || !world_dev || !Kinv_host || !E_host || h <= 0 || w <= 0) return OJDF_ERR_BADARG;
    Pose P;
    const double zero3[3] = {0, 0, 0};
    make_pose(P, Kinv_host, E_host, zero3, 1.0);
    return launch_extract(depth_dev, nullptr, nullptr, h, w, P, nullptr, nullptr, 1, 1, 1, 9, nullptr, nullptr, world_dev, nullptr,
                          nullptr, nullptr, nullptr, Pack{nullptr, nullptr, nullptr, nullptr, 0}, (cudaStream_t)stream);
}

extern "C" int ojdf_rays(const float *depth_dev, const float *world_in_dev, int h, int w, const float *Kinv_host,
                         const float *E_host, const double *origin_host, double resolution, float *out_world_dev,
                         double *out_ray_dev, void *stream)
{
    if ((!depth_dev && !world_in_dev) || !E_host || !origin_host || !out_ray_dev || h <= 0 || w <= 0 || !(resolution > 0.0) ||
        (!world_in_dev && !Kinv_host))
        return OJDF_ERR_BADARG;
    Pose pose;
    make_pose(pose, Kinv_host ? Kinv_host : kIdent, E_host, origin_host, resolution);
    return launch_extract(depth_dev, world_in_dev, nullptr, h, w, pose, nullptr, nullptr, 1, 1, 1, 9, nullptr, nullptr, out_world_dev,
                          out_ray_dev, nullptr, nullptr, nullptr, Pack{nullptr, nullptr, nullptr, nullptr, 0}, (cudaStream_t)stream);
}

static int gather_common(const double *ray_in, const float *depth_dev, const float *world_in_dev, int h, int w, const float *Kinv_host,
                         const float *E_host, const double *origin_host, double resolution, const void *tsdf_dev,
                         const void *wvol_dev, int X, int Y, int Z, int P, float *out_vals_dev, float *out_wts_dev,
                         float *out_world_dev, double *out_ray_dev, double *out_points_dev, int64_t *out_idx_dev, double *out_w_dev,
                         const Pack &pack, void *stream)
{
    if (!tsdf_dev || !wvol_dev || !out_vals_dev || !out_wts_dev || h <= 0 || w <= 0 || X <= 0 || Y <= 0 || Z <= 0 || P < 1 ||
        P > 33 || !(P & 1))
        return OJDF_ERR_BADARG;
    const bool full = out_points_dev || out_idx_dev || out_w_dev;
    if (full && !(out_points_dev && out_idx_dev && out_w_dev)) return OJDF_ERR_BADARG;
    if ((long long)X * Y * Z >= 0xFFFFFFFFll) return OJDF_ERR_TOOLARGE;
    if (pack.out_a && (!pack.last_a || pack.stride < 2 * P + 1 || (pack.out_b && !pack.last_b))) return OJDF_ERR_BADARG;
    Pose pose;
    if (!ray_in) {
        if ((!depth_dev && !world_in_dev) || !E_host || !origin_host || !out_ray_dev || !(resolution > 0.0) ||
            (!world_in_dev && !Kinv_host))
            return OJDF_ERR_BADARG;
        make_pose(pose, Kinv_host ? Kinv_host : kIdent, E_host, origin_host, resolution);
    } else {
        memset(&pose, 0, sizeof(pose));
    }
    return launch_extract(depth_dev, world_in_dev, ray_in, h, w, pose, (const __half *)tsdf_dev, (const __half *)wvol_dev, X, Y, Z, P,
                          out_vals_dev, out_wts_dev, out_world_dev, out_ray_dev, out_points_dev, (long long *)out_idx_dev, out_w_dev,
                          pack, (cudaStream_t)stream);
}

extern "C" int ojdf_extract(const float *depth_dev, const float *world_in_dev, int h, int w,
                            const float *Kinv_host, const float *E_host, const double *origin_host, double resolution,
                            const void *tsdf_dev, const void *wvol_dev, int X, int Y, int Z, int P,
                            float *out_vals_dev, float *out_wts_dev, float *out_world_dev, double *out_ray_dev,
                            double *out_points_dev, int64_t *out_idx_dev, double *out_w_dev, void *stream)
{
    return gather_common(nullptr, depth_dev, world_in_dev, h, w, Kinv_host, E_host, origin_host, resolution, tsdf_dev, wvol_dev, X, Y,
                         Z, P, out_vals_dev, out_wts_dev, out_world_dev, out_ray_dev, out_points_dev, out_idx_dev, out_w_dev,
                         Pack{nullptr, nullptr, nullptr, nullptr, 0}, stream);
}

extern "C" int ojdf_gather(const double *ray_dev, int h, int w, const void *tsdf_dev, const void *wvol_dev, int X, int Y, int Z,
                           int P, float *out_vals_dev, float *out_wts_dev, double *out_points_dev, int64_t *out_idx_dev,
                           double *out_w_dev, float *pack_a_dev, float *pack_b_dev, const float *last_a_dev,
                           const float *last_b_dev, int pack_stride, void *stream)
{
    if (!ray_dev) return OJDF_ERR_BADARG;
    return gather_common(ray_dev, nullptr, nullptr, h, w, nullptr, nullptr, nullptr, 1.0, tsdf_dev, wvol_dev, X, Y, Z, P, out_vals_dev,
                         out_wts_dev, nullptr, nullptr, out_points_dev, out_idx_dev, out_w_dev,
                         Pack{pack_a_dev, pack_b_dev, last_a_dev, last_b_dev, pack_stride}, stream);
}
