// Device-side ray / corner geometry shared by the gather and scatter kernels.
//
// The arithmetic restates modules/extractor.py:309-345 (ray samples) and :533-593
// (interpolation_weights) of the reference operation by operation: the reference runs
// these in float64 with every product and sum rounded separately (ATen elementwise
// kernels), so everything here uses the explicit round-to-nearest intrinsics and the
// library is additionally compiled with -fmad=false.  Voxel indices come out bit-exact.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ojdf {

struct Pose {
    float kinv[9];      // intrinsics.float().inverse(), row major
    float e[12];        // cam->world rows 0..2, row major (3x4)
    double origin[3];
    double ev[3];       // (double(eye) - origin) / resolution, computed on the host in IEEE f64
    double res;
};

// f32 FMA chain of the two small matmuls in Extractor.compute_coordinates
// (modules/extractor.py:113-117; order: SURVEY.md App. A.1).
__device__ __forceinline__ void unproject_pixel(const Pose &P, int r, int c, float z, float out[3])
{
    const float p0 = __fmul_rn((float)c, z), p1 = __fmul_rn((float)r, z), p2 = z;
    float q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float t = __fmul_rn(P.kinv[3 * i], p0);
        t = __fmaf_rn(P.kinv[3 * i + 1], p1, t);
        t = __fmaf_rn(P.kinv[3 * i + 2], p2, t);
        q[i] = t;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float t = __fmul_rn(P.e[4 * i], q[0]);
        t = __fmaf_rn(P.e[4 * i + 1], q[1], t);
        t = __fmaf_rn(P.e[4 * i + 2], q[2], t);
        t = __fmaf_rn(P.e[4 * i + 3], 1.0f, t);
        out[i] = t;
    }
}

// Per-ray record: centre point in voxel units and the unit eye->point direction
// (modules/extractor.py:314-318).  norm = sqrt((dx*dx + dy*dy) + dz*dz), eps clamp 1e-12.
__device__ __forceinline__ void ray_record(const Pose &P, const float world[3], double rec[6])
{
    double dl[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rec[a] = __ddiv_rn(__dsub_rn((double)world[a], P.origin[a]), P.res);
        dl[a] = __dsub_rn(rec[a], P.ev[a]);
    }
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(dl[0], dl[0]), __dmul_rn(dl[1], dl[1])), __dmul_rn(dl[2], dl[2]));
    const double nrm = __dsqrt_rn(s);
    const double den = nrm > 1e-12 ? nrm : 1e-12;
#pragma unroll
    for (int a = 0; a < 3; ++a) rec[3 + a] = __ddiv_rn(dl[a], den);
}

// Sample i (signed offset from the centre sample) of a ray: c + fl(i*d)
// (modules/extractor.py:326-330; cv - fl(i*d) == cv + fl(-i*d) bit for bit).
__device__ __forceinline__ double ray_sample(double c, double d, int i)
{
    return i == 0 ? c : __dadd_rn(c, __dmul_rn((double)i, d));
}

// Per-axis pieces of interpolation_weights (modules/extractor.py:535-556).
struct Axis {
    long long i0, i1;   // floor index, neighbour index
    double a, ai;       // alpha = |p - centre|, 1 - alpha
};

__device__ __forceinline__ Axis axis_setup(double p)
{
    Axis ax;
    const double fl = floor(p);
    const double ctr = __dadd_rn(fl, 0.5);
    const double df = __dsub_rn(ctr, p);
    const double nb = (double)((df > 0.0) - (df < 0.0));       // torch.sign
    ax.a = fabs(__dsub_rn(p, ctr));
    ax.ai = __dsub_rn(1.0, ax.a);
    // double -> int64 of a non-finite value: the reference's CPU conversion yields INT64_MIN (so the corner fails
    // get_index_mask and is dropped, modules/extractor.py:596-607), CUDA's yields 0 for NaN -- an in-grid voxel that a
    // NaN depth pixel would then poison.  Reproduce the CPU result.
    const bool fin = isfinite(fl);
    ax.i0 = fin ? (long long)fl : (long long)0x8000000000000000ull;
    ax.i1 = fin ? (long long)__dadd_rn(fl, nb) : (long long)0x8000000000000000ull;
    return ax;
}

// Corner c = 4i + 2j + k (i: x, slowest): weight ((w1*w2)*w3), modules/extractor.py:564-585.
__device__ __forceinline__ double corner_weight(const Axis &x, const Axis &y, const Axis &z, int c)
{
    const double w1 = (c & 4) ? x.a : x.ai, w2 = (c & 2) ? y.a : y.ai, w3 = (c & 1) ? z.a : z.ai;
    return __dmul_rn(__dmul_rn(w1, w2), w3);
}

__device__ __forceinline__ bool corner_index(const Axis &x, const Axis &y, const Axis &z, int c, int X, int Y, int Z,
                                             long long &ix, long long &iy, long long &iz)
{
    ix = (c & 4) ? x.i1 : x.i0;
    iy = (c & 2) ? y.i1 : y.i0;
    iz = (c & 1) ? z.i1 : z.i0;
    return ix >= 0 && ix < X && iy >= 0 && iy < Y && iz >= 0 && iz < Z;     // get_index_mask, :596-607
}

}  // namespace ojdf
