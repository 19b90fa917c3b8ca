// Internal (non-ABI) declarations shared by the translation units of libojdf.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ojdf.h"
#include "ojdf_geom.cuh"

namespace ojdf {

// Fill the by-value kernel parameter block.  The eye in voxel units is computed here in
// IEEE f64 (one rounding per op, exactly like the device intrinsics would).
void make_pose(Pose &P, const float *Kinv, const float *E, const double *origin, double res);

// Reduction half of a split-K convolution (implemented next to conv_reduce_kernel in ojdf_conv.cu).
struct SplitReduce {
    const float *scale, *shift, *residual;
    float *out, *partial;           // partial: [splits][npix][cpad]
    int out_stride, out_coff, res_stride;
};
int launch_split_reduce(const SplitReduce *problems, int n, int npix, int cout, int cpad, int splits, int act, float slope,
                        float out_mul, cudaStream_t s);

// Account `n` kernel launches and fold cudaGetLastError() into the ABI's return code.
int launched(int n);

}  // namespace ojdf
