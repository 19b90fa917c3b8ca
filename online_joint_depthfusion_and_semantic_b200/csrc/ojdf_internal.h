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

// Account `n` kernel launches and fold cudaGetLastError() into the ABI's return code.
int launched(int n);

}  // namespace ojdf
