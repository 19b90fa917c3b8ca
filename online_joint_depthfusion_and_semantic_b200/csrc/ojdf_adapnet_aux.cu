// The two ends of AdapNet++ that are not tap GEMMs (modules/adapnet.py:101,134-137; modules/pipeline.py:42-60,184):
//
//   stem_kernel       : conv1 7x7 / stride 2 / pad 3 (3 -> 64) + BatchNorm(eval) + ReLU + max-pool 3x3 / stride 2 /
//                       pad 1, NCHW image in, pixel-major (H/4 * W/4, 64) out -- one kernel instead of cuDNN conv + BN +
//                       ReLU + pool + a transpose.  A block owns a 6 x 16 tile of pooled pixels and 16 of the 64 channels:
//                       the 31 x 71 x 3 input patch and the 16 x 147 weights sit in shared memory, the 13 x 33 x 16
//                       convolution tile is computed once (fp32 FMA: cin = 3 is no tensor-core shape; a thread owns two
//                       adjacent positions x 16 channels), pooled from shared memory.
//   softmax_max_kernel: per pixel softmax over the class logits, its maximum (the score) and arg-max (the label),
//                       plus the normalised label frame (1 + id) / n_classes FusionNet's semantic head reads
//                       (modules/pipeline.py:57-58,96,184) -- one pass instead of softmax + permute + max + casts.
#include <cfloat>

#include "ojdf_internal.h"

namespace ojdf {

constexpr int kPT_H = 6, kPT_W = 16;                      // pooled tile
constexpr int kCT_H = 2 * kPT_H + 1, kCT_W = 2 * kPT_W + 1;   // conv tile 13 x 33
constexpr int kIT_H = 2 * kCT_H + 5, kIT_W = 2 * kCT_W + 5;   // input tile 31 x 71
constexpr int kIT_P = 76;                                 // its row pitch: the odd partner of the last column pair reads up to column 72
constexpr int kPairs = (kCT_W + 1) / 2;                   // 17 pairs of horizontally adjacent conv positions per row
constexpr int kStemThreads = 224;                         // 13 x 17 = 221 position pairs, one per thread
constexpr int kStemCout = 64, kStemTaps = 3 * 49, kStemCg = 16;   // a block computes 16 of the 64 output channels
constexpr int kStemIn = 3 * kIT_H * kIT_P;

struct StemProblem {
    const float *in;        // (3, H, W) f32
    const float *w;         // (147, 64): [ci*49 + ky*7 + kx][co]
    const float *scale, *shift;
    float *out;             // (H/4 * W/4, out_stride)
    int out_stride, out_coff;
};
struct StemBatch { StemProblem p[2]; };

// Grid (pooled tiles x, pooled tiles y, problems * 4 channel groups).  A thread owns two horizontally adjacent conv
// positions x 16 channels: per (ci, ky) it reads 9 input values and the 7 x 16 weights (a warp-wide broadcast) for
// 224 FMAs.  65 KB of shared memory per block: three blocks per SM, the whole 240x320 frame is one wave of 400 blocks.
// The summation order per output is (ci, ky, kx), one fmaf each.
__global__ void __launch_bounds__(kStemThreads, 3)
stem_kernel(StemBatch batch, int H, int W)
{
    extern __shared__ __align__(16) float smem[];
    float *s_in = smem;                                       // [3][kIT_H][kIT_P]
    float *s_w = s_in + kStemIn;                              // [147][16]
    float *s_conv = s_w + kStemTaps * kStemCg;                // [kCT_H * kCT_W][16]
    const StemProblem pr = batch.p[blockIdx.z >> 2];
    const int g = blockIdx.z & 3;
    const int Hc = H / 2, Wc = W / 2, Hp = H / 4, Wp = W / 4;
    const int py0 = blockIdx.y * kPT_H, px0 = blockIdx.x * kPT_W;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;           // first conv row / column of the tile
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;           // first input row / column
    for (int i = threadIdx.x; i < 3 * kIT_H * (kIT_P / 4); i += kStemThreads) {      // 19 float4 columns per patch row
        const int row = i / (kIT_P / 4), c4 = i - row * (kIT_P / 4);
        const int ci = row / kIT_H, r = row - ci * kIT_H;
        const int y = iy0 + r;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * c4 + e, x = ix0 + c;
            v[e] = (c < kIT_W && y >= 0 && y < H && x >= 0 && x < W) ? __ldg(pr.in + ((size_t)ci * H + y) * W + x) : 0.0f;
        }
        *reinterpret_cast<float4 *>(s_in + (size_t)row * kIT_P + 4 * c4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    for (int i = threadIdx.x; i < kStemTaps * (kStemCg / 4); i += kStemThreads)
        reinterpret_cast<float4 *>(s_w)[i] = __ldg(reinterpret_cast<const float4 *>(pr.w + (size_t)(i >> 2) * kStemCout + g * kStemCg) + (i & 3));
    __syncthreads();
    if (threadIdx.x < kCT_H * kPairs) {
        const int r = threadIdx.x / kPairs, cp = threadIdx.x - r * kPairs;
        float a0[kStemCg], a1[kStemCg];
#pragma unroll
        for (int j = 0; j < kStemCg; ++j) a0[j] = a1[j] = 0.0f;
#pragma unroll 1
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll 1
            for (int ky = 0; ky < 7; ++ky) {
                const float *irow = s_in + (size_t)(ci * kIT_H + 2 * r + ky) * kIT_P + 4 * cp;
                const float4 xa = *reinterpret_cast<const float4 *>(irow), xb = *reinterpret_cast<const float4 *>(irow + 4);
                const float x[9] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w, irow[8]};
                const float4 *w4 = reinterpret_cast<const float4 *>(s_w + (size_t)((ci * 7 + ky) * 7) * kStemCg);
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 wv = w4[kx * 4 + q];
                        a0[4 * q] = fmaf(x[kx], wv.x, a0[4 * q]); a0[4 * q + 1] = fmaf(x[kx], wv.y, a0[4 * q + 1]);
                        a0[4 * q + 2] = fmaf(x[kx], wv.z, a0[4 * q + 2]); a0[4 * q + 3] = fmaf(x[kx], wv.w, a0[4 * q + 3]);
                        a1[4 * q] = fmaf(x[kx + 2], wv.x, a1[4 * q]); a1[4 * q + 1] = fmaf(x[kx + 2], wv.y, a1[4 * q + 1]);
                        a1[4 * q + 2] = fmaf(x[kx + 2], wv.z, a1[4 * q + 2]); a1[4 * q + 3] = fmaf(x[kx + 2], wv.w, a1[4 * q + 3]);
                    }
                }
            }
        const int cy = cy0 + r;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c = 2 * cp + half, cx = cx0 + c;
            if (c >= kCT_W) continue;
            float *o = s_conv + (size_t)(r * kCT_W + c) * kStemCg;
            const bool inside = cy >= 0 && cy < Hc && cx >= 0 && cx < Wc;     // outside the conv map: never wins the max-pool
#pragma unroll
            for (int j = 0; j < kStemCg; ++j) {
                const int co = g * kStemCg + j;
                const float v = fmaf(half ? a1[j] : a0[j], __ldg(pr.scale + co), __ldg(pr.shift + co));   // BatchNorm (eval) folded
                o[j] = inside ? (v > 0.0f ? v : 0.0f) : -FLT_MAX;                                       // ReLU
            }
        }
    }
    __syncthreads();
    // max-pool 3x3 / 2 / pad 1 over the conv tile: pooled (py, px) <- conv rows 2py-1..2py+1 = tile rows 2*lpy..2*lpy+2
    for (int i = threadIdx.x; i < kPT_H * kPT_W * kStemCg; i += kStemThreads) {
        const int co = i % kStemCg, lp = i / kStemCg, lpy = lp / kPT_W, lpx = lp % kPT_W;
        const int py = py0 + lpy, px = px0 + lpx;
        if (py >= Hp || px >= Wp) continue;
        float m = -FLT_MAX;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) m = fmaxf(m, s_conv[(size_t)((2 * lpy + dy) * kCT_W + 2 * lpx + dx) * kStemCg + co]);
        pr.out[(size_t)(py * Wp + px) * pr.out_stride + pr.out_coff + g * kStemCg + co] = m;
    }
}

__global__ void __launch_bounds__(256)
softmax_max_kernel(const float *__restrict__ logits, int stride, int C, int npix, int n_classes, float *__restrict__ scores,
                   uint8_t *__restrict__ ids, float *__restrict__ sem_frame)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float *l = logits + (size_t)p * stride;
    float m = l[0];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
        const float v = l[c];
        if (v > m) { m = v; arg = c; }                     // first maximum wins, like torch.max on the CPU
    }
    float sum = 0.0f;
    for (int c = 0; c < C; ++c) sum += expf(l[c] - m);
    scores[p] = 1.0f / sum;                                // softmax value of the arg-max class: exp(0) / sum
    ids[p] = (uint8_t)arg;
    if (sem_frame) sem_frame[p] = (1.0f + (float)arg) / (float)n_classes;
}

}  // namespace ojdf

using namespace ojdf;

extern "C" int ojdf_adapnet_stem(const ojdf_stem_problem *problems_host, int n_problems, int H, int W, void *stream)
{
    if (!problems_host || n_problems < 1 || n_problems > 2 || H < 16 || W < 16 || (H & 3) || (W & 3)) return OJDF_ERR_BADARG;
    StemBatch b;
    for (int i = 0; i < n_problems; ++i) {
        const ojdf_stem_problem &q = problems_host[i];
        if (!q.in_dev || !q.weights_dev || !q.scale_dev || !q.shift_dev || !q.out_dev || q.out_stride < q.out_coffset + kStemCout ||
            ((uintptr_t)q.weights_dev & 15))
            return OJDF_ERR_BADARG;
        b.p[i] = StemProblem{q.in_dev, q.weights_dev, q.scale_dev, q.shift_dev, q.out_dev, q.out_stride, q.out_coffset};
    }
    const size_t smem = (size_t)(kStemIn + kStemTaps * kStemCg + kCT_H * kCT_W * kStemCg) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const dim3 grid((W / 4 + kPT_W - 1) / kPT_W, (H / 4 + kPT_H - 1) / kPT_H, n_problems * (kStemCout / kStemCg));
    stem_kernel<<<grid, kStemThreads, smem, (cudaStream_t)stream>>>(b, H, W);
    return launched(1);
}

extern "C" int ojdf_softmax_max(const float *logits_dev, int stride, int C, int npix, int n_classes, float *scores_dev,
                                uint8_t *ids_dev, float *sem_frame_dev, void *stream)
{
    if (!logits_dev || !scores_dev || !ids_dev || C < 1 || C > 256 || stride < C || npix < 1 || n_classes < 1) return OJDF_ERR_BADARG;
    softmax_max_kernel<<<(npix + 255) / 256, 256, 0, (cudaStream_t)stream>>>(logits_dev, stride, C, npix, n_classes, scores_dev, ids_dev,
                                                                           sem_frame_dev);
    return launched(1);
}
