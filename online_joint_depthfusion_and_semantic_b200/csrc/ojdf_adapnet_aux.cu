// The two ends of AdapNet++ that are not tap GEMMs (modules/adapnet.py:101,134-137; modules/pipeline.py:42-60,184):
//
//   stem_kernel       : conv1 7x7 / stride 2 / pad 3 (3 -> 64) + BatchNorm(eval) + ReLU + max-pool 3x3 / stride 2 /
//                       pad 1, NCHW image in, pixel-major (H/4 * W/4, 64) out -- one kernel instead of cuDNN conv + BN +
//                       ReLU + pool + a transpose.  A block owns a 4 x 8 tile of pooled pixels: the 23 x 39 x 3 input
//                       patch and the 64 x 147 weights sit in shared memory, the 9 x 17 x 64 convolution tile is
//                       computed once (fp32 FMA: cin = 3 is no tensor-core shape), pooled from shared memory.
//   softmax_max_kernel: per pixel softmax over the class logits, its maximum (the score) and arg-max (the label),
//                       plus the normalised label frame (1 + id) / n_classes FusionNet's semantic head reads
//                       (modules/pipeline.py:57-58,96,184) -- one pass instead of softmax + permute + max + casts.
#include <cfloat>

#include "ojdf_internal.h"

namespace ojdf {

constexpr int kStemThreads = 256;
constexpr int kPT_H = 4, kPT_W = 8;                       // pooled tile
constexpr int kCT_H = 2 * kPT_H + 1, kCT_W = 2 * kPT_W + 1;   // conv tile 9 x 17
constexpr int kIT_H = 2 * kCT_H + 5, kIT_W = 2 * kCT_W + 5;   // input tile 23 x 39
constexpr int kStemCout = 64, kStemTaps = 3 * 49;
constexpr int kStemIn = (3 * kIT_H * kIT_W + 3) & ~3;        // input patch, padded so the weights that follow stay 16-byte aligned

struct StemProblem {
    const float *in;        // (3, H, W) f32
    const float *w;         // (147, 64): [ci*49 + ky*7 + kx][co]
    const float *scale, *shift;
    float *out;             // (H/4 * W/4, out_stride)
    int out_stride, out_coff;
};
struct StemBatch { StemProblem p[2]; };

__global__ void __launch_bounds__(kStemThreads)
stem_kernel(StemBatch batch, int H, int W)
{
    extern __shared__ __align__(16) float smem[];
    float *s_in = smem;                                       // [3][kIT_H][kIT_W]
    float *s_w = s_in + kStemIn;                              // [147][64]
    float *s_conv = s_w + kStemTaps * kStemCout;              // [kCT_H * kCT_W][64]
    const StemProblem pr = batch.p[blockIdx.z];
    const int Hc = H / 2, Wc = W / 2, Hp = H / 4, Wp = W / 4;
    const int py0 = blockIdx.y * kPT_H, px0 = blockIdx.x * kPT_W;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;           // first conv row / column of the tile
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;           // first input row / column
    for (int i = threadIdx.x; i < 3 * kIT_H * kIT_W; i += kStemThreads) {
        const int ci = i / (kIT_H * kIT_W), r = (i / kIT_W) % kIT_H, c = i % kIT_W;
        const int y = iy0 + r, x = ix0 + c;
        s_in[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(pr.in + ((size_t)ci * H + y) * W + x) : 0.0f;
    }
    for (int i = threadIdx.x; i < kStemTaps * kStemCout / 4; i += kStemThreads)
        reinterpret_cast<float4 *>(s_w)[i] = __ldg(reinterpret_cast<const float4 *>(pr.w) + i);
    __syncthreads();
    // conv tile: work item = (position, group of 16 output channels)
#pragma unroll 1
    for (int item = threadIdx.x; item < kCT_H * kCT_W * 4; item += kStemThreads) {
        const int pos = item >> 2, g = item & 3;
        const int r = pos / kCT_W, c = pos % kCT_W;
        const int cy = cy0 + r, cx = cx0 + c;
        float *o = s_conv + (size_t)pos * kStemCout + g * 16;
        if (cy < 0 || cy >= Hc || cx < 0 || cx >= Wc) {       // outside the conv map: never wins the max-pool
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = -FLT_MAX;
            continue;
        }
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
#pragma unroll 1
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll 1                                           // (fully unrolled this nest is 650 KB of code: it lives in L2, not in the I-cache)
            for (int ky = 0; ky < 7; ++ky) {
                const float *irow = s_in + (ci * kIT_H + 2 * r + ky) * kIT_W + 2 * c;
                const float *wrow = s_w + (size_t)((ci * 7 + ky) * 7) * kStemCout + g * 16;
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const float x = irow[kx];
                    const float4 *w4 = reinterpret_cast<const float4 *>(wrow + kx * kStemCout);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 wv = w4[q];
                        acc[4 * q] = fmaf(x, wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(x, wv.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(x, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x, wv.w, acc[4 * q + 3]);
                    }
                }
            }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int co = g * 16 + j;
            const float v = fmaf(acc[j], __ldg(pr.scale + co), __ldg(pr.shift + co));   // BatchNorm (eval) folded
            o[j] = v > 0.0f ? v : 0.0f;                                                  // ReLU
        }
    }
    __syncthreads();
    // max-pool 3x3 / 2 / pad 1 over the conv tile: pooled (py, px) <- conv rows 2py-1..2py+1 = tile rows 2*lpy..2*lpy+2
    for (int i = threadIdx.x; i < kPT_H * kPT_W * kStemCout; i += kStemThreads) {
        const int co = i % kStemCout, lp = i / kStemCout, lpy = lp / kPT_W, lpx = lp % kPT_W;
        const int py = py0 + lpy, px = px0 + lpx;
        if (py >= Hp || px >= Wp) continue;
        float m = -FLT_MAX;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) m = fmaxf(m, s_conv[(size_t)((2 * lpy + dy) * kCT_W + 2 * lpx + dx) * kStemCout + co]);
        pr.out[(size_t)(py * Wp + px) * pr.out_stride + pr.out_coff + co] = m;
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
    const size_t smem = (size_t)(kStemIn + kStemTaps * kStemCout + kCT_H * kCT_W * kStemCout) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const dim3 grid((W / 4 + kPT_W - 1) / kPT_W, (H / 4 + kPT_H - 1) / kPT_H, n_problems);
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
