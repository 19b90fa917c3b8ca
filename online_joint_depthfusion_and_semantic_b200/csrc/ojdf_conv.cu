// FusionNet convolution stack (modules/model.py:4-283 of the reference) as fp32-exact sm_100a
// kernels on pixel-major (NHWC) activations.
//
// Every FusionNet layer is a small-channel "tap GEMM": out[p, co] = act(scale[co] * sum_tap sum_ci
// in[p + tap*dil, ci] * W[tap, ci, co] + shift[co]) with taps = 1 (1x1) or 9 (3x3, any dilation),
// Cin in {19..570} and Cout in {9, 19, .., 114}.  The reference issues one cuDNN convolution, one
// batch-norm and one activation kernel per layer and materialises every torch.cat; here
//   * BatchNorm (inference statistics), bias and the activation are the epilogue of the conv kernel,
//   * dense-block / vortex concatenations are channel offsets into one pixel-major buffer,
//   * the global-average-pool branch of VortexPooling collapses into a per-frame bias vector.
// The arithmetic stays fp32 FMA (parity within 1e-5 of the reference's fp32 convolutions); the
// tensor-core (tcgen05, bf16 / 3xTF32) version of the same tap-GEMM is the next step.
//
// conv kernel: one thread owns P pixels x 20 output channels in registers; the layer's weights for
// one 20-channel output group live in shared memory as [tap][ci][20] and are read as broadcast
// LDS.128; inputs are read straight from global/L1 as float4 (4 channels) per pixel and tap.
#include "ojdf_internal.h"

namespace ojdf {

constexpr int kGroup = 20;          // output channels per thread (19 padded to 20 for FusionNet)
constexpr int kConvThreads = 128;
constexpr int kPix = 2;             // pixels per thread

enum Act { kNone = 0, kRelu = 1, kLeaky = 2, kTanh = 3 };

__device__ __forceinline__ float activate(float v, int act, float slope)
{
    if (act == kRelu) return v > 0.0f ? v : 0.0f;
    if (act == kLeaky) return v > 0.0f ? v : v * slope;
    if (act == kTanh) return tanhf(v);
    return v;
}

// weights: [groups][taps][cin4*4][kGroup] fp32, zero padded.  Dynamic smem: taps*cin4*4*kGroup floats.
template <int TAPS>
__global__ void __launch_bounds__(kConvThreads)
conv_taps_kernel(const float *__restrict__ in, int in_stride, int cin, int H, int W, int dil,
                 const float *__restrict__ weights, const float *__restrict__ scale, const float *__restrict__ shift,
                 int cout, int act, float slope, float out_mul, float *__restrict__ out, int out_stride, int out_coff)
{
    extern __shared__ float4 s_w[];
    const int cin4 = (cin + 3) >> 2;
    const int g = blockIdx.y;
    const int wcount4 = TAPS * cin4 * 4 * (kGroup / 4);
    const float4 *wg = reinterpret_cast<const float4 *>(weights) + (size_t)g * wcount4;
    for (int i = threadIdx.x; i < wcount4; i += blockDim.x) s_w[i] = __ldg(wg + i);
    __syncthreads();

    const int npix = H * W;
    const int p0 = blockIdx.x * (kConvThreads * kPix) + threadIdx.x;
    int py[kPix], px[kPix];
    bool live[kPix];
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
        const int p = p0 + j * kConvThreads;
        live[j] = p < npix;
        py[j] = live[j] ? p / W : 0;
        px[j] = live[j] ? p - py[j] * W : 0;
    }
    float acc[kPix][kGroup];
#pragma unroll
    for (int j = 0; j < kPix; ++j)
#pragma unroll
        for (int c = 0; c < kGroup; ++c) acc[j][c] = 0.0f;

    const int tail = cin & 3;                      // channels in the last float4 that really exist (0 = all four)
#pragma unroll 1
    for (int tap = 0; tap < TAPS; ++tap) {
        const int dy = TAPS == 1 ? 0 : (tap / 3 - 1) * dil, dx = TAPS == 1 ? 0 : (tap % 3 - 1) * dil;
        const float4 *src[kPix];
        bool ok[kPix];
#pragma unroll
        for (int j = 0; j < kPix; ++j) {
            const int y = py[j] + dy, x = px[j] + dx;
            ok[j] = live[j] && y >= 0 && y < H && x >= 0 && x < W;         // zero padding
            src[j] = reinterpret_cast<const float4 *>(in + (size_t)(ok[j] ? y * W + x : 0) * in_stride);
        }
        const float4 *wt = s_w + tap * cin4 * 4 * (kGroup / 4);
#pragma unroll 2
        for (int c4 = 0; c4 < cin4; ++c4) {
            float4 xv[kPix];
#pragma unroll
            for (int j = 0; j < kPix; ++j) {
                xv[j] = ok[j] ? __ldg(src[j] + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (c4 == cin4 - 1 && tail) {      // never let a neighbouring tensor's channels in
                    if (tail < 2) xv[j].y = 0.f;
                    if (tail < 3) xv[j].z = 0.f;
                    xv[j].w = 0.f;
                }
            }
            const float4 *wr = wt + c4 * 4 * (kGroup / 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int q = 0; q < kGroup / 4; ++q) {
                    const float4 wv = wr[k * (kGroup / 4) + q];
#pragma unroll
                    for (int j = 0; j < kPix; ++j) {
                        const float xs = k == 0 ? xv[j].x : (k == 1 ? xv[j].y : (k == 2 ? xv[j].z : xv[j].w));
                        acc[j][4 * q + 0] = fmaf(xs, wv.x, acc[j][4 * q + 0]);
                        acc[j][4 * q + 1] = fmaf(xs, wv.y, acc[j][4 * q + 1]);
                        acc[j][4 * q + 2] = fmaf(xs, wv.z, acc[j][4 * q + 2]);
                        acc[j][4 * q + 3] = fmaf(xs, wv.w, acc[j][4 * q + 3]);
                    }
                }
            }
        }
    }
    const int co0 = g * kGroup;
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
        if (!live[j]) continue;
        float *o = out + (size_t)(p0 + j * kConvThreads) * out_stride + out_coff + co0;
#pragma unroll
        for (int c = 0; c < kGroup; ++c) {
            if (co0 + c < cout) {
                const float v = fmaf(acc[j][c], __ldg(scale + co0 + c), __ldg(shift + co0 + c));
                o[c] = activate(v, act, slope) * out_mul;
            }
        }
    }
}

// 3x3 average pool, stride 1, zero padding counted in the divisor (nn.AvgPool2d default), NHWC.
__global__ void __launch_bounds__(256)
avgpool3_kernel(const float *__restrict__ in, int in_stride, int H, int W, int C, float *__restrict__ out, int out_stride)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4n = C >> 2;
    if (i >= (long long)H * W * c4n) return;
    const int c4 = (int)(i % c4n);
    const int p = (int)(i / c4n), y = p / W, x = p - y * W;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(in + (size_t)(yy * W + xx) * in_stride) + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    reinterpret_cast<float4 *>(out + (size_t)p * out_stride)[c4] = make_float4(s.x / 9.0f, s.y / 9.0f, s.z / 9.0f, s.w / 9.0f);
}

// Per-channel sums over all pixels (global average pool numerator): partial sums per block into
// `partial` [gridDim.x][C]; the tiny second stage runs inside vortex_bias_kernel.
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float *__restrict__ in, int in_stride, int npix, int C, float *__restrict__ partial)
{
    // thread = channel (C <= 256), block strides over pixels
    const int c = threadIdx.x;
    if (c >= C) return;
    float s = 0.0f;
    for (int p = blockIdx.x; p < npix; p += gridDim.x) s += __ldg(in + (size_t)p * in_stride + c);
    partial[(size_t)blockIdx.x * C + c] = s;
}

// VortexPooling's global branch (modules/model.py:107-112): mean -> 1x1 conv -> (bilinear upsample
// of a 1x1 map = constant) -> BatchNorm gives a per-channel constant v1; the `final` 1x1 conv sees it
// as the bias  shift_out[co] = final_shift[co] + final_scale[co] * sum_c Wf[co, c] * v1[c].
__global__ void __launch_bounds__(256)
vortex_bias_kernel(const float *__restrict__ partial, int nblocks, int npix, int C,
                   const float *__restrict__ wg /* [Cg][C] */, const float *__restrict__ g_scale, const float *__restrict__ g_shift, int Cg,
                   const float *__restrict__ wf1 /* [Cout][Cg] */, const float *__restrict__ f_scale, const float *__restrict__ f_shift, int Cout,
                   float *__restrict__ shift_out)
{
    __shared__ float s_mean[256];
    __shared__ float s_v1[256];
    const int t = threadIdx.x;
    if (t < C) {
        float s = 0.0f;
        for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * C + t];
        s_mean[t] = s / (float)npix;
    }
    __syncthreads();
    if (t < Cg) {
        float a = 0.0f;
        for (int c = 0; c < C; ++c) a = fmaf(wg[(size_t)t * C + c], s_mean[c], a);
        s_v1[t] = fmaf(a, g_scale[t], g_shift[t]);
    }
    __syncthreads();
    if (t < Cout) {
        float a = 0.0f;
        for (int c = 0; c < Cg; ++c) a = fmaf(wf1[(size_t)t * Cg + c], s_v1[c], a);
        shift_out[t] = fmaf(a, f_scale[t], f_shift[t]);
    }
}

// Network input assembly (modules/pipeline.py:74-102 + modules/model.py:269,274): pixel-major
// [values(P) | weights(P) | last channel] into channel offset 0 of a buffer with `stride` floats per
// pixel; head A's last channel is the depth frame, head B's (optional) the normalised label frame.
__global__ void __launch_bounds__(256)
pack_input_kernel(const float *__restrict__ vals, const float *__restrict__ wts, const float *__restrict__ last_a,
                  const float *__restrict__ last_b, int npix, int P, float *__restrict__ out_a, float *__restrict__ out_b,
                  int stride)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float *a = out_a + (size_t)p * stride;
    float *b = out_b ? out_b + (size_t)p * stride : nullptr;
    for (int k = 0; k < P; ++k) {
        const float v = vals[(size_t)p * P + k], w = wts[(size_t)p * P + k];
        a[k] = v; a[P + k] = w;
        if (b) { b[k] = v; b[P + k] = w; }
    }
    a[2 * P] = last_a[p];
    if (b) b[2 * P] = last_b[p];
}

}  // namespace ojdf

using namespace ojdf;

extern "C" int ojdf_conv_nhwc(const float *in_dev, int in_stride, int cin, int H, int W, int taps, int dilation,
                              const float *weights_dev, const float *scale_dev, const float *shift_dev, int cout,
                              int act, float slope, float out_mul, float *out_dev, int out_stride, int out_coffset,
                              void *stream)
{
    if (!in_dev || !weights_dev || !scale_dev || !shift_dev || !out_dev || cin < 1 || cout < 1 || H < 1 || W < 1 ||
        (taps != 1 && taps != 9) || dilation < 1 || (in_stride & 3) || in_stride < ((cin + 3) & ~3) ||
        out_stride < out_coffset + cout || act < 0 || act > 3)
        return OJDF_ERR_BADARG;
    const int cin4 = (cin + 3) >> 2;
    const size_t smem = (size_t)taps * cin4 * 4 * kGroup * sizeof(float);
    if (smem > 200 * 1024) return OJDF_ERR_TOOLARGE;
    const int npix = H * W;
    dim3 grid((npix + kConvThreads * kPix - 1) / (kConvThreads * kPix), (cout + kGroup - 1) / kGroup);
    cudaStream_t s = (cudaStream_t)stream;
    if (taps == 1) {
        static bool attr1 = false;
        if (!attr1) { cudaFuncSetAttribute(conv_taps_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr1 = true; }
        conv_taps_kernel<1><<<grid, kConvThreads, smem, s>>>(in_dev, in_stride, cin, H, W, dilation, weights_dev, scale_dev,
                                                             shift_dev, cout, act, slope, out_mul, out_dev, out_stride, out_coffset);
    } else {
        static bool attr9 = false;
        if (!attr9) { cudaFuncSetAttribute(conv_taps_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr9 = true; }
        conv_taps_kernel<9><<<grid, kConvThreads, smem, s>>>(in_dev, in_stride, cin, H, W, dilation, weights_dev, scale_dev,
                                                             shift_dev, cout, act, slope, out_mul, out_dev, out_stride, out_coffset);
    }
    return launched(1);
}

extern "C" int ojdf_avgpool3_nhwc(const float *in_dev, int in_stride, int H, int W, int C, float *out_dev, int out_stride,
                                  void *stream)
{
    if (!in_dev || !out_dev || H < 1 || W < 1 || C < 4 || (C & 3) || (in_stride & 3) || (out_stride & 3) || in_stride < C ||
        out_stride < C)
        return OJDF_ERR_BADARG;
    const long long n = (long long)H * W * (C >> 2);
    avgpool3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in_dev, in_stride, H, W, C, out_dev, out_stride);
    return launched(1);
}

extern "C" int ojdf_vortex_bias(const float *in_dev, int in_stride, int npix, int C, const float *wg_dev,
                                const float *g_scale_dev, const float *g_shift_dev, int Cg, const float *wf1_dev,
                                const float *f_scale_dev, const float *f_shift_dev, int Cout, float *partial_dev,
                                int partial_blocks, float *shift_out_dev, void *stream)
{
    if (!in_dev || !wg_dev || !g_scale_dev || !g_shift_dev || !wf1_dev || !f_scale_dev || !f_shift_dev || !partial_dev ||
        !shift_out_dev || npix < 1 || C < 1 || C > 256 || Cg < 1 || Cg > 256 || Cout < 1 || Cout > 256 || in_stride < C ||
        partial_blocks < 1)
        return OJDF_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    channel_sum_kernel<<<partial_blocks, 256, 0, s>>>(in_dev, in_stride, npix, C, partial_dev);
    vortex_bias_kernel<<<1, 256, 0, s>>>(partial_dev, partial_blocks, npix, C, wg_dev, g_scale_dev, g_shift_dev, Cg, wf1_dev,
                                        f_scale_dev, f_shift_dev, Cout, shift_out_dev);
    return launched(2);
}

extern "C" int ojdf_pack_fusion_input(const float *vals_dev, const float *wts_dev, const float *last_a_dev,
                                      const float *last_b_dev, int npix, int P, float *out_a_dev, float *out_b_dev,
                                      int stride, void *stream)
{
    if (!vals_dev || !wts_dev || !last_a_dev || !out_a_dev || npix < 1 || P < 1 || stride < 2 * P + 1 ||
        (out_b_dev && !last_b_dev))
        return OJDF_ERR_BADARG;
    pack_input_kernel<<<(npix + 255) / 256, 256, 0, (cudaStream_t)stream>>>(vals_dev, wts_dev, last_a_dev, last_b_dev, npix, P,
                                                                            out_a_dev, out_b_dev, stride);
    return launched(1);
}
