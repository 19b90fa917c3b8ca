// Small kernels around the tensor-core convolutions of FusionNet / AdapNet++ (modules/model.py:4-283,
// modules/adapnet.py:152-216), pixel-major (NHWC) fp32 activations:
//   * conv_reduce_kernel     : second half of a split-K convolution (fixed summation order, the layer's epilogue);
//   * avgpool3 kernels       : VortexPooling's cascaded 3x3 average pools (modules/model.py:114-116), batched, with an
//                              optional scale/shift/ReLU epilogue (the pools commute with the branch's first 1x1 conv);
//   * channel_sum / gap_bias : the global-average-pool branches of VortexPooling and eASPP collapse into a per-frame bias
//                              vector of the block's final 1x1 convolution;
//   * pack_input             : FusionNet's input rows [values | weights | depth or label] (modules/pipeline.py:74-102);
//   * NCHW <-> NHWC transposes for the tail-only AdapNet++ engine.
// The convolutions themselves live in ojdf_conv_tc.cu (A operand in tensor memory) and ojdf_conv_ss.cu (both operands
// in shared memory).
#include <numeric>

#include "ojdf_internal.h"

namespace ojdf {

constexpr int kMaxBatch = 8;

enum Act { kNone = 0, kRelu = 1, kLeaky = 2, kTanh = 3, kSigmoid = 4, kSigmoidMul = 5 };

struct ConvProblem {
    const float *in; const float *weights; const float *scale; const float *shift; float *out;
    const float *residual;          // optional: added before the activation (ResNet shortcut), row stride res_stride
    float *partial;                 // split-K scratch of this problem: [splits][npix][groups*kGroup]
    int in_stride, out_stride, out_coff, dil, res_stride;
};
struct ConvBatch { ConvProblem p[kMaxBatch]; };

__device__ __forceinline__ float activate(float v, int act, float slope)
{
    if (act == kRelu) return v > 0.0f ? v : 0.0f;
    if (act == kLeaky) return v > 0.0f ? v : v * slope;
    if (act == kTanh) return tanhf(v);
    if (act == kSigmoid) return 1.0f / (1.0f + expf(-v));
    return v;
}

// Second half of a split-K convolution: sum the partials in split order, then the same epilogue.
__global__ void __launch_bounds__(256)
conv_reduce_kernel(ConvBatch batch, int npix, int cout, int cpad, int splits, int act, float slope, float out_mul)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");         // launched behind the convolution programmatically: its partials are complete from here on
    const ConvProblem pr = batch.p[blockIdx.y];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)npix * cout) return;
    const int p = (int)(i / cout), c = (int)(i - (long long)p * cout);
    float a = 0.0f;
    for (int s = 0; s < splits; ++s) a += pr.partial[((size_t)s * npix + p) * cpad + c];
    float v = fmaf(a, __ldg(pr.scale + c), __ldg(pr.shift + c));
    if (act == kSigmoidMul) {                                   // SSMA gate: sigmoid(conv) * gated tensor
        pr.out[(size_t)p * pr.out_stride + pr.out_coff + c] = pr.residual[(size_t)p * pr.res_stride + c] / (1.0f + expf(-v)) * out_mul;
        return;
    }
    if (pr.residual) v += pr.residual[(size_t)p * pr.res_stride + c];
    pr.out[(size_t)p * pr.out_stride + pr.out_coff + c] = activate(v, act, slope) * out_mul;
}

// The same with four channels per thread (128-bit loads; cout, cpad, the row strides and the channel offset multiples of 4)
// and the slice loads of a thread issued eight at a time before they are summed (in slice order, like the scalar kernel).
__global__ void __launch_bounds__(256)
conv_reduce4_kernel(ConvBatch batch, int npix, int cout, int cpad, int splits, int act, float slope, float out_mul)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const ConvProblem pr = batch.p[blockIdx.y];
    const int c4n = cout >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)npix * c4n) return;
    const int p = (int)(i / c4n), c = (int)(i - (long long)p * c4n) * 4;
    const float *src = pr.partial + (size_t)p * cpad + c;
    const size_t slab = (size_t)npix * cpad;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 8 <= splits; s += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)(s + j) * slab));
#pragma unroll
        for (int j = 0; j < 8; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
    }
    if (s + 4 <= splits) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)(s + j) * slab));
#pragma unroll
        for (int j = 0; j < 4; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
        s += 4;
    }
    for (; s < splits; ++s) {
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)s * slab));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(pr.scale + c)), sh = __ldg(reinterpret_cast<const float4 *>(pr.shift + c));
    float v[4] = {fmaf(a.x, sc.x, sh.x), fmaf(a.y, sc.y, sh.y), fmaf(a.z, sc.z, sh.z), fmaf(a.w, sc.w, sh.w)};
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pr.residual) r = *reinterpret_cast<const float4 *>(pr.residual + (size_t)p * pr.res_stride + c);
    const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (act == kSigmoidMul) v[j] = rr[j] / (1.0f + expf(-v[j])) * out_mul;
        else v[j] = activate(pr.residual ? v[j] + rr[j] : v[j], act, slope) * out_mul;
    }
    *reinterpret_cast<float4 *>(pr.out + (size_t)p * pr.out_stride + pr.out_coff + c) = make_float4(v[0], v[1], v[2], v[3]);
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

// Up to 8 independent 3x3 average pools of equal shape in one launch, with an optional per-channel
// scale/shift (+ ReLU) on the way out.  VortexPooling feeds branch b with pool^b(x) followed by a 1x1
// convolution + BatchNorm + ReLU (modules/model.py:114-135); pooling and the 1x1 convolution are both linear and
// the zero padding maps to zero, so W.pool^b(x) = pool^b(W.x): the engine pools the 19-channel product instead
// of the 114-channel input and applies bias / BatchNorm / ReLU here, after the last pool of the cascade.
struct PoolProblem {
    const float *in; float *out; const float *scale; const float *shift;      // scale == nullptr: plain pool
    int in_stride, out_stride;
    int identity;                                                             // 1: no pooling, only the scale / shift / ReLU epilogue
};
struct PoolBatch { PoolProblem p[kMaxBatch]; };

__global__ void __launch_bounds__(256)
avgpool3_batched_kernel(PoolBatch batch, int H, int W, int C, int relu)
{
    const PoolProblem pr = batch.p[blockIdx.y];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4n = C >> 2;
    if (i >= (long long)H * W * c4n) return;
    const int c4 = (int)(i % c4n);
    const int p = (int)(i / c4n), y = p / W, x = p - y * W;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    if (pr.identity) {
        o = __ldg(reinterpret_cast<const float4 *>(pr.in + (size_t)p * pr.in_stride) + c4);
    } else {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4 *>(pr.in + (size_t)(yy * W + xx) * pr.in_stride) + c4);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        o = make_float4(s.x / 9.0f, s.y / 9.0f, s.z / 9.0f, s.w / 9.0f);
    }
    if (pr.scale) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(pr.scale) + c4), b = __ldg(reinterpret_cast<const float4 *>(pr.shift) + c4);
        o.x = fmaf(o.x, a.x, b.x); o.y = fmaf(o.y, a.y, b.y); o.z = fmaf(o.z, a.z, b.z); o.w = fmaf(o.w, a.w, b.w);
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    }
    reinterpret_cast<float4 *>(pr.out + (size_t)p * pr.out_stride)[c4] = o;
}

// Per-channel sums over all pixels (global average pool numerator): block b sums its contiguous slice of
// pixels with float4 loads (thread = one float4 column of one of `lanes` interleaved pixels), folds the
// pixel lanes through shared memory in a fixed order and writes partial[b][C]; the tiny second stage runs
// inside gap_bias_kernel.  C % 4 == 0, C <= 1024 (wider inputs take the scalar kernel below).
__global__ void __launch_bounds__(256)
channel_sum4_kernel(const float *__restrict__ in, int in_stride, int npix, int C, float *__restrict__ partial)
{
    __shared__ float4 s_acc[256];
    const int c4n = C >> 2, lanes = 256 / c4n;
    const int c4 = threadIdx.x % c4n, lane = threadIdx.x / c4n;
    const int per = (npix + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(p0 + per, npix);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < lanes) {
        float4 b = a, c = a, d = a;                            // four independent chains: loads stay in flight
        int p = p0 + lane;
        for (; p + 3 * lanes < p1; p += 4 * lanes) {
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(in + (size_t)p * in_stride) + c4);
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(in + (size_t)(p + lanes) * in_stride) + c4);
            const float4 v2 = __ldg(reinterpret_cast<const float4 *>(in + (size_t)(p + 2 * lanes) * in_stride) + c4);
            const float4 v3 = __ldg(reinterpret_cast<const float4 *>(in + (size_t)(p + 3 * lanes) * in_stride) + c4);
            a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
            b.x += v1.x; b.y += v1.y; b.z += v1.z; b.w += v1.w;
            c.x += v2.x; c.y += v2.y; c.z += v2.z; c.w += v2.w;
            d.x += v3.x; d.y += v3.y; d.z += v3.z; d.w += v3.w;
        }
        for (; p < p1; p += lanes) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(in + (size_t)p * in_stride) + c4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        a.x = (a.x + b.x) + (c.x + d.x); a.y = (a.y + b.y) + (c.y + d.y);
        a.z = (a.z + b.z) + (c.z + d.z); a.w = (a.w + b.w) + (c.w + d.w);
    }
    s_acc[threadIdx.x] = a;
    __syncthreads();
    if (lane == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 v = s_acc[l * c4n + c4];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        reinterpret_cast<float4 *>(partial + (size_t)blockIdx.x * C)[c4] = a;
    }
}

__global__ void __launch_bounds__(256)
channel_sum_kernel(const float *__restrict__ in, int in_stride, int npix, int C, float *__restrict__ partial)
{
    const int c = blockIdx.y * blockDim.x + threadIdx.x;       // thread = channel, block row strides over pixels
    if (c >= C) return;
    float s = 0.0f;
    for (int p = blockIdx.x; p < npix; p += gridDim.x) s += __ldg(in + (size_t)p * in_stride + c);
    partial[(size_t)blockIdx.x * C + c] = s;
}

// Global-pool branch folded into a bias.  VortexPooling (modules/model.py:107-112): mean -> 1x1 conv ->
// (bilinear upsample of a 1x1 map = constant) -> BatchNorm; eASPP branch 5 (modules/adapnet.py:201-205):
// mean -> 1x1 conv -> ReLU -> upsample.  Either way the branch is a per-channel constant
// v[c] = act(g_scale[c] * (wg[c,:] . mean) + g_shift[c]) and the following 1x1 conv sees it as
// shift_out[co] = f_shift[co] + f_scale[co] * sum_c wf1[co,c] * v[c].   C <= 2048, Cg, Cout <= 256.
// One block: the partial sums are folded with 4 independent accumulators per channel, the two small
// matrix-vector products run one output per warp with lanes across the (contiguous) input index.
__global__ void __launch_bounds__(1024)
gap_bias_kernel(const float *__restrict__ partial, int nblocks, int npix, int C,
                const float *__restrict__ wg /* [Cg][C] */, const float *__restrict__ g_scale, const float *__restrict__ g_shift, int Cg,
                int v_relu, const float *__restrict__ wf1 /* [Cout][Cg] */, const float *__restrict__ f_scale,
                const float *__restrict__ f_shift, int Cout, float *__restrict__ shift_out)
{
    __shared__ float s_mean[2048];
    __shared__ float s_v[256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    __shared__ float s_red[1024];
    const int parts = C <= (int)blockDim.x ? (int)blockDim.x / C : 1;     // thread groups that share the partial blocks
    if (parts > 1) {
        const int c = t % C, part = t / C;
        if (part < parts) {
            float s0 = 0.f, s1 = 0.f;
            int b = part;
            for (; b + parts < nblocks; b += 2 * parts) {
                s0 += partial[(size_t)b * C + c];
                s1 += partial[(size_t)(b + parts) * C + c];
            }
            if (b < nblocks) s0 += partial[(size_t)b * C + c];
            s_red[part * C + c] = s0 + s1;
        }
        __syncthreads();
        if (t < C) {
            float a = 0.f;
            for (int q = 0; q < parts; ++q) a += s_red[q * C + t];     // fixed order: deterministic
            s_mean[t] = a / (float)npix;
        }
    } else {
        for (int c = t; c < C; c += blockDim.x) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int b = 0;
            for (; b + 3 < nblocks; b += 4) {
                s0 += partial[(size_t)b * C + c];
                s1 += partial[(size_t)(b + 1) * C + c];
                s2 += partial[(size_t)(b + 2) * C + c];
                s3 += partial[(size_t)(b + 3) * C + c];
            }
            for (; b < nblocks; ++b) s0 += partial[(size_t)b * C + c];
            s_mean[c] = ((s0 + s1) + (s2 + s3)) / (float)npix;
        }
    }
    __syncthreads();
    for (int o = warp; o < Cg; o += nwarps) {                  // one output per warp, four loads in flight per lane
        const float *w = wg + (size_t)o * C;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int c = lane;
        for (; c + 96 < C; c += 128) {
            a0 = fmaf(__ldg(w + c), s_mean[c], a0);
            a1 = fmaf(__ldg(w + c + 32), s_mean[c + 32], a1);
            a2 = fmaf(__ldg(w + c + 64), s_mean[c + 64], a2);
            a3 = fmaf(__ldg(w + c + 96), s_mean[c + 96], a3);
        }
        for (; c < C; c += 32) a0 = fmaf(__ldg(w + c), s_mean[c], a0);
        float a = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) {
            a = fmaf(a, g_scale[o], g_shift[o]);
            s_v[o] = v_relu ? fmaxf(a, 0.0f) : a;
        }
    }
    __syncthreads();
    for (int o = warp; o < Cout; o += nwarps) {
        float a = 0.0f;
        for (int c = lane; c < Cg; c += 32) a = fmaf(__ldg(wf1 + (size_t)o * Cg + c), s_v[c], a);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) shift_out[o] = fmaf(a, f_scale[o], f_shift[o]);
    }
}

// Wide pooled branches (eASPP: C = 2048, Cg = 256 -- 2 MB of weights): the first matrix-vector product on many blocks.
// Block b folds the partial sums into the mean vector (fixed order) and computes 8 outputs, one per warp:
// v[o] = act(g_scale[o] * (wg[o,:] . mean) + g_shift[o]).
__global__ void __launch_bounds__(256)
gap_v_kernel(const float *__restrict__ partial, int nblocks, int npix, int C, const float *__restrict__ wg,
             const float *__restrict__ g_scale, const float *__restrict__ g_shift, int Cg, int v_relu, float *__restrict__ v_out)
{
    __shared__ float s_mean[2048];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int c = t; c < C; c += blockDim.x) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int b = 0;
        for (; b + 3 < nblocks; b += 4) {
            s0 += partial[(size_t)b * C + c];
            s1 += partial[(size_t)(b + 1) * C + c];
            s2 += partial[(size_t)(b + 2) * C + c];
            s3 += partial[(size_t)(b + 3) * C + c];
        }
        for (; b < nblocks; ++b) s0 += partial[(size_t)b * C + c];
        s_mean[c] = ((s0 + s1) + (s2 + s3)) / (float)npix;
    }
    __syncthreads();
    const int o = blockIdx.x * 8 + warp;
    if (o >= Cg) return;
    const float *w = wg + (size_t)o * C;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int c = lane;
    for (; c + 96 < C; c += 128) {
        a0 = fmaf(__ldg(w + c), s_mean[c], a0);
        a1 = fmaf(__ldg(w + c + 32), s_mean[c + 32], a1);
        a2 = fmaf(__ldg(w + c + 64), s_mean[c + 64], a2);
        a3 = fmaf(__ldg(w + c + 96), s_mean[c + 96], a3);
    }
    for (; c < C; c += 32) a0 = fmaf(__ldg(w + c), s_mean[c], a0);
    float a = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) {
        a = fmaf(a, g_scale[o], g_shift[o]);
        v_out[o] = v_relu ? fmaxf(a, 0.0f) : a;
    }
}
// ... and the second one: shift_out[o] = f_shift[o] + f_scale[o] * (wf1[o,:] . v), one output per warp.
__global__ void __launch_bounds__(1024)
gap_shift_kernel(const float *__restrict__ v, int Cg, const float *__restrict__ wf1, const float *__restrict__ f_scale,
                 const float *__restrict__ f_shift, int Cout, float *__restrict__ shift_out)
{
    __shared__ float s_v[256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    if (t < Cg) s_v[t] = v[t];
    __syncthreads();
    for (int o = warp; o < Cout; o += nwarps) {
        float a = 0.0f;
        for (int c = lane; c < Cg; c += 32) a = fmaf(__ldg(wf1 + (size_t)o * Cg + c), s_v[c], a);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) shift_out[o] = fmaf(a, f_scale[o], f_shift[o]);
    }
}

// Decoder._join of AdapNet++ (modules/adapnet.py:305-315): gate[c] = relu(b[c] + W[c,:] . mean_pixels(x)) (24 channels from
// the 256 decoder features), out[p, c] = skip[p, c] * gate[c].  Every block folds the partial channel sums and computes
// the (tiny) gate itself, then scales its share of the pixels.  C <= 1024, Cg <= 32.
__global__ void __launch_bounds__(256)
skip_gate_kernel(const float *__restrict__ partial, int nblocks, int npix, int C, const float *__restrict__ w,
                 const float *__restrict__ b, int Cg, const float *__restrict__ skip, int skip_stride, float *__restrict__ out,
                 int out_stride)
{
    __shared__ float s_mean[1024];
    __shared__ float s_gate[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int c = t; c < C; c += blockDim.x) {
        float s0 = 0.f, s1 = 0.f;
        int k = 0;
        for (; k + 1 < nblocks; k += 2) {
            s0 += partial[(size_t)k * C + c];
            s1 += partial[(size_t)(k + 1) * C + c];
        }
        if (k < nblocks) s0 += partial[(size_t)k * C + c];
        s_mean[c] = (s0 + s1) / (float)npix;
    }
    __syncthreads();
    for (int o = warp; o < Cg; o += 8) {
        float a = 0.0f;
        for (int c = lane; c < C; c += 32) a = fmaf(__ldg(w + (size_t)o * C + c), s_mean[c], a);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) s_gate[o] = fmaxf(a + b[o], 0.0f);
    }
    __syncthreads();
    const long long n = (long long)npix * Cg;
    for (long long i = (long long)blockIdx.x * blockDim.x + t; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / Cg), c = (int)(i - (long long)p * Cg);
        out[(size_t)p * out_stride + c] = skip[(size_t)p * skip_stride + c] * s_gate[c];
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

// (C, npix) <-> (npix, stride) fp32 transposes through a 32x33 shared tile: the hand-over between
// the library-backed NCHW front of AdapNet++ and the pixel-major kernels.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float *__restrict__ in, int C, int npix, float *__restrict__ out, int out_stride, int out_coff)
{
    __shared__ float tile[32][33];
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && p < npix) ? in[(size_t)c * npix + p] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        if (p < npix && c < C) out[(size_t)p * out_stride + out_coff + c] = tile[threadIdx.x][r];
    }
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float *__restrict__ in, int in_stride, int in_coff, int C, int npix, float *__restrict__ out)
{
    __shared__ float tile[32][33];
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && p < npix) ? in[(size_t)p * in_stride + in_coff + c] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        if (c < C && p < npix) out[(size_t)c * npix + p] = tile[threadIdx.x][r];
    }
}

// Second half of a split-K convolution for other translation units (the tensor-core kernel): per problem
// out[p, coff+c] = out_mul * act(scale[c] * sum_s partial[s][p][c] + shift[c] (+ residual)), fixed summation order.
int launch_split_reduce(const SplitReduce *problems, int n, int npix, int cout, int cpad, int splits, int act, float slope,
                        float out_mul, cudaStream_t s)
{
    ConvBatch b;
    for (int i = 0; i < kMaxBatch; ++i) {
        const SplitReduce &q = problems[i < n ? i : 0];
        b.p[i] = ConvProblem{nullptr, nullptr, q.scale, q.shift, q.out, q.residual, q.partial, 0, q.out_stride, q.out_coff, 1,
                             q.res_stride};
    }
    bool vec = !(cout & 3) && !(cpad & 3);
    for (int i = 0; i < n && vec; ++i) {
        const SplitReduce &q = problems[i];
        vec = !(q.out_stride & 3) && !(q.out_coff & 3) && !((uintptr_t)q.out & 15) && !((uintptr_t)q.partial & 15) &&
              !((uintptr_t)q.scale & 15) && !((uintptr_t)q.shift & 15) && (!q.residual || (!(q.res_stride & 3) && !((uintptr_t)q.residual & 15)));
    }
    // programmatic dependent launch: the reduction's blocks are scheduled while the convolution (which released its
    // dependents at its first instruction) is still running and wait in griddepcontrol.wait for its completion
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t le;
    if (vec) {
        cfg.gridDim = dim3((unsigned)(((long long)npix * (cout >> 2) + 255) / 256), n);
        le = cudaLaunchKernelEx(&cfg, conv_reduce4_kernel, b, npix, cout, cpad, splits, act, slope, out_mul);
    } else {
        cfg.gridDim = dim3((unsigned)(((long long)npix * cout + 255) / 256), n);
        le = cudaLaunchKernelEx(&cfg, conv_reduce_kernel, b, npix, cout, cpad, splits, act, slope, out_mul);
    }
    if (le != cudaSuccess) { cudaGetLastError(); return (int)le; }
    return launched(1);
}

}  // namespace ojdf

using namespace ojdf;

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

extern "C" int ojdf_avgpool3_batched(const ojdf_pool_problem *problems_host, int n_problems, int H, int W, int C, int relu,
                                     void *stream)
{
    if (!problems_host || n_problems < 1 || n_problems > kMaxBatch || H < 1 || W < 1 || C < 4 || (C & 3)) return OJDF_ERR_BADARG;
    PoolBatch b;
    for (int i = 0; i < kMaxBatch; ++i) {
        const ojdf_pool_problem &q = problems_host[i < n_problems ? i : 0];
        if (i < n_problems && (!q.in_dev || !q.out_dev || (q.in_stride & 3) || (q.out_stride & 3) || q.in_stride < C ||
                               q.out_stride < C || (q.scale_dev && !q.shift_dev) || ((uintptr_t)q.in_dev & 15) ||
                               ((uintptr_t)q.out_dev & 15) || ((uintptr_t)q.scale_dev & 15) || ((uintptr_t)q.shift_dev & 15)))
            return OJDF_ERR_BADARG;
        b.p[i] = PoolProblem{q.in_dev, q.out_dev, q.scale_dev, q.shift_dev, q.in_stride, q.out_stride, q.identity ? 1 : 0};
    }
    const long long n = (long long)H * W * (C >> 2);
    avgpool3_batched_kernel<<<dim3((unsigned)((n + 255) / 256), n_problems), 256, 0, (cudaStream_t)stream>>>(b, H, W, C, relu);
    return launched(1);
}

extern "C" int ojdf_gap_bias(const float *in_dev, int in_stride, int npix, int C, const float *wg_dev,
                             const float *g_scale_dev, const float *g_shift_dev, int Cg, int v_relu, const float *wf1_dev,
                             const float *f_scale_dev, const float *f_shift_dev, int Cout, float *partial_dev,
                             int partial_blocks, float *shift_out_dev, void *stream)
{
    if (!in_dev || !wg_dev || !g_scale_dev || !g_shift_dev || !wf1_dev || !f_scale_dev || !f_shift_dev || !partial_dev ||
        !shift_out_dev || npix < 1 || C < 1 || C > 2048 || Cg < 1 || Cg > 256 || Cout < 1 || Cout > 256 || in_stride < C ||
        partial_blocks < 1)
        return OJDF_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    // few, fat partial blocks: the second stage (one block) walks them serially per channel
    int pb = npix / 128;
    if (pb < 8) pb = 8;
    if (pb > 148) pb = 148;
    // wide branches (eASPP: 2048 x 256 weights) spread the first matrix-vector product over Cg / 8 blocks; its result lives
    // in one more C-float block of the scratch
    const bool wide = (long long)C * Cg >= 65536 && partial_blocks >= 2 && Cg <= C;
    if (pb > partial_blocks - (wide ? 1 : 0)) pb = partial_blocks - (wide ? 1 : 0);
    if (pb > npix) pb = npix;
    if (!(C & 3) && C <= 1024 && !(in_stride & 3) && !((uintptr_t)in_dev & 15))
        channel_sum4_kernel<<<pb, 256, 0, s>>>(in_dev, in_stride, npix, C, partial_dev);
    else
        channel_sum_kernel<<<dim3(pb, (C + 255) / 256), 256, 0, s>>>(in_dev, in_stride, npix, C, partial_dev);
    if (wide) {
        float *v = partial_dev + (size_t)pb * C;                // the block of the scratch kept free above
        gap_v_kernel<<<(Cg + 7) / 8, 256, 0, s>>>(partial_dev, pb, npix, C, wg_dev, g_scale_dev, g_shift_dev, Cg, v_relu, v);
        gap_shift_kernel<<<1, 1024, 0, s>>>(v, Cg, wf1_dev, f_scale_dev, f_shift_dev, Cout, shift_out_dev);
        return launched(3);
    }
    gap_bias_kernel<<<1, 1024, 0, s>>>(partial_dev, pb, npix, C, wg_dev, g_scale_dev, g_shift_dev, Cg, v_relu, wf1_dev,
                                     f_scale_dev, f_shift_dev, Cout, shift_out_dev);
    return launched(2);
}

extern "C" int ojdf_adapnet_skip_join(const float *x_dev, int x_stride, int C, int npix, const float *w_dev, const float *b_dev,
                                      int Cg, const float *skip_dev, int skip_stride, float *out_dev, int out_stride,
                                      float *partial_dev, int partial_blocks, void *stream)
{
    if (!x_dev || !w_dev || !b_dev || !skip_dev || !out_dev || !partial_dev || npix < 1 || C < 4 || C > 1024 || (C & 3) ||
        (x_stride & 3) || x_stride < C || ((uintptr_t)x_dev & 15) || Cg < 1 || Cg > 32 || skip_stride < Cg || out_stride < Cg ||
        partial_blocks < 1)
        return OJDF_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    int pb = npix / 128;
    if (pb < 8) pb = 8;
    if (pb > 148) pb = 148;
    if (pb > partial_blocks) pb = partial_blocks;
    if (pb > npix) pb = npix;
    channel_sum4_kernel<<<pb, 256, 0, s>>>(x_dev, x_stride, npix, C, partial_dev);
    int blocks = (int)(((long long)npix * Cg + 2047) / 2048);
    if (blocks > 148) blocks = 148;
    skip_gate_kernel<<<blocks, 256, 0, s>>>(partial_dev, pb, npix, C, w_dev, b_dev, Cg, skip_dev, skip_stride, out_dev, out_stride);
    return launched(2);
}

extern "C" int ojdf_vortex_bias(const float *in_dev, int in_stride, int npix, int C, const float *wg_dev,
                                const float *g_scale_dev, const float *g_shift_dev, int Cg, const float *wf1_dev,
                                const float *f_scale_dev, const float *f_shift_dev, int Cout, float *partial_dev,
                                int partial_blocks, float *shift_out_dev, void *stream)
{
    return ojdf_gap_bias(in_dev, in_stride, npix, C, wg_dev, g_scale_dev, g_shift_dev, Cg, 0, wf1_dev, f_scale_dev,
                         f_shift_dev, Cout, partial_dev, partial_blocks, shift_out_dev, stream);
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

extern "C" int ojdf_nchw_to_nhwc(const float *in_dev, int C, int npix, float *out_dev, int out_stride, int out_coffset,
                                 void *stream)
{
    if (!in_dev || !out_dev || C < 1 || npix < 1 || out_stride < out_coffset + C || out_coffset < 0) return OJDF_ERR_BADARG;
    nchw_to_nhwc_kernel<<<dim3((npix + 31) / 32, (C + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        in_dev, C, npix, out_dev, out_stride, out_coffset);
    return launched(1);
}

extern "C" int ojdf_nhwc_to_nchw(const float *in_dev, int in_stride, int in_coffset, int C, int npix, float *out_dev,
                                 void *stream)
{
    if (!in_dev || !out_dev || C < 1 || npix < 1 || in_stride < in_coffset + C || in_coffset < 0) return OJDF_ERR_BADARG;
    nhwc_to_nchw_kernel<<<dim3((npix + 31) / 32, (C + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        in_dev, in_stride, in_coffset, C, npix, out_dev);
    return launched(1);
}
