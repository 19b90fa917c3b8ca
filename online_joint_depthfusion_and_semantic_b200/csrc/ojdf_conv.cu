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
// conv kernel (v4): one thread owns 4 pixels x 20 output channels in registers (80 FMAs per 6 LDS.128),
// one warp owns 128 consecutive pixels and runs its own barrier-free cp.async pipeline.
// A block covers a pixel tile of one convolution; up to 8 independent, equally shaped convolutions
// (the two FusionNet heads, the four VortexPooling branches) are batched along blockIdx.z and the
// host sizes tiles / block width (96..160 threads) so that the grid is a whole number of waves over
// the 148 SMs with several blocks co-resident per SM.  Inputs AND weights are streamed through
// shared memory in 8-channel chunks by a 3-stage cp.async pipeline (coalesced 16-byte copies, zero
// fill for the convolution padding): global memory is read once per tap with full-sector efficiency,
// weights arrive as [8 ci][20 co] slices read back as broadcast LDS.128.
#include <numeric>

#include "ojdf_internal.h"

namespace ojdf {

constexpr int kGroup = 20;          // output channels per thread (19 padded to 20 for FusionNet)
constexpr int kPix = 2;             // pixels per thread
constexpr int kMaxCT = 160;         // widest block (5 warps)
constexpr int kKC = 8;              // channels per pipeline chunk
constexpr int kRow4 = kKC / 4 + 1;  // float4 per staged pixel row (+1 pad: odd stride, conflict-free LDS.128)
constexpr int kWChunk4 = kKC * kGroup / 4;           // float4 of weights per chunk (8 ci x 20 co)
constexpr int kStages = 3;
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

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;                  // 0 -> the 16 bytes are zero filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// weights: [groups][taps][cin8][kGroup] fp32, zero padded (cin8 = cin rounded up to 8).
// dynamic smem: per warp, kStages x { 128 pixel rows of kRow4 float4, then kWChunk4 float4 of weights }.
// Every warp runs its own cp.async pipeline over its own 128 pixels (lane l owns pixels l, l+32, l+64,
// l+96 of the warp's slice), so the main loop has no block-wide barrier at all.
constexpr int kWarpPix = 32 * kPix;                            // 128
constexpr int kWarpStage4 = kWarpPix * kRow4 + kWChunk4;       // float4 per warp per stage

template <int TAPS>
__global__ void __launch_bounds__(kMaxCT, 3)
conv_tile_kernel(ConvBatch batch, int cin, int H, int W, int tile_px, int cout, int act, float slope, float out_mul,
                 int splits)
{
    extern __shared__ float4 smem4[];
    const int split = blockIdx.z % splits;
    const ConvProblem pr = batch.p[blockIdx.z / splits];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cin8 = (cin + kKC - 1) / kKC * kKC, nk = cin8 / kKC, cin4 = (cin + 3) >> 2;
    const int g = blockIdx.y;
    const float4 *wg = reinterpret_cast<const float4 *>(pr.weights) + (size_t)g * (TAPS * cin8 * (kGroup / 4));
    float4 *wsm = smem4 + (size_t)warp * (kStages * kWarpStage4);

    const int npix = H * W;
    const int tile0 = blockIdx.x * tile_px + warp * kWarpPix;
    const int tile_end = min(blockIdx.x * tile_px + tile_px, npix);
    // staging role: this lane copies float4 column (lane & 1) of the warp's pixels lane/2 + 16*i, i < 8.
    // Per staged pixel: element offset of the pixel row and a 9-bit mask of the taps that stay inside
    // the image (zero padding elsewhere).
    const int sc4 = lane & 1;
    int s_off[2 * kPix];
    unsigned s_ok[2 * kPix];
#pragma unroll
    for (int i = 0; i < 2 * kPix; ++i) {
        const int p = tile0 + (lane >> 1) + 16 * i;
        const int y = p / W, x = p - y * W;
        s_off[i] = p * pr.in_stride;
        unsigned m = 0;
        if (p < tile_end) {
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int yy = y + (TAPS == 1 ? 0 : (tap / 3 - 1) * pr.dil), xx = x + (TAPS == 1 ? 0 : (tap % 3 - 1) * pr.dil);
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) m |= 1u << tap;
            }
        }
        s_ok[i] = m;
    }
    const int nchunks_all = TAPS * nk;
    const int ch_begin = (int)((long long)nchunks_all * split / splits);       // this block's slice of the K loop
    const int nchunks = (int)((long long)nchunks_all * (split + 1) / splits) - ch_begin;
    int i_tap = ch_begin / nk, i_k8 = ch_begin - (ch_begin / nk) * nk;            // (tap, k8) of the next chunk to issue
    auto issue = [&](int ch) {
        if (ch < nchunks) {
            const int dy = TAPS == 1 ? 0 : (i_tap / 3 - 1) * pr.dil, dx = TAPS == 1 ? 0 : (i_tap % 3 - 1) * pr.dil;
            const int c4 = i_k8 * (kKC / 4) + sc4;
            const int shift = (dy * W + dx) * pr.in_stride + c4 * 4;
            const bool col_ok = c4 < cin4;
            float4 *dst = wsm + (size_t)(ch % kStages) * kWarpStage4;
#pragma unroll
            for (int i = 0; i < 2 * kPix; ++i) {
                const bool ok = col_ok && ((s_ok[i] >> i_tap) & 1u);
                cp_async16(dst + ((lane >> 1) + 16 * i) * kRow4 + sc4, pr.in + (ok ? s_off[i] + shift : 0), ok);
            }
            const float4 *wsrc = wg + (size_t)(i_tap * cin8 + i_k8 * kKC) * (kGroup / 4);
            cp_async16(dst + kWarpPix * kRow4 + lane, wsrc + lane, true);
            if (lane < kWChunk4 - 32) cp_async16(dst + kWarpPix * kRow4 + 32 + lane, wsrc + 32 + lane, true);
            if (++i_k8 == nk) { i_k8 = 0; ++i_tap; }
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);

    float acc[kPix][kGroup];
#pragma unroll
    for (int j = 0; j < kPix; ++j)
#pragma unroll
        for (int c = 0; c < kGroup; ++c) acc[j][c] = 0.0f;
    const int tail = cin & 3;                          // real channels in the last float4 (0 = all four)

    int k8 = ch_begin % nk;
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<1>();                            // this lane's copies of chunk ch have landed
        __syncwarp();                                  // ... and every lane's; stage (ch+2)%3 is free again
        issue(ch + 2);
        const float4 *xs = wsm + (size_t)(ch % kStages) * kWarpStage4;
        const float4 *wt = xs + kWarpPix * kRow4;
#pragma unroll
        for (int h4 = 0; h4 < kKC / 4; ++h4) {
            float4 xv[kPix];
#pragma unroll
            for (int j = 0; j < kPix; ++j) xv[j] = xs[(lane + 32 * j) * kRow4 + h4];
            if (tail && k8 * (kKC / 4) + h4 == cin4 - 1) {     // never let a neighbouring tensor's channels in
#pragma unroll
                for (int j = 0; j < kPix; ++j) {
                    if (tail < 2) xv[j].y = 0.f;
                    if (tail < 3) xv[j].z = 0.f;
                    xv[j].w = 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int q = 0; q < kGroup / 4; ++q) {
                    const float4 wv = wt[(h4 * 4 + k) * (kGroup / 4) + q];
#pragma unroll
                    for (int j = 0; j < kPix; ++j) {
                        const float x = k == 0 ? xv[j].x : (k == 1 ? xv[j].y : (k == 2 ? xv[j].z : xv[j].w));
                        acc[j][4 * q + 0] = fmaf(x, wv.x, acc[j][4 * q + 0]);
                        acc[j][4 * q + 1] = fmaf(x, wv.y, acc[j][4 * q + 1]);
                        acc[j][4 * q + 2] = fmaf(x, wv.z, acc[j][4 * q + 2]);
                        acc[j][4 * q + 3] = fmaf(x, wv.w, acc[j][4 * q + 3]);
                    }
                }
            }
        }
        if (++k8 == nk) k8 = 0;
    }
    cp_async_wait<0>();
    const int co0 = g * kGroup;
    if (splits > 1) {                                  // raw partial sums; conv_reduce_kernel finishes the layer
        const int cpad = gridDim.y * kGroup;
#pragma unroll
        for (int j = 0; j < kPix; ++j) {
            const int p = tile0 + lane + 32 * j;
            if (p >= tile_end) continue;
            float4 *o = reinterpret_cast<float4 *>(pr.partial + ((size_t)split * npix + p) * cpad + co0);
#pragma unroll
            for (int q = 0; q < kGroup / 4; ++q) o[q] = make_float4(acc[j][4 * q], acc[j][4 * q + 1], acc[j][4 * q + 2], acc[j][4 * q + 3]);
        }
        return;
    }
    float sc[kGroup], sh[kGroup];
#pragma unroll
    for (int c = 0; c < kGroup; ++c) {
        const bool live = co0 + c < cout;
        sc[c] = live ? __ldg(pr.scale + co0 + c) : 0.f;
        sh[c] = live ? __ldg(pr.shift + co0 + c) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
        const int p = tile0 + lane + 32 * j;
        if (p >= tile_end) continue;
        float *o = pr.out + (size_t)p * pr.out_stride + pr.out_coff + co0;
        const float *r = pr.residual ? pr.residual + (size_t)p * pr.res_stride + co0 : nullptr;
#pragma unroll
        for (int c = 0; c < kGroup; ++c)
            if (co0 + c < cout) {
                float v = fmaf(acc[j][c], sc[c], sh[c]);
                if (r) v += r[c];
                o[c] = activate(v, act, slope) * out_mul;
            }
    }
}

// Second half of a split-K convolution: sum the partials in split order, then the same epilogue.
__global__ void __launch_bounds__(256)
conv_reduce_kernel(ConvBatch batch, int npix, int cout, int cpad, int splits, int act, float slope, float out_mul)
{
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
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(pr.in + (size_t)(yy * W + xx) * pr.in_stride) + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    float4 o = make_float4(s.x / 9.0f, s.y / 9.0f, s.z / 9.0f, s.w / 9.0f);
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
    dim3 rgrid((unsigned)(((long long)npix * cout + 255) / 256), n);
    conv_reduce_kernel<<<rgrid, 256, 0, s>>>(b, npix, cout, cpad, splits, act, slope, out_mul);
    return launched(1);
}

}  // namespace ojdf

using namespace ojdf;

// Pick the tile count and block width: whole waves over the SMs, lanes as full as possible.
static void conv_geometry(int npix, int blocks_per_tile, int &tiles, int &tile_px, int &threads)
{
    const int sms = 148;
    double best = 1e30;
    const int tmin = (npix + kMaxCT * kPix - 1) / (kMaxCT * kPix);
    for (int t = tmin; t <= tmin * 4 + sms; ++t) {
        const int px = (npix + t - 1) / t;
        const int th = ((px + kPix - 1) / kPix + 31) / 32 * 32;
        if (th > kMaxCT) continue;
        const int real_tiles = (npix + px - 1) / px;
        const long long blocks = (long long)real_tiles * blocks_per_tile;
        const double cost = (double)((blocks + sms - 1) / sms) * th;   // rounds x time per round
        if (cost < best - 1e-9) { best = cost; tiles = real_tiles; tile_px = px; threads = th; }
    }
}

static int launch_conv(ConvBatch &batch, int n, int cin, int cout, int H, int W, int taps, int act, float slope,
                       float out_mul, float *scratch, size_t scratch_bytes, cudaStream_t s)
{
    const int npix = H * W;
    const int groups = (cout + kGroup - 1) / kGroup, cpad = groups * kGroup;
    const int nchunks = taps * ((cin + kKC - 1) / kKC);
    // split the K loop when the pixels alone cannot fill the machine (AdapNet's 15x20 feature maps)
    const long long warps = (long long)((npix + 32 * kPix - 1) / (32 * kPix)) * groups * n;
    int splits = 1;
    if (scratch && warps < 148 * 8) {
        splits = (int)((148 * 16 + warps - 1) / warps);
        if (splits > nchunks / 4) splits = nchunks / 4;          // keep >= 4 chunks per slice
        if (splits > 32) splits = 32;
        const size_t per_split = (size_t)n * npix * cpad * sizeof(float);
        if (per_split && (size_t)splits * per_split > scratch_bytes) splits = (int)(scratch_bytes / per_split);
        if (splits < 2) splits = 1;
    }
    if (splits > 1)
        for (int i = 0; i < n; ++i) batch.p[i].partial = scratch + (size_t)i * splits * npix * cpad;
    int tiles = 1, tile_px = npix, threads = 32;
    conv_geometry(npix, groups * n * splits, tiles, tile_px, threads);
    const size_t smem = (size_t)(threads / 32) * kStages * kWarpStage4 * sizeof(float4);
    dim3 grid(tiles, groups, n * splits);
    if (taps == 1) {
        static bool attr1 = false;
        if (!attr1) { cudaFuncSetAttribute(conv_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); attr1 = true; }
        conv_tile_kernel<1><<<grid, threads, smem, s>>>(batch, cin, H, W, tile_px, cout, act, slope, out_mul, splits);
    } else {
        static bool attr9 = false;
        if (!attr9) { cudaFuncSetAttribute(conv_tile_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); attr9 = true; }
        conv_tile_kernel<9><<<grid, threads, smem, s>>>(batch, cin, H, W, tile_px, cout, act, slope, out_mul, splits);
    }
    if (splits > 1) {
        dim3 rgrid((unsigned)(((long long)npix * cout + 255) / 256), n);
        conv_reduce_kernel<<<rgrid, 256, 0, s>>>(batch, npix, cout, cpad, splits, act, slope, out_mul);
        return launched(2);
    }
    return launched(1);
}

static bool problem_ok(const ojdf_conv_problem &q, int cin, int cout)
{
    return q.in_dev && q.weights_dev && q.scale_dev && q.shift_dev && q.out_dev && !(q.in_stride & 3) &&
           q.in_stride >= ((cin + 3) & ~3) && q.out_stride >= q.out_coffset + cout && q.out_coffset >= 0 && q.dilation >= 1 &&
           (!q.residual_dev || q.residual_stride >= cout) && q.in_step <= 1 && q.out_step <= 1;
}

extern "C" int ojdf_conv_nhwc_batched(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H,
                                      int W, int taps, int act, float slope, float out_mul, float *scratch_dev,
                                      size_t scratch_bytes, void *stream)
{
    if (!problems_host || n_problems < 1 || n_problems > kMaxBatch || cin < 1 || cout < 1 || H < 1 || W < 1 ||
        H > 32767 || W > 32767 || (taps != 1 && taps != 9) || act < 0 || act > 4)
        return OJDF_ERR_BADARG;
    ConvBatch b;
    for (int i = 0; i < kMaxBatch; ++i) {
        const ojdf_conv_problem &q = problems_host[i < n_problems ? i : 0];
        if (i < n_problems && !problem_ok(q, cin, cout)) return OJDF_ERR_BADARG;
        b.p[i] = ConvProblem{q.in_dev, q.weights_dev, q.scale_dev, q.shift_dev, q.out_dev, q.residual_dev, nullptr,
                             q.in_stride, q.out_stride, q.out_coffset, q.dilation, q.residual_stride};
    }
    return launch_conv(b, n_problems, cin, cout, H, W, taps, act, slope, out_mul, scratch_dev, scratch_bytes,
                       (cudaStream_t)stream);
}

extern "C" int ojdf_conv_nhwc(const float *in_dev, int in_stride, int cin, int H, int W, int taps, int dilation,
                              const float *weights_dev, const float *scale_dev, const float *shift_dev, int cout,
                              int act, float slope, float out_mul, float *out_dev, int out_stride, int out_coffset,
                              void *stream)
{
    ojdf_conv_problem q = {in_dev, weights_dev, scale_dev, shift_dev, out_dev, nullptr, in_stride, out_stride, out_coffset,
                           dilation, 0, 0, 0, 0, 0, 0};
    return ojdf_conv_nhwc_batched(&q, 1, cin, cout, H, W, taps, act, slope, out_mul, nullptr, 0, stream);
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
        b.p[i] = PoolProblem{q.in_dev, q.out_dev, q.scale_dev, q.shift_dev, q.in_stride, q.out_stride};
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
    if (pb > partial_blocks) pb = partial_blocks;
    if (pb > npix) pb = npix;
    if (!(C & 3) && C <= 1024 && !(in_stride & 3) && !((uintptr_t)in_dev & 15))
        channel_sum4_kernel<<<pb, 256, 0, s>>>(in_dev, in_stride, npix, C, partial_dev);
    else
        channel_sum_kernel<<<dim3(pb, (C + 255) / 256), 256, 0, s>>>(in_dev, in_stride, npix, C, partial_dev);
    gap_bias_kernel<<<1, 1024, 0, s>>>(partial_dev, pb, npix, C, wg_dev, g_scale_dev, g_shift_dev, Cg, v_relu, wf1_dev,
                                     f_scale_dev, f_shift_dev, Cout, shift_out_dev);
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
