// Tensor-core version of the tap GEMM of ojdf_conv.cu: tcgen05.mma (kind::tf32) with TMEM accumulators,
// TMA-staged operands and a split-precision (3xTF32) product so that the result stays within ~1e-6 of
// the fp32 convolution of the reference (modules/model.py:4-283, modules/adapnet.py:12-415).
//
//   out[p, coff+co] = out_mul * act(scale[co] * sum_tap sum_ci in[p + tap*dil, ci] * W[tap,ci,co] + shift[co] (+ residual))
//
// GEMM view, per CTA: M = 128 output pixels (a BH x BW rectangle of the image, BH*BW = 128), N = NPAD
// output channels (cout padded to 16, <= 128 per CTA), K = taps x cin in chunks of 32 channels.
//   * A (pixels x channels, K-major): the activation tensor is pixel-major (H, W, C) fp32; one TMA box
//     (32 channels, BW, BH) at the tap-shifted coordinate lands in shared memory as 128 rows of 128 bytes
//     in the SWIZZLE_128B pattern -- exactly the canonical K-major UMMA operand.  The convolution's zero
//     padding, the channel tail (cin not a multiple of 32) and partial tiles at the image border are all
//     TMA out-of-bounds zero fill: no masks, no im2col buffer.
//   * B (channels_out x channels_in, K-major): weights are packed on the host once
//     (ojdf_conv_tc_pack_weights) as ready-made swizzled shared-memory images [group][tap][kchunk][hi|lo]
//     [NPAD rows][32], fetched with one cp.async.bulk per chunk.
//   * 3xTF32: kind::tf32 reads fp32 containers and uses the top 19 bits.  x = hi + lo with hi = x with
//     the low 13 mantissa bits cleared and lo = x - hi (exact); D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo
//     drops only lo*lo (2^-22 relative).  The weight split is done by the packer; the activation split is
//     done in shared memory by the four epilogue warps while they wait (a second 16 KB tile per stage).
//   * Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
//     warps 2..5 = hi/lo splitter during the main loop, then epilogue (tcgen05.ld -> scale/shift/residual/
//     activation -> pixel-major stores).  mbarrier rings: full (TMA landed) -> split (lo tile written) ->
//     MMA -> empty (tcgen05.commit) ; acc (all MMAs retired) -> epilogue.
//   * One tile per CTA, 1-2 CTAs per SM (each owns its TMEM columns), up to 8 equally shaped problems
//     per launch along blockIdx.z (the two FusionNet heads, the four VortexPooling branches).
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "ojdf_internal.h"

namespace ojdf {
namespace tc {

constexpr int kBK = 32;                 // fp32 channels per K chunk: 128 bytes = one swizzle row
constexpr int kM = 128;                 // pixels per CTA tile (UMMA M)
constexpr int kATileBytes = kM * 128;   // 16 KB
constexpr int kThreads = 192;
constexpr int kMaxBatch = 8;
constexpr int kMaxStages = 6;
constexpr long long kSpinLimit = 4000000000LL;   // ~2 s of SM clocks: a wedged pipeline traps instead of hanging

enum Act { kNone = 0, kRelu = 1, kLeaky = 2, kTanh = 3, kSigmoid = 4 };

struct Problem {
    const float *weights;       // packed images of this problem
    const float *scale, *shift;
    float *out;
    const float *residual;
    int out_stride, out_coff, dil, res_stride;
};

struct Params {
    CUtensorMap tmap[kMaxBatch];
    Problem p[kMaxBatch];
    int H, W, cin, cout, taps, act, npad, nkc, stages, bw, bh, tiles_x, exact_hi;
    float slope, out_mul;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > kSpinLimit) {
            printf("ojdf conv_tc: mbarrier wait timed out (block %d,%d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, kind::tf32, issued by one thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);           // start address
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float activate(float v, int act, float slope)
{
    if (act == kRelu) return v > 0.0f ? v : 0.0f;
    if (act == kLeaky) return v > 0.0f ? v : v * slope;
    if (act == kTanh) return tanhf(v);
    if (act == kSigmoid) return 1.0f / (1.0f + expf(-v));
    return v;
}

// dynamic smem: [stages] x { A_hi 16 KB | A_lo 16 KB | B_hi npad*128 | B_lo npad*128 }, 1024-byte aligned,
// then the barriers.
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ Params prm)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bars[3 * kMaxStages + 1];
    __shared__ uint32_t s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const int S = prm.stages, npad = prm.npad;
    const uint32_t b_bytes = (uint32_t)npad * 128u;            // one of B_hi / B_lo
    const uint32_t stage_bytes = 2u * kATileBytes + 2u * b_bytes;
    const uint32_t bar0 = smem_u32(s_bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto split_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
    const uint32_t acc_bar = bar0 + 8u * (3 * kMaxStages);

    const Problem &pr = prm.p[blockIdx.z];
    const int tile_y = blockIdx.x / prm.tiles_x, tile_x = blockIdx.x - tile_y * prm.tiles_x;
    const int x0 = tile_x * prm.bw, y0 = tile_y * prm.bh;
    const int group = blockIdx.y;
    const int n_iter = prm.taps * prm.nkc;

    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)npad) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(split_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const CUtensorMap *map = &prm.tmap[blockIdx.z];
            const uint8_t *wbase = reinterpret_cast<const uint8_t *>(pr.weights) + (size_t)group * n_iter * 2u * b_bytes;
            int tap = 0, kc = 0;
            for (int it = 0; it < n_iter; ++it) {
                const int s = it % S;
                if (it >= S) mbar_wait(empty_bar(s), ((it / S) - 1) & 1);
                const int dy = prm.taps == 1 ? 0 : (tap / 3 - 1) * pr.dil, dx = prm.taps == 1 ? 0 : (tap % 3 - 1) * pr.dil;
                const uint32_t sa = base + (uint32_t)s * stage_bytes;
                mbar_expect_tx(full_bar(s), kATileBytes + 2u * b_bytes);
                tma_load_3d(sa, map, full_bar(s), kc * kBK, x0 + dx, y0 + dy);
                bulk_load(sa + 2u * kATileBytes, wbase + (size_t)it * 2u * b_bytes, 2u * b_bytes, full_bar(s));
                if (++kc == prm.nkc) { kc = 0; ++tap; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            // instruction descriptor: D f32, A/B tf32, both K-major, N = npad, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
            for (int it = 0; it < n_iter; ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                mbar_wait(full_bar(s), ph);
                mbar_wait(split_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = base + (uint32_t)s * stage_bytes;
                const uint64_t a_hi = smem_desc(sa), a_lo = smem_desc(sa + kATileBytes);
                const uint64_t b_hi = smem_desc(sa + 2u * kATileBytes), b_lo = smem_desc(sa + 2u * kATileBytes + b_bytes);
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {
                    const uint64_t ko = (uint64_t)(k * 32 >> 4);       // +32 bytes along K inside the swizzle atom
                    umma_tf32(tmem, a_lo + ko, b_hi + ko, idesc, (it | k) != 0);
                    umma_tf32(tmem, a_hi + ko, b_lo + ko, idesc, 1);
                    umma_tf32(tmem, a_hi + ko, b_hi + ko, idesc, 1);
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(acc_bar);
        }
    } else {
        // ------------------------------------------------------------ splitter, then epilogue
        const int t = threadIdx.x - 64;                        // 0..127
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % S;
            mbar_wait(full_bar(s), (it / S) & 1);
            uint4 *hi = reinterpret_cast<uint4 *>(smem + (size_t)s * stage_bytes);
            uint4 *lo = reinterpret_cast<uint4 *>(smem + (size_t)s * stage_bytes + kATileBytes);
#pragma unroll
            for (int j = 0; j < kATileBytes / 16 / 128; ++j) {
                const int i = t + 128 * j;
                const uint4 x = hi[i];
                uint4 h, l;
                h.x = x.x & 0xFFFFE000u; h.y = x.y & 0xFFFFE000u; h.z = x.z & 0xFFFFE000u; h.w = x.w & 0xFFFFE000u;
                l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x));
                l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y));
                l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z));
                l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w));
                lo[i] = l;
                if (prm.exact_hi) hi[i] = h;
            }
            fence_proxy_async();
            mbar_arrive(split_bar(s));
        }
        // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 ; lane = accumulator row = pixel of the tile
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int ty = m / prm.bw, tx = m - ty * prm.bw;
        const int y = y0 + ty, x = x0 + tx;
        const bool live = ty < prm.bh && y < prm.H && x < prm.W;
        const size_t pix = (size_t)y * prm.W + x;
        const int co_base = group * npad;
        float *orow = pr.out + pix * pr.out_stride + pr.out_coff + co_base;
        const float *rrow = pr.residual ? pr.residual + pix * pr.res_stride + co_base : nullptr;
        for (int n0 = 0; n0 < npad; n0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, v);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int co = co_base + n0 + c;
                    if (co < prm.cout) {
                        float r = fmaf(__uint_as_float(v[c]), __ldg(pr.scale + co), __ldg(pr.shift + co));
                        if (rrow) r += rrow[n0 + c];
                        orow[n0 + c] = activate(r, prm.act, prm.slope) * prm.out_mul;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

struct MapKey {
    const void *ptr;
    int cin, stride, H, W, bw, bh;
    bool operator==(const MapKey &o) const
    {
        return ptr == o.ptr && cin == o.cin && stride == o.stride && H == o.H && W == o.W && bw == o.bw && bh == o.bh;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const
    {
        size_t h = (size_t)k.ptr;
        const int v[6] = {k.cin, k.stride, k.H, k.W, k.bw, k.bh};
        for (int i = 0; i < 6; ++i) h = h * 1000003u ^ (size_t)v[i];
        return h;
    }
};

// (C, W, H) view of a pixel-major fp32 activation buffer; box = (32 channels, bw, bh), 128-byte swizzle,
// out-of-bounds elements (padding, channel tail, image border) read as zero.
static int activation_map(const float *in, int cin, int stride, int H, int W, int bw, int bh, CUtensorMap *out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{in, cin, stride, H, W, bw, bh};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[2] = {(cuuint64_t)stride * 4u, (cuuint64_t)stride * 4u * (cuuint64_t)W};
    const cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(in), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OJDF_ERR_BADARG;
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return 0;
}

static void layout(int cout, int *npad, int *groups)
{
    const int g = (cout + 127) / 128;
    const int per = (cout + g - 1) / g;
    *groups = g;
    *npad = (per + 15) / 16 * 16;
}

}  // namespace tc
}  // namespace ojdf

using namespace ojdf;

extern "C" int ojdf_conv_tc_layout(int cout, int *npad, int *groups)
{
    if (cout < 1 || !npad || !groups) return OJDF_ERR_BADARG;
    tc::layout(cout, npad, groups);
    return 0;
}

extern "C" size_t ojdf_conv_tc_weight_floats(int cin, int cout, int taps)
{
    if (cin < 1 || cout < 1 || taps < 1) return 0;
    int npad, groups;
    tc::layout(cout, &npad, &groups);
    const int nkc = (cin + tc::kBK - 1) / tc::kBK;
    return (size_t)groups * taps * nkc * 2 * npad * tc::kBK;
}

extern "C" int ojdf_conv_tc_pack_weights(const float *w_host, int cin, int cout, int taps, float *packed_host)
{
    if (!w_host || !packed_host || cin < 1 || cout < 1 || taps < 1) return OJDF_ERR_BADARG;
    int npad, groups;
    tc::layout(cout, &npad, &groups);
    const int nkc = (cin + tc::kBK - 1) / tc::kBK;
    memset(packed_host, 0, ojdf_conv_tc_weight_floats(cin, cout, taps) * sizeof(float));
    for (int g = 0; g < groups; ++g)
        for (int tap = 0; tap < taps; ++tap)
            for (int kc = 0; kc < nkc; ++kc) {
                float *img = packed_host + ((((size_t)g * taps + tap) * nkc + kc) * 2) * npad * tc::kBK;
                for (int r = 0; r < npad; ++r) {
                    const int co = g * npad + r;
                    if (co >= cout) continue;
                    for (int k = 0; k < tc::kBK; ++k) {
                        const int ci = kc * tc::kBK + k;
                        if (ci >= cin) continue;
                        const float w = w_host[((size_t)co * cin + ci) * taps + tap];      // (cout, cin, kh*kw)
                        uint32_t bits;
                        memcpy(&bits, &w, 4);
                        bits &= 0xFFFFE000u;
                        float hi;
                        memcpy(&hi, &bits, 4);
                        const float lo = w - hi;
                        // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
                        const int chunk = (k >> 2) ^ (r & 7);
                        const size_t off = (size_t)r * tc::kBK + chunk * 4 + (k & 3);
                        img[off] = hi;
                        img[(size_t)npad * tc::kBK + off] = lo;
                    }
                }
            }
    return 0;
}

extern "C" int ojdf_conv_tc_batched(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H,
                                    int W, int taps, int act, float slope, float out_mul, int flags, void *stream)
{
    if (!problems_host || n_problems < 1 || n_problems > tc::kMaxBatch || cin < 1 || cout < 1 || H < 1 || W < 1 ||
        H > 32767 || W > 32767 || (taps != 1 && taps != 9) || act < 0 || act > 4)
        return OJDF_ERR_BADARG;
    tc::Params prm;
    memset(&prm, 0, sizeof(prm));
    int npad, groups;
    tc::layout(cout, &npad, &groups);
    // tile rectangle: 128 pixels, as wide as the image allows (16 columns by default)
    int bw = 16, bh = 8;
    if (W <= 8) { bw = 8; bh = 16; }
    prm.H = H; prm.W = W; prm.cin = cin; prm.cout = cout; prm.taps = taps; prm.act = act; prm.npad = npad;
    prm.nkc = (cin + tc::kBK - 1) / tc::kBK;
    prm.bw = bw; prm.bh = bh;
    prm.tiles_x = (W + bw - 1) / bw;
    prm.exact_hi = (flags & 1) ? 0 : 1;
    prm.slope = slope; prm.out_mul = out_mul;
    const int tiles_y = (H + bh - 1) / bh;
    for (int i = 0; i < n_problems; ++i) {
        const ojdf_conv_problem &q = problems_host[i];
        if (!q.in_dev || !q.weights_dev || !q.scale_dev || !q.shift_dev || !q.out_dev || (q.in_stride & 3) || q.in_stride < cin ||
            ((uintptr_t)q.in_dev & 15) || ((uintptr_t)q.weights_dev & 15) || q.out_stride < q.out_coffset + cout ||
            q.out_coffset < 0 || q.dilation < 1 || (q.residual_dev && q.residual_stride < cout))
            return OJDF_ERR_BADARG;
        const int r = tc::activation_map(q.in_dev, cin, q.in_stride, H, W, bw, bh, &prm.tmap[i]);
        if (r) return r;
        prm.p[i] = tc::Problem{q.weights_dev, q.scale_dev, q.shift_dev, q.out_dev, q.residual_dev,
                               q.out_stride, q.out_coffset, q.dilation, q.residual_stride};
    }
    const size_t stage_bytes = 2 * (size_t)tc::kATileBytes + 2 * (size_t)npad * 128;
    const int n_iter = taps * prm.nkc;
    // two CTAs per SM when two stages each fit (the epilogue of one overlaps the main loop of the other)
    int stages = (int)((112 * 1024 - 1024) / stage_bytes);
    if (stages < 2) stages = (int)((224 * 1024 - 1024) / stage_bytes);
    if (stages > tc::kMaxStages) stages = tc::kMaxStages;
    if (stages > n_iter) stages = n_iter;
    if (stages < 1) return OJDF_ERR_BADARG;
    prm.stages = stages;
    const size_t smem = stages * stage_bytes + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(tc::conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
        attr = true;
    }
    dim3 grid(prm.tiles_x * tiles_y, groups, n_problems);
    tc::conv_tc_kernel<<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(prm);
    return launched(1);
}
