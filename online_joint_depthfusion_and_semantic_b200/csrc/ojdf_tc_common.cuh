// Shared pieces of the tcgen05 convolution kernels (ojdf_conv_tc.cu: A operand in tensor memory; ojdf_conv_ss.cu: both
// operands in shared memory): launch parameters, PTX wrappers, ring bookkeeping, the epilogue arithmetic and the
// host helpers that build tensor maps.
#pragma once
#include <cuda.h>

#include <cstdio>

#include "ojdf_internal.h"

namespace ojdf {
namespace tc {

using ojdf::SplitReduce;

constexpr int kBK = 32;                  // fp32 channels per K chunk (128 bytes)
constexpr int kBW = 16, kBH = 8;         // M-tile = 8 rows x 16 columns = 128 pixels
constexpr int kThreads = 608;              // 19 warps, see the role list above
constexpr int kEpi0 = 10, kEpiThreads = 256;   // epilogue warps 10..17
constexpr int kMma2 = 18;                      // second MMA issuer
constexpr int kMaxBatch = 8;
constexpr int kAS = 6;                   // most A stages in TMEM (64 columns each: hi 32 | lo 32); the ring of an
                                         // issuer/splitter pair has prm.a_slots of them, starting at column prm.acol0
constexpr int kMaxSrc = 4, kMaxB = 8;
constexpr int kSlabBytes = 128 * 128;    // staging slab: 128 pixels x 32 channels

enum Act { kNone = 0, kRelu = 1, kLeaky = 2, kTanh = 3, kSigmoid = 4, kSigmoidMul = 5 };   // 5: sigmoid(v) * residual (SSMA gate)

struct Problem {
    const float *weights, *scale, *shift;
    float *out;
    const float *residual;
    int out_stride, out_coff, dil, res_stride;
};

struct Params {
    CUtensorMap in_map[kMaxBatch];
    CUtensorMap out_map[kMaxBatch];
    Problem p[kMaxBatch];
    int H, W, cin, cout, taps, act, npad, groups, nkc, tiles_x, tiles_y, nprob;
    int mt, nacc, hd, src_stages, b_stages, src_bytes, box_bytes, bwid, store_mode, dbg, sub_rows, nsub, a_slots, acol0, ksplit, cpad;
    int tap_mask[kMaxBatch];            // live taps of each problem (bit t = tap t), never 0
    float *partial[kMaxBatch];          // split-K scratch per problem: [ksplit][H*W][cpad] raw partial sums
    float slope, out_mul;
    // split-K finished inside the kernel: the CTA that delivers the LAST K slice of a tile sums the slices in slice order
    // (deterministic) and applies the layer's real epilogue; `counters` (zero between launches) counts slices per tile
    SplitReduce red[kMaxBatch];
    unsigned int *counters;
    int red_act, red_cout;
    float red_out_mul;
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
// Suspend-time hint (ns) of the parked mbarrier waits.
#ifndef OJDF_PARK_NS
#define OJDF_PARK_NS 0x989680u
#endif
// Wait for a phase of an mbarrier.  try_wait with a suspend-time hint parks the warp in hardware (no issue
// slots burnt by the many waiting roles); a pipeline that is wedged for ~2 s traps instead of hanging.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t hint = OJDF_PARK_NS)
{
    uint32_t done = 0;
    for (int spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(hint)
            : "memory");
        if (done) break;
        if (spins > 20000000) {
            printf("ojdf conv_tc: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x,
                   bar, parity);
            __trap();
        }
    }
}
// Busy poll with the non-blocking test_wait: for the hot A-ring hand-offs, where parking the warp costs more
// than the few issue slots the poll takes.
__device__ __forceinline__ void mbar_poll(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (int spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (spins > 100000000) {
            printf("ojdf conv_tc: mbarrier poll timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x,
                   bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Fetch a tensor map (kernel parameter) into the descriptor cache ahead of its first use.
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1), "r"(c2)
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
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T ; A: 128 lanes x 8 columns of tf32, B: K-major SWIZZLE_128B descriptor.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float activate(float v, int act, float slope)
{
    if (act == kRelu) return v > 0.0f ? v : 0.0f;
    if (act == kLeaky) return v > 0.0f ? v : v * slope;
    if (act == kTanh) return tanhf(v);
    if (act == kSigmoid) return 1.0f / (1.0f + expf(-v));
    return v;
}

// 16 accumulator columns -> scale/shift (+ residual) -> activation.  `ncols` = real channels left in this chunk.
template <int ACT>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[16], float (&o)[16], const float2 *ss, const float *res, int ncols,
                                          float slope, float out_mul)
{
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float2 p = ss[c];
        float r = fmaf(__uint_as_float(v[c]), p.x, p.y);
        if (ACT != kSigmoidMul && res && c < ncols) r += res[c];
        if (ACT == kRelu) r = fmaxf(r, 0.0f);
        if (ACT == kLeaky) r = r > 0.0f ? r : r * slope;
        if (ACT == kTanh) r = tanhf(r);
        if (ACT == kSigmoid) r = 1.0f / (1.0f + expf(-r));
        if (ACT == kSigmoidMul) r = (res && c < ncols) ? res[c] / (1.0f + expf(-r)) : 0.0f;
        o[c] = r * out_mul;
    }
}

// Position in a ring of `n` stages plus the parity of the current lap (no runtime division in the hot loops).
// A consumer waits full(idx) with `phase`; a producer waits empty(idx) with `phase ^ 1`, which falls through
// on the first lap because a fresh mbarrier reports its preceding phase as complete.
struct Ring {
    int idx, phase, n;
    __device__ __forceinline__ explicit Ring(int n_) : idx(0), phase(0), n(n_) {}
    __device__ __forceinline__ void next()
    {
        if (++idx == n) { idx = 0; phase ^= 1; }
    }
};

// ---- host helpers (defined in ojdf_conv_tc.cu)
// (C, W, H) view of a pixel-major fp32 buffer with `c` visible channels; box = (32, bw, bh), 128-byte swizzle.
int pixel_map(const float *ptr, int c, int stride, int H, int W, int bw, int bh, CUtensorMap *out, long long row_stride = 0);
int sm_count();

}  // namespace tc
}  // namespace ojdf

// Output-channel group layout shared by the packer and both kernels.
void ojdf_tc_layout(int cout, int npad_req, int *npad, int *groups);
// ojdf_conv_ss.cu: the shared-memory-operand kernel behind ojdf_conv_tc_batched; returns OJDF_SS_DECLINED (nothing
// launched) for the shapes the tensor-memory kernel handles better, unless flag 32768 forces it.
#define OJDF_SS_DECLINED (-1000)
int ojdf_conv_ss_launch(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H, int W, int taps, int act,
                        float slope, float out_mul, int npad_req, int flags, float *scratch_dev, size_t scratch_bytes, void *stream);
// ojdf_conv_wt.cu: output channels as the M dimension, the whole (small) image as N -- feature maps of <= 304 pixels
// with >= 128 output channels (AdapNet++ at 15x20); OJDF_SS_DECLINED for everything else.
int ojdf_conv_wt_launch(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H, int W, int taps, int act,
                        float slope, float out_mul, int npad_req, int flags, float *scratch_dev, size_t scratch_bytes, void *stream);
