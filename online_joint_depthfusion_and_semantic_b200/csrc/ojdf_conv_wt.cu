// Tap GEMM for SMALL feature maps (AdapNet++ at 15x20: layer3[1:], layer4, eASPP, modules/adapnet.py:103-149,152-216)
// with the operand roles swapped -- the third kernel behind ojdf_conv_tc_batched:
//
//   out[p, coff+co] = out_mul * act(scale[co] * sum_tap sum_ci in[p + tap*dil, ci] * W[tap,ci,co] + shift[co] (+ residual))
//
// On a 300-pixel map the pixel-major kernels (ojdf_conv_tc.cu / ojdf_conv_ss.cu) fill 59 % of their 128-pixel M-tiles and
// issue N = 128 MMAs behind a long per-CTA K loop.  Here the OUTPUT CHANNELS are the M dimension (one CTA = 128 of them,
// the packed [W_hi | W_lo] images of ojdf_conv_tc_pack_weights are the A operand as they are) and the WHOLE IMAGE is the
// N dimension (<= 304 pixels = one or two MMAs of N <= 256 per K step): a TMA box of the full (W, H) map at the tap's
// offset (zero fill = convolution padding) is the K-major B operand, its tf32 `lo` half is computed once per box, and
// every MMA runs at the tensor pipe's full rate (N / 2 cycles at N >= 128).  fp32 accumulators: 128 lanes (channels) x
// <= 304 columns (pixels) of tensor memory; the epilogue thread of a lane owns one output channel, a warp stores 32
// consecutive channels of one pixel (128 bytes) per instruction.  The K loop (taps x 32-channel chunks) is split over
// CTAs; raw partial sums go to the caller's scratch and conv_reduce_kernel finishes the layer in a fixed order.
// Warps: 0 = TMA producer (activation box + weight stage per K step), 1 = MMA issuer (owns the TMEM allocation),
// 2..9 = epilogue, 10..17 = lo pass.  Same numerics as the other two kernels: x = hi + lo, hi*hi + hi*lo + lo*hi.
#include <cstdlib>
#include <cstring>

#include "ojdf_tc_common.cuh"

namespace ojdf {
namespace wt {

using namespace ojdf::tc;

constexpr int kWtThreads = 32 * 18;
constexpr int kEpiWarp0 = 2, kLoWarp0 = 10, kLoThreads = 256;
constexpr int kMaxStages = 4;
constexpr int kMaxPix = 304;

struct WtParams {
    CUtensorMap in_map[kMaxBatch];
    Problem p[kMaxBatch];
    int H, W, npix, npad_n, n1, n2, cin, cout, taps, act, groups, nkc, nprob, ksplit, cpad, stages, box_bytes, b_slot, fast;
    int th, ntiles;                                            // pixel tile: th whole image rows (th * W <= 304), ntiles of them
    int npad, a_bytes;                                         // output channels per CTA (64 or 128) and bytes of its [W_hi | W_lo] stage
    int cluster;                                               // 1: the K slices of an item are the CTAs of one cluster; they reduce through DSMEM
    int ntaps[kMaxBatch];
    int tap_list[kMaxBatch][9];
    float slope, out_mul;
};

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// fp32 load from the shared memory of CTA `rank` of this cluster, at the address `local` has in this CTA
__device__ __forceinline__ float ld_dsmem(uint32_t local, uint32_t rank)
{
    uint32_t ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

struct Item { int z, g, s0, s1, ks, tile; };
__device__ __forceinline__ Item decode(const WtParams &prm, int it)
{
    Item i;
    i.ks = it % prm.ksplit;
    int r = it / prm.ksplit;
    i.tile = r % prm.ntiles;
    r /= prm.ntiles;
    i.g = r % prm.groups;
    i.z = r / prm.groups;
    const int total = prm.ntaps[i.z] * prm.nkc;                 // K steps of this problem: live taps x 32-channel chunks
    i.s0 = total * i.ks / prm.ksplit;
    i.s1 = total * (i.ks + 1) / prm.ksplit;
    return i;
}

__global__ void __launch_bounds__(kWtThreads, 1) conv_wt_kernel(const __grid_constant__ WtParams prm)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bars[3 * kMaxStages + 2];
    __shared__ uint32_t s_tmem;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const int NS = prm.stages;
    const uint32_t kABytes = (uint32_t)prm.a_bytes;              // [W_hi | W_lo] of npad output channels x 32 input channels
    const uint32_t slot_bytes = 2u * (uint32_t)prm.b_slot + kABytes;        // [box | lo copy | weights]
    const uint32_t bar0 = smem_u32(s_bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto lo_full = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
    const uint32_t acc_full = bar0 + 8u * (3 * kMaxStages), acc_empty = acc_full + 8u;

    const int total = prm.nprob * prm.groups * prm.ntiles * prm.ksplit;
    const int begin = (int)((long long)total * blockIdx.x / gridDim.x);
    const int end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); mbar_init(lo_full(s), kLoThreads); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kEpiThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 64)
        for (int i = 0; i < prm.nprob; ++i) prefetch_map(&prm.in_map[i]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    asm volatile("griddepcontrol.wait;" ::: "memory");         // from here on this grid reads what the previous kernels wrote

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer: one activation box + one weight stage per K step
        Ring r(NS);
        for (int it = begin; it < end; ++it) {
            const Item I = decode(prm, it);
            const Problem &pr = prm.p[I.z];
            const uint8_t *wbase = reinterpret_cast<const uint8_t *>(pr.weights) + (size_t)I.g * prm.taps * prm.nkc * kABytes;
            for (int s = I.s0; s < I.s1; ++s, r.next()) {
                const int ti = s / prm.nkc, kc = s - ti * prm.nkc;
                const int tap = prm.tap_list[I.z][ti];
                const int dx = prm.taps == 9 ? (tap % 3 - 1) * pr.dil : 0, dy = prm.taps == 9 ? (tap / 3 - 1) * pr.dil : 0;
                mbar_wait(empty(r.idx), r.phase ^ 1);
                if (elect_one()) {
                    const uint32_t dst = base + (uint32_t)r.idx * slot_bytes;
                    mbar_expect_tx(full(r.idx), (uint32_t)prm.box_bytes + kABytes);
                    tma_load_3d(dst, &prm.in_map[I.z], full(r.idx), kc * kBK, dx, I.tile * prm.th + dy);
                    bulk_load(dst + 2u * (uint32_t)prm.b_slot, wbase + (size_t)(tap * prm.nkc + kc) * kABytes, kABytes, full(r.idx));
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
        const uint32_t id1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(prm.n1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(prm.n2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        Ring r(NS);
        uint32_t acc_phase = 0;
        for (int it = begin; it < end; ++it) {
            const Item I = decode(prm, it);
            mbar_wait(acc_empty, acc_phase ^ 1u);               // the epilogue of the previous item has read the accumulator
            tc_fence_after();
            acc_phase ^= 1u;
            for (int s = I.s0; s < I.s1; ++s, r.next()) {
                const int kc = s % prm.nkc;
                int ksteps = (prm.cin - kc * kBK + 7) >> 3;
                if (ksteps > kBK / 8) ksteps = kBK / 8;
                mbar_wait(lo_full(r.idx), r.phase);
                tc_fence_after();
                const uint32_t sb = base + (uint32_t)r.idx * slot_bytes;
                const uint64_t x_hi = smem_desc(sb), x_lo = smem_desc(sb + (uint32_t)prm.b_slot);
                const uint64_t w_hi = smem_desc(sb + 2u * (uint32_t)prm.b_slot), w_lo = smem_desc(sb + 2u * (uint32_t)prm.b_slot + kABytes / 2);
                const uint64_t half2 = (uint64_t)((prm.n1 * 128) >> 4);       // second pixel half: n1 rows further down the box
                if (elect_one()) {
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);   // +32 bytes along K inside the swizzle atom
                        const uint32_t first = (s == I.s0 && k == 0) ? 0u : 1u;
                        if (prm.fast) {
                            umma_tf32_ss(tmem, w_hi + ko, x_hi + ko, id1, first);
                            if (prm.n2) umma_tf32_ss(tmem + (uint32_t)prm.n1, w_hi + ko, x_hi + half2 + ko, id2, first);
                        } else {
                            umma_tf32_ss(tmem, w_hi + ko, x_lo + ko, id1, first);
                            umma_tf32_ss(tmem, w_lo + ko, x_hi + ko, id1, 1u);
                            umma_tf32_ss(tmem, w_hi + ko, x_hi + ko, id1, 1u);
                            if (prm.n2) {
                                umma_tf32_ss(tmem + (uint32_t)prm.n1, w_hi + ko, x_lo + half2 + ko, id2, first);
                                umma_tf32_ss(tmem + (uint32_t)prm.n1, w_lo + ko, x_hi + half2 + ko, id2, 1u);
                                umma_tf32_ss(tmem + (uint32_t)prm.n1, w_hi + ko, x_hi + half2 + ko, id2, 1u);
                            }
                        }
                    }
                    umma_commit(empty(r.idx));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(acc_full);
            __syncwarp();
        }
    } else if (warp >= kLoWarp0) {
        // ------------------------------------------------------------ lo pass: box -> lo = x - tf32(x), same layout
        const int lt = threadIdx.x - kLoWarp0 * 32;
        const int chunks = prm.box_bytes >> 4;
        Ring r(NS);
        for (int it = begin; it < end; ++it) {
            const Item I = decode(prm, it);
            for (int s = I.s0; s < I.s1; ++s, r.next()) {
                mbar_wait(full(r.idx), r.phase);
                const uint4 *src = reinterpret_cast<const uint4 *>(smem + (size_t)r.idx * slot_bytes);
                uint4 *dst = reinterpret_cast<uint4 *>(smem + (size_t)r.idx * slot_bytes + prm.b_slot);
                if (!prm.fast)
                    for (int i = lt; i < chunks; i += kLoThreads) {
                        const uint4 x = src[i];
                        uint4 l;
                        l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(x.x & 0xFFFFE000u));
                        l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(x.y & 0xFFFFE000u));
                        l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(x.z & 0xFFFFE000u));
                        l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(x.w & 0xFFFFE000u));
                        dst[i] = l;
                    }
                fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core's reads
                mbar_arrive(lo_full(r.idx));
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: lane = output channel, columns = pixels; warp pair
        // (q, half) owns TMEM lanes 32q..32q+31 and the 16-pixel chunks with index = half (mod 2)
        const int q = warp & 3, half = (warp - kEpiWarp0) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t acc_phase = 0;
        for (int it = begin; it < end; ++it) {
            const Item I = decode(prm, it);
            const Problem &pr = prm.p[I.z];
            const int co = I.g * prm.npad + q * 32 + lane;
            const bool live = q * 32 + lane < prm.npad && co < prm.cout;   // narrow groups: the MMA's upper rows read past W and are dropped
            float sc = 1.0f, sh = 0.0f;
            if (prm.ksplit == 1 && live && !prm.cluster) { sc = __ldg(pr.scale + co); sh = __ldg(pr.shift + co); }
            mbar_wait(acc_full, acc_phase);
            tc_fence_after();
            acc_phase ^= 1u;
            const int p0 = I.tile * prm.th * prm.W;             // first pixel of this tile
            int pend = p0 + prm.th * prm.W;                       // one past its last real pixel
            if (pend > prm.npix) pend = prm.npix;
            for (int n0 = half * 16; n0 < prm.npad_n; n0 += 32) {
                uint32_t v[16];
                tmem_ld16(lane_base + (uint32_t)n0, v);
                tmem_ld_wait();
                if (prm.cluster) {                               // this K slice's partial sums -> shared memory [pixel][128 channels]
                    float *part = reinterpret_cast<float *>(smem) + (size_t)n0 * 128 + q * 32 + lane;
#pragma unroll
                    for (int c = 0; c < 16; ++c) part[c * 128] = __uint_as_float(v[c]);
                    continue;
                }
                if (!live) continue;
                if (prm.ksplit > 1) {                           // raw partial sums: [ks][pixel][cpad]
                    float *o = pr.out + ((size_t)I.ks * prm.npix + p0 + n0) * prm.cpad + co;
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (p0 + n0 + c < pend) o[(size_t)c * prm.cpad] = __uint_as_float(v[c]);
                } else {
                    // the residual values of the chunk are read first: loads interleaved with the stores below would be
                    // serialised by the compiler (the output may alias them)
                    float res[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int p = p0 + n0 + c;
                        res[c] = (pr.residual && p < pend) ? __ldg(pr.residual + (size_t)p * pr.res_stride + co) : 0.0f;
                    }
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int p = p0 + n0 + c;
                        if (p >= pend) continue;
                        float r = fmaf(__uint_as_float(v[c]), sc, sh);
                        if (prm.act == kSigmoidMul) r = res[c] / (1.0f + expf(-r));
                        else r = activate(r + res[c], prm.act, prm.slope);
                        pr.out[(size_t)p * pr.out_stride + pr.out_coff + co] = r * prm.out_mul;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty);
        }
    }
    if (prm.cluster) {
        // Split-K inside a cluster: every CTA of the cluster holds the partial sums of its K slice in shared memory; CTA r
        // sums pixels [r * per, (r + 1) * per) over all slices in slice order (deterministic) through distributed shared
        // memory, applies the layer's epilogue and stores -- no scratch round trip, no reduction launch.
        cluster_sync();
        const Item I = decode(prm, begin);                       // grid == items: this CTA's only item
        const Problem &pr = prm.p[I.z];
        const int ks = prm.ksplit, per = (prm.npad_n + ks - 1) / ks;
        const int nb = I.ks * per, ne = min(nb + per, prm.npad_n);
        const int ch = threadIdx.x & 127, sub = threadIdx.x >> 7;
        const int co = I.g * prm.npad + ch;
        const int p0 = I.tile * prm.th * prm.W;
        int pend = p0 + prm.th * prm.W;
        if (pend > prm.npix) pend = prm.npix;
        if (threadIdx.x < 512 && ch < prm.npad && co < prm.cout) {
            const float sc = __ldg(pr.scale + co), sh = __ldg(pr.shift + co);
            for (int n = nb + sub; n < ne; n += 4) {
                const int p = p0 + n;
                if (p >= pend) break;
                const uint32_t la = base + (uint32_t)(n * 128 + ch) * 4u;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = j < ks ? ld_dsmem(la, (uint32_t)j) : 0.0f;
                const float res = pr.residual ? __ldg(pr.residual + (size_t)p * pr.res_stride + co) : 0.0f;
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) a += v[j];
                float r = fmaf(a, sc, sh);
                if (prm.act == kSigmoidMul) r = res / (1.0f + expf(-r));
                else r = activate(r + res, prm.act, prm.slope);
                pr.out[(size_t)p * pr.out_stride + pr.out_coff + co] = r * prm.out_mul;
            }
        }
        cluster_sync();                                          // nobody leaves while its shared memory is still being read
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace wt
}  // namespace ojdf

using namespace ojdf;

// Returns OJDF_SS_DECLINED (nothing launched) for shapes this kernel does not cover: output-channel groups narrower than
// 128, images wider than 256 pixels, strided writes.  Larger maps are cut into pixel tiles of whole image rows (<= 304
// pixels each).  flag 131072 forces the other kernels, flag 262144 keeps maps of more than one pixel tile on them.
int ojdf_conv_wt_launch(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H, int W, int taps, int act,
                        float slope, float out_mul, int npad_req, int flags, float *scratch_dev, size_t scratch_bytes, void *stream)
{
    static const bool env_off = [] { const char *e = getenv("OJDF_CONV_WT"); return e && !strcmp(e, "0"); }();
    if (env_off || (flags & 131072)) return OJDF_SS_DECLINED;
    int npad, groups;
    ojdf_tc_layout(cout, npad_req, &npad, &groups);
    const int npix = H * W;
    // 64-channel groups run as M = 128 MMAs whose upper 64 rows read whatever follows the weight image (finite or not: a row of
    // D depends on its own row of A only) and are never stored
    if (W > 256 || (npad != 128 && npad != 64)) return OJDF_SS_DECLINED;
    int th = wt::kMaxPix / W;                                   // image rows per pixel tile
    if (th > H) th = H;
    if (th > 256) th = 256;
    if (th < 1) return OJDF_SS_DECLINED;
    const int ntiles = (H + th - 1) / th;
    // wide maps: the pixel-major kernels share a weight stage between several pixel tiles, this one re-reads it per tile --
    // it wins while the whole layer is a handful of K steps per CTA (measured on B200, tools/tc_probe.py)
    if (ntiles > 1 && (flags & 262144)) return OJDF_SS_DECLINED;
    for (int i = 0; i < n_problems; ++i)
        if (problems_host[i].out_step > 1 || (problems_host[i].in_step > 1 && (problems_host[i].in_step != 2 || problems_host[i].in_width < (W - 1) * 2 + 1)))
            return OJDF_SS_DECLINED;
    wt::WtParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.H = H; prm.W = W; prm.npix = npix; prm.cin = cin; prm.cout = cout; prm.taps = taps; prm.act = act;
    prm.groups = groups; prm.nprob = n_problems;
    prm.nkc = (cin + tc::kBK - 1) / tc::kBK;
    prm.slope = slope; prm.out_mul = out_mul;
    prm.fast = (flags & 64) ? 1 : 0;
    prm.th = th; prm.ntiles = ntiles;
    prm.npad_n = (th * W + 15) & ~15;
    if (prm.npad_n <= 256) { prm.n1 = prm.npad_n; prm.n2 = 0; }
    else { prm.n1 = ((prm.npad_n / 2) + 15) & ~15; prm.n2 = prm.npad_n - prm.n1; }
    prm.box_bytes = th * W * 128;
    prm.b_slot = (prm.npad_n * 128 + 1023) / 1024 * 1024;       // the MMAs read npad_n rows: the tail rows are never stored
    prm.npad = npad;
    prm.a_bytes = 2 * npad * 128;
    const int slot = 2 * prm.b_slot + prm.a_bytes;
    const int tail_pad = (128 - npad) * 128;                    // what the last stage's W_lo descriptor reads past the stage
    const int budget = 227 * 1024 - 1024 - 1024 - tail_pad;
    prm.stages = budget / slot;
    if (prm.stages > wt::kMaxStages) prm.stages = wt::kMaxStages;
    if (prm.stages < 2) return OJDF_SS_DECLINED;
    int min_steps = 1 << 30;
    for (int i = 0; i < n_problems; ++i) {
        const ojdf_conv_problem &q = problems_host[i];
        if (!q.in_dev || !q.weights_dev || !q.scale_dev || !q.shift_dev || !q.out_dev || (q.in_stride & 3) || q.in_stride < cin ||
            ((uintptr_t)q.in_dev & 15) || ((uintptr_t)q.weights_dev & 15) || q.out_stride < q.out_coffset + cout ||
            q.out_coffset < 0 || q.dilation < 1 || (q.residual_dev && q.residual_stride < cout) || (act == 5 && !q.residual_dev))
            return OJDF_ERR_BADARG;
        const int step = q.in_step > 1 ? q.in_step : 1;          // stride-2 reads live in the tensor map
        const int r = tc::pixel_map(q.in_dev, cin, q.in_stride * step, H, W, W, th, &prm.in_map[i],
                                    step > 1 ? (long long)q.in_stride * q.in_width * step : 0);
        if (r) return r;
        int mask = taps == 9 ? (q.tap_mask ? (q.tap_mask & 511) : 511) : 1;
        if (taps == 9) {                                         // taps that only ever see the zero padding
            for (int t = 0; t < 9; ++t)
                if (abs(t / 3 - 1) * q.dilation >= H || abs(t % 3 - 1) * q.dilation >= W) mask &= ~(1 << t);
            if (!mask) mask = 1 << 4;
        }
        int n = 0;
        for (int t = 0; t < 9; ++t)
            if ((mask >> t) & 1) prm.tap_list[i][n++] = t;
        prm.ntaps[i] = n;
        if (n * prm.nkc < min_steps) min_steps = n * prm.nkc;
        prm.p[i] = tc::Problem{q.weights_dev, q.scale_dev, q.shift_dev, q.out_dev, q.residual_dev,
                                q.out_stride, q.out_coffset, q.dilation, q.residual_stride};
    }
    // split the K loop over CTAs: two K steps per CTA when the SMs allow it (both stage loads are in flight from the start:
    // a CTA is bound by the ~2 us latency of a 110 KB stage, not by its ~1 us of MMAs), at most one CTA per SM
    const long long items = (long long)n_problems * groups * ntiles;
    prm.ksplit = 1;
    prm.cpad = groups * npad;
    if (scratch_dev && !(flags & 4096)) {
        static const int steps_per_cta = [] { const char *e = getenv("OJDF_WT_STEPS"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
        int ks = min_steps / steps_per_cta;
        const int room = (int)(tc::sm_count() / items);
        if (ks > room) ks = room;
        if (ks > 32) ks = 32;
        const size_t per_split = (size_t)n_problems * npix * prm.cpad * sizeof(float);
        if ((size_t)ks * per_split > scratch_bytes) ks = (int)(scratch_bytes / per_split);
        if (ks >= 2) prm.ksplit = ks;
    }
    // Experiment (flag 524288): K slices as the CTAs of a cluster (<= 8, portable) that reduce through distributed shared
    // memory instead of the scratch + reduction launch.  Correct (tests/test_gpu_conv_tc.py runs it), but measured SLOWER on
    // B200 than the separate, programmatically launched reduction: 15x20 1024 -> 256: 16.9 vs 15.0 us, 30x40 128 -> 512: 27.5
    // vs 15.7 us, 3x3 30x40 256 -> 128 x4: 50.6 vs 28.1 us (clusters of 7 on 18-SM GPCs run in two waves; gang scheduling
    // of 226 KB CTAs waits for whole groups of free SMs), AdapNet++ 2.11 vs 1.88 ms.  Off by default.
    if ((flags & 524288) && !(flags & 4096)) {
        int kc = min_steps / 2;
        const int room = (int)(tc::sm_count() / items);
        if (kc > room) kc = room;
        if (kc > 8) kc = 8;
        if (kc >= 2 && ((min_steps + kc - 1) / kc <= 6 || prm.ksplit <= kc)) { prm.ksplit = kc; prm.cluster = 1; }
    }
    // long 1x1 K loops over many pixel tiles: the pixel-major kernel shares each weight stage between pixel tiles and wins
    // (measured: 60x80, 256 -> 256: 20 us there, 32 us here)
    if (ntiles > 1 && taps == 1 && min_steps / prm.ksplit > 6) return OJDF_SS_DECLINED;
    tc::SplitReduce red[tc::kMaxBatch];
    if (prm.ksplit > 1 && !prm.cluster)
        for (int i = 0; i < n_problems; ++i) {
            const ojdf_conv_problem &q = problems_host[i];
            float *part = scratch_dev + (size_t)i * prm.ksplit * npix * prm.cpad;
            red[i] = tc::SplitReduce{q.scale_dev, q.shift_dev, q.residual_dev, q.out_dev, part, q.out_stride, q.out_coffset, q.residual_stride};
            prm.p[i].out = part;
        }
    const size_t smem = (size_t)prm.stages * slot + 1024 + tail_pad;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(wt::conv_wt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
        attr = true;
    }
    const long long total = items * prm.ksplit;
    int grid = tc::sm_count();
    if (grid > total || prm.cluster) grid = (int)total;         // cluster mode: one item per CTA, consecutive CTAs = one cluster
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(wt::kWtThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = (flags & 8192) ? 0 : 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = (unsigned)prm.ksplit;
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = prm.cluster ? 2 : 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, wt::conv_wt_kernel, prm);
    if (le != cudaSuccess) { cudaGetLastError(); return (int)le; }
    if (prm.cluster) return launched(1);
    if (prm.ksplit > 1) {
        const int r = launched(1);
        if (r) return r;
        return launch_split_reduce(red, n_problems, npix, cout, prm.cpad, prm.ksplit, act, slope, out_mul, (cudaStream_t)stream);
    }
    return launched(1);
}
