// Persistent tensor-core tap GEMM of FusionNet / AdapNet++ (modules/model.py:4-283, modules/adapnet.py:12-415):
//
//   out[p, coff+co] = out_mul * act(scale[co] * sum_tap sum_ci in[p + tap*dil, ci] * W[tap,ci,co] + shift[co] (+ residual))
//
// on tcgen05.mma kind::tf32 with the 3xTF32 split (x = hi + lo; hi*hi + lo*hi + hi*lo) and fp32 TMEM
// accumulators (results within ~1e-6 of an fp32 convolution; kind::tf32 uses the top 19 bits of an fp32
// container, hi = x with the low 13 mantissa bits cleared, lo = x - hi exactly, only lo*lo is dropped).
// Design, each point measured on B200 against a first version that issued one 16 KB TMA box per tap and
// one tile per CTA (bound by L2->SM traffic and per-CTA prologue/epilogue):
//   * one CTA per SM, persistent over a contiguous range of 128-pixel M-tiles (8 rows x 16 columns each,
//     column-major over the image so that up to MT vertically adjacent M-tiles form one work group);
//   * the activation HALO of a group ((8*MT + 2d) x (16 + 2d) pixels x 32 channels) is fetched ONCE per
//     K chunk by a single TMA box (zero fill = convolution padding / channel tail / image border) and
//     serves all 9 taps; huge dilations (halo would not fit) fall back to one box per tap;
//   * the A operand lives in TENSOR MEMORY: eight "splitter" warps read the tap-shifted pixel rows out
//     of the swizzled halo tile (conflict-free LDS.128), split them into hi/lo in registers and
//     tcgen05.st both halves into a ring of TMEM stages (2-3 per issuer); tcgen05.mma reads A from TMEM and only the small
//     weight tiles (B, K-major SWIZZLE_128B images packed by the host) from shared memory -- shared
//     memory carries each activation byte once per tap instead of five times;
//   * one weight stage (tap, K chunk) is reused by all MT M-tiles of the group;
//   * accumulators are double buffered in TMEM (when 2*MT*NPAD <= 256 columns) so the epilogue of one
//     group (tcgen05.ld -> scale/shift/residual/activation -> swizzled staging tile -> TMA store, clipped
//     by the tensor map to the image and to channels < coff+cout) overlaps the main loop of the next;
//   * stride-2 layers read every other pixel through the input tensor map (in_step), the phases of a transposed
//     convolution write every s-th pixel through the output tensor map (out_step) with their dead taps masked;
//   * feature maps too small to give every SM an M-tile (AdapNet++ at 15x20) split the K loop over CTAs, raw
//     partial sums go to a scratch buffer and conv_reduce_kernel (ojdf_conv.cu) finishes the layer in a fixed order;
//   * programmatic dependent launch: barrier set-up and the TMEM allocation overlap the previous layer's tail.
// Warp roles (608 threads): 0 = TMA producer, 1 and 18 = MMA issuers (even / odd M-tiles of a group; warp 1
// also owns the TMEM allocation), 2..9 = splitters (two sets of four; set p feeds issuer p through its own
// in-order ring of A stages -- two consumers alternating on one mbarrier would alias its parity), 10..17 =
// epilogue (two warps per TMEM lane quarter, alternating 16-column chunks).  Every role walks the same
// deterministic sequence of work groups (decode()), so no tile scheduler state is shared.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "ojdf_tc_common.cuh"

namespace ojdf {
namespace tc {

// Optional role profile (flag 128, block 0 only): cycles each role spent waiting per barrier class.
__device__ long long g_prof[32];
__device__ __forceinline__ void mbar_wait_p(uint32_t bar, uint32_t parity, int cls, bool on, uint32_t hint = OJDF_PARK_NS)
{
    if (!on) {
        if (hint) mbar_wait(bar, parity, hint); else mbar_poll(bar, parity);
        return;
    }
    const long long t0 = clock64();
    if (hint) mbar_wait(bar, parity, hint); else mbar_poll(bar, parity);
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&g_prof[cls], (unsigned long long)(clock64() - t0));
}
// One work group: up to MT vertically adjacent M-tiles of one (problem, channel group).
struct Group { int z, g, col, row, n, ks, kc0, kc1; };
__device__ __forceinline__ Group decode(const Params &prm, int s, int end)
{
    const int tiles = prm.tiles_x * prm.tiles_y;
    Group gr;
    const int t = s % tiles, r = s / tiles;
    gr.ks = r % prm.ksplit;                                    // split-K slice: K chunks [kc0, kc1)
    const int zg = r / prm.ksplit;
    gr.kc0 = prm.nkc * gr.ks / prm.ksplit;
    gr.kc1 = prm.nkc * (gr.ks + 1) / prm.ksplit;
    gr.g = zg % prm.groups;
    gr.z = zg / prm.groups;
    gr.col = t / prm.tiles_y;
    gr.row = t - gr.col * prm.tiles_y;
    int n = prm.mt;
    if (n > end - s) n = end - s;
    if (n > prm.tiles_y - gr.row) n = prm.tiles_y - gr.row;
    gr.n = n;
    return gr;
}

__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ Params prm)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next layer's prologue start early
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bars[2 * kMaxSrc + 2 * kMaxB + 2 * kAS + 4];
    __shared__ uint32_t s_tmem;
    __shared__ float2 s_ss[128];                                 // (scale, shift) of the epilogue's current channel group
    __shared__ bool s_last;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // warp-uniform role index
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const int npad = prm.npad, HS = prm.src_stages, BS = prm.b_stages, MT = prm.mt, NACC = prm.nacc;
    const uint32_t b_bytes = (uint32_t)npad * 128u;             // one of B_hi / B_lo
    const uint32_t src0 = base;
    const uint32_t bst0 = src0 + (uint32_t)HS * prm.src_bytes;
    const uint32_t stg0 = bst0 + (uint32_t)BS * 2u * b_bytes;
    uint8_t *stg_ptr = smem + (size_t)HS * prm.src_bytes + (size_t)BS * 2u * b_bytes;
    const uint32_t bar0 = smem_u32(s_bars);
    auto src_full = [&](int s) { return bar0 + 8u * s; };
    auto src_empty = [&](int s) { return bar0 + 8u * (kMaxSrc + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (2 * kMaxSrc + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (2 * kMaxSrc + kMaxB + s); };
    auto a_full = [&](int s) { return bar0 + 8u * (2 * kMaxSrc + 2 * kMaxB + s); };
    auto a_empty = [&](int s) { return bar0 + 8u * (2 * kMaxSrc + 2 * kMaxB + kAS + s); };
    auto acc_full = [&](int s) { return bar0 + 8u * (2 * kMaxSrc + 2 * kMaxB + 2 * kAS + s); };
    auto acc_empty = [&](int s) { return bar0 + 8u * (2 * kMaxSrc + 2 * kMaxB + 2 * kAS + 2 + s); };

    const int tiles = prm.tiles_x * prm.tiles_y;
    const int total = prm.nprob * prm.groups * prm.ksplit * tiles;
    const int begin = (int)((long long)total * blockIdx.x / gridDim.x);
    const int end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
    const uint32_t hot_hint = (prm.dbg & 1024) ? 0u : OJDF_PARK_NS;   // experiment: do not park the A-ring waiters
    const bool halo_mode = prm.hd > 0 || prm.taps == 1;
    const int nbox = halo_mode ? 1 : prm.taps;                   // TMA boxes per K chunk
    const int taps_per_box = halo_mode ? prm.taps : 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxSrc; ++s) { mbar_init(src_full(s), 1); mbar_init(src_empty(s), 256); }
        for (int s = 0; s < kMaxB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }
        for (int s = 0; s < kAS; ++s) { mbar_init(a_full(s), 128); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 2); mbar_init(acc_empty(s), kEpiThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 64)
        for (int i = 0; i < prm.nprob; ++i) { prefetch_map(&prm.in_map[i]); if (prm.store_mode == 0) prefetch_map(&prm.out_map[i]); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // Programmatic dependent launch: everything above (barriers, TMEM) overlapped the tail of the previous
    // kernel in the stream; from here on this grid reads what that kernel wrote.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool prof = (prm.dbg & 128) && blockIdx.x == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 6 || warp == kEpi0);
    const long long t_role0 = prof ? clock64() : 0;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (whole warp walks the loop, one
        // elected lane issues: keeps addresses and barrier numbers in uniform registers)
        Ring rs(HS), rb(BS);
        for (int s = begin; s < end;) {
            const Group gr = decode(prm, s, end);
            const Problem &pr = prm.p[gr.z];
            const CUtensorMap *map = &prm.in_map[gr.z];
            const int x0 = gr.col * kBW, y0 = gr.row * kBH;
            const uint8_t *wbase = reinterpret_cast<const uint8_t *>(pr.weights) + (size_t)gr.g * prm.taps * prm.nkc * 2u * b_bytes;
            const int mask = prm.tap_mask[gr.z];
            for (int kc = gr.kc0; kc < gr.kc1; ++kc) {
                for (int box = 0; box < nbox; ++box) {
                    if (!halo_mode && !((mask >> box) & 1)) continue;        // dead tap: no box, no weights, no MMAs
                    const int slot = rs.idx;
                    mbar_wait_p(src_empty(slot), rs.phase ^ 1, 0, prof);
                    int cx = x0 - prm.hd, cy = y0 - prm.hd;
                    if (!halo_mode) { cx = x0 + (box % 3 - 1) * pr.dil; cy = y0 + (box / 3 - 1) * pr.dil; }
                    if (elect_one()) {
                        mbar_expect_tx(src_full(slot), (uint32_t)prm.box_bytes);
                        for (int i = 0; i < prm.nsub; ++i)
                            tma_load_3d(src0 + (uint32_t)slot * prm.src_bytes + (uint32_t)(i * prm.sub_rows * prm.bwid * 128), map,
                                        src_full(slot), kc * kBK, cx, cy + i * prm.sub_rows);
                    }
                    __syncwarp();
                    rs.next();
                    for (int tb = 0; tb < taps_per_box; ++tb) {
                        const int tap = halo_mode ? tb : box;
                        if (!((mask >> tap) & 1)) continue;
                        const int bs = rb.idx;
                        mbar_wait_p(b_empty(bs), rb.phase ^ 1, 1, prof);
                        if (elect_one()) {
                            mbar_expect_tx(b_full(bs), 2u * b_bytes);
                            bulk_load(bst0 + (uint32_t)bs * 2u * b_bytes, wbase + (size_t)(tap * prm.nkc + kc) * 2u * b_bytes,
                                      2u * b_bytes, b_full(bs));
                        }
                        __syncwarp();
                        rb.next();
                    }
                }
            }
            s += gr.n;
        }
    } else if (warp == 1 || warp == kMma2) {
        // ------------------------------------------------------------ MMA issuers (uniform loop, elected lane issues);
        // warp 1 owns the even M-tiles of every group, warp 18 the odd ones: each accumulator has one issuer
        const int par = warp == 1 ? 0 : 1;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        Ring ra(prm.a_slots), rb(BS), rc(NACC);                 // ra: the A stages of this issuer
        for (int s = begin; s < end; rc.next()) {
            const Group gr = decode(prm, s, end);
            const int buf = rc.idx;
            mbar_wait_p(acc_empty(buf), rc.phase ^ 1, 3, prof);
            tc_fence_after();
            const int mask = prm.tap_mask[gr.z], tap0 = __ffs(mask) - 1;
            for (int kc = gr.kc0; kc < gr.kc1; ++kc) {
                int ksteps = (prm.cin - kc * kBK + 7) >> 3;      // K steps of 8 channels that hold real channels
                if (ksteps > kBK / 8) ksteps = kBK / 8;
                for (int tap = 0; tap < prm.taps; ++tap) {
                    if (!((mask >> tap) & 1)) continue;
                    const int bs = rb.idx;
                    mbar_wait_p(b_full(bs), rb.phase, 4, prof);
                    const uint32_t sb = bst0 + (uint32_t)bs * 2u * b_bytes;
                    const uint64_t b_hi = smem_desc(sb), b_lo = smem_desc(sb + b_bytes);
                    for (int mt = par; mt < gr.n; mt += 2, ra.next()) {
                        // issuer `par` and splitter set `par` form an in-order pipeline over their own A slots
                        const int as = par * prm.a_slots + ra.idx;
                        mbar_wait_p(a_full(as), ra.phase, 5, prof, hot_hint);
                        if (!(prm.dbg & 512)) tc_fence_after();
                        const uint32_t acc = tmem + (uint32_t)((buf * MT + mt) * npad);
                        const uint32_t a_hi = tmem + (uint32_t)(prm.acol0 + as * 64), a_lo = a_hi + 32;
                        const uint32_t first = (uint32_t)((kc - gr.kc0) | (tap ^ tap0));   // 0 only for the group's first K step
                        if (elect_one()) {
                            if (!(prm.dbg & 16)) {
                                if (prm.dbg & 64) {                          // timing experiment: 1xTF32
#pragma unroll
                                    for (int k = 0; k < kBK / 8; ++k)
                                        if (k < ksteps) umma_tf32_ts(acc, a_hi + k * 8, b_hi + (uint64_t)(k * 2), idesc, first | (uint32_t)k);
                                } else {
#pragma unroll
                                    for (int k = 0; k < kBK / 8; ++k) {
                                        if (k >= ksteps) break;
                                        const uint64_t ko = (uint64_t)(k * 2);   // +32 bytes along K inside the swizzle atom
                                        umma_tf32_ts(acc, a_lo + k * 8, b_hi + ko, idesc, first | (uint32_t)k);
                                        umma_tf32_ts(acc, a_hi + k * 8, b_lo + ko, idesc, 1);
                                        umma_tf32_ts(acc, a_hi + k * 8, b_hi + ko, idesc, 1);
                                    }
                                }
                            }
                            if (prm.dbg & 256) mbar_arrive(a_empty(as)); else umma_commit(a_empty(as));
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(b_empty(bs));
                    __syncwarp();
                    rb.next();
                }
            }
            if (elect_one()) umma_commit(acc_full(buf));
            __syncwarp();
            s += gr.n;
        }
    } else if (warp < kEpi0) {
        // ------------------------------------------------------------ splitters: halo tile -> hi/lo -> TMEM A ring
        const int set = (warp - 2) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane, ty = m / kBW, tx = m % kBW;
        Ring ra(prm.a_slots), rs(HS);                           // ra: the A stages produced by this set
        for (int s = begin; s < end;) {
            const Group gr = decode(prm, s, end);
            const int dil = prm.p[gr.z].dil;
            const int mask = prm.tap_mask[gr.z];
            for (int kc = gr.kc0; kc < gr.kc1; ++kc) {
                // a chunk with <= 24 real channels (the 19-channel layers, the tail of a 114-channel input) is split and
                // stored as 24 columns: the splitter warps are the bottleneck of this kernel
                const bool narrow = prm.cin - kc * kBK <= 24;
                for (int box = 0; box < nbox; ++box) {
                    if (!halo_mode && !((mask >> box) & 1)) continue;
                    const int slot = rs.idx;
                    mbar_wait_p(src_full(slot), rs.phase, 7 + 3 * set, prof);
                    const uint8_t *src = smem + (size_t)slot * prm.src_bytes;
                    for (int tb = 0; tb < taps_per_box; ++tb) {
                        if (!((mask >> (halo_mode ? tb : box)) & 1)) continue;
                        int oy = 0, ox = 0;
                        if (halo_mode && prm.taps == 9) { oy = prm.hd + (tb / 3 - 1) * dil; ox = prm.hd + (tb % 3 - 1) * dil; }
                        for (int mt = set; mt < gr.n; mt += 2, ra.next()) {
                            const int as = set * prm.a_slots + ra.idx;
                            if (prm.dbg & 32) {                                  // timing experiment: no split work
                                mbar_wait(a_empty(as), ra.phase ^ 1);
                                mbar_arrive(a_full(as));
                                continue;
                            }
                            const int r = (mt * kBH + ty + oy) * prm.bwid + tx + ox;
                            const uint4 *row = reinterpret_cast<const uint4 *>(src + (size_t)r * 128);
                            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(prm.acol0 + as * 64);
                            if (narrow) {
                                uint32_t hi[24], lo[24];
#pragma unroll
                                for (int j = 0; j < 6; ++j) {
                                    const uint4 x = row[j ^ (r & 7)];
                                    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const uint32_t h = xs[e] & 0xFFFFE000u;
                                        hi[4 * j + e] = h;
                                        lo[4 * j + e] = __float_as_uint(__uint_as_float(xs[e]) - __uint_as_float(h));
                                    }
                                }
                                mbar_wait_p(a_empty(as), ra.phase ^ 1, 8 + 3 * set, prof, hot_hint);
                                tc_fence_after();
                                tmem_st16(ta, hi);
                                tmem_st8(ta + 16, hi + 16);
                                tmem_st16(ta + 32, lo);
                                tmem_st8(ta + 48, lo + 16);
                            } else {
                                uint32_t hi[32], lo[32];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint4 x = row[j ^ (r & 7)];
                                    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const uint32_t h = xs[e] & 0xFFFFE000u;
                                        hi[4 * j + e] = h;
                                        lo[4 * j + e] = __float_as_uint(__uint_as_float(xs[e]) - __uint_as_float(h));
                                    }
                                }
                                mbar_wait_p(a_empty(as), ra.phase ^ 1, 8 + 3 * set, prof, hot_hint);
                                tc_fence_after();
                                tmem_st32(ta, hi);
                                tmem_st32(ta + 32, lo);
                            }
                            tmem_st_wait();
                            tc_fence_before();
                            mbar_arrive(a_full(as));
                        }
                    }
                    mbar_arrive(src_empty(slot));
                    rs.next();
                }
            }
            s += gr.n;
        }
    } else {
        // ------------------------------------------------------------ epilogue: 8 warps; warp pair (q, half) owns TMEM
        // lanes 32q..32q+31 and the 16-column chunks with index = half (mod 2)
        const int q = warp & 3, half = (warp - kEpi0) >> 2;
        const int et = threadIdx.x - kEpi0 * 32;                // 0..255
        const int m = q * 32 + lane, ty = m / kBW, tx = m % kBW;
        const int nslab = (npad + 31) / 32;
        int cur_zg = -1;
        Ring rc(NACC);
        for (int s = begin; s < end; rc.next()) {
            const Group gr = decode(prm, s, end);
            const Problem &pr = prm.p[gr.z];
            const int buf = rc.idx;
            const int co_base = gr.g * npad;
            if (cur_zg != gr.z * prm.groups + gr.g) {            // (scale, shift) of this channel group -> smem
                named_bar(1, kEpiThreads);
                if (et < npad) {
                    const int co = co_base + et;
                    s_ss[et] = prm.ksplit > 1 ? make_float2(1.f, 0.f)        // split-K: raw partial sums
                               : co < prm.cout ? make_float2(__ldg(pr.scale + co), __ldg(pr.shift + co)) : make_float2(0.f, 0.f);
                }
                named_bar(1, kEpiThreads);
                cur_zg = gr.z * prm.groups + gr.g;
            }
            // split-K: pr.out is the problem's scratch, one [H*W][cpad] slab per K slice
            float *obase = pr.out + (prm.ksplit > 1 ? (size_t)gr.ks * prm.H * prm.W * prm.cpad : (size_t)0);
            mbar_wait_p(acc_full(buf), rc.phase, 13, prof);
            tc_fence_after();
            for (int mt = 0; mt < gr.n; ++mt) {
                const int x = gr.col * kBW + tx, y = (gr.row + mt) * kBH + ty;
                const bool live = y < prm.H && x < prm.W;
                const size_t pix = (size_t)y * prm.W + x;
                const float *rrow = (pr.residual && live) ? pr.residual + pix * pr.res_stride + co_base : nullptr;
                const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * MT + mt) * npad);
                if (prm.store_mode == 0 && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (prm.store_mode != 2) named_bar(1, kEpiThreads);      // the previous store pass has drained the staging tile
                for (int n0 = half * 16; n0 < npad; n0 += 32) {
                    uint32_t v[16];
                    tmem_ld16(tacc + (uint32_t)n0, v);
                    tmem_ld_wait();
                    float o[16];
                    switch (prm.act) {                           // uniform branch; the element loops are branch-free
                        case kRelu: epi_chunk<kRelu>(v, o, s_ss + n0, rrow ? rrow + n0 : nullptr, prm.cout - co_base - n0, prm.slope, prm.out_mul); break;
                        case kLeaky: epi_chunk<kLeaky>(v, o, s_ss + n0, rrow ? rrow + n0 : nullptr, prm.cout - co_base - n0, prm.slope, prm.out_mul); break;
                        case kTanh: epi_chunk<kTanh>(v, o, s_ss + n0, rrow ? rrow + n0 : nullptr, prm.cout - co_base - n0, prm.slope, prm.out_mul); break;
                        case kSigmoid: epi_chunk<kSigmoid>(v, o, s_ss + n0, rrow ? rrow + n0 : nullptr, prm.cout - co_base - n0, prm.slope, prm.out_mul); break;
                        case kSigmoidMul: epi_chunk<kSigmoidMul>(v, o, s_ss + n0, rrow ? rrow + n0 : nullptr, prm.cout - co_base - n0, prm.slope, prm.out_mul); break;
                        default: epi_chunk<kNone>(v, o, s_ss + n0, rrow ? rrow + n0 : nullptr, prm.cout - co_base - n0, prm.slope, prm.out_mul); break;
                    }
                    if (prm.store_mode == 2) {
                        if (live) {
                            float *orow = obase + pix * pr.out_stride + pr.out_coff + co_base;
#pragma unroll
                            for (int c = 0; c < 16; ++c)
                                if (co_base + n0 + c < prm.cout) orow[n0 + c] = o[c];
                        }
                    } else {
                        // staging slab (n0/32): 128 rows of 128 bytes, SWIZZLE_128B like the tensor map expects
                        uint8_t *slab = stg_ptr + (size_t)(n0 >> 5) * kSlabBytes + (size_t)m * 128;
                        const int j0 = (n0 & 16) >> 2;           // first 16-byte chunk of this half row
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4 *>(slab + (((j0 + j) ^ (m & 7)) << 4)) =
                                make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    }
                }
                if (prm.store_mode == 0) {
                    fence_proxy_async();
                    named_bar(1, kEpiThreads);
                    if (et == 0) {
                        for (int sl = 0; sl < nslab; ++sl)
                            tma_store_3d(&prm.out_map[gr.z], stg0 + (uint32_t)sl * kSlabBytes, pr.out_coff + co_base + sl * 32,
                                         gr.col * kBW, (gr.row + mt) * kBH);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else if (prm.store_mode == 1) {
                    // row-contiguous stores out of the staging tile: consecutive lanes write consecutive channels
                    // of one pixel (unaligned channel offsets / widths the TMA store cannot express)
                    named_bar(1, kEpiThreads);
                    int nco = prm.cout - co_base;
                    if (nco > npad) nco = npad;
                    const int rbeg = q * 32 + half * 16, rend = rbeg + 16;
                    if (nco <= 32) {
                        const int lanes_per_row = nco <= 16 ? 16 : 32, rpp = 32 / lanes_per_row;
                        const int c = lane % lanes_per_row, rsub = lane / lanes_per_row;
                        for (int r0 = rbeg; r0 < rend; r0 += rpp) {
                            const int rr = r0 + rsub;
                            const int px = gr.col * kBW + rr % kBW, py = (gr.row + mt) * kBH + rr / kBW;
                            if (c < nco && py < prm.H && px < prm.W) {
                                const float v = *reinterpret_cast<const float *>(stg_ptr + (size_t)rr * 128 + (((c >> 2) ^ (rr & 7)) << 4) + ((c & 3) << 2));
                                obase[((size_t)py * prm.W + px) * pr.out_stride + pr.out_coff + co_base + c] = v;
                            }
                        }
                    } else {
                        for (int rr = rbeg; rr < rend; ++rr) {
                            const int px = gr.col * kBW + rr % kBW, py = (gr.row + mt) * kBH + rr / kBW;
                            if (py >= prm.H || px >= prm.W) continue;
                            float *orow = obase + ((size_t)py * prm.W + px) * pr.out_stride + pr.out_coff + co_base;
                            for (int c = lane; c < nco; c += 32)
                                orow[c] = *reinterpret_cast<const float *>(stg_ptr + (size_t)(c >> 5) * kSlabBytes + (size_t)rr * 128 +
                                                                           ((((c & 31) >> 2) ^ (rr & 7)) << 4) + ((c & 3) << 2));
                        }
                    }
                }
                if (prm.ksplit > 1 && prm.counters) {
                    // this CTA's slice of the tile is written: count it; whoever delivers the last slice finishes the tile
                    __threadfence();
                    named_bar(1, kEpiThreads);
                    const int tile_id = (gr.z * prm.groups + gr.g) * tiles + gr.col * prm.tiles_y + gr.row + mt;
                    if (et == 0) {
                        const unsigned int old = atomicAdd(&prm.counters[tile_id], 1u);
                        s_last = old == (unsigned int)(prm.ksplit - 1);
                        if (s_last) prm.counters[tile_id] = 0u;          // back to idle for the next launch
                    }
                    named_bar(1, kEpiThreads);
                    if (s_last) {
                        __threadfence();
                        const SplitReduce &rd = prm.red[gr.z];
                        int nco = prm.red_cout - co_base;
                        if (nco > npad) nco = npad;
                        const size_t slab = (size_t)prm.H * prm.W * prm.cpad;
                        for (int i = et; i < 128 * nco; i += kEpiThreads) {
                            const int rr = i / nco, c = i - rr * nco;
                            const int px = gr.col * kBW + rr % kBW, py = (gr.row + mt) * kBH + rr / kBW;
                            if (py >= prm.H || px >= prm.W) continue;
                            const size_t pixi = (size_t)py * prm.W + px;
                            const float *pp = rd.partial + pixi * prm.cpad + co_base + c;
                            float a = 0.0f;
                            for (int k = 0; k < prm.ksplit; ++k) a += __ldcg(pp + (size_t)k * slab);
                            const int co = co_base + c;
                            float v = fmaf(a, __ldg(rd.scale + co), __ldg(rd.shift + co));
                            if (prm.red_act == kSigmoidMul) v = rd.residual[pixi * rd.res_stride + co] / (1.0f + expf(-v));
                            else {
                                if (rd.residual) v += rd.residual[pixi * rd.res_stride + co];
                                v = activate(v, prm.red_act, prm.slope);
                            }
                            rd.out[pixi * rd.out_stride + rd.out_coff + co] = v * prm.red_out_mul;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty(buf));
            s += gr.n;
        }
        // the staging slabs must have been READ before the CTA (and its shared memory) goes away; the global writes of the
        // bulk stores complete on their own before the grid does
        if (prm.store_mode == 0 && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (prof && lane == 0) {
        const int slot = warp == 0 ? 2 : warp == 1 ? 6 : warp == 2 ? 9 : warp == 6 ? 12 : 14;     // warp kEpi0 -> 14
        atomicAdd((unsigned long long *)&g_prof[slot], (unsigned long long)(clock64() - t_role0));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
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
    int c, stride, H, W, bw, bh;
    long long row_stride;
    bool operator==(const MapKey &o) const
    {
        return ptr == o.ptr && c == o.c && stride == o.stride && H == o.H && W == o.W && bw == o.bw && bh == o.bh &&
               row_stride == o.row_stride;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const
    {
        size_t h = (size_t)k.ptr;
        const int v[6] = {k.c, k.stride, k.H, k.W, k.bw, k.bh};
        for (int i = 0; i < 6; ++i) h = h * 1000003u ^ (size_t)v[i];
        return h;
    }
};

// (C, W, H) view of a pixel-major fp32 buffer with `c` visible channels; box = (32, bw, bh), 128-byte
// swizzle.  Loads read zeros outside the view, stores drop what falls outside.
// `stride` = floats between consecutive pixels of the view, `row_stride` = floats between its rows (0: W * stride).
int pixel_map(const float *ptr, int c, int stride, int H, int W, int bw, int bh, CUtensorMap *out, long long row_stride)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    if (!row_stride) row_stride = (long long)W * stride;
    const MapKey key{ptr, c, stride, H, W, bw, bh, row_stride};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[3] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[2] = {(cuuint64_t)stride * 4u, (cuuint64_t)row_stride * 4u};
    const cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OJDF_ERR_BADARG;
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return 0;
}

int sm_count()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    }
    return n;
}

}  // namespace tc
}  // namespace ojdf

using namespace ojdf;


// Channel-group layout: by default one group of round_up(cout,16) <= 128 columns (or ceil(cout/128) equal
// groups); `npad_req` (multiple of 16, <= 128) forces narrower groups -- more CTAs for the small feature
// maps of AdapNet++ where the pixels alone cannot fill the SMs.
void ojdf_tc_layout(int cout, int npad_req, int *npad, int *groups)
{
    if (npad_req >= 16 && npad_req <= 128 && !(npad_req & 15)) {
        *npad = npad_req < ((cout + 15) & ~15) ? npad_req : ((cout + 15) & ~15);
        *groups = (cout + *npad - 1) / *npad;
        return;
    }
    const int g = (cout + 127) / 128;
    const int per = (cout + g - 1) / g;
    *groups = g;
    *npad = (per + 15) / 16 * 16;
}

extern "C" int ojdf_conv_tc_layout(int cout, int npad_req, int *npad, int *groups)
{
    if (cout < 1 || !npad || !groups) return OJDF_ERR_BADARG;
    ojdf_tc_layout(cout, npad_req, npad, groups);
    return 0;
}

extern "C" size_t ojdf_conv_tc_weight_floats(int cin, int cout, int taps, int npad_req)
{
    if (cin < 1 || cout < 1 || taps < 1) return 0;
    int npad, groups;
    ojdf_tc_layout(cout, npad_req, &npad, &groups);
    const int nkc = (cin + tc::kBK - 1) / tc::kBK;
    return (size_t)groups * taps * nkc * 2 * npad * tc::kBK;
}

// Packed weights: [group][tap][K chunk of 32][hi | lo][npad rows][32 floats]; every [npad][32] block is a
// ready-made K-major SWIZZLE_128B shared-memory image (16-byte chunk c of row r stored at chunk c ^ (r & 7)),
// hi = w with the low 13 mantissa bits cleared (exactly representable in tf32), lo = w - hi.
extern "C" int ojdf_conv_tc_pack_weights(const float *w_host, int cin, int cout, int taps, int npad_req, float *packed_host)
{
    if (!w_host || !packed_host || cin < 1 || cout < 1 || taps < 1) return OJDF_ERR_BADARG;
    int npad, groups;
    ojdf_tc_layout(cout, npad_req, &npad, &groups);
    const int nkc = (cin + tc::kBK - 1) / tc::kBK;
    memset(packed_host, 0, ojdf_conv_tc_weight_floats(cin, cout, taps, npad_req) * sizeof(float));
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
                        const int chunk = (k >> 2) ^ (r & 7);
                        const size_t off = (size_t)r * tc::kBK + chunk * 4 + (k & 3);
                        img[off] = hi;
                        img[(size_t)npad * tc::kBK + off] = lo;
                    }
                }
            }
    return 0;
}

// Debug aid (not part of include/ojdf.h): read and clear the role profile filled by launches with flag 128.
extern "C" int ojdf_conv_tc_profile(long long *out_host32)
{
    long long zero[32] = {0};
    cudaError_t e = cudaMemcpyFromSymbol(out_host32, tc::g_prof, sizeof(zero));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(tc::g_prof, zero, sizeof(zero));
    return (int)e;
}

extern "C" int ojdf_conv_tc_batched(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H,
                                    int W, int taps, int act, float slope, float out_mul, int npad_req, int flags,
                                    float *scratch_dev, size_t scratch_bytes, void *stream)
{
    if (!problems_host || n_problems < 1 || n_problems > tc::kMaxBatch || cin < 1 || cout < 1 || H < 1 || W < 1 ||
        H > 32767 || W > 32767 || (taps != 1 && taps != 9) || act < 0 || act > 5)
        return OJDF_ERR_BADARG;
    {   // 3x3 layers whose halo box fits in shared memory go to the shared-memory-operand kernel (ojdf_conv_ss.cu), the
        // rest stays here.  OJDF_CONV_KERNEL=ts / flag 65536: always this file's kernel; =ss / flag 32768: always the other.
        static const int env = [] { const char *e = getenv("OJDF_CONV_KERNEL"); return !e ? 0 : !strcmp(e, "ts") ? 1 : !strcmp(e, "ss") ? 2 : 0; }();
        if (!(flags & (65536 | 32768)) && env == 0) {             // small maps with wide outputs: channels as M, the image as N
            const int r = ojdf_conv_wt_launch(problems_host, n_problems, cin, cout, H, W, taps, act, slope, out_mul, npad_req, flags,
                                              scratch_dev, scratch_bytes, stream);
            if (r != OJDF_SS_DECLINED) return r;
        }
        if (!(flags & 65536) && env != 1) {
            const int r = ojdf_conv_ss_launch(problems_host, n_problems, cin, cout, H, W, taps, act, slope, out_mul, npad_req,
                                              flags | (env == 2 ? 32768 : 0), scratch_dev, scratch_bytes, stream);
            if (r != OJDF_SS_DECLINED) return r;
        }
    }
    tc::Params prm;
    memset(&prm, 0, sizeof(prm));
    int npad, groups;
    ojdf_tc_layout(cout, npad_req, &npad, &groups);
    prm.H = H; prm.W = W; prm.cin = cin; prm.cout = cout; prm.taps = taps; prm.act = act; prm.npad = npad;
    prm.groups = groups; prm.nprob = n_problems;
    prm.nkc = (cin + tc::kBK - 1) / tc::kBK;
    prm.tiles_x = (W + tc::kBW - 1) / tc::kBW;
    prm.tiles_y = (H + tc::kBH - 1) / tc::kBH;
    prm.slope = slope; prm.out_mul = out_mul;
    // accumulators own 256 TMEM columns: MT M-tiles x NACC buffers x npad columns
    if (npad <= 32) { prm.mt = 4; prm.nacc = 2; }
    else if (npad <= 64) { prm.mt = 2; prm.nacc = 2; }
    else { prm.mt = 2; prm.nacc = 1; }
    if (flags & 2) prm.mt = 1;
    if (flags & 2048) prm.nacc = 1;                           // experiment: single accumulator set, deeper A ring
    if (prm.mt > prm.tiles_y) prm.mt = prm.tiles_y;
    int dil = problems_host[0].dilation;
    bool same_dil = true;
    for (int i = 0; i < n_problems; ++i) same_dil = same_dil && problems_host[i].dilation == dil;
    // shared-memory plan: staging slabs + weight stages + source (halo or per-tap) stages
    const int budget = 227 * 1024 - 1024 - 2048;
    const int stg = ((npad + 31) / 32) * tc::kSlabBytes;
    const int b_stage = 2 * npad * 128;
    int hd = 0;
    if (taps == 9 && same_dil && !(flags & 4)) {
        for (;;) {                                             // halo mode if two halo stages + two weight stages fit
            const int rows = (tc::kBH * prm.mt + 2 * dil) * (tc::kBW + 2 * dil);
            if (tc::kBW + 2 * dil <= 256 && tc::kBH * prm.mt + 2 * dil <= 256 &&
                2 * ((rows * 128 + 1023) / 1024 * 1024) + 2 * b_stage + stg <= budget) { hd = dil; break; }
            if (prm.mt == 1) break;
            prm.mt >>= 1;
        }
        if (!hd) {                                             // per-tap boxes: restore the TMEM-driven MT
            prm.mt = npad <= 32 ? 4 : 2;
            if (prm.mt > prm.tiles_y) prm.mt = prm.tiles_y;
        }
    }
    prm.hd = hd;
    {   // TMEM: accumulators first, the rest (up to 3 x 64 columns per issuer/splitter pair) is the A ring
        const int acc_cols = (prm.mt * prm.nacc * npad + 63) / 64 * 64;
        prm.acol0 = acc_cols;
        prm.a_slots = (512 - acc_cols) / 128;
        if (prm.a_slots > tc::kAS / 2) prm.a_slots = tc::kAS / 2;
        if (prm.a_slots < 1) return OJDF_ERR_BADARG;
    }
    prm.bwid = tc::kBW + 2 * hd;
    int bhid = tc::kBH * prm.mt + 2 * hd;
    // sub-boxes: a whole number of 1024-byte swizzle atoms each (sub_rows * bwid % 8 == 0)
    int sub_min = 1;
    while ((sub_min * prm.bwid) & 7) sub_min <<= 1;
    const int sub_mul = (flags >> 12) & 15;                     // experiment knob: 0 = one TMA operation per box
    prm.sub_rows = sub_mul ? sub_min * sub_mul : bhid;
    if (prm.sub_rows > bhid) prm.sub_rows = bhid;
    prm.nsub = (bhid + prm.sub_rows - 1) / prm.sub_rows;
    bhid = prm.nsub * prm.sub_rows;                             // extra rows (if any) are loaded and never read
    if (bhid > 256) return OJDF_ERR_BADARG;
    prm.box_bytes = prm.bwid * bhid * 128;
    prm.src_bytes = (prm.box_bytes + 1023) / 1024 * 1024;
    int bs = 2, hs = 2;
    if (hs * prm.src_bytes + bs * b_stage + stg > budget) hs = 1;
    if (hs * prm.src_bytes + bs * b_stage + stg > budget) return OJDF_ERR_BADARG;
    for (bool grew = true; grew;) {                             // grow the rings while they fit: weights first
        grew = false;
        if (bs < tc::kMaxB && hs * prm.src_bytes + (bs + 1) * b_stage + stg <= budget) { ++bs; grew = true; }
        if (hs < tc::kMaxSrc && hs < 3 && (hs + 1) * prm.src_bytes + bs * b_stage + stg <= budget) { ++hs; grew = true; }
    }
    prm.src_stages = hs; prm.b_stages = bs;
    // store mode 0: TMA store (needs 16-byte aligned channel offset and a width that is a multiple of 4, or
    // pad channels the caller does not care about); 1: coalesced stores from the staging tile; 2: per-thread stores
    prm.store_mode = (flags & 8) ? 2 : 0;
    prm.dbg = flags & (16 | 32 | 64 | 128 | 256 | 512 | 1024);
    for (int i = 0; i < n_problems; ++i) {
        const ojdf_conv_problem &q = problems_host[i];
        if (!q.in_dev || !q.weights_dev || !q.scale_dev || !q.shift_dev || !q.out_dev || (q.in_stride & 3) || q.in_stride < cin ||
            ((uintptr_t)q.in_dev & 15) || ((uintptr_t)q.weights_dev & 15) || q.out_stride < q.out_coffset + cout ||
            q.out_coffset < 0 || q.dilation < 1 || (q.residual_dev && q.residual_stride < cout))
            return OJDF_ERR_BADARG;
        // TMA stores move whole 16-byte units: the channel offset must be a multiple of 4 and a width that is not
        // is rounded up -- allowed only when the caller owns those pad channels (flag 1); they receive zeros
        if (prm.store_mode == 0 && ((q.out_stride & 3) || ((uintptr_t)q.out_dev & 15) || (q.out_coffset & 3) ||
                                    ((cout & 3) && (!(flags & 1) || q.out_coffset + ((cout + 3) & ~3) > q.out_stride))))
            prm.store_mode = 1;
    }
    for (int i = 0; i < n_problems; ++i) {
        const ojdf_conv_problem &q = problems_host[i];
        const int step = q.in_step > 1 ? q.in_step : 1;
        if (step > 1 && (step != 2 || q.in_width < (W - 1) * step + 1)) return OJDF_ERR_BADARG;
        int r = tc::pixel_map(q.in_dev, cin, q.in_stride * step, H, W, prm.bwid, prm.sub_rows, &prm.in_map[i],
                              step > 1 ? (long long)q.in_stride * q.in_width * step : 0);
        if (r) return r;
        const int ostep = q.out_step > 1 ? q.out_step : 1;
        if (ostep > 1 && (prm.store_mode != 0 || q.out_width < (W - 1) * ostep + 1)) return OJDF_ERR_BADARG;
        if (prm.store_mode == 0) {
            r = tc::pixel_map(q.out_dev, q.out_coffset + ((cout + 3) & ~3), q.out_stride * ostep, H, W, tc::kBW, tc::kBH,
                              &prm.out_map[i], ostep > 1 ? (long long)q.out_stride * q.out_width * ostep : 0);
            if (r) return r;
        }
        int mask = taps == 9 ? (q.tap_mask ? (q.tap_mask & 511) : 511) : 1;
        if (taps == 9) {                                         // taps that only ever see the zero padding
            for (int t = 0; t < 9; ++t)
                if (abs(t / 3 - 1) * q.dilation >= H || abs(t % 3 - 1) * q.dilation >= W) mask &= ~(1 << t);
            if (!mask) mask = 1 << 4;
        }
        prm.tap_mask[i] = mask;
        prm.p[i] = tc::Problem{q.weights_dev, q.scale_dev, q.shift_dev, q.out_dev, q.residual_dev,
                                q.out_stride, q.out_coffset, q.dilation, q.residual_stride};
    }
    const size_t smem = (size_t)hs * prm.src_bytes + (size_t)bs * b_stage + stg + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(tc::conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        attr = true;
    }
    // Split the K loop over several CTAs when the pixels alone cannot fill the SMs (AdapNet++'s 15x20 maps with
    // 256..2048 input channels): raw partial sums go to the caller's scratch, launch_split_reduce() finishes the layer.
    const long long items = (long long)n_problems * groups * prm.tiles_x * prm.tiles_y;
    prm.ksplit = 1;
    prm.cpad = groups * npad;
    bool strided_out = false;
    for (int i = 0; i < n_problems; ++i) strided_out = strided_out || problems_host[i].out_step > 1;
    if (scratch_dev && !(flags & 4096) && !strided_out && items * 2 <= tc::sm_count() && prm.nkc >= 8) {
        int ks = (int)((tc::sm_count() + items - 1) / items);
        if (ks > prm.nkc / 4) ks = prm.nkc / 4;                   // at least 4 K chunks per slice
        if (ks > 16) ks = 16;
        const size_t per_split = (size_t)n_problems * H * W * prm.cpad * sizeof(float);
        if ((size_t)ks * per_split > scratch_bytes) ks = (int)(scratch_bytes / per_split);
        if (ks >= 2) prm.ksplit = ks;
    }
    tc::SplitReduce red[tc::kMaxBatch];
    // flag 16384: finish the split-K layer inside this kernel (the CTA that delivers the last slice of a tile sums the
    // slices; the first 16 KB of the scratch are then per-tile slice counters, zero before the first launch and left zero
    // by every launch).  Measured on B200: one SM reducing a whole tile (1 MB of partials at 15x20 / 16 slices) takes
    // longer than the separate, chip-wide conv_reduce_kernel launch it saves (85 vs 19 us per layer), so the default is
    // the separate reduction.
    const bool fused_reduce = (flags & 16384) != 0;
    const size_t counter_floats = 4096;
    if (prm.ksplit > 1 && fused_reduce) {
        if (scratch_bytes < counter_floats * 4 + (size_t)2 * n_problems * H * W * prm.cpad * sizeof(float) || items > (long long)counter_floats) {
            prm.ksplit = 1;
        } else {
            const size_t per_split = (size_t)n_problems * H * W * prm.cpad * sizeof(float);
            const int fit = (int)((scratch_bytes - counter_floats * 4) / per_split);
            if (prm.ksplit > fit) prm.ksplit = fit;
            if (prm.ksplit < 2) prm.ksplit = 1;
        }
    }
    if (prm.ksplit > 1) {
        float *part0 = scratch_dev + (fused_reduce ? counter_floats : 0);
        prm.counters = fused_reduce ? reinterpret_cast<unsigned int *>(scratch_dev) : nullptr;
        prm.red_act = act; prm.red_cout = cout; prm.red_out_mul = out_mul;
        for (int i = 0; i < n_problems; ++i) {
            const ojdf_conv_problem &q = problems_host[i];
            float *part = part0 + (size_t)i * prm.ksplit * H * W * prm.cpad;
            red[i] = tc::SplitReduce{q.scale_dev, q.shift_dev, q.residual_dev, q.out_dev, part, q.out_stride, q.out_coffset,
                                     q.residual_stride};
            prm.red[i] = red[i];
            prm.p[i].out = part;
            prm.p[i].residual = nullptr;
            prm.p[i].out_stride = prm.cpad;
            prm.p[i].out_coff = 0;
        }
        prm.cout = prm.cpad;                                     // every accumulator column is stored
        prm.act = 0;
        prm.out_mul = 1.0f;
        prm.store_mode = 1;
    }
    const long long total = items * prm.ksplit;
    int grid = tc::sm_count();
    if (grid > total) grid = (int)total;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(tc::kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = (flags & 8192) ? 0 : 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const cudaError_t le = cudaLaunchKernelEx(&cfg, tc::conv_tc_kernel, prm);
        if (le != cudaSuccess) { cudaGetLastError(); return (int)le; }
    }
    if (prm.ksplit > 1 && !prm.counters) {
        const int r = launched(1);
        if (r) return r;
        return launch_split_reduce(red, n_problems, H * W, cout, prm.cpad, prm.ksplit, act, slope, out_mul, (cudaStream_t)stream);
    }
    return launched(1);
}
