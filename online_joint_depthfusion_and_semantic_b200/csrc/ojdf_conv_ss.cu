// Tap GEMM of FusionNet / AdapNet++ (modules/model.py:4-283, modules/adapnet.py:12-415) with BOTH operands in
// shared memory -- the kernel behind ojdf_conv_tc_batched:
//
//   out[p, coff+co] = out_mul * act(scale[co] * sum_tap sum_ci in[p + tap*dil, ci] * W[tap,ci,co] + shift[co] (+ residual))
//
// on tcgen05.mma kind::tf32 with the 3xTF32 split (x = hi + lo; hi*hi + hi*lo + lo*hi, fp32 accumulators in tensor
// memory; ~1e-6 of an fp32 convolution).  What the measurements on B200 (tools/probe/umma_probe.cu) say and what the
// design does with it:
//   * an A descriptor (K-major, SWIZZLE_128B) may start at ANY 128-byte pixel row of a swizzled tile and may step
//     between its 8-row groups by ANY multiple of 128 bytes: the swizzle is a function of the absolute shared-memory
//     address.  So a 16-row x 8-column pixel tile of a TMA-loaded HALO box is one descriptor, and each of the nine
//     taps of a 3x3 convolution is the same box with a different start row -- no per-tap copy, no per-tap split;
//   * kind::tf32 reads the top 19 bits of the fp32 container, so the TMA-written box itself is the `hi` operand;
//     only `lo = x - hi` is computed, ONCE per box (not per tap), by eight warps into a second buffer of the same
//     layout;
//   * an SS-mode MMA of M = 128, K = 8 costs max(N/2, 32 + N/4) cycles (the 32 are the A read): with 19..32 output
//     channels the A read dominates, so hi*W_hi and hi*W_lo are ONE instruction against the [W_hi | W_lo] rows
//     (N = 2*npad, its two column halves are added in the epilogue) and lo*W_hi is the second: 88 instead of 120
//     cycles per K step at npad = 32;
//   * the issuing warp has no per-tile hand-off left: it waits once per box (lo ready) and once per weight stage.
// One persistent CTA per SM; warp roles: 0..7 = lo pass, 8..15 = epilogue, 16 = TMA producer of the activation boxes,
// 17 = producer of the weight stages, 18.. = MMA issuers (18 owns the TMEM allocation) (tcgen05.ld -> scale/shift/residual/activation -> swizzled staging -> TMA store).
// Work item = (problem, channel group, K slice, 16x8 pixel tile); a group of up to MT tiles shares its weight
// stages (and, in halo mode, one box).  Boxes too large for shared memory (dilation 9, 27) and 1x1 convolutions use
// one box per (tile, tap).  Stride-2 reads / phase-strided writes live in the tensor maps, split-K and the three
// store modes are those of the first kernel (ojdf_conv_tc.cu).
#include <cstdlib>
#include <cstring>

#include "ojdf_tc_common.cuh"

namespace ojdf {
namespace ss {

using namespace ojdf::tc;

constexpr int kTW = 8, kTH = 16;               // pixel tile: 16 rows x 8 columns (an 8-pixel row = one swizzle atom)
constexpr int kMaxIssuers = 2;
constexpr int kSsThreads = 32 * (18 + kMaxIssuers);
// warp roles; the issuer and the producer get the HIGHEST warp ids: the SM sub-partition arbiter serves the highest
// eligible warp first, and a low-numbered issuer starves behind the waiting loops of the other roles
constexpr int kSplit0 = 0, kSplitThreads = 256;   // warps 0..7: lo pass
constexpr int kEpiW0 = 8;                       // warps 8..15: epilogue
constexpr int kProducerWarp = 16, kProducerBWarp = 17, kIssuerWarp = 18;   // issuers: warps 18 .. 18 + NI - 1

struct SsParams {
    CUtensorMap in_map[kMaxBatch];
    CUtensorMap out_map[kMaxBatch];
    Problem p[kMaxBatch];
    int H, W, cin, cout, taps, act, npad, groups, nkc, tiles_x, tiles_y, nprob;
    int mt, nacc, hd, halo, fused, acc_cols, fast;
    int src_stages, b_stages, src_bytes, box_bytes, bwid, stg_slabs, store_mode, ksplit, cpad, dbg, ni;
    int tap_mask[kMaxBatch];
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
// K-major SWIZZLE_128B operand: `sbo` bytes between consecutive 8-row groups.
__device__ __forceinline__ uint64_t ss_desc(uint32_t addr, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// tcgen05.mma with both descriptors given as (low word, high word).
__device__ __forceinline__ void umma_ss2(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All K steps of one (tile, tap, K chunk): a / b are the low descriptor words of the raw box window and of W_hi;
// +lo_delta = the lo copy of the box, +wl_delta = W_lo; +2 per K step (32 bytes inside the swizzle atom).
template <int KS, bool FUSED>
__device__ __forceinline__ void issue_k(uint32_t acc, uint32_t a, uint32_t lo_delta, uint32_t b, uint32_t wl_delta, uint32_t AH, uint32_t BH,
                                        uint32_t idw, uint32_t idn, uint32_t accum)
{
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        if (FUSED) {
            umma_ss2(acc, a + 2 * k, AH, b + 2 * k, BH, idw, k ? 1u : accum);                  // hi * [W_hi | W_lo]
            umma_ss2(acc, a + lo_delta + 2 * k, AH, b + 2 * k, BH, idn, 1u);                  // lo * W_hi
        } else {
            umma_ss2(acc, a + lo_delta + 2 * k, AH, b + 2 * k, BH, idn, k ? 1u : accum);
            umma_ss2(acc, a + 2 * k, AH, b + wl_delta + 2 * k, BH, idn, 1u);
            umma_ss2(acc, a + 2 * k, AH, b + 2 * k, BH, idn, 1u);
        }
    }
}
__device__ __forceinline__ void issue_tile(bool fused, int ksteps, uint32_t acc, uint32_t a, uint32_t lo_delta, uint32_t b, uint32_t wl_delta,
                                           uint32_t AH, uint32_t BH, uint32_t idw, uint32_t idn, uint32_t accum, bool fast = false)
{
    if (fast) {                                                 // 1xTF32: hi * W_hi only (precision mode `fast`)
        for (int k = 0; k < ksteps; ++k) umma_ss2(acc, a + 2 * k, AH, b + 2 * k, BH, idn, k ? 1u : accum);
        return;
    }
    if (fused) {
        switch (ksteps) {
            case 4: issue_k<4, true>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
            case 3: issue_k<3, true>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
            case 2: issue_k<2, true>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
            default: issue_k<1, true>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
        }
    } else {
        switch (ksteps) {
            case 4: issue_k<4, false>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
            case 3: issue_k<3, false>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
            case 2: issue_k<2, false>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
            default: issue_k<1, false>(acc, a, lo_delta, b, wl_delta, AH, BH, idw, idn, accum); break;
        }
    }
}

// Lean mbarrier wait for the single-warp roles: one try_wait (parks the warp in hardware) and a branch when the
// phase is already complete; the spin counter / trap only exist on the retry path.
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(OJDF_PARK_NS)
        : "memory");
    if (!done) mbar_wait(bar, parity);
}
// Optional role profile (compile with -DOJDF_SS_PROFILE; flag 128, block 0): total cycles per role and cycles spent
// waiting per barrier class.  Off by default: the bookkeeping alone slows the single-warp roles measurably.
__device__ long long g_ss_prof[32];
__device__ __forceinline__ void wait_p(uint32_t bar, uint32_t parity, int cls, bool on)
{
#ifdef OJDF_SS_PROFILE
    if (on) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&g_ss_prof[cls], (unsigned long long)(clock64() - t0));
        return;
    }
#endif
    mbar_wait_fast(bar, parity);
}

// One work group: up to MT tiles of one (problem, channel group, K slice); halo mode: adjacent along x.
struct Group { int z, g, t0, n, ks, kc0, kc1; };
__device__ __forceinline__ Group decode(const SsParams &prm, int s, int end)
{
    const int tiles = prm.tiles_x * prm.tiles_y;
    Group gr;
    const int t = s % tiles, r = s / tiles;
    gr.ks = r % prm.ksplit;
    const int zg = r / prm.ksplit;
    gr.kc0 = prm.nkc * gr.ks / prm.ksplit;
    gr.kc1 = prm.nkc * (gr.ks + 1) / prm.ksplit;
    gr.g = zg % prm.groups;
    gr.z = zg / prm.groups;
    gr.t0 = t;
    int n = prm.mt;
    if (n > end - s) n = end - s;
    const int room = prm.halo ? prm.tiles_x - t % prm.tiles_x : tiles - t;
    if (n > room) n = room;
    gr.n = n;
    return gr;
}

__global__ void __launch_bounds__(kSsThreads, 1) conv_ss_kernel(const __grid_constant__ SsParams prm)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bars[3 * kMaxSrc + 2 * kMaxB + 4];
    __shared__ uint32_t s_tmem;
    __shared__ float2 s_ss[128];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const int npad = prm.npad, HS = prm.src_stages, BS = prm.b_stages, MT = prm.mt, NACC = prm.nacc;
    const uint32_t b_bytes = (uint32_t)npad * 128u;             // one of W_hi / W_lo
    const uint32_t src0 = base;                                  // raw boxes (the `hi` operand)
    const uint32_t lo0 = src0 + (uint32_t)HS * prm.src_bytes;    // their `lo` copies
    const uint32_t bst0 = lo0 + (uint32_t)HS * prm.src_bytes;
    const uint32_t stg0 = bst0 + (uint32_t)BS * 2u * b_bytes;
    uint8_t *stg_ptr = smem + (size_t)2 * HS * prm.src_bytes + (size_t)BS * 2u * b_bytes;
    const uint32_t bar0 = smem_u32(s_bars);
    auto src_full = [&](int s) { return bar0 + 8u * s; };
    auto src_empty = [&](int s) { return bar0 + 8u * (kMaxSrc + s); };
    auto lo_full = [&](int s) { return bar0 + 8u * (2 * kMaxSrc + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (3 * kMaxSrc + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (3 * kMaxSrc + kMaxB + s); };
    auto acc_full = [&](int s) { return bar0 + 8u * (3 * kMaxSrc + 2 * kMaxB + s); };
    auto acc_empty = [&](int s) { return bar0 + 8u * (3 * kMaxSrc + 2 * kMaxB + 2 + s); };

    const bool halo = prm.halo != 0;
    const int tiles = prm.tiles_x * prm.tiles_y;
    const int total = prm.nprob * prm.groups * prm.ksplit * tiles;
    const int begin = (int)((long long)total * blockIdx.x / gridDim.x);
    const int end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxSrc; ++s) { mbar_init(src_full(s), 1); mbar_init(src_empty(s), halo ? prm.ni : 1); mbar_init(lo_full(s), kSplitThreads); }
        for (int s = 0; s < kMaxB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), prm.ni); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), prm.ni); mbar_init(acc_empty(s), kEpiThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kIssuerWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32)
        for (int i = 0; i < prm.nprob; ++i) { prefetch_map(&prm.in_map[i]); if (prm.store_mode == 0) prefetch_map(&prm.out_map[i]); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    asm volatile("griddepcontrol.wait;" ::: "memory");         // from here on this grid reads what the previous kernel wrote
    const bool prof = (prm.dbg & 128) && blockIdx.x == 0 && (warp == kProducerWarp || warp == kIssuerWarp || warp == kSplit0 || warp == kEpiW0);
    const long long t_role0 = prof ? clock64() : 0;

    if (warp == kProducerWarp) {
        // ------------------------------------------------------------ TMA producer of the activation boxes (its own
        // in-order stream: a box can be requested as soon as its slot frees up, independent of the weight ring)
        Ring rs(HS);
        for (int s = begin; s < end;) {
            const Group gr = decode(prm, s, end);
            const Problem &pr = prm.p[gr.z];
            const CUtensorMap *map = &prm.in_map[gr.z];
            const int mask = prm.tap_mask[gr.z];
            for (int kc = gr.kc0; kc < gr.kc1; ++kc) {
                if (halo) {
                    const int x0 = (gr.t0 % prm.tiles_x) * kTW, y0 = (gr.t0 / prm.tiles_x) * kTH;
                    wait_p(src_empty(rs.idx), rs.phase ^ 1, 0, prof);
                    if (elect_one()) {
                        if (prm.dbg & 512) {                      // timing experiment: no activation traffic
                            mbar_arrive(src_full(rs.idx));
                        } else {
                            mbar_expect_tx(src_full(rs.idx), (uint32_t)prm.box_bytes);
                            tma_load_3d(src0 + (uint32_t)rs.idx * prm.src_bytes, map, src_full(rs.idx), kc * kBK, x0 - prm.hd, y0 - prm.hd);
                        }
                    }
                    __syncwarp();
                    rs.next();
                    continue;
                }
                for (int tap = 0; tap < prm.taps; ++tap) {
                    if (!((mask >> tap) & 1)) continue;          // dead tap: no weights, no boxes, no MMAs
                    const int dx = prm.taps == 9 ? (tap % 3 - 1) * pr.dil : 0, dy = prm.taps == 9 ? (tap / 3 - 1) * pr.dil : 0;
                    for (int t = 0; t < gr.n; ++t) {
                        const int tt = gr.t0 + t;
                        const int x0 = (tt % prm.tiles_x) * kTW, y0 = (tt / prm.tiles_x) * kTH;
                        wait_p(src_empty(rs.idx), rs.phase ^ 1, 0, prof);
                        if (elect_one()) {
                            mbar_expect_tx(src_full(rs.idx), (uint32_t)prm.box_bytes);
                            tma_load_3d(src0 + (uint32_t)rs.idx * prm.src_bytes, map, src_full(rs.idx), kc * kBK, x0 + dx, y0 + dy);
                        }
                        __syncwarp();
                        rs.next();
                    }
                }
            }
            s += gr.n;
        }
    } else if (warp == kProducerBWarp) {
        // ------------------------------------------------------------ producer of the weight stages
        Ring rb(BS);
        for (int s = begin; s < end;) {
            const Group gr = decode(prm, s, end);
            const uint8_t *wbase = reinterpret_cast<const uint8_t *>(prm.p[gr.z].weights) + (size_t)gr.g * prm.taps * prm.nkc * 2u * b_bytes;
            const int mask = prm.tap_mask[gr.z];
            for (int kc = gr.kc0; kc < gr.kc1; ++kc)
                for (int tap = 0; tap < prm.taps; ++tap) {
                    if (!((mask >> tap) & 1)) continue;
                    wait_p(b_empty(rb.idx), rb.phase ^ 1, 1, prof);
                    if (elect_one()) {
                        if (prm.dbg & 256) {                      // timing experiment: no weight traffic
                            mbar_arrive(b_full(rb.idx));
                        } else {
                            mbar_expect_tx(b_full(rb.idx), 2u * b_bytes);
                            bulk_load(bst0 + (uint32_t)rb.idx * 2u * b_bytes, wbase + (size_t)(tap * prm.nkc + kc) * 2u * b_bytes, 2u * b_bytes,
                                      b_full(rb.idx));
                        }
                    }
                    __syncwarp();
                    rb.next();
                }
            s += gr.n;
        }
    } else if (warp >= kIssuerWarp) {
        // ------------------------------------------------------------ MMA issuers (warp-uniform loops, elected lane issues).
        // A lone warp issues one dependent instruction every ~6 cycles (ncu: the issuer of the first version ran
        // 215 instructions per tap at 8 cycles each and the tensor pipe starved), so (a) the stream is kept short --
        // descriptors are 32-bit words advanced by small adds (the high word: group stride, version, swizzle mode is
        // constant, the low word is (address >> 4) | LBO), tap offsets advance incrementally, K steps are straight
        // line code -- and (b) NI issuer warps share the work: issuer p owns tiles p, p + NI, ... of every group.
        const int p = warp - kIssuerWarp, NI = prm.ni;
        const uint32_t idw = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((prm.fused ? 2 * npad : npad) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idn = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t AH = (((uint32_t)prm.bwid * 128u) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t BH = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_base = (src0 >> 4) | (1u << 16), b_base = (bst0 >> 4) | (1u << 16);
        const uint32_t lo_delta = (lo0 - src0) >> 4;             // raw box -> its lo copy
        const uint32_t a_step = (uint32_t)prm.src_bytes >> 4, b_step = (2u * b_bytes) >> 4, wl_delta = b_bytes >> 4;
        const bool fused = prm.fused != 0, no_mma = (prm.dbg & 16) != 0, fast = prm.fast != 0;
        const uint32_t acc_cols = (uint32_t)prm.acc_cols;
        Ring rs(HS), rb(BS), rc(NACC);
        if (p < NI)
        for (int s = begin; s < end; rc.next()) {
            const Group gr = decode(prm, s, end);
            const int buf = rc.idx;
            wait_p(acc_empty(buf), rc.phase ^ 1, 3, prof);
            tc_fence_after();
            const int mask = prm.tap_mask[gr.z];
            const uint32_t acc0 = tmem + (uint32_t)(buf * MT) * acc_cols;
            uint32_t accum = 0;                                  // 0 only for the first K step of the group
            for (int kc = gr.kc0; kc < gr.kc1; ++kc) {
                int ksteps = (prm.cin - kc * kBK + 7) >> 3;      // K steps of 8 channels that hold real channels
                if (ksteps > kBK / 8) ksteps = kBK / 8;
                if (halo) {
                    const int slot = rs.idx;
                    wait_p(lo_full(slot), rs.phase, 5, prof);
                    tc_fence_after();
                    // tap (ty, tx) reads the box window that starts (hd + (ty-1) d) rows down, (hd + (tx-1) d) pixels right
                    const int dil = prm.p[gr.z].dil;
                    const uint32_t colstep = (uint32_t)dil * 8u, rowstep = (uint32_t)(dil * prm.bwid) * 8u;
                    uint32_t a_row = a_base + (uint32_t)slot * a_step + (uint32_t)((prm.hd - dil) * (prm.bwid + 1)) * 8u + (uint32_t)(p * kTW) * 8u;
                    int tap = 0;
                    const int ntap_y = prm.taps == 9 ? 3 : 1;
                    for (int ty = 0; ty < ntap_y; ++ty, a_row += rowstep) {
                        uint32_t a = a_row;
                        for (int tx = 0; tx < ntap_y; ++tx, ++tap, a += colstep) {
                            if (!((mask >> tap) & 1)) continue;
                            const int bs = rb.idx;
                            wait_p(b_full(bs), rb.phase, 4, prof);
                            const uint32_t b = b_base + (uint32_t)bs * b_step;
                            if (elect_one()) {
                                if (!no_mma) {
                                    uint32_t at = prm.taps == 9 ? a : a_base + (uint32_t)slot * a_step + (uint32_t)(p * kTW) * 8u;
                                    uint32_t acc = acc0 + (uint32_t)p * acc_cols;
                                    for (int t = p; t < gr.n; t += NI, at += (uint32_t)(NI * kTW) * 8u, acc += (uint32_t)NI * acc_cols)
                                        issue_tile(fused, ksteps, acc, at, lo_delta, b, wl_delta, AH, BH, idw, idn, accum, fast);
                                }
                                umma_commit(b_empty(bs));
                            }
                            __syncwarp();
                            accum = 1;
                            rb.next();
                        }
                    }
                    if (elect_one()) umma_commit(src_empty(slot));
                    __syncwarp();
                    rs.next();
                } else {
                    for (int tap = 0; tap < prm.taps; ++tap) {
                        if (!((mask >> tap) & 1)) continue;
                        const int bs = rb.idx;
                        wait_p(b_full(bs), rb.phase, 4, prof);
                        const uint32_t b = b_base + (uint32_t)bs * b_step;
                        for (int t = 0; t < gr.n; ++t, rs.next()) {
                            if (t % NI != p) continue;           // every issuer walks the whole ring, acts on its own tiles
                            const int slot = rs.idx;
                            wait_p(lo_full(slot), rs.phase, 5, prof);
                            tc_fence_after();
                            if (elect_one()) {
                                if (!no_mma)
                                    issue_tile(fused, ksteps, acc0 + (uint32_t)t * acc_cols, a_base + (uint32_t)slot * a_step, lo_delta, b, wl_delta, AH, BH, idw, idn, accum, fast);
                                umma_commit(src_empty(slot));
                            }
                            __syncwarp();
                        }
                        if (elect_one()) umma_commit(b_empty(bs));
                        __syncwarp();
                        accum = 1;
                        rb.next();
                    }
                }
            }
            if (elect_one()) umma_commit(acc_full(buf));
            __syncwarp();
            s += gr.n;
        }
    } else if (warp < kEpiW0) {
        // ------------------------------------------------------------ lo pass: box -> lo = x - tf32(x), same layout
        const int st = threadIdx.x - kSplit0 * 32;              // 0..255
        const int chunks = prm.box_bytes >> 4;
        Ring rs(HS);
        for (int s = begin; s < end;) {
            const Group gr = decode(prm, s, end);
            int boxes;
            if (halo) boxes = gr.kc1 - gr.kc0;
            else boxes = (gr.kc1 - gr.kc0) * __popc(prm.tap_mask[gr.z]) * gr.n;
            for (int b = 0; b < boxes; ++b, rs.next()) {
                const int slot = rs.idx;
                wait_p(src_full(slot), rs.phase, 7, prof);
                const uint4 *src = reinterpret_cast<const uint4 *>(smem + (size_t)slot * prm.src_bytes);
                uint4 *dst = reinterpret_cast<uint4 *>(smem + (size_t)(HS + slot) * prm.src_bytes);
                for (int i = st; i < (((prm.dbg & 32) || prm.fast) ? 0 : chunks); i += kSplitThreads) {
                    const uint4 x = src[i];
                    uint4 l;
                    l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(x.x & 0xFFFFE000u));
                    l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(x.y & 0xFFFFE000u));
                    l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(x.z & 0xFFFFE000u));
                    l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(x.w & 0xFFFFE000u));
                    dst[i] = l;
                }
                fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core's reads
                mbar_arrive(lo_full(slot));
            }
            s += gr.n;
        }
    } else {
        // ------------------------------------------------------------ epilogue: warp pair (q, half) owns TMEM lanes
        // 32q..32q+31 and the 16-column chunks with index = half (mod 2)
        const int q = warp & 3, half = (warp - kEpiW0) >> 2;
        const int et = threadIdx.x - kEpiW0 * 32;               // 0..255
        const int m = q * 32 + lane, ty = m / kTW, tx = m % kTW;
        const int nslab = (npad + 31) / 32;
        const int SS_ = prm.stg_slabs;
        int cur_zg = -1;
        Ring rc(NACC);
        for (int s = begin; s < end; rc.next()) {
            const Group gr = decode(prm, s, end);
            const Problem &pr = prm.p[gr.z];
            const int buf = rc.idx;
            const int co_base = gr.g * npad;
            if (cur_zg != gr.z * prm.groups + gr.g) {
                named_bar(1, kEpiThreads);
                if (et < npad) {
                    const int co = co_base + et;
                    s_ss[et] = prm.ksplit > 1 ? make_float2(1.f, 0.f)
                               : co < prm.cout ? make_float2(__ldg(pr.scale + co), __ldg(pr.shift + co)) : make_float2(0.f, 0.f);
                }
                named_bar(1, kEpiThreads);
                cur_zg = gr.z * prm.groups + gr.g;
            }
            float *obase = pr.out + (prm.ksplit > 1 ? (size_t)gr.ks * prm.H * prm.W * prm.cpad : (size_t)0);
            wait_p(acc_full(buf), rc.phase, 13, prof);
            tc_fence_after();
            for (int t = 0; t < gr.n; ++t) {
                const int tt = gr.t0 + t;
                const int tx0 = (tt % prm.tiles_x) * kTW, ty0 = (tt / prm.tiles_x) * kTH;
                const int x = tx0 + tx, y = ty0 + ty;
                const bool live = y < prm.H && x < prm.W;
                const size_t pix = (size_t)y * prm.W + x;
                const float *rrow = (pr.residual && live) ? pr.residual + pix * pr.res_stride + co_base : nullptr;
                const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * MT + t) * prm.acc_cols);
                for (int sl0 = 0; sl0 < nslab; sl0 += SS_) {    // rounds of SS_ staging slabs
                    if (prm.store_mode == 0 && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    if (prm.store_mode != 2) named_bar(1, kEpiThreads);      // the previous round has drained the staging slabs
                    for (int sl = sl0; sl < sl0 + SS_ && sl < nslab; ++sl) {
                        const int n0 = sl * 32 + half * 16;
                        if (n0 >= npad) continue;
                        uint32_t v[16];
                        tmem_ld16(tacc + (uint32_t)n0, v);
                        if (prm.fused) {
                            uint32_t v2[16];
                            tmem_ld16(tacc + (uint32_t)(npad + n0), v2);
                            tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(v2[c]));
                        } else {
                            tmem_ld_wait();
                        }
                        float o[16];
                        const float *rr = rrow ? rrow + n0 : nullptr;
                        const int nleft = prm.cout - co_base - n0;
                        switch (prm.act) {
                            case kRelu: epi_chunk<kRelu>(v, o, s_ss + n0, rr, nleft, prm.slope, prm.out_mul); break;
                            case kLeaky: epi_chunk<kLeaky>(v, o, s_ss + n0, rr, nleft, prm.slope, prm.out_mul); break;
                            case kTanh: epi_chunk<kTanh>(v, o, s_ss + n0, rr, nleft, prm.slope, prm.out_mul); break;
                            case kSigmoid: epi_chunk<kSigmoid>(v, o, s_ss + n0, rr, nleft, prm.slope, prm.out_mul); break;
                            case kSigmoidMul: epi_chunk<kSigmoidMul>(v, o, s_ss + n0, rr, nleft, prm.slope, prm.out_mul); break;
                            default: epi_chunk<kNone>(v, o, s_ss + n0, rr, nleft, prm.slope, prm.out_mul); break;
                        }
                        if (prm.store_mode == 2) {
                            if (live) {
                                float *orow = obase + pix * pr.out_stride + pr.out_coff + co_base;
#pragma unroll
                                for (int c = 0; c < 16; ++c)
                                    if (co_base + n0 + c < prm.cout) orow[n0 + c] = o[c];
                            }
                        } else {
                            uint8_t *slab = stg_ptr + (size_t)(sl - sl0) * kSlabBytes + (size_t)m * 128;
                            const int j0 = (n0 & 16) >> 2;
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
                            for (int sl = sl0; sl < sl0 + SS_ && sl < nslab; ++sl)
                                tma_store_3d(&prm.out_map[gr.z], stg0 + (uint32_t)(sl - sl0) * kSlabBytes, pr.out_coff + co_base + sl * 32, tx0, ty0);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    } else if (prm.store_mode == 1) {
                        // row-contiguous stores out of the staging slabs: consecutive lanes write consecutive channels of one pixel
                        named_bar(1, kEpiThreads);
                        int nco = prm.cout - co_base;
                        if (nco > npad) nco = npad;
                        const int c_lo = sl0 * 32, c_hi = (sl0 + SS_) * 32 < nco ? (sl0 + SS_) * 32 : nco;
                        const int rbeg = q * 32 + half * 16, rend = rbeg + 16;
                        for (int rr = rbeg; rr < rend; ++rr) {
                            const int px = tx0 + rr % kTW, py = ty0 + rr / kTW;
                            if (py >= prm.H || px >= prm.W) continue;
                            float *orow = obase + ((size_t)py * prm.W + px) * pr.out_stride + pr.out_coff + co_base;
                            for (int c = c_lo + lane; c < c_hi; c += 32)
                                orow[c] = *reinterpret_cast<const float *>(stg_ptr + (size_t)((c >> 5) - sl0) * kSlabBytes + (size_t)rr * 128 +
                                                                           ((((c & 31) >> 2) ^ (rr & 7)) << 4) + ((c & 3) << 2));
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
        const int slot = warp == kProducerWarp ? 2 : warp == kIssuerWarp ? 6 : warp == kSplit0 ? 9 : 14;
        atomicAdd((unsigned long long *)&g_ss_prof[slot], (unsigned long long)(clock64() - t_role0));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kIssuerWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace ss
}  // namespace ojdf

using namespace ojdf;

// Debug aid (not part of include/ojdf.h): read and clear the role profile filled by launches with flag 128.
extern "C" int ojdf_conv_ss_profile(long long *out_host32)
{
    long long zero[32] = {0};
    cudaError_t e = cudaMemcpyFromSymbol(out_host32, ss::g_ss_prof, sizeof(zero));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(ss::g_ss_prof, zero, sizeof(zero));
    return (int)e;
}

int ojdf_conv_ss_launch(const ojdf_conv_problem *problems_host, int n_problems, int cin, int cout, int H, int W, int taps, int act,
                        float slope, float out_mul, int npad_req, int flags, float *scratch_dev, size_t scratch_bytes, void *stream)
{
    ss::SsParams prm;
    memset(&prm, 0, sizeof(prm));
    int npad, groups;
    ojdf_tc_layout(cout, npad_req, &npad, &groups);
    prm.H = H; prm.W = W; prm.cin = cin; prm.cout = cout; prm.taps = taps; prm.act = act; prm.npad = npad;
    prm.groups = groups; prm.nprob = n_problems;
    prm.nkc = (cin + tc::kBK - 1) / tc::kBK;
    prm.tiles_x = (W + ss::kTW - 1) / ss::kTW;
    prm.tiles_y = (H + ss::kTH - 1) / ss::kTH;
    prm.slope = slope; prm.out_mul = out_mul;
    prm.fast = (flags & 64) ? 1 : 0;                            // 1xTF32 (~1e-3 relative): BASELINE.json configs[2]'s fast mode
    prm.fused = (npad <= 64 && !prm.fast) ? 1 : 0;
    prm.acc_cols = prm.fused ? 2 * npad : npad;
    // tensor memory: MT tiles x NACC accumulator sets x acc_cols columns <= 512.  Wide channel groups with a long K
    // loop trade the second accumulator set for twice the weight-stage reuse (their epilogue is a small share).
    prm.nacc = 2;
    prm.mt = 512 / (2 * prm.acc_cols);
    if (npad > 64 && prm.nkc >= 4 && !(flags & 2048)) { prm.nacc = 1; prm.mt = 512 / prm.acc_cols; }
    if (prm.mt > 4) prm.mt = 4;
    if (flags & 2) prm.mt = 1;
    int dil = problems_host[0].dilation;
    bool same_dil = true;
    for (int i = 0; i < n_problems; ++i) same_dil = same_dil && problems_host[i].dilation == dil;
    const int budget = 227 * 1024 - 1024 - 2048;
    const int nslab = (npad + 31) / 32;
    const int b_stage = 2 * npad * 128;
    int stg_slabs = nslab;
    auto src_bytes_for = [&](int bw, int bh) { return (bw * bh * 128 + 1023) / 1024 * 1024; };
    // halo mode: one box of (16 + 2d) x (8 MT + 2d) pixels per K chunk serves all nine taps and all MT tiles
    prm.halo = 0; prm.hd = 0;
    if (taps == 9 && same_dil && !(flags & 4)) {
        for (int mt = prm.mt; mt >= 1; --mt) {
            const int bw = ss::kTW * mt + 2 * dil, bh = ss::kTH + 2 * dil;
            if (bw > 256 || bh > 256) continue;
            const int sb = src_bytes_for(bw, bh);
            int slabs = nslab;
            while (slabs > 1 && 4 * sb + 3 * b_stage + slabs * tc::kSlabBytes > budget) slabs = (slabs + 1) / 2;
            if (4 * sb + 3 * b_stage + slabs * tc::kSlabBytes <= budget) {
                prm.halo = 1; prm.hd = dil; prm.mt = mt; stg_slabs = slabs;
                break;
            }
        }
    }
    // Measured on B200 (tools/tc_probe.py): this kernel wins for 3x3 convolutions whose halo box fits in shared memory
    // (FusionNet's 19-channel blocks, AdapNet++'s 60x80 / 30x40 layers); 1x1 convolutions, huge dilations and the
    // split-K layers run faster on the tensor-memory kernel.  flag 32768 forces this kernel.
    if (!(flags & 32768)) {
        const long long items0 = (long long)n_problems * groups * prm.tiles_x * prm.tiles_y;
        const bool would_split = scratch_dev && !(flags & 4096) && items0 * 2 <= tc::sm_count() && prm.nkc >= 8;
        if (!prm.halo || would_split) return OJDF_SS_DECLINED;
    }
    if (prm.mt > prm.tiles_x * prm.tiles_y) prm.mt = prm.tiles_x * prm.tiles_y;
    if (prm.halo && prm.mt > prm.tiles_x) prm.mt = prm.tiles_x;
    prm.bwid = prm.halo ? ss::kTW * prm.mt + 2 * prm.hd : ss::kTW;
    const int bhid = prm.halo ? ss::kTH + 2 * prm.hd : ss::kTH;
    prm.box_bytes = prm.bwid * bhid * 128;
    prm.src_bytes = src_bytes_for(prm.bwid, bhid);
    int hs = 2, bs = 3;
    while (stg_slabs > 1 && 2 * hs * prm.src_bytes + bs * b_stage + stg_slabs * tc::kSlabBytes > budget) stg_slabs = (stg_slabs + 1) / 2;
    if (2 * hs * prm.src_bytes + bs * b_stage + stg_slabs * tc::kSlabBytes > budget) bs = 2;
    if (2 * hs * prm.src_bytes + bs * b_stage + stg_slabs * tc::kSlabBytes > budget) return OJDF_ERR_BADARG;
    for (bool grew = true; grew;) {                             // grow the rings while they fit
        grew = false;
        const int hs_max = prm.halo ? 2 : tc::kMaxSrc;
        if (bs < tc::kMaxB && 2 * hs * prm.src_bytes + (bs + 1) * b_stage + stg_slabs * tc::kSlabBytes <= budget) { ++bs; grew = true; }
        if (hs < hs_max && 2 * (hs + 1) * prm.src_bytes + bs * b_stage + stg_slabs * tc::kSlabBytes <= budget) { ++hs; grew = true; }
    }
    prm.src_stages = hs; prm.b_stages = bs; prm.stg_slabs = stg_slabs;
    prm.ni = (flags & 16384) ? 1 : (prm.mt >= 2 ? 2 : 1);
    prm.store_mode = (flags & 8) ? 2 : 0;
    prm.dbg = flags & (16 | 32 | 128 | 256 | 512);
    for (int i = 0; i < n_problems; ++i) {
        const ojdf_conv_problem &q = problems_host[i];
        if (!q.in_dev || !q.weights_dev || !q.scale_dev || !q.shift_dev || !q.out_dev || (q.in_stride & 3) || q.in_stride < cin ||
            ((uintptr_t)q.in_dev & 15) || ((uintptr_t)q.weights_dev & 15) || q.out_stride < q.out_coffset + cout ||
            q.out_coffset < 0 || q.dilation < 1 || (q.residual_dev && q.residual_stride < cout))
            return OJDF_ERR_BADARG;
        if (prm.store_mode == 0 && ((q.out_stride & 3) || ((uintptr_t)q.out_dev & 15) || (q.out_coffset & 3) ||
                                    ((cout & 3) && (!(flags & 1) || q.out_coffset + ((cout + 3) & ~3) > q.out_stride))))
            prm.store_mode = 1;
    }
    for (int i = 0; i < n_problems; ++i) {
        const ojdf_conv_problem &q = problems_host[i];
        const int step = q.in_step > 1 ? q.in_step : 1;
        if (step > 1 && (step != 2 || q.in_width < (W - 1) * step + 1)) return OJDF_ERR_BADARG;
        int r = tc::pixel_map(q.in_dev, cin, q.in_stride * step, H, W, prm.bwid, bhid, &prm.in_map[i],
                              step > 1 ? (long long)q.in_stride * q.in_width * step : 0);
        if (r) return r;
        const int ostep = q.out_step > 1 ? q.out_step : 1;
        if (ostep > 1 && (prm.store_mode != 0 || q.out_width < (W - 1) * ostep + 1)) return OJDF_ERR_BADARG;
        if (prm.store_mode == 0) {
            r = tc::pixel_map(q.out_dev, q.out_coffset + ((cout + 3) & ~3), q.out_stride * ostep, H, W, ss::kTW, ss::kTH,
                              &prm.out_map[i], ostep > 1 ? (long long)q.out_stride * q.out_width * ostep : 0);
            if (r) return r;
        }
        int mask = taps == 9 ? (q.tap_mask ? (q.tap_mask & 511) : 511) : 1;
        if (taps == 9) {
            for (int t = 0; t < 9; ++t)
                if (abs(t / 3 - 1) * q.dilation >= H || abs(t % 3 - 1) * q.dilation >= W) mask &= ~(1 << t);
            if (!mask) mask = 1 << 4;
        }
        prm.tap_mask[i] = mask;
        prm.p[i] = tc::Problem{q.weights_dev, q.scale_dev, q.shift_dev, q.out_dev, q.residual_dev,
                                q.out_stride, q.out_coffset, q.dilation, q.residual_stride};
    }
    const size_t smem = (size_t)2 * hs * prm.src_bytes + (size_t)bs * b_stage + (size_t)stg_slabs * tc::kSlabBytes + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(ss::conv_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        attr = true;
    }
    const long long items = (long long)n_problems * groups * prm.tiles_x * prm.tiles_y;
    prm.ksplit = 1;
    prm.cpad = groups * npad;
    bool strided_out = false;
    for (int i = 0; i < n_problems; ++i) strided_out = strided_out || problems_host[i].out_step > 1;
    if (scratch_dev && !(flags & 4096) && !strided_out && items * 2 <= tc::sm_count() && prm.nkc >= 8) {
        int ks = (int)((tc::sm_count() + items - 1) / items);
        if (ks > prm.nkc / 4) ks = prm.nkc / 4;
        if (ks > 16) ks = 16;
        const size_t per_split = (size_t)n_problems * H * W * prm.cpad * sizeof(float);
        if ((size_t)ks * per_split > scratch_bytes) ks = (int)(scratch_bytes / per_split);
        if (ks >= 2) prm.ksplit = ks;
    }
    tc::SplitReduce red[tc::kMaxBatch];
    if (prm.ksplit > 1) {
        for (int i = 0; i < n_problems; ++i) {
            const ojdf_conv_problem &q = problems_host[i];
            float *part = scratch_dev + (size_t)i * prm.ksplit * H * W * prm.cpad;
            red[i] = tc::SplitReduce{q.scale_dev, q.shift_dev, q.residual_dev, q.out_dev, part, q.out_stride, q.out_coffset,
                                     q.residual_stride};
            prm.p[i].out = part;
            prm.p[i].residual = nullptr;
            prm.p[i].out_stride = prm.cpad;
            prm.p[i].out_coff = 0;
        }
        prm.cout = prm.cpad;
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
        cfg.blockDim = dim3(ss::kSsThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = (flags & 8192) ? 0 : 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const cudaError_t le = cudaLaunchKernelEx(&cfg, ss::conv_ss_kernel, prm);
        if (le != cudaSuccess) { cudaGetLastError(); return (int)le; }
    }
    if (prm.ksplit > 1) {
        const int r = launched(1);
        if (r) return r;
        return launch_split_reduce(red, n_problems, H * W, cout, prm.cpad, prm.ksplit, act, slope, out_mul, (cudaStream_t)stream);
    }
    return launched(1);
}
