// Chains of 1x1 convolutions kept on chip (modules/model.py:24-52 Pred, :118-141,157-159 VortexPooling's branch-out +
// final convolution): the kernel behind ojdf_conv_chain.
//
// FusionNet ends in eleven 1x1 convolutions (Pred x5: 114 -> 95 -> 95 -> 76 -> ... -> 19 -> 9, each + BatchNorm +
// LeakyReLU) and every VortexPooling block ends in four 19 -> 114 convolutions (+ BN + ReLU) whose concatenation is
// the input of a 456 -> 114 convolution.  Run layer by layer these are memory round trips of 35 MB (Pred) up to
// 285 MB (the concatenated branch outputs) per layer for a few hundred MFLOP each.  Here a 128-pixel tile walks the
// whole chain inside one CTA: the accumulator of a layer (tensor memory) is read by the epilogue warps, scaled /
// shifted / activated, split into its tf32 hi / lo halves and written BACK to tensor memory as the A operand of the
// next layer's tcgen05.mma -- activations never leave the SM between layers.  A step is one GEMM
//     D[acc] (+)= A . W^T     A: a TMA box of a global pixel-major buffer (shared memory, 32-channel chunks)
//                                or the previous step's activated output (tensor memory)
// followed by nothing (a later step adds to the same accumulator: sum over the vortex branches = the concatenation
// folded into the final convolution), by "activate -> next A", or by the output epilogue (TMA store / plain stores).
// Same numerics as the layer kernels: kind::tf32, x = hi + lo, hi*hi + lo*hi + hi*lo, fp32 accumulation.
//
// Tensor memory (512 columns): D0 = [0,128), D1 = [128,256), A_hi = [256,384), A_lo = [384,512): one tile in flight
// per CTA.  When consecutive steps use different accumulators the issuer starts the next step's MMAs on K chunk kc
// (32 columns of A) as soon as the epilogue has written that chunk (four a_ready barriers), so the tensor pipe works
// on layer l+1 while the epilogue warps are still converting layer l; otherwise the two alternate.  Warps: 0 = TMA producer of the input boxes, 1 = producer
// of the weight stages (one stage = [W_hi | W_lo] of one 32-channel K chunk, streamed from L2 for every tile), 2 = MMA
// issuer (owns the TMEM allocation), 4..11 = epilogue / hi-lo split, 12..15 = lo pass over the input boxes.
#include <cstring>

#include "ojdf_tc_common.cuh"

namespace ojdf {
namespace chain {

using namespace ojdf::tc;

constexpr int kMaxSteps = 12, kMaxIn = 4, kMaxZ = 2;
constexpr int kCThreads = 32 * 16;
constexpr int kEpiWarp0 = 4, kLoWarp0 = 12, kLoThreads = 128;
constexpr int kIS = 2, kBS = 3;                               // input-box and weight-stage ring depths
constexpr uint32_t kBoxBytes = 128 * 128;                     // 8 x 16 pixels x 32 channels fp32
constexpr uint32_t kInSlot = 2 * kBoxBytes;                   // raw box (= hi operand) + its lo copy
constexpr uint32_t kBSlot = 2 * 128 * 128;                    // [W_hi | W_lo] at npad = 128
constexpr int kStgSlabs = 2;
constexpr uint32_t kColD = 0, kColAhi = 256, kColAlo = 384;

struct Step {
    const float *w[kMaxZ], *scale[kMaxZ], *shift[kMaxZ];
    int src;                 // >= 0: global input `src` through its tensor map; -1: the activated output of the previous epilogue
    int cin, nkc, npad, cout;
    int acc, fresh;          // accumulator (0 / 1); fresh = 1: the first MMA overwrites it
    int epi;                 // 0: none, 1: activation -> next A, 2: output
    int act;
    float slope;
};
struct ChainParams {
    CUtensorMap in_map[kMaxIn][kMaxZ];
    CUtensorMap out_map[kMaxZ];
    Step st[kMaxSteps];
    float *out[kMaxZ];
    int out_coff[kMaxZ];
    int nsteps, nz, H, W, tiles_x, tiles_y, out_stride, out_cout, store_mode, fast;
    float out_mul;
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
template <int ACT>
__device__ __forceinline__ void chunk_act(const uint32_t (&v)[16], float (&o)[16], const float2 *ss, float slope, float out_mul)
{
    epi_chunk<ACT>(v, o, ss, nullptr, 16, slope, out_mul);
}
__device__ __forceinline__ void chunk_any(int act, const uint32_t (&v)[16], float (&o)[16], const float2 *ss, float slope, float out_mul)
{
    switch (act) {                                              // uniform branch; the element loops are branch-free
        case kRelu: chunk_act<kRelu>(v, o, ss, slope, out_mul); break;
        case kLeaky: chunk_act<kLeaky>(v, o, ss, slope, out_mul); break;
        case kTanh: chunk_act<kTanh>(v, o, ss, slope, out_mul); break;
        case kSigmoid: chunk_act<kSigmoid>(v, o, ss, slope, out_mul); break;
        default: chunk_act<kNone>(v, o, ss, slope, out_mul); break;
    }
}

__global__ void __launch_bounds__(kCThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams prm)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bars[3 * kIS + 2 * kBS + 2 + 4];
    __shared__ uint32_t s_tmem;
    __shared__ float2 s_ss[128];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t in0 = base;                                   // input slots: [raw box | lo copy]
    const uint32_t bst0 = in0 + kIS * kInSlot;                   // weight stages
    const uint32_t stg0 = bst0 + kBS * kBSlot;                   // output staging slabs
    uint8_t *stg_ptr = smem + (size_t)kIS * kInSlot + (size_t)kBS * kBSlot;
    const uint32_t bar0 = smem_u32(s_bars);
    auto src_full = [&](int s) { return bar0 + 8u * s; };
    auto src_empty = [&](int s) { return bar0 + 8u * (kIS + s); };
    auto lo_full = [&](int s) { return bar0 + 8u * (2 * kIS + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (3 * kIS + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (3 * kIS + kBS + s); };
    const uint32_t acc_full = bar0 + 8u * (3 * kIS + 2 * kBS), epi_done = acc_full + 8u;
    auto a_ready = [&](int kc) { return epi_done + 8u + 8u * kc; };   // K chunk kc (32 columns) of the next step's A operand is in tensor memory

    const int tiles = prm.tiles_x * prm.tiles_y;
    const int total = prm.nz * tiles;
    const int begin = (int)((long long)total * blockIdx.x / gridDim.x);
    const int end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kIS; ++s) { mbar_init(src_full(s), 1); mbar_init(src_empty(s), 1); mbar_init(lo_full(s), kLoThreads); }
        for (int s = 0; s < kBS; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        mbar_init(acc_full, 1);
        mbar_init(epi_done, kEpiThreads);
        for (int kc = 0; kc < 4; ++kc) mbar_init(a_ready(kc), kEpiThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 96)
        for (int z = 0; z < prm.nz; ++z) {
            for (int s = 0; s < prm.nsteps; ++s)
                if (prm.st[s].src >= 0) prefetch_map(&prm.in_map[prm.st[s].src][z]);
            if (prm.store_mode == 0) prefetch_map(&prm.out_map[z]);
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    asm volatile("griddepcontrol.wait;" ::: "memory");         // from here on this grid reads what the previous kernels wrote

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer of the input boxes
        Ring rs(kIS);
        for (int it = begin; it < end; ++it) {
            const int z = it / tiles, t = it - z * tiles;
            const int col = t / prm.tiles_y, row = t - col * prm.tiles_y;
            for (int s = 0; s < prm.nsteps; ++s) {
                const Step &S = prm.st[s];
                if (S.src < 0) continue;
                for (int kc = 0; kc < S.nkc; ++kc) {
                    mbar_wait(src_empty(rs.idx), rs.phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(src_full(rs.idx), kBoxBytes);
                        tma_load_3d(in0 + (uint32_t)rs.idx * kInSlot, &prm.in_map[S.src][z], src_full(rs.idx), kc * kBK, col * kBW, row * kBH);
                    }
                    __syncwarp();
                    rs.next();
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ producer of the weight stages
        Ring rb(kBS);
        for (int it = begin; it < end; ++it) {
            const int z = it / tiles;
            for (int s = 0; s < prm.nsteps; ++s) {
                const Step &S = prm.st[s];
                const uint32_t bytes = 2u * (uint32_t)S.npad * 128u;
                const uint8_t *wbase = reinterpret_cast<const uint8_t *>(S.w[z]);
                for (int kc = 0; kc < S.nkc; ++kc) {
                    mbar_wait(b_empty(rb.idx), rb.phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(b_full(rb.idx), bytes);
                        bulk_load(bst0 + (uint32_t)rb.idx * kBSlot, wbase + (size_t)kc * bytes, bytes, b_full(rb.idx));
                    }
                    __syncwarp();
                    rb.next();
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
        Ring rs(kIS), rb(kBS);
        uint32_t epi_phase = 0, a_phase = 0;
        bool epi_pending = false;                                // an "activation -> next A" epilogue is running: this step's K chunks
        int a_chunks = 0;                                        // start as soon as theirs is written (a_ready), not after all of it
        for (int it = begin; it < end; ++it) {
            for (int s = 0; s < prm.nsteps; ++s) {
                const Step &S = prm.st[s];
                if (epi_pending && S.src >= 0) {                 // a shared-memory step does not read A, but it may overwrite the
                    mbar_wait(epi_done, epi_phase);              // accumulator that epilogue is still reading
                    tc_fence_after();
                    epi_phase ^= 1u; a_phase ^= 1u; epi_pending = false;
                }
                const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(S.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint32_t acc = tmem + kColD + (uint32_t)S.acc * 128u;
                const uint32_t b_lo_off = (uint32_t)S.npad * 128u;
                for (int kc = 0; kc < S.nkc; ++kc) {
                    int ksteps = (S.cin - kc * kBK + 7) >> 3;
                    if (ksteps > kBK / 8) ksteps = kBK / 8;
                    const int bs = rb.idx;
                    mbar_wait(b_full(bs), rb.phase);
                    const uint32_t sb = bst0 + (uint32_t)bs * kBSlot;
                    const uint64_t b_hi = smem_desc(sb), b_lo = smem_desc(sb + b_lo_off);
                    const uint32_t fresh = (S.fresh && kc == 0) ? 0u : 1u;
                    if (S.src >= 0) {
                        const int slot = rs.idx;
                        mbar_wait(lo_full(slot), rs.phase);
                        tc_fence_after();
                        const uint64_t a_hi = smem_desc(in0 + (uint32_t)slot * kInSlot), a_lo = smem_desc(in0 + (uint32_t)slot * kInSlot + kBoxBytes);
                        if (elect_one()) {
                            for (int k = 0; k < ksteps; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);   // +32 bytes along K inside the swizzle atom
                                if (prm.fast) {
                                    umma_tf32_ss(acc, a_hi + ko, b_hi + ko, idesc, k ? 1u : fresh);
                                } else {
                                    umma_tf32_ss(acc, a_lo + ko, b_hi + ko, idesc, k ? 1u : fresh);
                                    umma_tf32_ss(acc, a_hi + ko, b_lo + ko, idesc, 1u);
                                    umma_tf32_ss(acc, a_hi + ko, b_hi + ko, idesc, 1u);
                                }
                            }
                            umma_commit(src_empty(slot));
                        }
                        __syncwarp();
                        rs.next();
                    } else {
                        if (epi_pending) {
                            if (kc < a_chunks) mbar_wait(a_ready(kc), a_phase);
                            else mbar_wait(epi_done, epi_phase ^ 0u);   // (never taken: a step reads no more chunks than were written)
                            tc_fence_after();
                        }
                        const uint32_t a_hi = tmem + kColAhi + (uint32_t)(kc * kBK), a_lo = tmem + kColAlo + (uint32_t)(kc * kBK);
                        if (elect_one()) {
                            for (int k = 0; k < ksteps; ++k) {
                                const uint64_t ko = (uint64_t)(k * 2);
                                if (prm.fast) {
                                    umma_tf32_ts(acc, a_hi + k * 8, b_hi + ko, idesc, k ? 1u : fresh);
                                } else {
                                    umma_tf32_ts(acc, a_lo + k * 8, b_hi + ko, idesc, k ? 1u : fresh);
                                    umma_tf32_ts(acc, a_hi + k * 8, b_lo + ko, idesc, 1u);
                                    umma_tf32_ts(acc, a_hi + k * 8, b_hi + ko, idesc, 1u);
                                }
                            }
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(b_empty(bs));
                    __syncwarp();
                    rb.next();
                }
                if (epi_pending) {                                // all K chunks of this step were issued: that epilogue is complete
                    mbar_wait(epi_done, epi_phase);
                    tc_fence_after();
                    epi_phase ^= 1u; a_phase ^= 1u; epi_pending = false;
                }
                if (S.epi) {
                    if (elect_one()) umma_commit(acc_full);
                    __syncwarp();
                    if (S.epi == 1 && s + 1 < prm.nsteps && prm.st[s + 1].src < 0 && prm.st[s + 1].acc != S.acc) {
                        // the next step reads this epilogue's output chunk by chunk into the OTHER accumulator
                        epi_pending = true;
                        a_chunks = (S.npad + 31) / 32;
                    } else {
                        // the epilogue warps read this accumulator (and may rewrite A): nothing is issued until they are done
                        mbar_wait(epi_done, epi_phase);
                        tc_fence_after();
                        epi_phase ^= 1u;
                        if (S.epi == 1) a_phase ^= 1u;           // the four a_ready barriers complete one phase per "-> next A" epilogue
                    }
                }
            }
        }
    } else if (warp >= kLoWarp0) {
        // ------------------------------------------------------------ lo pass: box -> lo = x - tf32(x), same layout
        const int lt = threadIdx.x - kLoWarp0 * 32;             // 0..127
        Ring rs(kIS);
        for (int it = begin; it < end; ++it) {
            for (int s = 0; s < prm.nsteps; ++s) {
                const Step &S = prm.st[s];
                if (S.src < 0) continue;
                for (int kc = 0; kc < S.nkc; ++kc, rs.next()) {
                    const int slot = rs.idx;
                    mbar_wait(src_full(slot), rs.phase);
                    const uint4 *src = reinterpret_cast<const uint4 *>(smem + (size_t)slot * kInSlot);
                    uint4 *dst = reinterpret_cast<uint4 *>(smem + (size_t)slot * kInSlot + kBoxBytes);
                    if (!prm.fast) {
#pragma unroll
                        for (int i = 0; i < (int)(kBoxBytes / 16) / kLoThreads; ++i) {
                            const uint4 x = src[lt + i * kLoThreads];
                            uint4 l;
                            l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(x.x & 0xFFFFE000u));
                            l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(x.y & 0xFFFFE000u));
                            l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(x.z & 0xFFFFE000u));
                            l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(x.w & 0xFFFFE000u));
                            dst[lt + i * kLoThreads] = l;
                        }
                    }
                    fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core's reads
                    mbar_arrive(lo_full(slot));
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ------------------------------------------------------------ epilogue / split: warp pair (q, half) owns TMEM lanes
        // 32q..32q+31 and the 16-column chunks with index = half (mod 2)
        const int q = warp & 3, half = (warp - kEpiWarp0) >> 2;
        const int et = threadIdx.x - kEpiWarp0 * 32;            // 0..255
        const int m = q * 32 + lane, ty = m / kBW, tx = m % kBW;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t acc_phase = 0;
        for (int it = begin; it < end; ++it) {
            const int z = it / tiles, t = it - z * tiles;
            const int col = t / prm.tiles_y, row = t - col * prm.tiles_y;
            for (int s = 0; s < prm.nsteps; ++s) {
                const Step &S = prm.st[s];
                if (!S.epi) continue;
                const int npad = S.npad;
                named_bar(1, kEpiThreads);                       // the previous epilogue is done with s_ss
                if (et < npad) s_ss[et] = et < S.cout ? make_float2(__ldg(S.scale[z] + et), __ldg(S.shift[z] + et)) : make_float2(0.f, 0.f);
                named_bar(1, kEpiThreads);
                mbar_wait(acc_full, acc_phase);
                tc_fence_after();
                acc_phase ^= 1u;
                const uint32_t tacc = lane_base + kColD + (uint32_t)S.acc * 128u;
                if (S.epi == 1) {
                    // all accumulator chunks of this thread are requested at once; then chunk by chunk: activation, hi / lo
                    // split, store, and the K chunk is announced (the issuer starts the next layer's MMAs on it)
                    uint32_t v[4][16];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (half * 16 + 32 * i < npad) tmem_ld16(tacc + (uint32_t)(half * 16 + 32 * i), v[i]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int n0 = half * 16 + 32 * i;
                        if (32 * i >= npad) break;                       // (uniform: both halves of a 32-column chunk arrive together)
                        if (n0 < npad) {
                            float o[16];
                            chunk_any(S.act, v[i], o, s_ss + n0, S.slope, 1.0f);
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const uint32_t x = __float_as_uint(o[c]);
                                hi[c] = x & 0xFFFFE000u;
                                lo[c] = __float_as_uint(o[c] - __uint_as_float(hi[c]));
                            }
                            tmem_st16(lane_base + kColAhi + (uint32_t)n0, hi);
                            tmem_st16(lane_base + kColAlo + (uint32_t)n0, lo);
                            tmem_st_wait();
                        }
                        tc_fence_before();
                        mbar_arrive(a_ready(i));
                    }
                    for (int i = (npad + 31) / 32; i < 4; ++i) mbar_arrive(a_ready(i));   // unused chunks: keep the four phases in step
                    mbar_arrive(epi_done);
                } else {
                    const int x = col * kBW + tx, y = row * kBH + ty;
                    const bool live = y < prm.H && x < prm.W;
                    const size_t pix = (size_t)y * prm.W + x;
                    const int nslab = (npad + 31) / 32;
                    for (int sl0 = 0; sl0 < nslab; sl0 += kStgSlabs) {       // rounds of kStgSlabs staging slabs
                        if (prm.store_mode == 0) {
                            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            named_bar(1, kEpiThreads);           // the previous round has drained the staging slabs
                        }
                        for (int sl = sl0; sl < sl0 + kStgSlabs && sl < nslab; ++sl) {
                            const int n0 = sl * 32 + half * 16;
                            if (n0 >= npad) continue;
                            uint32_t v[16];
                            tmem_ld16(tacc + (uint32_t)n0, v);
                            tmem_ld_wait();
                            float o[16];
                            chunk_any(S.act, v, o, s_ss + n0, S.slope, prm.out_mul);
                            if (prm.store_mode == 2) {
                                if (live) {
                                    float *orow = prm.out[z] + pix * prm.out_stride + prm.out_coff[z];
#pragma unroll
                                    for (int c = 0; c < 16; ++c)
                                        if (n0 + c < prm.out_cout) orow[n0 + c] = o[c];
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
                                for (int sl = sl0; sl < sl0 + kStgSlabs && sl < nslab; ++sl)
                                    tma_store_3d(&prm.out_map[z], stg0 + (uint32_t)(sl - sl0) * kSlabBytes, prm.out_coff[z] + sl * 32, col * kBW, row * kBH);
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(epi_done);                       // the accumulator is free; the stores drain behind
                }
            }
        }
        // the staging slabs must have been READ before the CTA (and its shared memory) goes away; the global writes of the
        // bulk stores complete on their own before the grid does
        if (prm.store_mode == 0 && et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace chain
}  // namespace ojdf

using namespace ojdf;

extern "C" int ojdf_conv_chain(const ojdf_chain_input *inputs_host, int n_inputs, const ojdf_chain_step *steps_host, int n_steps,
                               int n_problems, int H, int W, float *const *out_dev_host, const int *out_coffset_host, int out_stride,
                               float out_mul, int flags, void *stream)
{
    if (!inputs_host || !steps_host || !out_dev_host || !out_coffset_host || n_inputs < 1 || n_inputs > chain::kMaxIn || n_steps < 1 ||
        n_steps > chain::kMaxSteps || n_problems < 1 || n_problems > chain::kMaxZ || H < 1 || W < 1 || H > 32767 || W > 32767)
        return OJDF_ERR_BADARG;
    chain::ChainParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.nsteps = n_steps; prm.nz = n_problems; prm.H = H; prm.W = W;
    prm.tiles_x = (W + tc::kBW - 1) / tc::kBW;
    prm.tiles_y = (H + tc::kBH - 1) / tc::kBH;
    prm.out_stride = out_stride; prm.out_mul = out_mul;
    prm.fast = (flags & 64) ? 1 : 0;
    int prev_cout = -1;                                         // width of the activation the last "-> next A" epilogue left in TMEM
    bool acc_live[2] = {false, false};
    for (int s = 0; s < n_steps; ++s) {
        const ojdf_chain_step &q = steps_host[s];
        chain::Step &S = prm.st[s];
        if (q.cout < 1 || q.cout > 128 || q.acc < 0 || q.acc > 1 || q.epi < 0 || q.epi > 2 || q.act < 0 || q.act > 4) return OJDF_ERR_BADARG;
        if (q.src >= n_inputs) return OJDF_ERR_BADARG;
        if (q.src < 0) {
            if (prev_cout < 1 || q.cin != prev_cout) return OJDF_ERR_BADARG;       // reads what the previous epilogue wrote
        } else if (q.cin != inputs_host[q.src].cin) {
            return OJDF_ERR_BADARG;
        }
        if (q.cin < 1 || (q.src < 0 && q.cin > 128) || q.cin > 4096) return OJDF_ERR_BADARG;
        if (!q.fresh && !acc_live[q.acc]) return OJDF_ERR_BADARG;              // adding to an accumulator nobody started
        if ((q.epi == 2) != (s == n_steps - 1)) return OJDF_ERR_BADARG;        // the output epilogue is the last step, and only it
        int npad, groups;
        ojdf_tc_layout(q.cout, 0, &npad, &groups);
        S.src = q.src < 0 ? -1 : q.src;
        S.cin = q.cin; S.cout = q.cout; S.npad = npad; S.nkc = (q.cin + tc::kBK - 1) / tc::kBK;
        S.acc = q.acc; S.fresh = q.fresh ? 1 : 0; S.epi = q.epi; S.act = q.act; S.slope = q.slope;
        for (int z = 0; z < n_problems; ++z) {
            if (!q.weights_dev[z] || ((uintptr_t)q.weights_dev[z] & 15) || (q.epi && (!q.scale_dev[z] || !q.shift_dev[z]))) return OJDF_ERR_BADARG;
            S.w[z] = q.weights_dev[z]; S.scale[z] = q.scale_dev[z]; S.shift[z] = q.shift_dev[z];
        }
        acc_live[q.acc] = true;
        if (q.epi == 1) { prev_cout = q.cout; acc_live[q.acc] = false; }
        if (q.epi == 2) prm.out_cout = q.cout;
    }
    // store mode 0: TMA store (16-byte aligned rows / channel offsets; a width that is not a multiple of 4 is rounded up when
    // the caller owns those pad channels: flag 1); 2: per-thread stores
    prm.store_mode = (flags & 8) ? 2 : 0;
    for (int z = 0; z < n_problems; ++z) {
        float *o = out_dev_host[z];
        const int coff = out_coffset_host[z];
        if (!o || coff < 0 || out_stride < coff + prm.out_cout) return OJDF_ERR_BADARG;
        if ((out_stride & 3) || ((uintptr_t)o & 15) || (coff & 3) ||
            ((prm.out_cout & 3) && (!(flags & 1) || coff + ((prm.out_cout + 3) & ~3) > out_stride)))
            prm.store_mode = 2;
        prm.out[z] = o;
        prm.out_coff[z] = coff;
    }
    for (int i = 0; i < n_inputs; ++i) {
        const ojdf_chain_input &q = inputs_host[i];
        if (q.cin < 1 || (q.in_stride & 3) || q.in_stride < q.cin) return OJDF_ERR_BADARG;
        for (int z = 0; z < n_problems; ++z) {
            if (!q.in_dev[z] || ((uintptr_t)q.in_dev[z] & 15)) return OJDF_ERR_BADARG;
            const int r = tc::pixel_map(q.in_dev[z], q.cin, q.in_stride, H, W, tc::kBW, tc::kBH, &prm.in_map[i][z]);
            if (r) return r;
        }
    }
    if (prm.store_mode == 0)
        for (int z = 0; z < n_problems; ++z) {
            const int r = tc::pixel_map(prm.out[z], prm.out_coff[z] + ((prm.out_cout + 3) & ~3), out_stride, H, W, tc::kBW, tc::kBH, &prm.out_map[z]);
            if (r) return r;
        }
    const size_t smem = (size_t)chain::kIS * chain::kInSlot + (size_t)chain::kBS * chain::kBSlot + (size_t)chain::kStgSlabs * tc::kSlabBytes + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(chain::conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        attr = true;
    }
    const long long total = (long long)n_problems * prm.tiles_x * prm.tiles_y;
    int grid = tc::sm_count();
    if (grid > total) grid = (int)total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(chain::kCThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = (flags & 8192) ? 0 : 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, chain::conv_chain_kernel, prm);
    if (le != cudaSuccess) { cudaGetLastError(); return (int)le; }
    return launched(1);
}
