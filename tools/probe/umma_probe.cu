// Micro-probe for tcgen05.mma kind::tf32 on sm_100a (design input for csrc/ojdf_conv_tc.cu):
//   1. cycles per MMA instruction (M = 128, K = 8) as a function of N, with A in tensor memory (TS) or in shared
//      memory (SS), B always a K-major SWIZZLE_128B shared-memory tile;
//   2. whether an SS-mode A descriptor may start at ANY 128-byte row of a swizzled tile (row offset not a multiple
//      of 8), with and without the descriptor's base-offset field -- the "flat shifted window" a tap GEMM wants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (long long spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > 50000000) { printf("probe: mbarrier timeout\n"); __trap(); }
    }
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, int base_off, int sbo = 1024)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(sbo >> 4) << 32;        // SBO: bytes between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
    d |= (uint64_t)(base_off & 7) << 49;    // matrix base offset
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

constexpr int kARows = 640;      // A region: 640 rows of 128 bytes (80 KB)
constexpr int kBRows = 256;      // B region: 256 rows of 128 bytes (32 KB)

// mode 0: timing TS, 1: timing SS, 2: correctness of a shifted SS window (rowoff, use_base_off), 3: as 2 but TS reference
// kspan: k-steps cycled through per MMA (1 = same operands every time, 4 = walk the 4 K=8 slices of the 128-byte rows)
__global__ void __launch_bounds__(128, 1) probe(int mode, int N, int iters, int rowoff, int use_base, int kspan, int nacc, float *out, long long *cycles, int bw)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    float *A = reinterpret_cast<float *>(smem), *B = reinterpret_cast<float *>(smem + kARows * 128);
    const int tid = threadIdx.x, warp = tid >> 5;
    // A[r][k] = (r % 251) + 1 for k == r % 32 ... simple: A[r][k] = float((r * 7 + k * 3) % 509); exact in tf32 (< 1024)
    for (int i = tid; i < kARows * 32; i += 128) {
        const int r = i >> 5, k = i & 31;
        const int chunk = (k >> 2) ^ (r & 7);
        A[r * 32 + chunk * 4 + (k & 3)] = (float)((r * 7 + k * 3) % 509);
    }
    for (int i = tid; i < kBRows * 32; i += 128) {          // B[n][k] = 1 if k == n % 32: D[m][n] = A[m][n % 32]
        const int n = i >> 5, k = i & 31;
        const int chunk = (k >> 2) ^ (n & 7);
        B[n * 32 + chunk * 4 + (k & 3)] = (k == (n & 31)) ? 1.0f : 0.0f;
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // A operand in TMEM (columns 448..479): row m of the window, 32 columns
    {
        const int m = tid;
        for (int c0 = 0; c0 < 32; c0 += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint((float)(((m + rowoff) * 7 + (c0 + j) * 3) % 509));
            tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 448u + c0, v);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_addr = base + (uint32_t)rowoff * 128u;
    const uint64_t adesc = smem_desc(a_addr, use_base ? (int)((a_addr >> 7) & 7) : 0, bw * 128);
    const uint64_t bdesc = smem_desc(base + kARows * 128, 0);
    if (warp == 0) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        long long t0 = 0, t1 = 0;
        if (pred) {
            t0 = clock64();
            if (mode == 3) {
                for (int i = 0; i < iters; ++i) umma_ts(tmem, tmem + 448u + (i & 3) * 8, bdesc + (uint64_t)((i & 3) * 2), idesc, (uint32_t)(i > 0));
            } else if (mode == 2) {
                for (int i = 0; i < iters; ++i) umma_ss(tmem, adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, (uint32_t)(i > 0));
            } else if (mode == 0) {                           // timing: 8 MMAs per trip, operands fixed at compile time
                const uint32_t d1 = tmem + (uint32_t)((nacc - 1) * N);
                for (int i = 0; i < iters; i += 8) {
                    if (kspan == 4) {
                        umma_ts(tmem, tmem + 448u, bdesc, idesc, 1u);          umma_ts(tmem, tmem + 456u, bdesc + 2, idesc, 1u);
                        umma_ts(tmem, tmem + 464u, bdesc + 4, idesc, 1u);      umma_ts(tmem, tmem + 472u, bdesc + 6, idesc, 1u);
                        umma_ts(d1, tmem + 448u, bdesc, idesc, 1u);            umma_ts(d1, tmem + 456u, bdesc + 2, idesc, 1u);
                        umma_ts(d1, tmem + 464u, bdesc + 4, idesc, 1u);        umma_ts(d1, tmem + 472u, bdesc + 6, idesc, 1u);
                    } else {
                        umma_ts(tmem, tmem + 448u, bdesc, idesc, 1u); umma_ts(d1, tmem + 448u, bdesc, idesc, 1u);
                        umma_ts(tmem, tmem + 448u, bdesc, idesc, 1u); umma_ts(d1, tmem + 448u, bdesc, idesc, 1u);
                        umma_ts(tmem, tmem + 448u, bdesc, idesc, 1u); umma_ts(d1, tmem + 448u, bdesc, idesc, 1u);
                        umma_ts(tmem, tmem + 448u, bdesc, idesc, 1u); umma_ts(d1, tmem + 448u, bdesc, idesc, 1u);
                    }
                }
            } else {
                const uint32_t d1 = tmem + (uint32_t)((nacc - 1) * N);
                for (int i = 0; i < iters; i += 8) {
                    if (kspan == 4) {
                        umma_ss(tmem, adesc, bdesc, idesc, 1u);          umma_ss(tmem, adesc + 2, bdesc + 2, idesc, 1u);
                        umma_ss(tmem, adesc + 4, bdesc + 4, idesc, 1u);  umma_ss(tmem, adesc + 6, bdesc + 6, idesc, 1u);
                        umma_ss(d1, adesc, bdesc, idesc, 1u);            umma_ss(d1, adesc + 2, bdesc + 2, idesc, 1u);
                        umma_ss(d1, adesc + 4, bdesc + 4, idesc, 1u);    umma_ss(d1, adesc + 6, bdesc + 6, idesc, 1u);
                    } else {
                        umma_ss(tmem, adesc, bdesc, idesc, 1u); umma_ss(d1, adesc, bdesc, idesc, 1u);
                        umma_ss(tmem, adesc, bdesc, idesc, 1u); umma_ss(d1, adesc, bdesc, idesc, 1u);
                        umma_ss(tmem, adesc, bdesc, idesc, 1u); umma_ss(d1, adesc, bdesc, idesc, 1u);
                        umma_ss(tmem, adesc, bdesc, idesc, 1u); umma_ss(d1, adesc, bdesc, idesc, 1u);
                    }
                }
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        if (pred) { t1 = clock64(); cycles[0] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode >= 2) {                                         // D[m][0..31] -> out
        for (int c0 = 0; c0 < 32; c0 += 8) {
            uint32_t v[8];
            tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int j = 0; j < 8; ++j) out[tid * 32 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}


// ---- cost of the synchronisation primitives as seen by ONE issuing thread (cycles per loop trip)
// what: 0 = tcgen05.commit only; 1 = 4 MMAs (N = 32, SS) + commit; 2 = try_wait on an already completed phase;
//       3 = mbarrier.arrive (self) + try_wait of that phase; 4 = 16 MMAs + commit; 5 = elect.sync + __syncwarp only;
//       6 = commit + try_wait for ITS completion (commit -> mbarrier latency); 7 = 4 MMAs + commit + wait for completion
__global__ void __launch_bounds__(128, 1) sync_probe(int what, int iters, long long *cycles)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t s_tmem;
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = smem_desc(base, 0), bdesc = smem_desc(base + kARows * 128, 0);
    if (warp == 0) {
        const uint32_t b0 = smem_u32(&bars[0]);
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (what == 0) { if (pred) umma_commit(b0 + 8 * (i & 3)); }
            else if (what == 1 || what == 4) {
                if (pred) {
                    for (int r = 0; r < (what == 4 ? 4 : 1); ++r) {
                        umma_ss(tmem, adesc, bdesc, idesc, 1u); umma_ss(tmem, adesc + 2, bdesc + 2, idesc, 1u);
                        umma_ss(tmem, adesc + 4, bdesc + 4, idesc, 1u); umma_ss(tmem, adesc + 6, bdesc + 6, idesc, 1u);
                    }
                    umma_commit(b0 + 8 * (i & 3));
                }
            } else if (what == 2) { mbar_wait(b0 + 32, 1); }              // a fresh barrier reports its preceding phase as complete
            else if (what == 3) {
                if (pred) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0 + 40) : "memory");
                mbar_wait(b0 + 40, phase); phase ^= 1;
            } else if (what == 6 || what == 7) {
                if (pred) {
                    if (what == 7) { umma_ss(tmem, adesc, bdesc, idesc, 1u); umma_ss(tmem, adesc + 2, bdesc + 2, idesc, 1u);
                                     umma_ss(tmem, adesc + 4, bdesc + 4, idesc, 1u); umma_ss(tmem, adesc + 6, bdesc + 6, idesc, 1u); }
                    umma_commit(b0 + 48);
                }
                mbar_wait(b0 + 48, phase); phase ^= 1;
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (tid == 0) cycles[0] = t1 - t0;
        // drain: make sure every commit has landed before the CTA exits
        if (what == 0 || what == 1 || what == 4) { __nanosleep(20000); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 128 * 32 * 4); cudaMalloc(&cyc, 8);
    const size_t smem = (kARows + kBRows) * 128 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    auto run = [&](int mode, int N, int iters, int rowoff, int use_base, int kspan, int nacc, int bw = 8) -> long long {
        probe<<<1, 128, smem>>>(mode, N, iters, rowoff, use_base, kspan, nacc, out, cyc, bw);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); return c;
    };
    printf("== correctness of shifted SS windows (K = 32 as 4 MMAs, N = 32, D[m][n] must equal A[rowoff + m][n])\n");
    std::vector<float> h(128 * 32);
    for (int use_base = 0; use_base < 2; ++use_base)
        for (int rowoff : {0, 8, 1, 2, 3, 5, 7, 9, 13, 34, 77, 333}) {
            run(2, 32, 4, rowoff, use_base, 4, 1);
            cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0, first = -1;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
                const float want = (float)(((m + rowoff) * 7 + n * 3) % 509);
                if (h[m * 32 + n] != want) { if (first < 0) first = m * 32 + n; ++bad; }
            }
            printf("  SS rowoff %3d base_offset %s: %s (%d wrong%s)\n", rowoff, use_base ? "set" : "0  ", bad ? "MISMATCH" : "ok", bad, "");
            if (bad && first >= 0) printf("     first wrong: m %d n %d got %.0f want %.0f\n", first / 32, first % 32, h[first], (float)((((first / 32) + rowoff) * 7 + (first % 32) * 3) % 509));
        }
    printf("== 8-pixel row groups at a stride of bw rows (SBO = bw * 128 bytes): D[m][n] must equal A[rowoff + (m / 8) * bw + m %% 8][n]\n");
    for (int bw : {8, 10, 18, 22, 13, 34})
        for (int rowoff : {0, 1, 5, 19, 40}) {
            run(2, 32, 4, rowoff, 0, 4, 1, bw);
            cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
                const int r = rowoff + (m / 8) * bw + m % 8;
                bad += h[m * 32 + n] != (float)((r * 7 + n * 3) % 509);
            }
            printf("  SS bw %2d rowoff %3d: %s (%d wrong)\n", bw, rowoff, bad ? "MISMATCH" : "ok", bad);
        }
    run(3, 32, 4, 5, 0, 4, 1);
    cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
    { int bad = 0; for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) bad += h[m * 32 + n] != (float)(((m + 5) * 7 + n * 3) % 509);
      printf("  TS reference rowoff 5: %s\n", bad ? "MISMATCH" : "ok"); }
    printf("== cycles per MMA (M 128, K 8, tf32), 2000 MMAs back to back, one issuing thread\n");
    printf("  %-4s %-5s %-6s %-5s %10s\n", "mode", "N", "kspan", "nacc", "cyc/MMA");
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {8, 24, 32, 64, 96, 128, 160, 192, 216, 256})
            for (int kspan : {1, 4})
                for (int nacc : {1, 2}) {
                    if (nacc * N > 440) continue;
                    run(mode, N, 200, 0, 0, kspan, nacc);
                    const long long c = run(mode, N, 4000, 0, 0, kspan, nacc);
                    printf("  %-4s %-5d %-6d %-5d %10.1f\n", mode ? "SS" : "TS", N, kspan, nacc, (double)c / 4000.0);
                }

    printf("== SS cycles per MMA with halo-style windows (N 64 then N 32 alternating like the conv kernel: reported per pair / 2)\n");
    for (int bw : {8, 18, 22})
        for (int rowoff : {0, 19, 37})
            for (int N : {32, 64}) {
                run(1, N, 200, rowoff, 0, 4, 1, bw);
                const long long c = run(1, N, 4000, rowoff, 0, 4, 1, bw);
                printf("  SS N %3d bw %2d rowoff %2d: %6.1f\n", N, bw, rowoff, (double)c / 4000.0);
            }
    printf("== synchronisation primitives, cycles per loop trip (one issuing warp, elect + syncwarp included)\n");
    cudaFuncSetAttribute(sync_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char *names[8] = {"commit only", "4 MMAs (N 32, SS) + commit", "try_wait, phase already complete", "arrive + try_wait", "16 MMAs + commit",
                            "elect + syncwarp only", "commit + wait for its arrival", "4 MMAs + commit + wait for its arrival"};
    for (int what = 0; what < 8; ++what) {
        sync_probe<<<1, 128, smem>>>(what, 64, cyc);
        cudaDeviceSynchronize();
        sync_probe<<<1, 128, smem>>>(what, 512, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  %s: launch failed: %s\n", names[what], cudaGetErrorString(e)); return 1; }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  %-40s %8.1f\n", names[what], (double)c / 512.0);
    }
    return 0;
}
