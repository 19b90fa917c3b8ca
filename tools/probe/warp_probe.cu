// What does a lone warp pay for the "elected lane does the work" pattern?  Cycles per loop trip (sm_100a).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int WHAT>
__global__ void __launch_bounds__(128, 1) k(int iters, long long *out, int *sink)
{
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_x[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp != 0) return;
    int acc = lane;
    const uint32_t b = smem_u32(&bar[0]);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        if (WHAT == 0) { acc += i; }                                                  // empty loop
        if (WHAT == 1) { if (elect_one()) acc += i; }                                 // elect + predicated add
        if (WHAT == 2) { acc += i; __syncwarp(); }                                    // syncwarp only
        if (WHAT == 3) { if (elect_one()) s_x[i & 31] = acc; __syncwarp(); }          // elect + store + syncwarp
        if (WHAT == 4) { if (lane == 0) s_x[i & 31] = acc; __syncwarp(); }            // lane 0 + store + syncwarp
        if (WHAT == 5) { if (lane == 0) s_x[i & 31] = acc; }                          // lane 0 + store, no syncwarp
        if (WHAT == 6) { if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); __syncwarp(); }
        if (WHAT == 7) { if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
        if (WHAT == 8) {                                                               // arrive by the elected lane, then everybody waits that phase
            if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(i & 1), "r"(0x989680u) : "memory");
        }
        if (WHAT == 9) {                                                               // same, a single thread does both
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(i & 1), "r"(0x989680u) : "memory");
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = acc + s_x[lane];
}
template <int W> void run(const char *name, long long *d, int *sink)
{
    k<W><<<1, 128>>>(64, d, sink); cudaDeviceSynchronize();
    k<W><<<1, 128>>>(4096, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("  %-58s %7.1f %s\n", name, (double)c / 4096.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main()
{
    long long *d; int *sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 512);
    printf("== lone warp, cycles per loop trip\n");
    run<0>("empty loop (add)", d, sink);
    run<1>("elect.sync + predicated add", d, sink);
    run<2>("add + __syncwarp", d, sink);
    run<3>("if (elect) st.shared; __syncwarp", d, sink);
    run<4>("if (lane == 0) st.shared; __syncwarp", d, sink);
    run<5>("if (lane == 0) st.shared", d, sink);
    run<6>("if (elect) mbarrier.arrive; __syncwarp", d, sink);
    run<7>("if (elect) mbarrier.arrive", d, sink);
    run<8>("if (elect) arrive; all lanes try_wait that phase", d, sink);
    run<9>("lane 0: arrive + try_wait that phase", d, sink);
    return 0;
}
