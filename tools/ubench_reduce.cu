// What the warp-level part of k_track_run's sum reduction costs when NINE warps do it at the same time (as they do behind
// phase 2): three packed sums per thread -> three totals per warp, by
//   redux   3 x __reduce_add_sync                                   (REDUX.SUM, what the kernel does)
//   shfl    butterfly of 5 x 3 SHFL.BFLY + adds
//   mma     ONE mma.sync.m16n8k32.s8 per warp: every thread's six counts (|v| <= 64) as signed bytes in the B fragment, A a
//           constant 0/1 matrix that routes byte position -> component, then two shuffle steps over the four threads of a row
// Reported: clock ticks from a CTA barrier to the next one around the reduction (all warps), minus an empty pair of barriers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_reduce tools/ubench_reduce.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 2000;
constexpr int kWarps = 12, kProducers = 9;

__device__ __forceinline__ void mma_s8(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(0), "r"(0), "r"(0), "r"(0));
}

template <int kMode>
__global__ void k(unsigned long long* out, uint32_t* sink, uint32_t seed)
{
    __shared__ uint4 slot[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = seed * 2654435761u + threadIdx.x * 40503u;
    unsigned long long total = 0;
    uint32_t keep = 0;
    // A fragment of m16n8k32 (row-major 16 x 32 s8): thread t holds rows g = t / 4 and g + 8, k = 4 (t % 4) .. + 3 and + 16.
    // Row m selects component m: byte position k belongs to component (k & 3) for k < 16 and 4 + (k & 3) for k >= 16.
    uint32_t afrag[4] = {0u, 0u, 0u, 0u};
    {
        const int g = lane >> 2;                       // row (component) this thread's A registers describe: g and g + 8
        // a0: row g, k = 4(t%4)+i (i = 0..3): 1 where i == g (g < 4).  a2: row g, k = 16 + 4(t%4)+i: 1 where 4 + i == g.
        if (g < 4) afrag[0] = 1u << (8 * g);
        else if (g < 8) afrag[2] = 1u << (8 * (g - 4));
    }
    for (int it = 0; it < N; it++) {
        x = x * 1664525u + 1013904223u;
        int c[6];
        for (int i = 0; i < 6; i++) c[i] = (int)((x >> (5 * i)) & 63u);          // six counts of this thread, 0 .. 63
        uint32_t p0 = (uint32_t)c[0] | (uint32_t)c[1] << 16, p1 = (uint32_t)c[2] | (uint32_t)c[3] << 16, p2 = (uint32_t)c[4] | (uint32_t)c[5] << 16;
        __syncthreads();
        const long long t0 = clock64();
        uint32_t r0 = 0, r1 = 0, r2 = 0;
        if (warp < kProducers) {
            if (kMode == 0) {
                r0 = __reduce_add_sync(0xFFFFFFFFu, p0);
                r1 = __reduce_add_sync(0xFFFFFFFFu, p1);
                r2 = __reduce_add_sync(0xFFFFFFFFu, p2);
            } else if (kMode == 1) {
                r0 = p0; r1 = p1; r2 = p2;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    r0 += __shfl_xor_sync(0xFFFFFFFFu, r0, d);
                    r1 += __shfl_xor_sync(0xFFFFFFFFu, r1, d);
                    r2 += __shfl_xor_sync(0xFFFFFFFFu, r2, d);
                }
            } else if (kMode == 2) {
                const uint32_t b[2] = {(uint32_t)c[0] | (uint32_t)c[1] << 8 | (uint32_t)c[2] << 16 | (uint32_t)c[3] << 24,
                                       (uint32_t)c[4] | (uint32_t)c[5] << 8};
                int d[4];
                mma_s8(d, afrag, b);
                // thread t: d[0], d[1] = row t/4 (component), columns 2(t%4), 2(t%4)+1: add, then over the four threads of the row
                int s = d[0] + d[1];
                s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
                s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
                // component m in lanes 4m .. 4m+3: hand the six totals to lane 0's packed form
                const int e0 = __shfl_sync(0xFFFFFFFFu, s, 0), e1 = __shfl_sync(0xFFFFFFFFu, s, 4), e2 = __shfl_sync(0xFFFFFFFFu, s, 8);
                const int e3 = __shfl_sync(0xFFFFFFFFu, s, 12), e4 = __shfl_sync(0xFFFFFFFFu, s, 16), e5 = __shfl_sync(0xFFFFFFFFu, s, 20);
                r0 = (uint32_t)e0 | (uint32_t)e1 << 16; r1 = (uint32_t)e2 | (uint32_t)e3 << 16; r2 = (uint32_t)e4 | (uint32_t)e5 << 16;
            } else if (kMode == 3) {     // mma, lanes 0,4,..20 store their own total: no gather shuffles
                const uint32_t b[2] = {(uint32_t)c[0] | (uint32_t)c[1] << 8 | (uint32_t)c[2] << 16 | (uint32_t)c[3] << 24,
                                       (uint32_t)c[4] | (uint32_t)c[5] << 8};
                int d[4];
                mma_s8(d, afrag, b);
                int s = d[0] + d[1];
                s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
                s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
                if ((lane & 3) == 0 && lane < 24) reinterpret_cast<uint32_t*>(&slot[warp])[0] = 0u, ((uint16_t*)&slot[warp])[lane >> 2] = (uint16_t)s;
                r0 = (uint32_t)s;
            }
            if (kMode != 3 && lane == 0) slot[warp] = make_uint4(r0, r1, r2, 0u);
        }
        __syncthreads();
        const long long t1 = clock64();
        if (threadIdx.x == 0) total += (unsigned long long)(t1 - t0);
        keep += slot[it % kProducers].x + r0;
    }
    if (threadIdx.x == 0) out[0] = total;
    sink[threadIdx.x] = keep;
}

template <int kMode>
static double run(unsigned long long* d_out, uint32_t* d_sink)
{
    k<kMode><<<1, kWarps * 32>>>(d_out, d_sink, 1);
    cudaDeviceSynchronize();
    k<kMode><<<1, kWarps * 32>>>(d_out, d_sink, 2);
    cudaDeviceSynchronize();
    unsigned long long t = 0;
    cudaMemcpy(&t, d_out, 8, cudaMemcpyDeviceToHost);
    return (double)t / N;
}

// correctness of the mma routing: totals of mode 2 must equal mode 0
__global__ void check(uint32_t* out)
{
    const int lane = threadIdx.x & 31;
    uint32_t afrag[4] = {0u, 0u, 0u, 0u};
    const int g = lane >> 2;
    if (g < 4) afrag[0] = 1u << (8 * g);
    else if (g < 8) afrag[2] = 1u << (8 * (g - 4));
    int c[6];
    for (int i = 0; i < 6; i++) c[i] = (lane * (7 + i) + 3 * i) % 61 - (i == 5 ? 30 : 0);      // one signed column
    const uint32_t b[2] = {(uint32_t)(uint8_t)c[0] | (uint32_t)(uint8_t)c[1] << 8 | (uint32_t)(uint8_t)c[2] << 16 | (uint32_t)(uint8_t)c[3] << 24,
                           (uint32_t)(uint8_t)c[4] | (uint32_t)(uint8_t)c[5] << 8};
    int d[4];
    mma_s8(d, afrag, b);
    int s = d[0] + d[1];
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
    for (int i = 0; i < 6; i++) {
        const int want = __reduce_add_sync(0xFFFFFFFFu, c[i]);
        const int got = __shfl_sync(0xFFFFFFFFu, s, 4 * i);
        if (lane == 0) { out[2 * i] = (uint32_t)want; out[2 * i + 1] = (uint32_t)got; }
    }
}

int main()
{
    unsigned long long* d_out;
    uint32_t* d_sink;
    cudaMalloc(&d_out, 64);
    cudaMalloc(&d_sink, 4096);
    check<<<1, 32>>>(d_sink);
    uint32_t h[12];
    cudaMemcpy(h, d_sink, sizeof h, cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int i = 0; i < 6; i++) ok &= h[2 * i] == h[2 * i + 1];
    printf("mma routing: component totals %s (%d %d | %d %d | ... | %d %d)\n", ok ? "equal the REDUX totals" : "WRONG", (int)h[0], (int)h[1],
           (int)h[2], (int)h[3], (int)h[10], (int)h[11]);
    printf("nine warps reduce three packed sums each, barrier to barrier, clock ticks:\n");
    printf("  3 x REDUX per warp + lane-0 store         %7.1f\n", run<0>(d_out, d_sink));
    printf("  SHFL butterfly (15 SHFL) + lane-0 store   %7.1f\n", run<1>(d_out, d_sink));
    printf("  1 mma.s8 + 2 SHFL + 6 gather SHFL + store %7.1f\n", run<2>(d_out, d_sink));
    printf("  1 mma.s8 + 2 SHFL, six lanes store        %7.1f\n", run<3>(d_out, d_sink));
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
