// Integer-pipe microbenchmark for sm_100a: warp-instruction issue rates of the ops the correlator is
// built from ("INT32 LOP3/POPC/IADD issue rate x 148 SMs - microbenchmark on
// the box").  Prints lane-ops per clock per SM for each op, measured with clock64() inside the kernel
// so that DVFS does not matter.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_int tools/ubench_int.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

enum Op { OP_LOP3, OP_POPC, OP_IADD3, OP_DP4A, OP_IMAD, OP_SHF, OP_PRMT, OP_POPC_LOP3_ADD, OP_DP4A_LOP3, OP_LDS, OP_COUNT };
static const char* kNames[] = {"lop3", "popc", "iadd3", "dp4a", "imad", "shf(funnel)", "prmt", "xor+popc+add", "dp4a+lop3 (dual)", "lds.32"};
static const int kOpsPerIter[] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 1};

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, long long* cycles, uint32_t seed)
{
    __shared__ uint32_t sm[1024];
    sm[threadIdx.x] = seed * threadIdx.x;
    __syncthreads();
    uint32_t a[CHAINS], b = seed ^ threadIdx.x, c = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) a[i] = seed + i * 977 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == OP_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
            if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == OP_DP4A) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == OP_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == OP_POPC_LOP3_ADD) {
                uint32_t x, p;
                asm volatile("xor.b32 %0, %1, %2;" : "=r"(x) : "r"(b + i), "r"(c));
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(p));
            }
            if (OP == OP_DP4A_LOP3) {
                asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(a[(i + 4) % CHAINS]), "r"(c));
            }
            if (OP == OP_LDS) a[i] = sm[(a[i] + i) & 1023];
        }
    }
    long long t1 = clock64();
    uint32_t r = b;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, uint32_t* d_out, long long* d_cyc)
{
    const int blocks = sms * 2, threads = 1024;  // 2048 threads per SM = full occupancy
    k<OP><<<blocks, threads>>>(d_out, d_cyc, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d_out, d_cyc, 999u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h[i];
    avg /= blocks;
    delete[] h;
    // per SM: 2 blocks x 1024 threads x ITERS x CHAINS x ops in `avg` cycles
    double lane_ops = 2.0 * 1024.0 * ITERS * CHAINS * kOpsPerIter[OP];
    printf("%-20s %8.2f lane-ops/clk/SM   (%.3f ms, %.0f cycles, %.0f MHz)\n", kNames[OP], lane_ops / avg, ms, avg,
           avg / (ms * 1e3));
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("no device\n"); return 1; }
    printf("device: %s, %d SMs, sm_%d%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    uint32_t* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 2 * 1024 * 4);
    cudaMalloc(&d_cyc, (size_t)p.multiProcessorCount * 2 * 8);
    run<OP_LOP3>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_POPC>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_IADD3>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_DP4A>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_IMAD>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_SHF>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_PRMT>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_POPC_LOP3_ADD>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_DP4A_LOP3>(p.multiProcessorCount, d_out, d_cyc);
    run<OP_LDS>(p.multiProcessorCount, d_out, d_cyc);
    return 0;
}
