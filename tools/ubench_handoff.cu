// Hand-off latency between two warps of one CTA through shared memory: what one link of k_track_run's per-millisecond
// chain costs (control thread -> workers: NCO words / code offsets; workers -> control threads: the six sums).
// A "ping-pong": warp 0 lane 0 publishes a word, warp 1 lane 0 (or all lanes of N waiting warps) sees it and answers;
// reported is clock64 ticks per ONE-WAY hand-off (round trip / 2), for several mechanisms:
//   mbar      mbarrier.arrive  ->  mbarrier.try_wait.parity loop           (what k_track_run used in round 1)
//   spin      st.volatile.shared  ->  ld.volatile.shared spin              (sequence number in the payload word)
//   spin128   the same with a 16-byte payload (st.shared.v4 / ld.shared.v4: sequence + three data words in one access)
//   bar       bar.sync of the whole CTA (both sides arrive)                (barrier A)
//   named     bar.arrive / bar.sync on a named barrier with only the two warps taking part
//   red+bar   9 warps atomicAdd to one shared word, bar.sync, one LDS      (how the sums reached the control threads)
//   redux     __reduce_add_sync of three registers                          (the warp-level part of the same)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_handoff tools/ubench_handoff.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4000;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_vol4(const uint4* p)
{
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol4(uint4* p, uint4 v)
{
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

enum Mode { M_MBAR, M_SPIN, M_SPIN128, M_BAR, M_NAMED, M_REDBAR, M_REDUX, M_SPIN_ALLLANES, M_FANOUT8, M_FANIN8, M_COUNT };
static const char* kNames[] = {"mbarrier arrive -> try_wait (1 -> 1 thread)", "volatile st -> ld spin (1 -> 1 thread)",
                               "volatile st.v4 -> ld.v4 spin (16-byte payload)", "bar.sync of the whole CTA (12 warps)",
                               "named barrier, two warps", "9 x atomicAdd.shared + bar.sync + LDS (sums to a control thread)",
                               "3 x __reduce_add_sync (dependent on the inputs, result used)",
                               "volatile st -> ld spin, all 32 lanes of the waiting warp read the word",
                               "fan-out: 1 thread publishes 16 B, 8 warps (all lanes) spin on it, last one answers",
                               "fan-in: 8 warps each publish 16 B + bump a counter (atomic), 1 thread spins on the counter, reads 8 x 16 B"};

__global__ void __launch_bounds__(384, 1) k(int mode, long long* out, uint32_t* sink)
{
    __shared__ unsigned long long bar[2];
    __shared__ __align__(16) uint32_t word[2][32];
    __shared__ __align__(16) uint4 slot[2];
    __shared__ __align__(16) uint4 part[16];
    __shared__ uint32_t counter[2];
    __shared__ uint32_t accum[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        word[0][0] = word[1][0] = 0;
        slot[0] = slot[1] = make_uint4(0, 0, 0, 0);
        counter[0] = counter[1] = 0;
        accum[0] = 0;
    }
    __syncthreads();
    long long t0 = 0, t1 = 0;
    uint32_t acc = 0;
    if (mode == M_MBAR) {
        if (tid == 0) {
            t0 = clock64();
            for (int i = 0; i < N; i++) { mbar_arrive(&bar[0]); mbar_wait(&bar[1], i & 1); }
            t1 = clock64();
        } else if (tid == 32) {
            for (int i = 0; i < N; i++) { mbar_wait(&bar[0], i & 1); mbar_arrive(&bar[1]); }
        }
    } else if (mode == M_SPIN) {
        if (tid == 0) {
            t0 = clock64();
            for (int i = 1; i <= N; i++) { st_vol(&word[0][0], i); while (ld_vol(&word[1][0]) != (uint32_t)i) {} }
            t1 = clock64();
        } else if (tid == 32) {
            for (int i = 1; i <= N; i++) { while (ld_vol(&word[0][0]) != (uint32_t)i) {} st_vol(&word[1][0], i); }
        }
    } else if (mode == M_SPIN128) {
        if (tid == 0) {
            t0 = clock64();
            for (int i = 1; i <= N; i++) {
                st_vol4(&slot[0], make_uint4(i, acc, i * 3, i * 5));
                uint4 v;
                do v = ld_vol4(&slot[1]); while (v.x != (uint32_t)i);
                acc += v.y + v.z;
            }
            t1 = clock64();
        } else if (tid == 32) {
            for (int i = 1; i <= N; i++) {
                uint4 v;
                do v = ld_vol4(&slot[0]); while (v.x != (uint32_t)i);
                st_vol4(&slot[1], make_uint4(i, v.y + v.z, v.w, 0));
            }
        }
    } else if (mode == M_BAR) {
        if (tid == 0) t0 = clock64();
        for (int i = 0; i < N; i++) __syncthreads();
        if (tid == 0) t1 = clock64();
    } else if (mode == M_NAMED) {
        if (warp == 0) {
            if (tid == 0) t0 = clock64();
            for (int i = 0; i < N; i++) {
                asm volatile("bar.arrive 1, 64;" ::: "memory");
                asm volatile("bar.sync 2, 64;" ::: "memory");
            }
            if (tid == 0) t1 = clock64();
        } else if (warp == 1) {
            for (int i = 0; i < N; i++) {
                asm volatile("bar.sync 1, 64;" ::: "memory");
                asm volatile("bar.arrive 2, 64;" ::: "memory");
            }
        }
    } else if (mode == M_REDBAR) {
        if (tid == 0) t0 = clock64();
        for (int i = 0; i < N; i++) {
            if (warp < 9 && lane == 0) atomicAdd(&accum[0], (uint32_t)(i + warp + acc));
            __syncthreads();
            if (tid == 352) acc += *(volatile uint32_t*)&accum[0];      // a "control thread" reads the total
            acc += i;
        }
        if (tid == 0) t1 = clock64();
    } else if (mode == M_REDUX) {
        uint32_t a = tid, b = tid * 3, c = tid * 7;
        if (tid == 0) t0 = clock64();
        if (warp == 0)
            for (int i = 0; i < N; i++) {
                const uint32_t x = __reduce_add_sync(0xFFFFFFFFu, a), y = __reduce_add_sync(0xFFFFFFFFu, b),
                               z = __reduce_add_sync(0xFFFFFFFFu, c);
                a = (x & 1023u) + lane; b = (y & 1023u) + lane; c = (z & 1023u) + lane;
            }
        if (tid == 0) t1 = clock64();
        acc = a + b + c;
    } else if (mode == M_SPIN_ALLLANES) {
        if (tid == 0) {
            t0 = clock64();
            for (int i = 1; i <= N; i++) { st_vol(&word[0][0], i); while (ld_vol(&word[1][0]) != (uint32_t)i) {} }
            t1 = clock64();
        } else if (warp == 1) {
            for (int i = 1; i <= N; i++) {
                while (ld_vol(&word[0][0]) != (uint32_t)i) {}
                __syncwarp();
                if (lane == 0) st_vol(&word[1][0], i);
            }
        }
    } else if (mode == M_FANOUT8) {
        if (tid == 352) {
            t0 = clock64();
            for (int i = 1; i <= N; i++) {
                st_vol4(&slot[0], make_uint4(i, acc, i * 3, i * 5));
                while (ld_vol(&counter[0]) != (uint32_t)(8 * i)) {}
            }
            t1 = clock64();
        } else if (warp < 8) {
            for (int i = 1; i <= N; i++) {
                uint4 v;
                do v = ld_vol4(&slot[0]); while (v.x != (uint32_t)i);
                acc += v.z;
                __syncwarp();
                if (lane == 0) atomicAdd(&counter[0], 1u);
            }
        }
    } else if (mode == M_FANIN8) {
        // the return path of the sums: 8 worker warps each store their 16-byte partial, then bump a counter; the
        // control thread spins on the counter and reads the eight partials with vector loads
        if (tid == 352) {
            t0 = clock64();
            for (int i = 1; i <= N; i++) {
                st_vol(&word[0][0], i);                                   // "go"
                while (ld_vol(&counter[1]) != (uint32_t)(8 * i)) {}
                uint32_t s = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) { const uint4 v = ld_vol4(&part[w]); s += v.x + v.y + v.z; }
                acc += s;
            }
            t1 = clock64();
        } else if (warp < 8) {
            for (int i = 1; i <= N; i++) {
                while (ld_vol(&word[0][0]) != (uint32_t)i) {}
                __syncwarp();
                if (lane == 0) {
                    st_vol4(&part[warp], make_uint4(i + warp, i, warp, 0));
                    __threadfence_block();
                    atomicAdd(&counter[1], 1u);
                }
            }
        }
    }
    if (t1) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}

int main()
{
    long long* d_out;
    uint32_t* d_sink;
    cudaMalloc(&d_out, 8);
    cudaMalloc(&d_sink, 4);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("tools/bin/ubench_handoff on %s - source tools/ubench_handoff.cu\n", prop.name);
    printf("clock64 ticks per hand-off (one way where a round trip is timed), 384-thread CTA, %d iterations:\n", N);
    for (int m = 0; m < M_COUNT; m++) {
        long long best = 1ll << 60;
        for (int rep = 0; rep < 3; rep++) {
            long long h = 0;
            cudaMemset(d_out, 0, 8);
            k<<<1, 384>>>(m, d_out, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  %s: %s\n", kNames[m], cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            if (h < best) best = h;
        }
        const bool round_trip = m == M_MBAR || m == M_SPIN || m == M_SPIN128 || m == M_NAMED || m == M_SPIN_ALLLANES;
        const bool two_hops = m == M_FANOUT8 || m == M_FANIN8;
        printf("  %-100s %7.1f%s\n", kNames[m], (double)best / N / (round_trip ? 2.0 : 1.0),
               two_hops ? "  (out and back)" : "");
    }
    return 0;
}
