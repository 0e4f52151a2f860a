/*
 * gpsb_track_loop.cuh - k_track_run: the device-resident closed tracking loop.
 *
 * The reference tracks one satellite per millisecond slot with a strictly serial dependency: the NCO word
 * and the code offsets of millisecond t+1 come out of the DLL / PLL / FLL fed with the six sums of
 * millisecond t (Firmware/project_main/GPS/tracking.c:92-170).  With the loop filters on the host every
 * millisecond costs a PCIe round trip (several microseconds) for a fraction of a microsecond of arithmetic.
 * Here the whole loop of a channel lives in ONE CTA for a whole run of milliseconds:
 *
 *   workers   (kLoopWorkers threads)  integrate-and-dump of the current millisecond straight from the raw
 *             frame in shared memory (core/gpsb_epl_core.h), REDUX per warp, three shared atomics per warp;
 *   control   (one thread of an extra warp) runs the reference's loop filters, false-lock check, NCO / code
 *             offset planning, nav-bit synchronisation and word assembly (core/gpsb_loop_core.h - the same
 *             source libgpsb_host.so is built from) on the channel record held in shared memory;
 *   frames    are fetched two milliseconds ahead by TMA bulk copies (cp.async.bulk -> mbarrier), issued by the
 *             control thread: they do not depend on the loop, so HBM/L2 latency never sits on the serial path.
 *
 * Per millisecond:  workers correlate(m) | control finishes tail(m-1)   -> barrier A ->
 *                   control: sums, DLL/PLL/FLL, plan(m+1)               -> barrier B -> ...
 * so the serial path is correlate + filters + two CTA barriers, and the nav-bit / SNR bookkeeping of a
 * millisecond overlaps the next millisecond's correlation.
 *
 * Bound: latency of one SM (a dependent chain of ~10^3 instructions per millisecond per channel); channels are
 * independent, one CTA each, so throughput scales with the channel count up to the SM count at no extra time.
 * Algorithmic HBM bytes per channel-millisecond: 2046 (frame, shared by all channels through L2) + 12 + 1 (logs).
 */
#pragma once

#include "../core/gpsb_epl_core.h"
#include "../core/gpsb_loop_core.h"
#include "gpsb_kernels.cuh"

namespace gpsb {

constexpr int kLoopWorkers = 256;
constexpr int kLoopNw = kWords / kLoopWorkers;        // replica words per worker
constexpr int kLoopThreads = kLoopWorkers + 32;       // + the control warp
static_assert(kLoopNw >= 1 && kLoopNw <= EC_NW_MAX && kLoopNw * kLoopWorkers == kWords, "work split");

struct LoopSmem {
    uint32_t S[2][kWords];          // raw frames m, m+1 (TMA destinations, 16-byte aligned)
    uint32_t E[kWords];             // chip-expanded code of this channel's satellite
    unsigned long long full[2];     // mbarriers: frame buffer b has landed
    gps_ch_t ch;
    gpsb_aux aux;
    gpsb_epl_req rq;                // what the workers correlate this millisecond
    uint32_t sums[2][4];            // packed I | Q << 16 per arm, double buffered by millisecond parity
    int stop;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// One thread: expect `bytes` on the barrier and start the bulk copy global -> shared that completes it.
__device__ __forceinline__ void tma_load_frame(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

template <typename T>
__device__ __forceinline__ void copy_words(T* dst, const T* src, int tid, int nthreads)
{
    static_assert(sizeof(T) % 4 == 0, "word copy");
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    for (int i = tid; i < (int)(sizeof(T) / 4); i += nthreads) d[i] = s[i];
}

__global__ void __launch_bounds__(kLoopThreads, 1)
k_track_run(gps_ch_t* __restrict__ chans, gpsb_aux* __restrict__ auxs, const uint32_t* __restrict__ codes,
            const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t ms0, uint32_t n_ms,
            int16_t* __restrict__ iq_log, int8_t* __restrict__ nav_log, gpsb_loop_result* __restrict__ results)
{
    __shared__ __align__(128) LoopSmem sm;
    const int tid = threadIdx.x;
    const uint32_t chn = blockIdx.x, n_ch = gridDim.x;
    const bool worker = tid < kLoopWorkers;
    const bool control = tid == kLoopWorkers;

    copy_words(&sm.ch, chans + chn, tid, kLoopThreads);
    copy_words(&sm.aux, auxs + chn, tid, kLoopThreads);
    {
        const uint32_t* __restrict__ e = codes + (size_t)chans[chn].prn * kWords;
        for (int i = tid; i < kWords; i += kLoopThreads) sm.E[i] = __ldg(e + i);
    }
    if (control) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.stop = LC_STOP_NONE;
        for (int b = 0; b < 2; b++)
            for (int a = 0; a < 4; a++) sm.sums[b][a] = 0u;
    }
    __syncthreads();
    if (control) {
        if (sm.ch.tracking_data.state == GPS_PRE_TRACK_DONE) sm.ch.tracking_data.state = GPS_TRACKING_RUN;   // tracking.c:74-78
        if (sm.ch.tracking_data.state != GPS_TRACKING_RUN) {
            sm.stop = LC_STOP_STATE;
        } else if (n_ms) {
            for (uint32_t k = 0; k < 2 && k < n_ms; k++)
                tma_load_frame(sm.S[k], signal + (size_t)((ms0 + k) % ring_ms) * kWords, GPSB_FRAME_BYTES, &sm.full[k]);
            lc_trk_plan_run(&sm.ch, ms0, ms0, &sm.rq);
        }
    }
    __syncthreads();

    uint32_t m = 0;
    int stop = sm.stop;
    for (; m < n_ms && stop == LC_STOP_NONE; m++) {
        const uint32_t ms = ms0 + m;
        const uint32_t b = m & 1u;
        const uint8_t index = (uint8_t)(ms % LC_SLOT_LEN);
        if (worker) {
            const gpsb_epl_req rq = sm.rq;
            const uint32_t off[3] = {rq.off_e, rq.off_p, rq.off_l};
            uint32_t acc[3] = {0u, 0u, 0u};
            mbar_wait(&sm.full[b], (m >> 1) & 1u);
            ec_epl_partial(sm.S[b], sm.E, rq.acc0, rq.step32, off, rq.off_bits, tid * kLoopNw, kLoopNw, acc);
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, acc[a]);
                if ((tid & 31) == 0) atomicAdd(&sm.sums[b][a], v);
            }
        }
        __syncthreads();   // A: the six sums of millisecond m are complete, frame buffer b is free
        int16_t iq[6];
        if (control) {
            const uint32_t packed[3] = {sm.sums[b][0], sm.sums[b][1], sm.sums[b][2]};
            sm.sums[b][0] = sm.sums[b][1] = sm.sums[b][2] = 0u;
            if (m + 2 < n_ms)
                tma_load_frame(sm.S[b], signal + (size_t)((ms + 2) % ring_ms) * kWords, GPSB_FRAME_BYTES, &sm.full[b]);
            ec_unpack_sums(packed, iq);
            if (iq_log) {
                uint32_t* o = reinterpret_cast<uint32_t*>(iq_log + ((size_t)m * n_ch + chn) * 6);
                o[0] = (uint16_t)iq[0] | ((uint32_t)(uint16_t)iq[1] << 16);
                o[1] = (uint16_t)iq[2] | ((uint32_t)(uint16_t)iq[3] << 16);
                o[2] = (uint16_t)iq[4] | ((uint32_t)(uint16_t)iq[5] << 16);
            }
            if (lc_dll_is_degenerate(iq)) {          // 0/0 in the DLL: x86 and the GPU disagree on NaN bits, host finishes this ms
                sm.stop = LC_STOP_DLL_NAN;
                gpsb_loop_result r;
                r.done_ms = m;
                r.stop = LC_STOP_DLL_NAN;
                for (int k = 0; k < 6; k++) r.iq[k] = iq[k];
                r.reserved = 0;
                results[chn] = r;
                if (nav_log) nav_log[(size_t)m * n_ch + chn] = -1;
            } else {
                lc_finish_loops(&sm.ch, &sm.aux, index, iq);
                if (m + 1 < n_ms) lc_trk_plan_run(&sm.ch, ms + 1, ms + 1, &sm.rq);
            }
        }
        __syncthreads();   // B: next request published
        stop = sm.stop;
        if (control && stop == LC_STOP_NONE) {       // overlaps the workers' next correlation
            lc_finish_tail(&sm.ch, &sm.aux, index, iq[2], iq[3], ms);
            if (nav_log) nav_log[(size_t)m * n_ch + chn] = sm.aux.last_nav_bit;
        }
    }
    __syncthreads();
    copy_words(chans + chn, &sm.ch, tid, kLoopThreads);
    copy_words(auxs + chn, &sm.aux, tid, kLoopThreads);
    if (control && stop != LC_STOP_DLL_NAN) {
        gpsb_loop_result r;
        r.done_ms = m;
        r.stop = stop;
        for (int k = 0; k < 6; k++) r.iq[k] = 0;
        r.reserved = 0;
        results[chn] = r;
    }
}

// Level-0 view of the loop's float discriminators for the self-test against the host libm:
// out[(ip - ip_lo) * 16369 + (qp + 8184)], kind 0 = Costas error (tracking.c:180-183), 1 = FLL angle (:232).
__global__ void k_l0_loop_math(int kind, int ip_lo, int n_ip, float* __restrict__ out)
{
    const size_t n = (size_t)n_ip * 16369u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ip = ip_lo + (int)(i / 16369u);
        const int qp = (int)(i % 16369u) - 8184;
        out[i] = kind == 0 ? lc_costas_err((int16_t)ip, (int16_t)qp) : lc_fll_angle((int16_t)ip, (int16_t)qp);
    }
}

}  // namespace gpsb
