/*
 * gpsb_track_loop.cuh - k_track_run: the device-resident closed tracking loop.
 *
 * The reference tracks one satellite per millisecond slot with a strictly serial dependency: the NCO word
 * and the code offsets of millisecond t+1 come out of the DLL / PLL / FLL fed with the six sums of
 * millisecond t (Firmware/project_main/GPS/tracking.c:92-170).  With the loop filters on the host every
 * millisecond costs a PCIe round trip (several microseconds) for a fraction of a microsecond of arithmetic.
 * Here the whole loop of a channel lives in ONE CTA for a whole run of milliseconds, and the serial chain of a
 * millisecond is cut into pieces that run side by side on different warps:
 *
 *   workers  (255 threads, two data words each, + 12 "edge" lanes for words 0, 511 and the four bytes an odd offset
 *            skips)  integrate-and-dump in two phases (core/gpsb_epl_core.h):
 *            phase 1 needs only the code offsets and forms raw word ^ replica window + byte masks in registers;
 *            phase 2 needs the carrier NCO words: the carrier phase of a word selects its I and Q count among the four
 *            quadrant-pattern counts phase 1 left behind (one PRMT per arm), then REDUX per warp and one shared-memory
 *            RED per warp and arm - or, in the builds it is faster in, one 16-byte store per warp into a slot of its own, the
 *            nine slots added up by every warp behind the barrier (kSlots).  Nothing mixed is ever staged in memory.
 *   code     (1 thread)  DLL + code-offset planning (tracking.c:333-393, 115-130); releases phase 1 through an
 *            mbarrier as soon as the offsets exist, while the carrier thread is still busy.
 *   carrier  (1 thread)  Costas PLL / FLL / false-lock check + NCO planning (tracking.c:175-327, gps_misc.c:250);
 *            releases phase 2 through its own mbarrier.
 *   nav      (1 thread)  bit synchronisation, word assembly, parity, SNR bookkeeping (nav_data.c:46-453,
 *            tracking.c:154-169); waits for the DLL only for the rare bit-edge refinement that reads the code phase.
 *   frames   arrive by TMA bulk copies (cp.async.bulk -> mbarrier) a millisecond ahead (code thread, after its DLL);
 *            in a streaming run not before the producer's watermark has passed them; the eight sub-byte
 *            shifted, periodically extended replica streams are built once per run.  Neither HBM/L2 latency nor
 *            the period-2046-byte seam handling sits on the serial path.
 *
 * The control threads run the very functions libgpsb_host.so is built from (core/gpsb_loop_core.h); each field of
 * the channel record is written by exactly one of them.  The one field read across threads,
 * nav_data.period_sync_ok_flag, is written at slot index 3 only and read by the PLL at slot index 0 only.
 *
 * Per millisecond:  phase 2 -> barrier A -> { carrier -> nco_ready | code -> offs_ready | nav | workers: wait offs_ready,
 *                    phase 1(m+1), wait nco_ready }  - one full barrier; everything else is handed over by mbarriers
 * Bound: latency of one SM; channels are independent, one CTA each, so throughput scales with the channel count
 * up to the SM count at no extra time.  Algorithmic HBM bytes per channel-millisecond: 2046 (frame, shared by all
 * channels through L2) + 12 + 1 (logs).
 */
#pragma once

#include "../core/gpsb_epl_core.h"
#include "../core/gpsb_loop_core.h"
#include "gpsb_kernels.cuh"

namespace gpsb {

#ifndef GPSB_LOOP_WORKERS
#define GPSB_LOOP_WORKERS 256
#endif
#ifndef GPSB_LOOP_WORKER_DLL
#define GPSB_LOOP_WORKER_DLL 1
#endif
#ifndef GPSB_LOOP_SLOTS             // -1: per build (see kSlots); 0 / 1: all builds without / with the per-warp slots (experiments)
#define GPSB_LOOP_SLOTS (-1)
#endif
#ifndef GPSB_LOOP_WORKER_DLL_WALK
#define GPSB_LOOP_WORKER_DLL_WALK 0
#endif
constexpr int kLoopWorkers = GPSB_LOOP_WORKERS;
constexpr int kLoopNw = kWords / kLoopWorkers;        // data words per worker thread (words 1..510; 0 and 511 are edge words)
constexpr int kWorkerWarps = kLoopWorkers / 32;
constexpr int kLoopWarps = kWorkerWarps + 4;          // + code, edge, nav, carrier
constexpr int kLoopThreads = kLoopWarps * 32;
static_assert(kLoopNw >= 1 && kLoopNw <= EC_NW_MAX && kLoopNw * kLoopWorkers == kWords, "work split");

struct LoopSmem {
    uint32_t S[2][kWords];              // raw frames m, m+1 (TMA destinations, 16-byte aligned)
    uint32_t RX[8][EC_RX_WORDS + 3];    // replica stream of this satellite for sub-byte shifts 0..7, extended (ec_rx_word)
    uint32_t E[kWords];                 // chip-expanded code of this channel's satellite
    unsigned long long full[2];         // mbarriers: raw frame buffer b has landed
    unsigned long long offs_ready;      // mbarrier: the code thread has published the next offsets
    unsigned long long nco_ready;       // mbarrier: the carrier thread has published the next NCO words
    gps_ch_t ch;
    gpsb_aux aux;
    gpsb_epl_req rq;                    // what the workers correlate next
    uint32_t top_lut[16];               // ec_top_nibble_counts(0..15), see EC_COUNTS_FULL
    uint4 sums[2];                      // packed I | Q << 16 of the three arms, accumulated by one shared-memory RED per warp
                                        // and arm; double buffered by ms parity, re-zeroed by the code thread one ms later
    uint4 partial[2][16];               // kSlots builds instead: per warp (slots 0..7 the workers, 9 the edge warp; the others
                                        // stay zero) its packed partial sums of a millisecond, double buffered by ms parity
    int stop;                           // set before the loop: the run does not start (state, no frames)
    int2 ctl[2];                        // .x = stop code, set DURING millisecond m into slot (m+1)&1, read after barrier A of
                                        // millisecond m+1: two slots, so it is never written in the barrier interval in which it
                                        // is read.  .y (streaming) = the run's length: n_ms, or m+2 once the producer has missed
                                        // frame m+2 - millisecond m+1 is then the last (nobody plans a successor) and the loop
                                        // ends as at the end of a shorter run.  One 8-byte load per thread and millisecond.
};

// What the code thread and the carrier thread own of gps_tracking_t (PM/GPS/gps_misc.h:62-99), under the record's own
// field names so that core/gpsb_loop_core.h instantiates on them: scalars only, so they live in registers for a whole run.
struct CodeRegs {
    float code_phase_fine, dll_code_err;
#if (ENABLE_CODE_FILTER)
    uint16_t code_filt_cnt;
    float code_phase_fine_filt;
#endif
};
struct CarrierRegs {
    float if_freq_offset_hz;
    uint32_t if_freq_accum, prev_track_timestamp;
    float pll_code_err;
    int16_t fll_old_i, fll_old_q;
    float fll_err;
    int16_t pll_check_buf[TRACKING_CH_LENGTH];
    uint8_t pll_bad_state_cnt;
    uint16_t pll_bad_state_master_cnt;
};


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

__device__ __forceinline__ void load_sums(const uint4* sums, int16_t iq[6])
{
    const uint4 v = *sums;
    const uint32_t packed[3] = {v.x, v.y, v.z};
    ec_unpack_sums(packed, iq);
}

// Streaming ingest: while a run is in flight the host keeps DMA-ing frames into the ring and, after each chunk, the
// number of the first millisecond NOT yet uploaded into `*watermark` (device memory, same copy stream, so the frames
// land first).  The code thread looks at it only when the frame it is about to fetch is not known to be there yet -
// once per chunk in steady state - and never waits longer than timeout_ns in total for one frame (a stalled producer
// ends the run with LC_STOP_STARVED instead of wedging the GPU).  watermark == nullptr: everything is resident.
struct StreamGate {
    const uint32_t* watermark;
    uint32_t* progress;              // mapped host memory, [n_ch]: millisecond the channel has reached (flow control)
    unsigned long long timeout_ns;
    const uint32_t* abort;           // mapped host memory: non-zero = the producer has given up (gpsb_stream_abort): a wait
                                     // for a frame ends at once with "starved" instead of running into the time-out
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// 1 when frame `ms` is in the ring (known_upto caches the last watermark seen), 0 when the producer timed out
__device__ __forceinline__ int frame_present(const StreamGate& gate, uint32_t ms, uint32_t& known_upto)
{
    if (!gate.watermark || (int32_t)(known_upto - ms) > 0) return 1;
    unsigned long long t0 = 0;
    for (;;) {
        known_upto = ld_acquire_u32(gate.watermark);
        if ((int32_t)(known_upto - ms) > 0) return 1;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (!t0) t0 = now;
        else if (now - t0 > gate.timeout_ns) return 0;
        if (gate.abort && *(volatile const uint32_t*)gate.abort) return 0;
        __nanosleep(200);
    }
}

// kProf: diagnostic build that accumulates clock64 ticks per phase into prof[chn * 16 ..] (GPSB_LOOP_PROFILE=1):
//   0 workers: phase 2 + reduce   1 workers: A -> phase 1 of the next ms complete   2 code thread: DLL + offsets
//   3 carrier thread              4 nav thread                                     5 whole loop
//   6 carrier thread: wait at barrier A   7..9 phase 2 split   10, 11 carrier thread at slot index 0   12 phase-1 redos
// kExp: timing experiments only (results wrong): bit 0 skips phase 1, bit 1 skips the carrier filters, bit 2 the DLL
// kStream: streaming-ingest build (the resident build carries none of its checks)
// kWalk: the slot-phase walk (lc_walk_*) is compiled in.  A run whose channels all have it switched off, sit at slot
// phase 0 and have no idle gap pending takes the build without it (the host checks the records; a record that does
// not qualify makes that build refuse the channel with LC_STOP_STATE).
template <bool kProf, int kExp = 0, bool kStream = false, bool kWalk = true>
#ifndef GPSB_LOOP_STREAM_WDLL
#define GPSB_LOOP_STREAM_WDLL 1
#endif
// Register cap per build, measured (tools/ab_run.sh, us per ms of signal; profiles/loop_experiments_r2.txt):
//   resident, no walk (per-warp sum slots)    88: 0.977   96: 0.951   104: 0.962   112: 0.960   128: 0.964      -> 96
//   streaming, no walk (per-warp sum slots)   96: 1.007   104: 0.967  112: 0.947   120: 0.955   128: 0.960      -> 112
//   resident, walk                            96 + atomics: 1.016   112 + atomics: 1.014   104 / 112 / 120 + slots: 1.034 / 1.030 / 1.021   -> 96, atomics
//   streaming, walk                           96 + atomics: 1.056   112 + atomics: 1.038   104 / 112 / 120 + slots: 1.015 / 1.042 / 1.044   -> 104, slots
#ifndef GPSB_LOOP_REGS_STREAM
#define GPSB_LOOP_REGS_STREAM 112
#endif
#ifndef GPSB_LOOP_REGS_RESIDENT
#define GPSB_LOOP_REGS_RESIDENT 96
#endif
__global__ void __maxnreg__(kWalk ? (kStream ? 104 : 96) : kStream ? GPSB_LOOP_REGS_STREAM : GPSB_LOOP_REGS_RESIDENT)      // measured: 96 registers 1.068 us per ms, uncapped (123) 1.087, 80: 1.124
k_track_run(gps_ch_t* __restrict__ chans, gpsb_aux* __restrict__ auxs, const uint32_t* __restrict__ codes,
            const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t ms0, uint32_t n_ms,
            int16_t* __restrict__ iq_log, int8_t* __restrict__ nav_log, gpsb_loop_result* __restrict__ results,
            unsigned long long* __restrict__ prof, StreamGate gate, const uint32_t* __restrict__ first_ms)
{
    __shared__ __align__(128) LoopSmem sm;
    // first_ms != nullptr: channel c starts at millisecond first_ms[c] >= ms0 instead of ms0 (k_pretrack_run, launched ahead of
    // this kernel, left it there); rows of the logs and done_ms stay relative to ms0.
    uint32_t done_base = 0;
    if (first_ms) {
        uint32_t skip = first_ms[blockIdx.x] - ms0;
        if (skip > n_ms) skip = n_ms;
        done_base = skip;
        ms0 += skip;
        n_ms -= skip;
        if (iq_log) iq_log += (size_t)skip * gridDim.x * 6;
        if (nav_log) nav_log += (size_t)skip * gridDim.x;
    }
    long long pt[15] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long c0 = 0;
    const long long loop_begin = kProf ? clock64() : 0;
    // kProf: raw clock64 stamps of four consecutive milliseconds (kTlFirst ..) per warp, eight per millisecond, behind the
    // per-phase totals: prof[n_ch * 16 + ((chn * kLoopWarps + warp) * 4 + k) * 8 + point]
    constexpr uint32_t kTlFirst = 500u;
    __shared__ uint32_t tl_buf[kProf ? kLoopWarps * 4 * 8 : 1];      // stamps go to shared memory (one STS), dumped after the loop
#define GPSB_TL(point)                                                                                                  \
    do {                                                                                                                \
        if (kProf && (threadIdx.x & 31) == 0 && m >= kTlFirst && m < kTlFirst + 4u)                                      \
            tl_buf[((threadIdx.x >> 5) * 4 + (m - kTlFirst)) * 8 + (point)] = (uint32_t)clock();                         \
    } while (0)
    const int tid = threadIdx.x;
    const uint32_t chn = blockIdx.x, n_ch = gridDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    // Warp w issues from scheduler w % 4.  The eight worker warps are spread over all four schedulers (two each: phase
    // 1 is bound by the 16-lane integer pipes of a scheduler); the control threads and the edge lanes get one scheduler
    // each, sharing it with two worker warps.  (With a second full barrier per millisecond the control threads were
    // better off alone on one scheduler; now that the workers' phase 1 has to beat the carrier thread, they are not.)
    const int widx = warp < kWorkerWarps ? warp : -1;      // worker warp index, -1 otherwise
    const bool worker = widx >= 0;
    const bool code_thr = warp == kWorkerWarps && lane == 0;
    const bool edge_warp = warp == kWorkerWarps + 1;                // lanes 0..11: the irregular words and bytes (ec_epl_edge_phase1)
    const bool nav_thr = warp == kWorkerWarps + 2 && lane == 0;
    const bool carrier_thr = warp == kWorkerWarps + 3 && lane == 0;
    const bool edge = edge_warp && lane < EC_EDGE_LANES;
    const int wtid = widx * 32 + lane;                      // worker thread index 0..255
    const int w0 = wtid * kLoopNw + 1;                      // first data word of a worker: words 1..510
    const bool plain = worker && wtid < (kWords - 2) / kLoopNw;   // words 1..510; the last worker thread(s) have no words left
    uint32_t edge_counts = 0u;
    int edge_w = 0, edge_neg = 0;
    const int edge_role = lane / 3, edge_arm = lane % 3;    // an edge lane's entry: role 0..3, arm 0..2 (ec_epl_edge_entry)

    if (kProf) for (int i = tid; i < kLoopWarps * 4 * 8; i += kLoopThreads) tl_buf[i] = 0u;
    copy_words(&sm.ch, chans + chn, tid, kLoopThreads);
    copy_words(&sm.aux, auxs + chn, tid, kLoopThreads);
    {
        const uint32_t* __restrict__ e = codes + (size_t)chans[chn].prn * kWords;
        for (int i = tid; i < kWords; i += kLoopThreads) sm.E[i] = __ldg(e + i);
    }
    if (code_thr) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        mbar_init(&sm.offs_ready, 1);
        mbar_init(&sm.nco_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.stop = LC_STOP_NONE;
        sm.ctl[0] = sm.ctl[1] = make_int2(LC_STOP_NONE, (int)n_ms);
        sm.sums[0] = make_uint4(0u, 0u, 0u, 0u);
        sm.sums[1] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = 0; i < 16; i++) sm.partial[0][i] = sm.partial[1][i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    if (tid < 16) sm.top_lut[tid] = ec_top_nibble_counts((uint32_t)tid);
    for (int i = tid; i < 8 * EC_RX_WORDS; i += kLoopThreads)        // tracking only uses shifts 0..7, tracking.c:116
        sm.RX[i / EC_RX_WORDS][i % EC_RX_WORDS] = ec_rx_word(sm.E, i % EC_RX_WORDS, (uint32_t)(i / EC_RX_WORDS));
    uint32_t known_upto = ms0;          // code thread: frames below this are known to be in the ring
    uint32_t issued = 0, consumed = 0;  // code thread: frames fetched by bulk copy / frames the workers have waited for
    if (code_thr) {
        // tracking.c:74-78 - on the channel's NEXT millisecond, so not when it has none left in this run (k_pretrack_run
        // may have used them all)
        if (n_ms && sm.ch.tracking_data.state == GPS_PRE_TRACK_DONE) sm.ch.tracking_data.state = GPS_TRACKING_RUN;
        if (n_ms == 0) {
            sm.stop = LC_STOP_NONE;                         // nothing to do: done_ms = what k_pretrack_run used
        } else if (sm.ch.tracking_data.state != GPS_TRACKING_RUN) {
            sm.stop = LC_STOP_STATE;
        } else if (n_ms) {
            const uint32_t first = n_ms < 2 ? n_ms : 2;
            if (kStream && !frame_present(gate, ms0 + first - 1, known_upto)) sm.stop = LC_STOP_STARVED;    // nothing fetched, nothing done
            else {
                for (uint32_t k = 0; k < first; k++)
                    tma_load_frame(sm.S[k], signal + (size_t)((ms0 + k) % ring_ms) * kWords, GPSB_FRAME_BYTES, &sm.full[k]);
                issued = first;
                consumed = 1;           // phase 1 of millisecond 0, below
            }
            // a run that starts inside an idle gap of the slot-phase walk (lc_walk_*): no carrier plan for that
            // millisecond - the gap's last millisecond plans the first one behind it, in the loop
            if (!kWalk && (sm.aux.walk_enable || sm.aux.slot_phase || sm.aux.skip_len)) {
                sm.stop = LC_STOP_STATE;                    // the host picked the build without the walk for a walking channel
            } else if (kWalk && lc_walk_idle(sm.aux.skip_ms, sm.aux.skip_len, ms0)) {
                sm.rq.sv_slot = sm.ch.prn;
                sm.rq.ms_index = ms0;
                sm.rq.acc0 = sm.rq.step32 = 0u;
                lc_plan_code(&sm.ch.tracking_data, &sm.rq);
            } else {
                lc_trk_plan_run(&sm.ch, ms0, ms0, &sm.rq);
            }
        }
    }
    __syncthreads();
    int stop = sm.stop;
    // Slot-phase walk (core/gpsb_loop_core.h, lc_walk_*).  The workers know nothing of it: in a millisecond the channel
    // leaves out they correlate as ever and nobody looks at the sums.  Each CONTROL thread keeps its own copy of the
    // channel's slot phase and of the idle gap the nav thread may have decided; the nav thread writes a decision at slot
    // index 3, the others re-read it at the end of their slot-index-2 work (barriers in between), a whole slot before
    // it takes effect.  Everything the walk adds sits inside the control threads' own branches, after their chains.
    uint32_t w_phase = sm.aux.slot_phase, w_skip_ms = sm.aux.skip_ms, w_skip_len = sm.aux.skip_len;
#define GPSB_WALK_STEP_END()                                                                     \
    do {                                                                                         \
        if (kWalk && idle && !idle_next) w_phase = lc_walk_phase_after(ms);                      \
        if (kWalk && index == LC_SLOT_LEN - 2) {                                                 \
            w_skip_ms = *(volatile uint32_t*)&sm.aux.skip_ms;                                    \
            w_skip_len = *(volatile uint8_t*)&sm.aux.skip_len;                                   \
        }                                                                                        \
    } while (0)
    ec_partial part;
    if (stop == LC_STOP_NONE && n_ms && (plain || edge)) {  // phase 1 of millisecond 0
        const uint32_t off[3] = {sm.rq.off_e, sm.rq.off_p, sm.rq.off_l};
        mbar_wait(&sm.full[0], 0u);
        if (plain) ec_epl_phase1(sm.S[0], sm.RX[sm.rq.off_bits & 7u], off, w0, kLoopNw, &part, sm.top_lut);
        else edge_counts = ec_epl_edge_entry(sm.S[0], sm.RX[sm.rq.off_bits & 7u], (&sm.rq.off_e)[edge_arm], edge_role, &edge_w, &edge_neg);
    }
    __syncthreads();                                        // raw buffer 0 has been consumed
    // Streaming runs: a frame that is not there in time ends the run.  The code thread shortens the run (sm.ctl[].y) so
    // that the next millisecond is the last one; nobody plans a successor, so the records leave the kernel exactly as
    // after a shorter run.
    lc_angle_cache angle_cache;
    angle_cache.valid = 0;
    // The code and the carrier thread keep private copies of the channel record for the whole run, so that the
    // fields they own live in registers instead of taking a shared-memory round trip per access; what another
    // thread reads (the code phase for the nav thread's edge refinement, the bit-sync flag for the PLL gains) is
    // exchanged through sm.ch, and the owned fields are written back when the loop ends.
    CodeRegs cod;
    CarrierRegs car;
    {
        const gps_tracking_t* t = &sm.ch.tracking_data;
        cod.code_phase_fine = t->code_phase_fine;
        cod.dll_code_err = t->dll_code_err;
#if (ENABLE_CODE_FILTER)
        cod.code_filt_cnt = t->code_filt_cnt;
        cod.code_phase_fine_filt = t->code_phase_fine_filt;
#endif
        car.if_freq_offset_hz = t->if_freq_offset_hz;
        car.if_freq_accum = t->if_freq_accum;
        car.prev_track_timestamp = t->prev_track_timestamp;
        car.pll_code_err = t->pll_code_err;
        car.fll_old_i = t->fll_old_i;
        car.fll_old_q = t->fll_old_q;
        car.fll_err = t->fll_err;
        car.pll_check_buf[0] = t->pll_check_buf[0];
        car.pll_check_buf[1] = t->pll_check_buf[1];
        car.pll_check_buf[2] = t->pll_check_buf[2];
        car.pll_check_buf[3] = t->pll_check_buf[3];
        car.pll_bad_state_cnt = t->pll_bad_state_cnt;
        car.pll_bad_state_master_cnt = t->pll_bad_state_master_cnt;
    }
    // kWorkerDll: the workers do not wait for the code thread's offsets - every worker thread runs the DLL itself on the
    // six sums it reads after barrier A (same IEEE arithmetic, same result in every thread) and goes straight into
    // phase 1: the hand-over code thread -> shared memory -> mbarrier -> workers leaves the serial path.
    constexpr bool kWorkerDll = GPSB_LOOP_WORKER_DLL && (!kWalk || GPSB_LOOP_WORKER_DLL_WALK) && (!kStream || GPSB_LOOP_STREAM_WDLL) && !kProf && kExp == 0;
    CodeRegs wcod = cod;
    const uint8_t prn = sm.ch.prn;
    const int16_t found_freq_offset_hz = sm.ch.acq_data.found_freq_offset_hz;

    // One barrier per millisecond.  Phase 2 of millisecond m+1 needs two things, each handed over by its own
    // mbarrier: the phase-1 partials (offs_ready: the DLL's offsets, then the warp's own phase 1) and the NCO words
    // (nco_ready: the carrier thread).  A worker warp goes straight from its phase 1 into phase 2 the moment the NCO
    // words exist; nobody waits for the slowest warp's phase 1 or for the nav thread in between.  Barrier A - all six
    // sums complete - is the only full barrier; it also orders everything that is reused one millisecond later (the
    // request record, the sum buffers, the frame buffers, the stop flag).
    // How the six sums get from nine warps to everybody.  kSlots: one 16-byte store per warp into a slot of its own, every warp
    // adds the slots up after barrier A (three REDUX).  Otherwise: 27 shared-memory atomics on three words in front of the
    // barrier, one 16-byte load behind it.  Measured per build (tools/ab_run.sh, us per ms of signal, atomics -> slots):
    // resident without the walk 0.985 -> 0.950; streaming without the walk 0.979 -> 1.007 at 96 registers, but 0.962 -> 0.947 at
    // 112; resident with the walk 1.012 -> 1.021.  So each build takes what is faster for it (the difference is instruction
    // scheduling and register allocation, not the algorithm): slots everywhere but in the resident build with the walk.
    constexpr bool kSlots = GPSB_LOOP_SLOTS >= 0 ? (GPSB_LOOP_SLOTS != 0) : ((!kWalk || kStream) && !kProf && kExp == 0);
    uint32_t m = 0;
    uint32_t limit = n_ms;              // streaming: re-read every millisecond (sm.ctl[].y)
    for (; m < limit && stop == LC_STOP_NONE; m++) {
        const uint32_t ms = ms0 + m;
        const uint32_t b = m & 1u;
        const uint8_t index = kWalk ? (uint8_t)((ms + w_phase) & (LC_SLOT_LEN - 1u))   // control threads only (w_phase is theirs)
                                    : (uint8_t)(ms % LC_SLOT_LEN);
        if (kProf) c0 = clock64();
        if (worker || edge_warp) {                          // phase 2: the carrier phase of each word selects its I and Q count
            GPSB_TL(0);
            uint32_t acc[3] = {0u, 0u, 0u};
            if (plain) ec_epl_phase2(sm.rq.acc0, sm.rq.step32, w0, kLoopNw, &part, acc);
            else if (edge) {
                const uint32_t v = ec_epl_edge_phase2(sm.rq.acc0, sm.rq.step32, edge_w, edge_neg, edge_counts);
                const int a = edge_arm;
                acc[0] = a == 0 ? v : 0u;
                acc[1] = a == 1 ? v : 0u;
                acc[2] = a == 2 ? v : 0u;
            }
            long long c1 = 0, c2 = 0;
            if (kProf) { c1 = clock64(); c1 += (long long)(acc[0] & 0u); }
            GPSB_TL(1);
            uint32_t v[3];
#pragma unroll
            for (int a = 0; a < 3; a++) v[a] = __reduce_add_sync(0xFFFFFFFFu, acc[a]);
            if (kProf) { c2 = clock64(); c2 += (long long)(v[0] & 0u); }
            GPSB_TL(2);
            if (kSlots) {
                if (lane == 0) sm.partial[b][warp] = make_uint4(v[0], v[1], v[2], 0u);     // one 16-byte store into the warp's own slot
            } else if (lane == 0) {
                uint32_t* acc_s = reinterpret_cast<uint32_t*>(&sm.sums[b]);
                atomicAdd(acc_s + 0, v[0]);
                atomicAdd(acc_s + 1, v[1]);
                atomicAdd(acc_s + 2, v[2]);
            }
            if (kProf && wtid == 0 && worker) { const long long c3 = clock64(); pt[0] += c3 - c0; pt[7] += c1 - c0; pt[8] += c2 - c1; pt[9] += c3 - c2; }
            GPSB_TL(3);
        }
        if (kProf && carrier_thr) c0 = clock64();
        __syncthreads();   // A: the six sums of millisecond m are complete
        // kSlots: every warp forms the six sums for itself - lane l < 16 takes slot l, three REDUX - so they sit in the registers
        // of every thread that needs them (control threads, and the workers for their own DLL): no contending atomics in front of
        // the barrier, no shared-memory round trip behind it, no buffer to clear.
        int16_t sums[6] = {0, 0, 0, 0, 0, 0};
        if (kSlots) {
            const uint4 pv = sm.partial[b][lane & 15];
            const bool mine = lane < 16;
            const uint32_t packed[3] = {__reduce_add_sync(0xFFFFFFFFu, mine ? pv.x : 0u), __reduce_add_sync(0xFFFFFFFFu, mine ? pv.y : 0u),
                                        __reduce_add_sync(0xFFFFFFFFu, mine ? pv.z : 0u)};
            ec_unpack_sums(packed, sums);
        }
        if (m > 0) {                    // the previous millisecond ended the run (written before this barrier) or, streaming, shortened it
            if (kStream) {
                const int2 f = sm.ctl[b];
                stop = f.x;
                limit = (uint32_t)f.y;
            } else {
                stop = sm.ctl[b].x;
            }
            if (stop != LC_STOP_NONE) break;
        }
        // `more`: the control threads plan a successor millisecond.  The workers only ask whether a successor frame
        // exists (it has been fetched, so waiting for it and forming its phase 1 is harmless when the run is about to
        // end for lack of LATER frames) - they do not read the starvation flag, which keeps it off their path.
        GPSB_TL(4);
        const bool next_frame = m + 1 < limit;
        if (worker || edge_warp) {
            if (kProf) c0 = clock64();
            if (kWorkerDll) {
                if (next_frame && (plain || edge)) {
                    int16_t iq[6];
                    if (kSlots) { for (int k = 0; k < 6; k++) iq[k] = sums[k]; } else load_sums(&sm.sums[b], iq);
                    // Walk build: a millisecond the channel leaves out moves no code phase (the sums are nobody's).  The gap is
                    // read from the record itself: the nav thread writes it at least LC_WALK_LEAD_MS before it begins and after
                    // the previous one has ended, so whichever of the two values a worker sees - or a mix of them - says "not
                    // idle" for the millisecond at hand.
                    const bool w_idle = kWalk && lc_walk_idle(*(volatile uint32_t*)&sm.aux.skip_ms, *(volatile uint8_t*)&sm.aux.skip_len, ms);
                    if (w_idle || !lc_dll_is_degenerate(iq)) {        // (a degenerate millisecond ends the run at the next barrier)
                        if (!w_idle) lc_dll_update(&wcod, iq[0], iq[1], iq[4], iq[5]);
                        gpsb_epl_req wrq;
                        lc_arm_offsets(wcod.code_phase_fine, &wrq);
                        mbar_wait(&sm.full[b ^ 1u], ((m + 1) >> 1) & 1u);
                        if (plain) {
                            const uint32_t off[3] = {wrq.off_e, wrq.off_p, wrq.off_l};
                            ec_epl_phase1(sm.S[b ^ 1u], sm.RX[wrq.off_bits & 7u], off, w0, kLoopNw, &part, sm.top_lut);
                        } else {
                            const uint32_t o = edge_arm == 0 ? wrq.off_e : edge_arm == 1 ? wrq.off_p : wrq.off_l;
                            edge_counts = ec_epl_edge_entry(sm.S[b ^ 1u], sm.RX[wrq.off_bits & 7u], o, edge_role, &edge_w, &edge_neg);
                        }
                    }
                    mbar_wait(&sm.nco_ready, m & 1u);       // the NCO words of millisecond m+1: on to its phase 2
                }
            } else if (next_frame && (plain || edge)) {     // phase 1 of millisecond m+1 as soon as its offsets exist
                mbar_wait(&sm.full[b ^ 1u], ((m + 1) >> 1) & 1u);
                mbar_wait(&sm.offs_ready, m & 1u);
                long long c1 = 0;
                if (kProf) c1 = clock64();
                GPSB_TL(5);
                if (sm.ctl[b ^ 1u].x == LC_STOP_NONE && !(kExp & 1)) {      // ordered after the code thread's write by offs_ready
                    if (plain) {
                        const uint32_t off[3] = {sm.rq.off_e, sm.rq.off_p, sm.rq.off_l};
                        ec_epl_phase1(sm.S[b ^ 1u], sm.RX[sm.rq.off_bits & 7u], off, w0, kLoopNw, &part, sm.top_lut);
                    } else {                                // one load: the offset of this lane's own arm (off_e, off_p, off_l are consecutive)
                        edge_counts = ec_epl_edge_entry(sm.S[b ^ 1u], sm.RX[sm.rq.off_bits & 7u], (&sm.rq.off_e)[edge_arm], edge_role,
                                                        &edge_w, &edge_neg);
                    }
                }
                if (kProf && lane == 0) { long long c2 = clock64(); c2 += (long long)((part.C[0][0] + edge_counts) & 0u); pt[13] += c2 - c1; }
                if (kProf && wtid == 0 && worker) pt[1] += clock64() - c0;
                GPSB_TL(6);
                mbar_wait(&sm.nco_ready, m & 1u);           // the NCO words of millisecond m+1: on to its phase 2
                GPSB_TL(7);
            }
        } else if (code_thr) {
            if (kProf) c0 = clock64();
            int16_t iq[6];
            if (kSlots) { for (int k = 0; k < 6; k++) iq[k] = sums[k]; } else load_sums(&sm.sums[b], iq);
            if (kProf) iq[0] += (int16_t)(clock64() & 0);
            GPSB_TL(5);
            if (!kSlots) sm.sums[b ^ 1u] = make_uint4(0u, 0u, 0u, 0u);      // read one ms ago by everybody; filled again after the workers have seen offs_ready
            if (next_frame) consumed++;                         // the workers wait for frame m+1 this millisecond
            const bool idle = kWalk && lc_walk_idle(w_skip_ms, w_skip_len, ms);            // the channel leaves this millisecond out
            const bool idle_next = kWalk && lc_walk_idle(w_skip_ms, w_skip_len, ms + 1u);
            if (idle) iq[0] = iq[1] = iq[2] = iq[3] = iq[4] = iq[5] = 0;          // ... the sums are nobody's
            const bool degenerate = !idle && lc_dll_is_degenerate(iq);   // 0/0 in the DLL: x86 and the GPU disagree on NaN bits, host finishes this ms
            if (degenerate) sm.ctl[b ^ 1u].x = LC_STOP_DLL_NAN;
            else if (!idle) {
                if (!(kExp & 4)) lc_dll_update(&cod, iq[0], iq[1], iq[4], iq[5]);
                if (next_frame) lc_plan_code(&cod, &sm.rq);
                sm.ch.tracking_data.code_phase_fine = cod.code_phase_fine;   // for lc_refine_edge
            }
            mbar_arrive(&sm.offs_ready);                    // DLL done: releases the offset check and the nav thread's edge refinement
            if (kProf) pt[2] += clock64() - c0;
            GPSB_TL(6);
            // frame buffer b (millisecond m) was consumed before this barrier: fetch millisecond m+2 into it.  A frame
            // that is missing is fetched all the same (the ring memory is there; nothing will use the result).
            if (m + 2 < limit) {
                if (kStream && !frame_present(gate, ms + 2, known_upto)) sm.ctl[b ^ 1u].y = (int)(m + 2);   // millisecond m+1 is the last
                tma_load_frame(sm.S[b], signal + (size_t)((ms + 2) % ring_ms) * kWords, GPSB_FRAME_BYTES, &sm.full[b]);
                issued = m + 3;
            }
            if (kStream && gate.progress && (m & 63u) == 63u) *(volatile uint32_t*)(gate.progress + chn) = ms;

            if (iq_log) {
                uint32_t* o = reinterpret_cast<uint32_t*>(iq_log + ((size_t)m * n_ch + chn) * 6);
                o[0] = (uint16_t)iq[0] | ((uint32_t)(uint16_t)iq[1] << 16);
                o[1] = (uint16_t)iq[2] | ((uint32_t)(uint16_t)iq[3] << 16);
                o[2] = (uint16_t)iq[4] | ((uint32_t)(uint16_t)iq[5] << 16);
            }
            if (degenerate) {
                gpsb_loop_result r;
                r.done_ms = m + done_base;
                r.stop = LC_STOP_DLL_NAN;
                for (int k = 0; k < 6; k++) r.iq[k] = iq[k];
                r.reserved = 0;
                results[chn] = r;
            }
            GPSB_TL(7);
            GPSB_WALK_STEP_END();
        } else if (carrier_thr) {
            if (kProf) { const long long c1 = clock64(); pt[6] += c1 - c0; c0 = c1; }
            int16_t iq[6];
            if (kSlots) { for (int k = 0; k < 6; k++) iq[k] = sums[k]; } else load_sums(&sm.sums[b], iq);
            if (kProf) iq[2] += (int16_t)(clock64() & 0);
            GPSB_TL(5);
            const bool idle = kWalk && lc_walk_idle(w_skip_ms, w_skip_len, ms);            // the channel leaves this millisecond out
            const bool idle_next = kWalk && lc_walk_idle(w_skip_ms, w_skip_len, ms + 1u);
            const bool live = !idle && !lc_dll_is_degenerate(iq);
            // no plan for a millisecond the channel leaves out: the last millisecond of the gap plans the one behind it
            // (now - prev_track_timestamp = gap + 1: lc_plan_carrier catches the NCO up, tracking.c:102-113).
            const bool plan = (live || idle) && next_frame && !idle_next;
            if (live) {
                // period_sync_ok_flag is written by the nav thread at slot index 3 and read here at slot index 0
                if (!(kExp & 2)) {
                    lc_pll_update(&car, sm.ch.nav_data.period_sync_ok_flag, index, iq[2], iq[3]);
                    lc_fll_update(&car, &sm.aux, found_freq_offset_hz, index, iq[2], iq[3], &angle_cache);
                }
            }
            if (kProf) car.if_freq_accum += (uint32_t)(clock64() & 0);
            GPSB_TL(6);
            if (plan) lc_plan_carrier(&car, prn, ms + 1, ms + 1, &sm.rq);
            if (next_frame) mbar_arrive(&sm.nco_ready);
            GPSB_TL(7);     // always: the workers wait for it whether or not a plan was made
            // Off the serial path: at slot index 1 the FLL needs the angle of THIS prompt sample as its "before" value;
            // evaluate it now, while the workers are busy, instead of next to the new angle in the next millisecond.
            if (index == 0 && live) {
                angle_cache.i = iq[2];
                angle_cache.q = iq[3];
                angle_cache.angle = lc_fll_angle(iq[2], iq[3]);
                angle_cache.valid = 1;
            }
            if (kProf) { const long long d = clock64() - c0; pt[3] += d; if (index == 0) { pt[10] += d; pt[11] += (iq[2] > 0); } }
            GPSB_WALK_STEP_END();
        } else if (nav_thr) {                               // nav bits and SNR of this millisecond (nav_data.c:46-453, tracking.c:154-169)
            if (kProf) c0 = clock64();
            int16_t iq[6];
            if (kSlots) { for (int k = 0; k < 6; k++) iq[k] = sums[k]; } else load_sums(&sm.sums[b], iq);
            int8_t bit = -1;
            const bool idle = kWalk && lc_walk_idle(w_skip_ms, w_skip_len, ms);            // the channel leaves this millisecond out
            const bool idle_next = kWalk && lc_walk_idle(w_skip_ms, w_skip_len, ms + 1u);
            if (!idle && !lc_dll_is_degenerate(iq)) {
                const int refine = lc_nav_new_code(&sm.ch, &sm.aux, index, iq[2], ms);
                bit = sm.aux.last_nav_bit;
                if (refine) {                               // reads the code phase the DLL has just produced
                    mbar_wait(&sm.offs_ready, m & 1u);
                    lc_refine_edge(&sm.ch, &sm.aux);
                }
                lc_snr_update(&sm.ch, &sm.aux, iq[2], iq[3]);
                if (kWalk && index == LC_SLOT_LEN - 1) lc_walk_policy(&sm.ch, &sm.aux, ms);     // end of a slot: move the slots?
            }
            if (idle && !idle_next) {                       // last millisecond of an idle gap: the next one starts a slot
                sm.aux.slot_phase = lc_walk_phase_after(ms);
                sm.aux.phase_since_ms = ms + 1u;
            }
            if (nav_log) nav_log[(size_t)m * n_ch + chn] = bit;
            if (kProf) pt[4] += clock64() - c0;
            GPSB_TL(7);
            GPSB_WALK_STEP_END();
        }
    }
    __syncthreads();                                        // the last millisecond's control work is done
    if (stop == LC_STOP_NONE) stop = sm.ctl[m & 1u].x;      // a stop raised by the very last millisecond (m == limit here)
    if (kStream && stop == LC_STOP_NONE && m < n_ms) stop = LC_STOP_STARVED;      // the run was shortened: the producer missed a frame
    if (kStream && code_thr && gate.progress) *(volatile uint32_t*)(gate.progress + chn) = ms0 + n_ms;   // needs no more frames
    if (code_thr)      // early exit: bulk copies nobody waited for may still be in flight - let them land before the CTA retires
        for (uint32_t f = consumed; f < issued; f++) mbar_wait(&sm.full[f & 1u], (f >> 1) & 1u);
    if (code_thr) {                                         // owned fields back into the shared record
        gps_tracking_t* t = &sm.ch.tracking_data;
        t->code_phase_fine = cod.code_phase_fine;
        t->dll_code_err = cod.dll_code_err;
#if (ENABLE_CODE_FILTER)
        t->code_filt_cnt = cod.code_filt_cnt;
        t->code_phase_fine_filt = cod.code_phase_fine_filt;
#endif
    } else if (carrier_thr) {
        gps_tracking_t* t = &sm.ch.tracking_data;
        t->if_freq_offset_hz = car.if_freq_offset_hz;
        t->if_freq_accum = car.if_freq_accum;
        t->prev_track_timestamp = car.prev_track_timestamp;
        t->pll_code_err = car.pll_code_err;
        t->fll_old_i = car.fll_old_i;
        t->fll_old_q = car.fll_old_q;
        t->fll_err = car.fll_err;
        t->pll_check_buf[0] = car.pll_check_buf[0];
        t->pll_check_buf[1] = car.pll_check_buf[1];
        t->pll_check_buf[2] = car.pll_check_buf[2];
        t->pll_check_buf[3] = car.pll_check_buf[3];
        t->pll_bad_state_cnt = car.pll_bad_state_cnt;
        t->pll_bad_state_master_cnt = car.pll_bad_state_master_cnt;
    }
    __syncthreads();
    if (kProf) {
        for (int i = tid; i < kLoopWarps * 4 * 8; i += kLoopThreads)
            prof[gridDim.x * 16 + (size_t)blockIdx.x * kLoopWarps * 4 * 8 + i] = tl_buf[i];
        if (worker && wtid == 0) {
            prof[chn * 16 + 0] = (unsigned long long)pt[0]; prof[chn * 16 + 1] = (unsigned long long)pt[1];
            prof[chn * 16 + 7] = (unsigned long long)pt[7]; prof[chn * 16 + 8] = (unsigned long long)pt[8];
            prof[chn * 16 + 9] = (unsigned long long)pt[9]; prof[chn * 16 + 12] = (unsigned long long)pt[13];
        }
        if (edge_warp && lane == 0) prof[chn * 16 + 13] = (unsigned long long)pt[13];
        if (code_thr) {
            prof[chn * 16 + 2] = (unsigned long long)pt[2];
            prof[chn * 16 + 5] = (unsigned long long)(clock64() - loop_begin);
        }
        if (carrier_thr) {
            prof[chn * 16 + 3] = (unsigned long long)pt[3]; prof[chn * 16 + 6] = (unsigned long long)pt[6];
            prof[chn * 16 + 10] = (unsigned long long)pt[10]; prof[chn * 16 + 11] = (unsigned long long)pt[11];
        }
        if (nav_thr) prof[chn * 16 + 4] = (unsigned long long)pt[4];
    }
    copy_words(chans + chn, &sm.ch, tid, kLoopThreads);
    copy_words(auxs + chn, &sm.aux, tid, kLoopThreads);
    if (code_thr && stop != LC_STOP_DLL_NAN) {
        gpsb_loop_result r;
        r.done_ms = m + done_base;
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
