/*
 * gpsb_cuda.cu - kernels + C ABI of libgpsb_cuda.so (see include/gpsb.h).
 *
 * Kernels (all sm_100a, integer LOP3/SHF/POPC work on bit-packed samples):
 *   k_expand_code   chips (1 byte each) -> 512-word expanded replica table      gps_misc.c:282-300
 *   k_gen_code      C/A Gold code from the PRN number on the device            gps_misc.c:317-372
 *   k_pack_iq2      ingest adaptor: byte-per-sample 2-bit I/Q -> packed I sign  signal_capture.c:9-16
 *   k_epl           one tracking integrate-and-dump per CTA                     tracking.c:115-138
 *   k_search        one acquisition / pre-track cell per CTA                    acquisition.c:282-294,
 *                                                                               :198-209, tracking.c:403-426
 *   k_l0_*          the bare reference primitives on explicit buffers           gps_misc.h:198-216
 *
 * There is deliberately no CPU implementation in this file: if CUDA is unavailable every compute
 * entry point returns GPSB_ERR_CUDA.
 */
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <pthread.h>
#include <new>

#include "gpsb_kernels.cuh"

namespace gpsb {
struct SweepParams {
    const uint32_t* sv_slots;
    const uint32_t* step32;
    uint32_t n_bins, ms0, n_ms, off_bits;
    // multi-GPU sharding of the (bin, ms) cell groups: CTA group k of this launch is group group0 + k * (group_skip + 1)
    // of the sweep; dense != 0: results go to res[k * n_sv + sv] (this rank's block of the all-gather) instead of the
    // (sv, bin, ms) grid.  All zero = the whole sweep on one GPU.
    uint32_t group0, group_skip, dense;
};
}  // namespace gpsb

#include "gpsb_acq_dp4a.cuh"
#include "gpsb_epl_batch.cuh"
#include "gpsb_track_loop.cuh"
#include "../core/gpsb_acq_core.h"

using namespace gpsb;

/* ============================================================================ kernels ========= */

// chips[1023] (0/1 bytes) -> E[512]; E[w] low half = chip 2w, high half = chip 2w+1 (0xFFFF each).
// S[k] (optional) = chips 4k..4k+3 as int8 +1 (chip 0) / -1 (chip 1), zero from chip 1022 on: the multiplier
// table of the dp4a search, whose whole-chip sum covers chips 0..1021 (gpsb_acq_dp4a.cuh).
__device__ __forceinline__ uint32_t signed_chip_word(const uint8_t* chips, int k)
{
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) {
        const int c = 4 * k + i;
        if (c < (int)GPSB_CHIPS - 1) v |= (chips[c] ? 0xFFu : 0x01u) << (8 * i);
    }
    return v;
}

__global__ void k_expand_code(const uint8_t* __restrict__ chips, uint32_t* __restrict__ E, uint32_t* __restrict__ S)
{
    for (int w = threadIdx.x; w < kWords; w += blockDim.x) {
        uint32_t lo = chips[2 * w] ? 0x0000FFFFu : 0u;
        uint32_t hi = (2 * w + 1 < (int)GPSB_CHIPS && chips[2 * w + 1]) ? 0xFFFF0000u : 0u;
        E[w] = lo | hi;
    }
    if (S)
        for (int k = threadIdx.x; k < kChipSteps; k += blockDim.x) S[k] = signed_chip_word(chips, k);
}

// G2 delays per PRN (IS-GPS-200 code phase assignments as chip delays; reference table
// gps_misc.c:319-341).
__constant__ uint16_t c_g2_delay[210] = {
    5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258, 469, 470, 471, 472,
    473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862, 863, 950, 947, 948, 950, 67, 103,
    91, 19, 679, 225, 625, 946, 638, 161, 1001, 554, 280, 710, 709, 775, 864, 558, 220, 397, 55,
    898, 759, 367, 299, 1018, 729, 695, 780, 801, 788, 732, 34, 320, 327, 389, 407, 525, 405, 221,
    761, 260, 326, 955, 653, 699, 422, 188, 438, 959, 539, 879, 677, 586, 153, 792, 814, 446, 264,
    1015, 278, 536, 819, 156, 957, 159, 712, 885, 461, 248, 713, 126, 807, 279, 122, 197, 693, 632,
    771, 467, 647, 203, 145, 175, 52, 21, 237, 235, 886, 657, 634, 762, 355, 1012, 176, 603, 130,
    359, 595, 68, 386, 797, 456, 499, 883, 307, 127, 211, 121, 118, 163, 628, 853, 484, 289, 811,
    202, 1021, 463, 568, 904, 670, 230, 911, 684, 309, 644, 932, 12, 314, 891, 212, 185, 675, 503,
    150, 395, 345, 846, 798, 992, 357, 995, 877, 112, 144, 476, 193, 109, 445, 291, 87, 399, 292,
    901, 339, 208, 711, 189, 263, 537, 663, 942, 173, 900, 30, 500, 935, 556, 373, 85, 652, 310};

// One CTA: two 10-stage LFSRs run by thread 0 (1023 serial steps), then all threads combine
// chip[i] = G1[i] ^ G2[(i - delay) mod 1023] and expand.
__global__ void k_gen_code(uint32_t prn, uint8_t* __restrict__ chips_out, uint32_t* __restrict__ E,
                           uint32_t* __restrict__ S)
{
    __shared__ uint8_t g1[GPSB_CHIPS], g2[GPSB_CHIPS], chip[GPSB_CHIPS + 1];
    if (threadIdx.x == 0) {
        uint32_t r1 = 0x3FFu, r2 = 0x3FFu;
        for (int i = 0; i < (int)GPSB_CHIPS; i++) {
            g1[i] = (r1 >> 9) & 1u;
            g2[i] = (r2 >> 9) & 1u;
            uint32_t f1 = ((r1 >> 2) ^ (r1 >> 9)) & 1u;
            uint32_t f2 = ((r2 >> 1) ^ (r2 >> 2) ^ (r2 >> 5) ^ (r2 >> 7) ^ (r2 >> 8) ^ (r2 >> 9)) & 1u;
            r1 = ((r1 << 1) | f1) & 0x3FFu;
            r2 = ((r2 << 1) | f2) & 0x3FFu;
        }
    }
    __syncthreads();
    const int d = c_g2_delay[prn - 1];
    for (int i = threadIdx.x; i <= (int)GPSB_CHIPS; i += blockDim.x) {
        uint8_t c = 0;
        if (i < (int)GPSB_CHIPS) {
            c = g1[i] ^ g2[(i + GPSB_CHIPS - d) % GPSB_CHIPS];
            chips_out[i] = c;
        }
        chip[i] = c;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < kWords; w += blockDim.x)
        E[w] = (chip[2 * w] ? 0x0000FFFFu : 0u) | (chip[2 * w + 1] ? 0xFFFF0000u : 0u);
    for (int k = threadIdx.x; k < kChipSteps; k += blockDim.x) S[k] = signed_chip_word(chip, k);
}

// E table -> chips (inverse of k_expand_code, for read-back).
__global__ void k_unexpand_code(const uint32_t* __restrict__ E, uint8_t* __restrict__ chips)
{
    for (int k = threadIdx.x; k < (int)GPSB_CHIPS; k += blockDim.x)
        chips[k] = (E[k >> 1] >> ((k & 1) * 16)) & 1u;
}

// Ingest adaptor: one byte per sample (bit0 = I sign) -> packed LSB-first frames.  Each thread builds
// one 32-bit output word from 32 consecutive sample bytes (two 16-byte loads); a warp therefore reads
// 1 KiB contiguous and writes 128 B contiguous.  HBM-bound: 8.125 bytes moved per output byte.
__global__ void k_pack_iq2(const uint4* __restrict__ samples, uint32_t* __restrict__ frames,
                           uint32_t frame0, uint32_t ring_ms, uint32_t n_ms)
{
    const uint32_t words_per_ms = kMixWords + 1;  // 512 words, the last one holds 16 samples + pad
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total = n_ms * words_per_ms;
    for (; gid < total; gid += gridDim.x * blockDim.x) {
        uint32_t m = gid / words_per_ms, w = gid % words_per_ms;
        uint32_t nsamp = (w == words_per_ms - 1) ? 16u : 32u;
        const uint4* src = samples + ((size_t)m * GPSB_MS_SAMPLES + (size_t)w * 32u) / 16u;
        uint32_t out = 0;
#pragma unroll
        for (uint32_t q = 0; q < 2; q++) {
            if (q * 16u < nsamp) {
                uint4 v = __ldg(src + q);
                uint32_t lanes[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t x = lanes[j] & 0x01010101u;          // bit0 of each of 4 bytes
                    x = (x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xFu;  // gather to 4 bits
                    out |= x << (q * 16u + j * 4u);
                }
            }
        }
        frames[(size_t)((frame0 + m) % ring_ms) * kWords + w] = out;
    }
}

// ---------------------------------------------------------------------------------- tracking
// One CTA (128 threads) per request.  Thread t owns replica words t, t+128, t+256, t+384 for all
// three arms; sums are reduced with REDUX (warp) + a 4x6 shared-memory stage.
__device__ __forceinline__ void epl_cell(const gpsb_epl_req& rq, int16_t* __restrict__ out6,
                                         const uint32_t* __restrict__ codes, const uint32_t* __restrict__ signal,
                                         uint32_t ring_ms)
{
    __shared__ CellSmem s;
    __shared__ int partial[kEplThreads / 32][6];
    const int tid = threadIdx.x;

    stage_replica(s.R, codes + (size_t)rq.sv_slot * kWords, rq.off_bits & 15u, tid, kEplThreads);
    stage_mix(s.I, s.Q, signal + (size_t)(rq.ms_index % ring_ms) * kWords, rq.acc0, rq.step32, tid,
              kEplThreads);
    extend_period<kEplThreads>(s.I, s.Q, tid);

    const uint32_t offs[3] = {rq.off_e, rq.off_p, rq.off_l};
    int acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const uint32_t off = offs[a];
        const int x0 = (int)(off >> 2);
        const uint32_t sh = (off & 3u) * 8u;
#pragma unroll
        for (int i = 0; i < kWords / kEplThreads; i++) {
            const int W = tid + i * kEplThreads;
            const uint32_t m = word_mask(W, off);
            const uint32_t r = s.R[W];
            uint32_t vi = __funnelshift_r(s.I[x0 + W], s.I[x0 + W + 1], sh);
            uint32_t vq = __funnelshift_r(s.Q[x0 + W], s.Q[x0 + W + 1], sh);
            acc[2 * a] += __popc((vi ^ r) & m);
            acc[2 * a + 1] += __popc((vq ^ r) & m);
        }
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
        int v = __reduce_add_sync(0xFFFFFFFFu, acc[j]);
        if ((tid & 31) == 0) partial[tid >> 5][j] = v;
    }
    __syncthreads();
    if (tid < 6) {
        int v = 0;
#pragma unroll
        for (int w = 0; w < kEplThreads / 32; w++) v += partial[w][tid];
        out6[tid] = (int16_t)(v - kHalfSum);  // gps_misc.c:140-141
    }
}

__global__ void __launch_bounds__(kEplThreads)
k_epl(const gpsb_epl_req* __restrict__ reqs, int16_t* __restrict__ out,
      const uint32_t* __restrict__ codes, const uint32_t* __restrict__ signal, uint32_t ring_ms)
{
    const gpsb_epl_req rq = reqs[blockIdx.x];
    epl_cell(rq, out + (size_t)blockIdx.x * 6, codes, signal, ring_ms);
}

// ---------------------------------------------------------------------------------- closed-loop (1 kHz) paths
// The tracking loop is serial per channel: the NCO words of millisecond t+1 come out of the host loop
// filters fed with the sums of millisecond t (tracking.c:140-143).  What matters is therefore the round
// trip, not bandwidth.  Two forms, both exchanging with the host through MAPPED PINNED memory only:
//
//  k_epl_rt          one launch per millisecond; the requests ride in the kernel parameter space, every CTA
//                    writes its six sums plus the batch sequence number as ONE 16-byte store to host memory;
//                    the host spins on the sequence numbers (no memcpy nodes, no stream synchronise).
//  k_epl_session     persistent: one CTA per channel slot stays resident for a whole tracking run, polls its
//                    32-byte command slot in host memory, answers with the same 16-byte record.  No launch on
//                    the critical path at all.  Bounded by an idle time-out and a maximum lifetime (global
//                    timer), so it can never outlive a stalled or dead host.
constexpr int kRtMaxCells = 128;
constexpr uint32_t kRtStop = 0xFFFFFFFFu;
struct EplRtBatch {
    gpsb_epl_req rq[kRtMaxCells];
};
struct RtCmd {   // host -> device, 32 bytes; both 16-byte halves carry the sequence number
    uint32_t w[8];   // sv_slot, ms_index, acc0, seq | step32, off_e|off_p<<16, off_l|off_bits<<16, seq
};
struct RtRsp {   // device -> host, 16 bytes, written with a single store
    int16_t sums[6];
    uint32_t seq;
};

__device__ __forceinline__ void publish_sums(RtRsp* __restrict__ slot, const int16_t* sums6, uint32_t seq)
{
    uint4 v;
    v.x = (uint16_t)sums6[0] | ((uint32_t)(uint16_t)sums6[1] << 16);
    v.y = (uint16_t)sums6[2] | ((uint32_t)(uint16_t)sums6[3] << 16);
    v.z = (uint16_t)sums6[4] | ((uint32_t)(uint16_t)sums6[5] << 16);
    v.w = seq;
    // one 16-byte transaction: sums and tag arrive together
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(slot), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(kEplThreads)
k_epl_rt(const __grid_constant__ EplRtBatch batch, RtRsp* __restrict__ host_rsp,
         const uint32_t* __restrict__ codes, const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t seq)
{
    __shared__ int16_t sums[8];
    epl_cell(batch.rq[blockIdx.x], sums, codes, signal, ring_ms);
    __syncthreads();
    if (threadIdx.x == 0) publish_sums(host_rsp + blockIdx.x, sums, seq);
}

__device__ __forceinline__ uint64_t global_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint4 ld_host16(const volatile void* p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kEplThreads)
k_epl_session(const volatile RtCmd* __restrict__ host_cmd, RtRsp* __restrict__ host_rsp,
              const uint32_t* __restrict__ codes, const uint32_t* __restrict__ signal, uint32_t ring_ms,
              uint64_t idle_ns, uint64_t life_ns, uint32_t stagger_ns)
{
    __shared__ int16_t sums[8];
    __shared__ uint32_t cw[8];
    __shared__ uint32_t seen_seq;      // newest command tag any poller has seen
    __shared__ int go;
    const volatile RtCmd* my_cmd = host_cmd + blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Every slot carries its own sequence; what this slot answered last is in its response record, so a
    // re-launched kernel picks up exactly the commands that are still unanswered.
    uint32_t last_seq = 0;
    if (threadIdx.x == 0) {
        const uint4 r = ld_host16(host_rsp + blockIdx.x);
        seen_seq = r.w;
        go = 1;
    }
    __syncthreads();
    last_seq = seen_seq;
    const uint64_t t_begin = global_ns();
    uint64_t t_last = t_begin;
    for (;;) {
        // one poller per warp, staggered by a fraction of the PCIe read latency, so a new command is noticed
        // about a quarter of a read period after it lands instead of half a period
        if (lane == 0) {
            if (warp) {      // phase offset of this poller; gives up at once when another poller has the command
                const uint64_t t0 = global_ns();
                while (global_ns() - t0 < (uint64_t)stagger_ns * warp)
                    if (*(volatile uint32_t*)&seen_seq != last_seq) break;
            }
            for (;;) {
                if (*(volatile uint32_t*)&seen_seq != last_seq) break;       // another poller has it
                const uint4 a = ld_host16(&my_cmd->w[0]);
                const uint4 b = ld_host16(&my_cmd->w[4]);
                if (a.w == b.w && a.w != last_seq) {
                    if (atomicCAS(&seen_seq, last_seq, a.w) == last_seq) {
                        cw[0] = a.x; cw[1] = a.y; cw[2] = a.z; cw[3] = a.w;
                        cw[4] = b.x; cw[5] = b.y; cw[6] = b.z;
                        go = a.w != kRtStop;
                    }
                    break;
                }
                const uint64_t now = global_ns();
                if (now - t_last > idle_ns || now - t_begin > life_ns) {      // host gone quiet / lease over
                    if (atomicCAS(&seen_seq, last_seq, last_seq + 1u) == last_seq) go = 0;
                    break;
                }
            }
        }
        __syncthreads();
        if (!go) return;
        gpsb_epl_req rq;
        rq.sv_slot = cw[0];
        rq.ms_index = cw[1];
        rq.acc0 = cw[2];
        rq.step32 = cw[4];
        rq.off_e = (uint16_t)(cw[5] & 0xFFFFu);
        rq.off_p = (uint16_t)(cw[5] >> 16);
        rq.off_l = (uint16_t)(cw[6] & 0xFFFFu);
        rq.off_bits = (uint16_t)(cw[6] >> 16);
        last_seq = cw[3];
        if (rq.sv_slot != 0xFFFFFFFFu) {            // 0xFFFFFFFF = keep-alive for a slot without work this ms
            epl_cell(rq, sums, codes, signal, ring_ms);
            __syncthreads();
            if (threadIdx.x == 0) publish_sums(host_rsp + blockIdx.x, sums, last_seq);
        }
        t_last = global_ns();
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- search
// Block-wide (max key, sum) reduction; result valid in thread 0.
__device__ __forceinline__ void block_reduce_search(uint32_t& key, int& total, uint32_t* sk, int* st)
{
    const int tid = threadIdx.x;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        uint32_t ok = __shfl_xor_sync(0xFFFFFFFFu, key, d);
        key = ok > key ? ok : key;
    }
    total = __reduce_add_sync(0xFFFFFFFFu, total);
    if ((tid & 31) == 0) {
        sk[tid >> 5] = key;
        st[tid >> 5] = total;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kSearchThreads / 32; w++) {
            key = sk[w] > key ? sk[w] : key;
            total += st[w];
        }
    }
}

// Shared tail of the search kernels: sweep offsets [start, stop), emit the reference's triple.
__device__ __forceinline__ void search_window(const CellSmem& s, uint32_t start, uint32_t stop,
                                              gpsb_search_res* res, int16_t* iq, uint32_t* sk, int* st)
{
    const int tid = threadIdx.x;
    uint32_t key = 0;  // (value << 16) | (0xFFFF - offset); 0 == "nothing above zero yet"
    int total = 0;
    for (uint32_t off = start + tid; off < stop; off += kSearchThreads) {
        int si, sq;
        corr_offset(s, off, si, sq);
        if (iq) {
            iq[2 * (off - start)] = (int16_t)(si - kHalfSum);
            iq[2 * (off - start) + 1] = (int16_t)(sq - kHalfSum);
        }
        int c = detector(si, sq);
        total += c;
        if (c > 0) {  // strict '>' from 0: first (lowest) offset wins ties, gps_misc.c:170
            uint32_t k = ((uint32_t)c << 16) | (0xFFFFu - off);
            key = k > key ? k : key;
        }
    }
    block_reduce_search(key, total, sk, st);
    if (tid == 0 && res) {
        gpsb_search_res r;
        r.max = (uint16_t)(key >> 16);
        r.phase = key ? (uint16_t)(0xFFFFu - (key & 0xFFFFu)) : 0;  // all-zero window -> phase 0
        r.avg = (uint16_t)(total / (2 * (int)GPSB_CHIPS));           // always /2046, gps_misc.c:178
        r.reserved = 0;
        *res = r;
    }
}

template <bool kSweep>
__global__ void __launch_bounds__(kSearchThreads)
k_search(const gpsb_search_req* __restrict__ reqs, SweepParams sp, gpsb_search_res* __restrict__ res,
         const uint32_t* __restrict__ res_map, int16_t* __restrict__ iq, const uint32_t* __restrict__ codes,
         const uint32_t* __restrict__ signal, uint32_t ring_ms)
{
    __shared__ CellSmem s;
    __shared__ uint32_t sk[kSearchThreads / 32];
    __shared__ int st[kSearchThreads / 32];
    const int tid = threadIdx.x;

    gpsb_search_req rq;
    if (kSweep) {  // cell index -> (sv, bin, ms)
        uint32_t c = blockIdx.x;
        uint32_t m = c % sp.n_ms;
        uint32_t b = (c / sp.n_ms) % sp.n_bins;
        uint32_t v = c / (sp.n_ms * sp.n_bins);
        rq.sv_slot = sp.sv_slots[v];
        rq.ms_index = sp.ms0 + m;
        rq.acc0 = 0;
        rq.step32 = sp.step32[b];
        rq.off_bits = (uint16_t)sp.off_bits;
        rq.start = 0;
        rq.stop = GPSB_OFFSETS;
        rq.flags = 0;
    } else {
        rq = reqs[blockIdx.x];
    }

    stage_replica(s.R, codes + (size_t)rq.sv_slot * kWords, rq.off_bits & 15u, tid, kSearchThreads);
    stage_mix(s.I, s.Q, signal + (size_t)(rq.ms_index % ring_ms) * kWords, rq.acc0, rq.step32, tid,
              kSearchThreads);
    extend_period<kSearchThreads>(s.I, s.Q, tid);
    search_window(s, rq.start, rq.stop, res + (res_map ? res_map[blockIdx.x] : blockIdx.x), iq, sk, st);
}

// ---------------------------------------------------------------------------------- device-resident pre-track
// k_pretrack_run: gps_pre_track_process (tracking.c:398-450) for a whole run of milliseconds in ONE launch, one CTA per
// channel: per millisecond the seven-offset search cell of k_search (replica at shift 0 staged once, stateless mixer at
// the acquisition's Doppler, gps_correlation8 over code_search_start + 7 * index ..), then the slot's running maximum,
// the collected winners and their mode (core/gpsb_loop_core.h, lc_pre_*: the same source the host path runs) by thread 0.
// A channel whose pre-track settles in millisecond k is left in GPS_PRE_TRACK_DONE with first[chn] = ms0 + k + 1, which is
// where k_track_run - launched behind this kernel on the same stream - takes it up; channels that are not in pre-track
// are not touched (first[chn] = ms0).  Replaces about a hundred per-millisecond launches and host round trips per cold start.
__global__ void __launch_bounds__(kSearchThreads)
k_pretrack_run(gps_ch_t* __restrict__ chans, gpsb_aux* __restrict__ auxs, const uint32_t* __restrict__ codes,
               const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t ms0, uint32_t n_ms,
               uint32_t* __restrict__ first)
{
    __shared__ CellSmem s;
    __shared__ uint32_t sk[kSearchThreads / 32];
    __shared__ int st[kSearchThreads / 32];
    __shared__ gps_ch_t ch;
    __shared__ gpsb_aux aux;
    __shared__ gpsb_search_res cell;
    __shared__ uint32_t win[4];          // start, stop, step32, go
    const int tid = threadIdx.x;
    const uint32_t chn = blockIdx.x;
    {
        const uint32_t state = chans[chn].tracking_data.state;      // uniform over the CTA
        if ((state != GPS_NEED_PRE_TRACK && state != GPS_PRE_TRACK_RUN) || auxs[chn].skip_len != 0) {
            if (tid == 0) first[chn] = ms0;
            return;
        }
    }
    copy_words(&ch, chans + chn, tid, kSearchThreads);
    copy_words(&aux, auxs + chn, tid, kSearchThreads);
    __syncthreads();
    stage_replica(s.R, codes + (size_t)ch.prn * kWords, 0u, tid, kSearchThreads);      // tracking.c:404: shift 0
    uint32_t done = n_ms;
    for (uint32_t m = 0; m < n_ms; m++) {
        const uint32_t ms = ms0 + m;
        const uint8_t index = (uint8_t)((ms + aux.slot_phase) & (LC_SLOT_LEN - 1u));
        if (tid == 0) {
            if (ch.tracking_data.state == GPS_NEED_PRE_TRACK) lc_pre_arm(&ch);
            uint16_t lo, hi;
            lc_pre_window(&ch.tracking_data, index, &lo, &hi);
            win[0] = lo;
            win[1] = hi;
            win[2] = lc_nco_step32((float)IF_FREQ_HZ + ch.tracking_data.if_freq_offset_hz);
        }
        __syncthreads();
        stage_mix(s.I, s.Q, signal + (size_t)(ms % ring_ms) * kWords, 0u, win[2], tid, kSearchThreads);
        extend_period<kSearchThreads>(s.I, s.Q, tid);
        if (tid == 0) cell = gpsb_search_res{0, 0, 0, 0};           // empty window: max 0, phase 0 (gps_misc.c:161-181)
        if (win[0] < win[1]) search_window(s, win[0], win[1], &cell, nullptr, sk, st);
        __syncthreads();
        if (tid == 0) {
            lc_pre_finish(&ch, &aux, index, cell.max, cell.phase);
            win[3] = ch.tracking_data.state == GPS_PRE_TRACK_RUN;
        }
        __syncthreads();
        if (!win[3]) {                    // settled (GPS_PRE_TRACK_DONE): tracking starts with the next millisecond
            done = m + 1;
            break;
        }
    }
    copy_words(chans + chn, &ch, tid, kSearchThreads);
    copy_words(auxs + chn, &aux, tid, kSearchThreads);
    if (tid == 0) first[chn] = ms0 + done;
}

// ---------------------------------------------------------------------------------- level 0
// ---------------------------------------------------------------------------------- device-resident code-phase rounds
// k_code_rounds_run: the code-phase narrowing rounds of the acquisition (acquisition.c:134-275) over a whole span of
// snapshots in ONE launch, one CTA per channel: per snapshot thread 0 runs the state machine (core/gpsb_acq_core.h, ac_*:
// the same source the host path runs), all threads the window's search cell (replica at shift 0 staged once, stateless
// mixer at the Doppler the sweep found, gps_correlation8 over [code_search_start, code_search_stop)), thread 0 the
// histogram vote.  mode[chn]: 0 = not this channel; 1 = snapshots ms0 .. until the channel's state has left busy_mask
// (at most n_ms); 2 = exactly n_ms snapshots (a channel that is served but does not keep the round alive).
// used[chn] = snapshots consumed.  Replaces one search launch + host round trip per look-ahead window.
__global__ void __launch_bounds__(kSearchThreads)
k_code_rounds_run(gps_ch_t* __restrict__ chans, gpsb_aux* __restrict__ auxs, const uint32_t* __restrict__ codes,
                  const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t ms0, uint32_t n_ms, uint32_t busy_mask,
                  const uint8_t* __restrict__ mode, uint32_t* __restrict__ used)
{
    __shared__ CellSmem s;
    __shared__ uint32_t sk[kSearchThreads / 32];
    __shared__ int st[kSearchThreads / 32];
    __shared__ gps_ch_t ch;
    __shared__ gpsb_aux aux;
    __shared__ gpsb_search_res cell;
    __shared__ uint32_t win[4];          // start, stop, search wanted, go on
    const int tid = threadIdx.x;
    const uint32_t chn = blockIdx.x;
    const uint32_t how = mode[chn];      // uniform over the CTA
    if (how == 0 || n_ms == 0) {
        if (tid == 0) used[chn] = 0;
        return;
    }
    copy_words(&ch, chans + chn, tid, kSearchThreads);
    copy_words(&aux, auxs + chn, tid, kSearchThreads);
    __syncthreads();
    stage_replica(s.R, codes + (size_t)ch.prn * kWords, 0u, tid, kSearchThreads);      // off_bits 0, acquisition.c:289
    // the Doppler does not change during the code rounds: int -> float at the call, acquisition.c:288
    const uint32_t step32 = lc_nco_step32((float)(IF_FREQ_HZ + ch.acq_data.found_freq_offset_hz));
    uint32_t done = 0;
    if (how == 1 && !((busy_mask >> (uint32_t)ch.acq_data.state) & 1u)) n_ms = 0;      // nothing to do for this one
    for (uint32_t m = 0; m < n_ms; m++) {
        const uint32_t ms = ms0 + m;
        if (tid == 0) {
            win[2] = (uint32_t)ac_code_plan(&ch, &aux, ms);
            win[0] = ch.acq_data.code_search_start;
            win[1] = ch.acq_data.code_search_stop;
            cell = gpsb_search_res{0, 0, 0, 0};                     // empty window: max 0, phase 0 (gps_misc.c:161-181)
        }
        __syncthreads();
        if (win[2] && win[0] < win[1]) {
            stage_mix(s.I, s.Q, signal + (size_t)(ms % ring_ms) * kWords, 0u, step32, tid, kSearchThreads);
            extend_period<kSearchThreads>(s.I, s.Q, tid);
            search_window(s, win[0], win[1], &cell, nullptr, sk, st);
        }
        __syncthreads();
        if (tid == 0) {
            if (win[2]) ac_finish_code_window(&ch, cell.phase, ms);
            win[3] = how == 2 || ((busy_mask >> (uint32_t)ch.acq_data.state) & 1u);
        }
        __syncthreads();
        done = m + 1;
        if (!win[3]) break;
    }
    copy_words(chans + chn, &ch, tid, kSearchThreads);
    copy_words(auxs + chn, &aux, tid, kSearchThreads);
    if (tid == 0) used[chn] = done;
}

__global__ void k_l0_replica(const uint32_t* __restrict__ E, uint32_t bits, uint32_t* __restrict__ out)
{
    for (int W = threadIdx.x; W < kWords; W += blockDim.x) out[W] = replica_word(E, W, bits & 15u);
}

__global__ void k_l0_mix(const uint32_t* __restrict__ frame, uint32_t acc0, uint32_t step32,
                         uint32_t* __restrict__ out_i, uint32_t* __restrict__ out_q)
{
    for (int w = threadIdx.x; w < kMixWords; w += blockDim.x) {
        uint32_t sgn = frame[w];
        uint32_t ph = (acc0 + (uint32_t)w * step32) >> 30;
        out_i[w] = cos_pattern(ph) ^ sgn;
        out_q[w] = sin_pattern(ph) ^ sgn;
    }
}

// Explicit-buffer search: prn / data_i / data_q are 512-word device copies of the caller's 2046-byte
// arrays (last two bytes zero-padded).  Bytes 2044..2045 of the data are honoured as given.
__global__ void __launch_bounds__(kSearchThreads)
k_l0_search(const uint32_t* __restrict__ prn, const uint32_t* __restrict__ data_i,
            const uint32_t* __restrict__ data_q, uint32_t start, uint32_t stop,
            gpsb_search_res* __restrict__ res, int16_t* __restrict__ iq)
{
    __shared__ CellSmem s;
    __shared__ uint32_t sk[kSearchThreads / 32];
    __shared__ int st[kSearchThreads / 32];
    const int tid = threadIdx.x;
    for (int w = tid; w < kWords; w += kSearchThreads) {
        s.R[w] = prn[w];
        s.I[w] = data_i[w];
        s.Q[w] = data_q[w];
    }
    extend_period<kSearchThreads>(s.I, s.Q, tid);
    search_window(s, start, stop, res, iq, sk, st);
}

/* ============================================================================ host side ======= */

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(GPSB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

struct gpsb_ctx {
    int device = 0;
    uint32_t max_sv = 0, ring_ms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    uint32_t* d_codes = nullptr;   // max_sv x 512 words
    uint32_t* d_schips = nullptr;  // max_sv x 256 words: +-1 chip bytes for the dp4a search
    uint32_t* d_rxt = nullptr;     // max_sv x 16 x 16 x 1040 words (1 MB per slot): extended replica streams for k_epl_batch
    uint32_t epl_batch_min = 512;  // gpsb_track_epl_dev batches of at least this many cells go to k_epl_batch
    int epl_batch_kernel = GPSB_BATCH_TMA;   // which of the two batch kernels (gpsb_set_epl_batch_kernel)
    int n_sm = 148;
    int sweep_method = GPSB_SWEEP_DP4A;
    // closed-loop mailbox (mapped pinned host memory, see k_epl_rt / k_epl_session)
    RtCmd* h_cmd = nullptr;        // kRtMaxCells command slots, host view
    RtCmd* d_cmd = nullptr;        // device alias
    RtRsp* h_rsp = nullptr;        // kRtMaxCells response slots, host view
    RtRsp* d_rsp = nullptr;
    uint32_t rt_seq = 0;                 // tag of the launch-per-ms path (k_epl_rt)
    uint32_t slot_seq[128] = {};         // per-slot tags of the session path; each slot is driven by ONE thread
    pthread_mutex_t relaunch_lock = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_t call_lock = PTHREAD_MUTEX_INITIALIZER;   // serialises the staged (non-session) entry points
    int realtime = 1;
    cudaStream_t rt_stream = nullptr;   // the session kernel lives on its own stream
    uint32_t session_slots = 0;         // > 0 while a session kernel is (supposed to be) resident
    uint64_t session_relaunches = 0;
    uint8_t* d_code_set = nullptr; // not used on device; host mirror below
    uint8_t* code_set = nullptr;   // host: slot has a code
    uint32_t* d_signal = nullptr;  // ring_ms x 512 words
    // staging (grown on demand)
    void* h_stage = nullptr;       // pinned
    void* d_stage = nullptr;
    size_t stage_cap = 0;
    uint8_t* d_chips = nullptr;    // 1024 B scratch
    uint32_t* d_l0 = nullptr;      // 4 x 512 words scratch for level-0 calls
    cudaEvent_t ev_start[8] = {}, ev_stop[8] = {};
    uint64_t launches = 0;
    // streaming ingest (gpsb_stream_*): frames are DMA-ed on copy_stream while a k_track_run launch is in flight
    cudaStream_t copy_stream = nullptr;
    uint32_t* d_watermark = nullptr;     // first millisecond NOT yet uploaded (device memory, written by the copy stream)
    uint32_t* h_wm_ring = nullptr;       // pinned source values of the watermark copies (one slot per push, cycled)
    uint32_t wm_next = 0;
    uint32_t* h_progress = nullptr;      // mapped pinned: millisecond each channel of the running loop has reached
    uint32_t* d_progress = nullptr;
    uint32_t stream_timeout_ms = 2000;
    cudaEvent_t ev_reset = nullptr;
    AcqScratch* d_acq_scratch = nullptr; // where the two parity halves of a dp4a search cell meet
    size_t acq_scratch_cap = 0;          // records
    void* d_iq2 = nullptr;               // staging of gpsb_stream_push_iq2
    size_t iq2_cap = 0;
    // multi-GPU (gpsb_comm_*): one context per rank, NCCL communicator over NVLink / NVSwitch
    void* comm = nullptr;                // ncclComm_t
    int comm_rank = 0, comm_size = 1;
    gpsb_search_res* d_part = nullptr;   // this rank's dense block of a sharded sweep
    gpsb_search_res* d_all = nullptr;    // all ranks' blocks after the all-gather
    gpsb_search_res* d_grid = nullptr;   // the whole sweep in (sv, bin, ms) order
    size_t part_cap = 0, grid_cap = 0;   // records
    bool loop_open = false;              // between gpsb_track_loop_begin and _end (call_lock held)
    struct {
        void* channels; void* aux; gpsb_loop_result* results; int16_t* iq_log; int8_t* nav_log;
        size_t ch_b, aux_b, res_b, iq_b, nav_b, o_aux, o_res, o_iq, o_nav;
    } open_loop = {};
};
static const uint32_t kWmSlots = 4096;
static const uint32_t kAbortWord = 256;      // h_progress[kAbortWord] != 0: a streaming loop stops waiting for frames

static int ensure_stage(gpsb_ctx* c, size_t bytes)
{
    // Between gpsb_track_loop_begin and _end the staging buffers hold the loop's channel records and its queued copy back:
    // no other entry point may write or reallocate them (only gpsb_stream_* may be called then, include/gpsb.h).
    if (c->loop_open) return fail(GPSB_ERR_STATE, "a tracking loop is open on this context: only gpsb_stream_* until gpsb_track_loop_end");
    if (bytes <= c->stage_cap) return GPSB_OK;
    size_t cap = c->stage_cap ? c->stage_cap : (1u << 16);
    while (cap < bytes) cap *= 2;
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_stage) cudaFree(c->d_stage);
    c->h_stage = nullptr;
    c->d_stage = nullptr;
    c->stage_cap = 0;
    CU(cudaMallocHost(&c->h_stage, cap));
    CU(cudaMalloc(&c->d_stage, cap));
    c->stage_cap = cap;
    return GPSB_OK;
}

static int check_launch(gpsb_ctx* c, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GPSB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    c->launches++;
    return GPSB_OK;
}

struct CallGuard {   // the staged entry points share one pinned/device staging buffer and one stream
    pthread_mutex_t* m;
    explicit CallGuard(gpsb_ctx* c) : m(&c->call_lock) { pthread_mutex_lock(m); }
    ~CallGuard() { pthread_mutex_unlock(m); }
};

/* ------------------------------------------------------------------ closed-loop mailbox helpers */
static int session_launch(gpsb_ctx* c)
{
    // 250 ms without a command, or 20 s of life, and the kernel leaves on its own (it is re-launched on demand)
    k_epl_session<<<c->session_slots, kEplThreads, 0, c->rt_stream>>>(c->d_cmd, c->d_rsp, c->d_codes, c->d_signal,
                                                                     c->ring_ms, 250000000ull, 20000000000ull, 400u);
    return check_launch(c, "k_epl_session");
}

static uint32_t bump(uint32_t s)
{
    ++s;
    if (s == 0 || s == kRtStop) s = 1;
    return s;
}

// Wait until the response record of `slot` carries `seq`, then copy the six sums out.  `watch` is the stream
// whose work produces it; in session mode a finished stream means the resident kernel timed out (idle / lease)
// and a new one is started - by exactly one of the waiting threads.
static int wait_slot(gpsb_ctx* c, uint32_t slot, uint32_t seq, int16_t* out6, cudaStream_t watch, bool session)
{
    volatile RtRsp* r = c->h_rsp + slot;
    uint64_t spins = 0;
    while (r->seq != seq) {
        if ((++spins & 0x3FFFF) != 0) continue;
        if (session) pthread_mutex_lock(&c->relaunch_lock);
        cudaError_t q = cudaStreamQuery(watch);
        int rc = GPSB_OK;
        if (q != cudaErrorNotReady && r->seq != seq) {
            if (q != cudaSuccess) rc = fail(GPSB_ERR_CUDA, "closed-loop kernel failed: %s", cudaGetErrorString(q));
            else if (!session) rc = fail(GPSB_ERR_CUDA, "k_epl_rt finished without publishing cell %u", slot);
            else {
                c->session_relaunches++;
                rc = session_launch(c);
            }
        }
        if (session) pthread_mutex_unlock(&c->relaunch_lock);
        if (rc) return rc;
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    memcpy(out6, (const void*)r->sums, 12);
    return GPSB_OK;
}

// Write one command into a slot: payload first, then the tag in both 16-byte halves.
static uint32_t post_slot(gpsb_ctx* c, uint32_t slot, const gpsb_epl_req* rq)
{
    volatile uint32_t* w = c->h_cmd[slot].w;
    const uint32_t seq = c->slot_seq[slot] = bump(c->slot_seq[slot]);
    if (rq) {
        w[0] = rq->sv_slot;
        w[1] = rq->ms_index;
        w[2] = rq->acc0;
        w[4] = rq->step32;
        w[5] = (uint32_t)rq->off_e | ((uint32_t)rq->off_p << 16);
        w[6] = (uint32_t)rq->off_l | ((uint32_t)rq->off_bits << 16);
    } else {
        w[0] = 0xFFFFFFFFu;                          // keep-alive: the slot has no work this millisecond
    }
    __atomic_thread_fence(__ATOMIC_RELEASE);         // payload before tags (x86 keeps store order anyway)
    w[3] = seq;
    w[7] = seq;
    return seq;
}

static int session_exchange(gpsb_ctx* c, uint32_t n, const gpsb_epl_req* req, int16_t* out)
{
    uint32_t seq[kRtMaxCells];
    for (uint32_t i = 0; i < c->session_slots; i++) seq[i] = post_slot(c, i, i < n ? &req[i] : nullptr);
    for (uint32_t i = 0; i < n; i++) {
        int rc = wait_slot(c, i, seq[i], out + 6u * i, c->rt_stream, true);
        if (rc) return rc;
    }
    return GPSB_OK;
}

// Scratch records of the parity-split dp4a search, zeroed on the stream ahead of the launch that uses them.
static int acq_scratch(gpsb_ctx* c, size_t n_results)
{
    if (n_results > c->acq_scratch_cap) {
        size_t cap = c->acq_scratch_cap ? c->acq_scratch_cap : 8192;
        while (cap < n_results) cap *= 2;
        CU(cudaStreamSynchronize(c->stream));
        if (c->d_acq_scratch) cudaFree(c->d_acq_scratch);
        c->d_acq_scratch = nullptr;
        c->acq_scratch_cap = 0;
        CU(cudaMalloc(&c->d_acq_scratch, cap * sizeof(AcqScratch)));
        c->acq_scratch_cap = cap;
    }
    CU(cudaMemsetAsync(c->d_acq_scratch, 0, n_results * sizeof(AcqScratch), c->stream));
    return GPSB_OK;
}

// One launch of the byte-popcount / dp4a search: parity-split form (two 256-thread CTAs per cell group and per SM) by
// default, the undivided 512-thread form with GPSB_SWEEP_DP4A_FULL.  n_results = size of the result array.
template <int NSV, bool kSweep>
static int launch_acq(gpsb_ctx* c, dim3 grid, const AcqGroup* dg, SweepParams sp, uint32_t n_sv, gpsb_search_res* d_res,
                      size_t n_results)
{
    if (c->sweep_method == GPSB_SWEEP_DP4A_FULL) {
        k_acq_dp4a<NSV, kSweep, false><<<grid, kAcqThreads, sizeof(AcqSmem<NSV, false>), c->stream>>>(
            dg, sp, n_sv, d_res, c->d_codes, c->d_schips, c->d_signal, c->ring_ms, nullptr);
        return check_launch(c, "k_acq_dp4a(full)");
    }
    int rc = acq_scratch(c, n_results);
    if (rc) return rc;
    grid.x *= 2;                                           // blockIdx.x & 1 = parity
    k_acq_dp4a<NSV, kSweep, true><<<grid, kAcqThreads / 2, sizeof(AcqSmem<NSV, true>), c->stream>>>(
        dg, sp, n_sv, d_res, c->d_codes, c->d_schips, c->d_signal, c->ring_ms, c->d_acq_scratch);
    return check_launch(c, "k_acq_dp4a");
}

extern "C" {

uint32_t gpsb_abi_version(void) { return 1u; }
const char* gpsb_last_error(void) { return g_err; }
uint64_t gpsb_launch_count(const gpsb_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint32_t gpsb_ring_ms(const gpsb_ctx* ctx) { return ctx ? ctx->ring_ms : 0; }

int gpsb_create(gpsb_ctx** out, int device, uint32_t max_sv, uint32_t ring_ms)
{
    if (!out || max_sv == 0 || ring_ms == 0) return fail(GPSB_ERR_ARG, "gpsb_create: bad argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(GPSB_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(GPSB_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(GPSB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    gpsb_ctx* c = new (std::nothrow) gpsb_ctx();
    if (!c) return fail(GPSB_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->max_sv = max_sv;
    c->ring_ms = ring_ms;
    c->code_set = (uint8_t*)calloc(max_sv, 1);
    if (!c->code_set) { delete c; return fail(GPSB_ERR_NOMEM, "out of host memory"); }
    *out = c;  // from here on gpsb_destroy can clean up partial state
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(cudaMalloc(&c->d_codes, (size_t)max_sv * kWords * 4));
    CU(cudaMemset(c->d_codes, 0, (size_t)max_sv * kWords * 4));
    c->n_sm = prop.multiProcessorCount;
    CU(cudaMalloc(&c->d_rxt, (size_t)max_sv * kRxtShifts * kRxtCopies * kRxtWords * 4));
    CU(cudaMemset(c->d_rxt, 0, (size_t)max_sv * kRxtShifts * kRxtCopies * kRxtWords * 4));
    CU(cudaMalloc(&c->d_schips, (size_t)max_sv * kChipSteps * 4));
    CU(cudaMemset(c->d_schips, 0, (size_t)max_sv * kChipSteps * 4));
#define GPSB_ACQ_ATTR(N, SW)                                                                                                      \
    CU(cudaFuncSetAttribute(k_acq_dp4a<N, SW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AcqSmem<N, false>))); \
    CU(cudaFuncSetAttribute(k_acq_dp4a<N, SW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AcqSmem<N, true>)));
    GPSB_ACQ_ATTR(8, true) GPSB_ACQ_ATTR(8, false) GPSB_ACQ_ATTR(4, true) GPSB_ACQ_ATTR(4, false) GPSB_ACQ_ATTR(1, true) GPSB_ACQ_ATTR(1, false)
#undef GPSB_ACQ_ATTR
    CU(cudaMalloc(&c->d_signal, (size_t)ring_ms * GPSB_FRAME_BYTES));
    CU(cudaMemset(c->d_signal, 0, (size_t)ring_ms * GPSB_FRAME_BYTES));
    CU(cudaMalloc(&c->d_chips, 1024));
    {
        void* hp = nullptr;
        const size_t bytes = kRtMaxCells * (sizeof(RtCmd) + sizeof(RtRsp));
        CU(cudaHostAlloc(&hp, bytes, cudaHostAllocMapped));
        memset(hp, 0, bytes);
        void* dp = nullptr;
        CU(cudaHostGetDevicePointer(&dp, hp, 0));
        c->h_cmd = (RtCmd*)hp;
        c->d_cmd = (RtCmd*)dp;
        c->h_rsp = (RtRsp*)((uint8_t*)hp + kRtMaxCells * sizeof(RtCmd));
        c->d_rsp = (RtRsp*)((uint8_t*)dp + kRtMaxCells * sizeof(RtCmd));
        CU(cudaStreamCreateWithFlags(&c->rt_stream, cudaStreamNonBlocking));
    }
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_reset, cudaEventDisableTiming));
    CU(cudaMalloc(&c->d_watermark, 64));
    CU(cudaMemset(c->d_watermark, 0, 64));
    CU(cudaMallocHost(&c->h_wm_ring, kWmSlots * sizeof(uint32_t)));
    {
        void* hp = nullptr;
        CU(cudaHostAlloc(&hp, 320 * sizeof(uint32_t), cudaHostAllocMapped));       // 256 progress words + the abort word
        memset(hp, 0, 320 * sizeof(uint32_t));
        void* dp = nullptr;
        CU(cudaHostGetDevicePointer(&dp, hp, 0));
        c->h_progress = (uint32_t*)hp;
        c->d_progress = (uint32_t*)dp;
    }
    CU(cudaMalloc(&c->d_l0, 4 * kWords * 4 + 64));
    for (int i = 0; i < 8; i++) {
        CU(cudaEventCreate(&c->ev_start[i]));
        CU(cudaEventCreate(&c->ev_stop[i]));
    }
    int rc = ensure_stage(c, 1u << 16);
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    return GPSB_OK;
}

void gpsb_destroy(gpsb_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 8; i++) {
        if (c->ev_start[i]) cudaEventDestroy(c->ev_start[i]);
        if (c->ev_stop[i]) cudaEventDestroy(c->ev_stop[i]);
    }
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->d_codes) cudaFree(c->d_codes);
    if (c->d_schips) cudaFree(c->d_schips);
    if (c->d_rxt) cudaFree(c->d_rxt);
    if (c->d_signal) cudaFree(c->d_signal);
    if (c->d_chips) cudaFree(c->d_chips);
    if (c->session_slots) gpsb_session_end(c);
    if (c->rt_stream) cudaStreamDestroy(c->rt_stream);
    if (c->h_cmd) cudaFreeHost((void*)c->h_cmd);
    if (c->d_l0) cudaFree(c->d_l0);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_reset) cudaEventDestroy(c->ev_reset);
    if (c->d_iq2) cudaFree(c->d_iq2);
    if (c->comm) gpsb_comm_destroy(c);
    if (c->d_part) cudaFree(c->d_part);
    if (c->d_all) cudaFree(c->d_all);
    if (c->d_grid) cudaFree(c->d_grid);
    if (c->d_acq_scratch) cudaFree(c->d_acq_scratch);
    if (c->d_watermark) cudaFree(c->d_watermark);
    if (c->h_wm_ring) cudaFreeHost(c->h_wm_ring);
    if (c->h_progress) cudaFreeHost(c->h_progress);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    free(c->code_set);
    delete c;
}

int gpsb_set_stream(gpsb_ctx* c, void* cuda_stream)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return GPSB_OK;
}

int gpsb_synchronize(gpsb_ctx* c)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    CU(cudaStreamSynchronize(c->stream));
    return GPSB_OK;
}

int gpsb_timer_start(gpsb_ctx* c, uint32_t slot)
{
    if (!c || slot >= 8) return fail(GPSB_ERR_ARG, "bad timer slot");
    CU(cudaEventRecord(c->ev_start[slot], c->stream));
    return GPSB_OK;
}
int gpsb_timer_stop(gpsb_ctx* c, uint32_t slot)
{
    if (!c || slot >= 8) return fail(GPSB_ERR_ARG, "bad timer slot");
    CU(cudaEventRecord(c->ev_stop[slot], c->stream));
    return GPSB_OK;
}
int gpsb_timer_elapsed_ms(gpsb_ctx* c, uint32_t slot, float* ms)
{
    if (!c || slot >= 8 || !ms) return fail(GPSB_ERR_ARG, "bad timer slot");
    CU(cudaEventSynchronize(c->ev_stop[slot]));
    CU(cudaEventElapsedTime(ms, c->ev_start[slot], c->ev_stop[slot]));
    return GPSB_OK;
}

/* ------------------------------------------------------------------ resident data */
int gpsb_set_code(gpsb_ctx* c, uint32_t slot, const uint8_t chips[GPSB_CHIPS])
{
    if (!c || !chips) return fail(GPSB_ERR_ARG, "gpsb_set_code: null argument");
    if (slot >= c->max_sv) return fail(GPSB_ERR_ARG, "sv_slot %u out of range (max_sv %u)", slot, c->max_sv);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->d_chips, chips, GPSB_CHIPS, cudaMemcpyHostToDevice, c->stream));
    k_expand_code<<<1, 256, 0, c->stream>>>(c->d_chips, c->d_codes + (size_t)slot * kWords,
                                            c->d_schips + (size_t)slot * kChipSteps);
    int rc = check_launch(c, "k_expand_code");
    if (rc) return rc;
    k_build_rxt<<<32, 256, 0, c->stream>>>(c->d_codes + (size_t)slot * kWords, c->d_rxt + (size_t)slot * kRxtShifts * kRxtCopies * kRxtWords);
    rc = check_launch(c, "k_build_rxt");
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    c->code_set[slot] = 1;
    return GPSB_OK;
}

int gpsb_set_code_prn(gpsb_ctx* c, uint32_t slot, uint32_t prn)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    if (slot >= c->max_sv) return fail(GPSB_ERR_ARG, "sv_slot %u out of range (max_sv %u)", slot, c->max_sv);
    if (prn < 1 || prn > 210) return fail(GPSB_ERR_ARG, "prn %u out of range 1..210", prn);
    CU(cudaSetDevice(c->device));
    k_gen_code<<<1, 256, 0, c->stream>>>(prn, c->d_chips, c->d_codes + (size_t)slot * kWords,
                                         c->d_schips + (size_t)slot * kChipSteps);
    int rc = check_launch(c, "k_gen_code");
    if (rc) return rc;
    k_build_rxt<<<32, 256, 0, c->stream>>>(c->d_codes + (size_t)slot * kWords, c->d_rxt + (size_t)slot * kRxtShifts * kRxtCopies * kRxtWords);
    rc = check_launch(c, "k_build_rxt");
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    c->code_set[slot] = 1;
    return GPSB_OK;
}

int gpsb_get_code(gpsb_ctx* c, uint32_t slot, uint8_t chips[GPSB_CHIPS])
{
    if (!c || !chips) return fail(GPSB_ERR_ARG, "gpsb_get_code: null argument");
    if (slot >= c->max_sv) return fail(GPSB_ERR_ARG, "sv_slot %u out of range", slot);
    if (!c->code_set[slot]) return fail(GPSB_ERR_STATE, "no code set for slot %u", slot);
    CU(cudaSetDevice(c->device));
    k_unexpand_code<<<1, 256, 0, c->stream>>>(c->d_codes + (size_t)slot * kWords, c->d_chips);
    int rc = check_launch(c, "k_unexpand_code");
    if (rc) return rc;
    CU(cudaMemcpyAsync(chips, c->d_chips, GPSB_CHIPS, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return GPSB_OK;
}

static int upload_impl(gpsb_ctx* c, uint32_t ms0, uint32_t n_ms, const uint8_t* packed, bool sync)
{
    if (!c || !packed) return fail(GPSB_ERR_ARG, "gpsb_upload_signal: null argument");
    if (n_ms == 0) return GPSB_OK;
    if (n_ms > c->ring_ms) return fail(GPSB_ERR_ARG, "n_ms %u exceeds ring capacity %u", n_ms, c->ring_ms);
    CU(cudaSetDevice(c->device));
    uint32_t f0 = ms0 % c->ring_ms;
    uint32_t first = (f0 + n_ms <= c->ring_ms) ? n_ms : c->ring_ms - f0;
    // 2046-byte host rows -> 2048-byte device frames; the two pad bytes stay zero from creation.
    CU(cudaMemcpy2DAsync((uint8_t*)c->d_signal + (size_t)f0 * GPSB_FRAME_BYTES, GPSB_FRAME_BYTES, packed,
                         GPSB_MS_BYTES, GPSB_MS_BYTES, first, cudaMemcpyHostToDevice, c->stream));
    if (first < n_ms)
        CU(cudaMemcpy2DAsync((uint8_t*)c->d_signal, GPSB_FRAME_BYTES, packed + (size_t)first * GPSB_MS_BYTES,
                             GPSB_MS_BYTES, GPSB_MS_BYTES, n_ms - first, cudaMemcpyHostToDevice, c->stream));
    if (sync) CU(cudaStreamSynchronize(c->stream));
    return GPSB_OK;
}

int gpsb_upload_signal(gpsb_ctx* c, uint32_t ms0, uint32_t n_ms, const uint8_t* packed)
{
    return upload_impl(c, ms0, n_ms, packed, true);
}
int gpsb_upload_signal_async(gpsb_ctx* c, uint32_t ms0, uint32_t n_ms, const uint8_t* packed)
{
    return upload_impl(c, ms0, n_ms, packed, false);
}

int gpsb_upload_signal_iq2(gpsb_ctx* c, uint32_t ms0, uint32_t n_ms, const uint8_t* samples)
{
    if (!c || !samples) return fail(GPSB_ERR_ARG, "gpsb_upload_signal_iq2: null argument");
    if (n_ms == 0) return GPSB_OK;
    if (n_ms > c->ring_ms) return fail(GPSB_ERR_ARG, "n_ms %u exceeds ring capacity %u", n_ms, c->ring_ms);
    CU(cudaSetDevice(c->device));
    const uint32_t chunk_ms = 256;  // 4 MiB of samples per staging pass
    int rc = ensure_stage(c, (size_t)chunk_ms * GPSB_MS_SAMPLES + 64);
    if (rc) return rc;
    for (uint32_t done = 0; done < n_ms; done += chunk_ms) {
        uint32_t n = (n_ms - done < chunk_ms) ? n_ms - done : chunk_ms;
        CU(cudaMemcpyAsync(c->d_stage, samples + (size_t)done * GPSB_MS_SAMPLES, (size_t)n * GPSB_MS_SAMPLES,
                           cudaMemcpyHostToDevice, c->stream));
        uint32_t total = n * kWords;
        k_pack_iq2<<<(total + 255) / 256, 256, 0, c->stream>>>((const uint4*)c->d_stage, c->d_signal,
                                                               (ms0 + done) % c->ring_ms, c->ring_ms, n);
        rc = check_launch(c, "k_pack_iq2");
        if (rc) return rc;
        CU(cudaStreamSynchronize(c->stream));
    }
    return GPSB_OK;
}

int gpsb_download_signal(gpsb_ctx* c, uint32_t ms0, uint32_t n_ms, uint8_t* packed)
{
    if (!c || !packed) return fail(GPSB_ERR_ARG, "gpsb_download_signal: null argument");
    if (n_ms > c->ring_ms) return fail(GPSB_ERR_ARG, "n_ms %u exceeds ring capacity %u", n_ms, c->ring_ms);
    CU(cudaSetDevice(c->device));
    for (uint32_t m = 0; m < n_ms; m++) {
        uint32_t f = (ms0 + m) % c->ring_ms;
        CU(cudaMemcpyAsync(packed + (size_t)m * GPSB_MS_BYTES, (uint8_t*)c->d_signal + (size_t)f * GPSB_FRAME_BYTES,
                           GPSB_MS_BYTES, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return GPSB_OK;
}

/* ------------------------------------------------------------------ request validation */
static int check_epl(const gpsb_ctx* c, uint32_t n, const gpsb_epl_req* rq)
{
    for (uint32_t i = 0; i < n; i++) {
        if (rq[i].sv_slot >= c->max_sv) return fail(GPSB_ERR_ARG, "request %u: sv_slot %u out of range", i, rq[i].sv_slot);
        if (!c->code_set[rq[i].sv_slot]) return fail(GPSB_ERR_STATE, "request %u: no code set for slot %u", i, rq[i].sv_slot);
        if (rq[i].off_e >= GPSB_OFFSETS || rq[i].off_p >= GPSB_OFFSETS || rq[i].off_l >= GPSB_OFFSETS)
            return fail(GPSB_ERR_ARG, "request %u: byte offset out of range 0..2045", i);
        if (rq[i].off_bits > 15) return fail(GPSB_ERR_ARG, "request %u: off_bits %u > 15", i, rq[i].off_bits);
    }
    return GPSB_OK;
}

static int check_prompt(const gpsb_ctx* c, uint32_t n, const gpsb_epl_req* rq)     // only the prompt arm's fields matter
{
    for (uint32_t i = 0; i < n; i++) {
        if (rq[i].sv_slot >= c->max_sv) return fail(GPSB_ERR_ARG, "request %u: sv_slot %u out of range", i, rq[i].sv_slot);
        if (!c->code_set[rq[i].sv_slot]) return fail(GPSB_ERR_STATE, "request %u: no code set for slot %u", i, rq[i].sv_slot);
        if (rq[i].off_p >= GPSB_OFFSETS) return fail(GPSB_ERR_ARG, "request %u: byte offset out of range 0..2045", i);
        if (rq[i].off_bits > 15) return fail(GPSB_ERR_ARG, "request %u: off_bits %u > 15", i, rq[i].off_bits);
    }
    return GPSB_OK;
}

static int check_search(const gpsb_ctx* c, uint32_t n, const gpsb_search_req* rq)
{
    for (uint32_t i = 0; i < n; i++) {
        if (rq[i].sv_slot >= c->max_sv) return fail(GPSB_ERR_ARG, "request %u: sv_slot %u out of range", i, rq[i].sv_slot);
        if (!c->code_set[rq[i].sv_slot]) return fail(GPSB_ERR_STATE, "request %u: no code set for slot %u", i, rq[i].sv_slot);
        if (rq[i].stop > GPSB_OFFSETS) return fail(GPSB_ERR_ARG, "request %u: stop %u > 2046", i, rq[i].stop);
        if (rq[i].off_bits > 15) return fail(GPSB_ERR_ARG, "request %u: off_bits %u > 15", i, rq[i].off_bits);
    }
    return GPSB_OK;
}

/* ------------------------------------------------------------------ level 1 */
}   // extern "C"
// One warp per cell on a persistent grid: frames through the TMA ring (default) or staged in registers.
template <int kArms>
static int launch_batch(gpsb_ctx* c, uint32_t n, const gpsb_epl_req* d_req, int16_t* d_out)
{
    const uint32_t want = (n + kBatchThreads / 32 - 1) / (kBatchThreads / 32);
    if (c->epl_batch_kernel == GPSB_BATCH_TMA) {
        static bool attr_set[2] = {false, false};
        if (!attr_set[kArms == 3]) {
            CU(cudaFuncSetAttribute(k_epl_batch_tma<kArms>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BatchTmaSmem)));
            attr_set[kArms == 3] = true;
        }
        const uint32_t cap = (uint32_t)c->n_sm * (kArms == 3 ? 2 : kTmaCtasPerSm);
        k_epl_batch_tma<kArms><<<want < cap ? want : cap, kBatchThreads, sizeof(BatchTmaSmem), c->stream>>>(
            d_req, d_out, c->d_rxt, c->d_signal, c->ring_ms, n);
        return check_launch(c, kArms == 3 ? "k_epl_batch_tma" : "k_epl_batch_tma<prompt>");
    }
    const uint32_t cap = (uint32_t)c->n_sm * kBatchCtasPerSm;
    k_epl_batch<kArms><<<want < cap ? want : cap, kBatchThreads, 0, c->stream>>>(d_req, d_out, c->d_rxt, c->d_signal, c->ring_ms, n);
    return check_launch(c, kArms == 3 ? "k_epl_batch" : "k_epl_batch<prompt>");
}
extern "C" {
int gpsb_track_epl_dev(gpsb_ctx* c, uint32_t n, const gpsb_epl_req* d_req, int16_t* d_out)
{
    if (!c || !d_req || !d_out) return fail(GPSB_ERR_ARG, "gpsb_track_epl_dev: null argument");
    if (n == 0) return GPSB_OK;
    CU(cudaSetDevice(c->device));
    if (n >= c->epl_batch_min) {            // large batches: one warp per cell, persistent grid (gpsb_epl_batch.cuh)
        return launch_batch<3>(c, n, d_req, d_out);
    }
    k_epl<<<n, kEplThreads, 0, c->stream>>>(d_req, d_out, c->d_codes, c->d_signal, c->ring_ms);
    return check_launch(c, "k_epl");
}

int gpsb_set_epl_batch_kernel(gpsb_ctx* c, int kernel)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    if (kernel != GPSB_BATCH_TMA && kernel != GPSB_BATCH_REGISTERS) return fail(GPSB_ERR_ARG, "unknown batch kernel %d", kernel);
    c->epl_batch_kernel = kernel;
    return GPSB_OK;
}

int gpsb_set_epl_batch_min(gpsb_ctx* c, uint32_t n_cells)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    c->epl_batch_min = n_cells;
    return GPSB_OK;
}

int gpsb_prompt_iq_dev(gpsb_ctx* c, uint32_t n, const gpsb_epl_req* d_req, int16_t* d_out)
{
    if (!c || !d_req || !d_out) return fail(GPSB_ERR_ARG, "gpsb_prompt_iq_dev: null argument");
    if (n == 0) return GPSB_OK;
    CU(cudaSetDevice(c->device));
    return launch_batch<1>(c, n, d_req, d_out);
}

int gpsb_prompt_iq(gpsb_ctx* c, uint32_t n, const gpsb_epl_req* req, int16_t* out)
{
    if (!c || !req || !out) return fail(GPSB_ERR_ARG, "gpsb_prompt_iq: null argument");
    if (n == 0) return GPSB_OK;
    int rc = check_prompt(c, n, req);
    if (rc) return rc;
    CallGuard guard(c);
    const size_t req_b = (size_t)n * sizeof(gpsb_epl_req), out_b = (size_t)n * 4;
    const size_t out_off = (req_b + 255) & ~(size_t)255;
    rc = ensure_stage(c, out_off + out_b);
    if (rc) return rc;
    memcpy(c->h_stage, req, req_b);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, req_b, cudaMemcpyHostToDevice, c->stream));
    rc = gpsb_prompt_iq_dev(c, n, (const gpsb_epl_req*)c->d_stage, (int16_t*)((uint8_t*)c->d_stage + out_off));
    if (rc) return rc;
    CU(cudaMemcpyAsync((uint8_t*)c->h_stage + out_off, (uint8_t*)c->d_stage + out_off, out_b, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(out, (uint8_t*)c->h_stage + out_off, out_b);
    return GPSB_OK;
}

int gpsb_track_epl(gpsb_ctx* c, uint32_t n, const gpsb_epl_req* req, int16_t* out)
{
    if (!c || !req || !out) return fail(GPSB_ERR_ARG, "gpsb_track_epl: null argument");
    if (n == 0) return GPSB_OK;
    int rc = check_epl(c, n, req);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (c->session_slots && n <= c->session_slots) return session_exchange(c, n, req, out);
    if (c->realtime && !c->session_slots && n <= (uint32_t)kRtMaxCells) {
        // closed-loop fast path: one launch, no copies, completion by sequence numbers in mapped host memory
        static_assert(sizeof(EplRtBatch) <= 4096, "request batch must fit the kernel parameter space");
        EplRtBatch batch;
        memcpy(batch.rq, req, (size_t)n * sizeof(gpsb_epl_req));
        // tags must differ from whatever the response records hold: continue each slot's own sequence
        uint32_t seq = 0;
        for (uint32_t i = 0; i < n; i++) seq = c->slot_seq[i] > seq ? c->slot_seq[i] : seq;
        seq = bump(seq);
        for (uint32_t i = 0; i < n; i++) c->slot_seq[i] = seq;
        k_epl_rt<<<n, kEplThreads, 0, c->stream>>>(batch, c->d_rsp, c->d_codes, c->d_signal, c->ring_ms, seq);
        rc = check_launch(c, "k_epl_rt");
        if (rc) return rc;
        for (uint32_t i = 0; i < n; i++) {
            rc = wait_slot(c, i, seq, out + 6u * i, c->stream, false);
            if (rc) return rc;
        }
        return GPSB_OK;
    }
    CallGuard guard(c);
    size_t req_b = (size_t)n * sizeof(gpsb_epl_req), out_b = (size_t)n * 12;
    size_t out_off = (req_b + 255) & ~(size_t)255;
    rc = ensure_stage(c, out_off + out_b);
    if (rc) return rc;
    memcpy(c->h_stage, req, req_b);
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, req_b, cudaMemcpyHostToDevice, c->stream));
    rc = gpsb_track_epl_dev(c, n, (const gpsb_epl_req*)c->d_stage, (int16_t*)((uint8_t*)c->d_stage + out_off));
    if (rc) return rc;
    CU(cudaMemcpyAsync((uint8_t*)c->h_stage + out_off, (uint8_t*)c->d_stage + out_off, out_b,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(out, (uint8_t*)c->h_stage + out_off, out_b);
    return GPSB_OK;
}

/* ------------------------------------------------------------------ device-resident tracking loop */
void gpsb_track_loop_record_bytes(uint32_t* channel_bytes, uint32_t* aux_bytes)
{
    if (channel_bytes) *channel_bytes = (uint32_t)sizeof(gps_ch_t);
    if (aux_bytes) *aux_bytes = (uint32_t)sizeof(gpsb_aux);
}

static int track_loop_launch(gpsb_ctx* c, uint32_t n_ch, void* d_channels, void* d_aux, uint32_t ms0, uint32_t n_ms,
                             int16_t* d_iq_log, int8_t* d_nav_log, gpsb_loop_result* d_results, uint32_t flags,
                             const uint32_t* d_first = nullptr)
{
    if (!c || !d_channels || !d_aux || !d_results) return fail(GPSB_ERR_ARG, "gpsb_track_loop_dev: null argument");
    if (n_ch == 0) return GPSB_OK;
    // a resident run reads frames that are all in the ring already; a streaming run may be longer than the ring (the
    // producer refills it behind the consumer, gpsb_stream_progress)
    if (!(flags & GPSB_LOOP_STREAMING) && n_ms > c->ring_ms)
        return fail(GPSB_ERR_ARG, "run of %u ms exceeds the signal ring (%u ms)", n_ms, c->ring_ms);
    if ((flags & GPSB_LOOP_STREAMING) && c->ring_ms < 8) return fail(GPSB_ERR_ARG, "a streaming run needs a ring of at least 8 ms");
    if ((flags & GPSB_LOOP_STREAMING) && n_ch > 256) return fail(GPSB_ERR_ARG, "a streaming run carries at most 256 channels");
    CU(cudaSetDevice(c->device));
    StreamGate gate = {nullptr, nullptr, 0ull, nullptr};
    if (flags & GPSB_LOOP_STREAMING) {
        gate.watermark = c->d_watermark;
        gate.progress = c->d_progress;
        gate.abort = c->d_progress + kAbortWord;
        gate.timeout_ns = (unsigned long long)c->stream_timeout_ms * 1000000ull;
    }
    static const bool profile = getenv("GPSB_LOOP_PROFILE") != nullptr;     // diagnostic: per-phase clock64 ticks to stderr
    if (profile) {
        unsigned long long* d_prof = nullptr;
        const size_t prof_words = (size_t)n_ch * (16 + (size_t)kLoopWarps * 4 * 8);
        CU(cudaMalloc(&d_prof, prof_words * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(d_prof, 0, prof_words * sizeof(unsigned long long), c->stream));
        k_track_run<true><<<n_ch, kLoopThreads, 0, c->stream>>>((gps_ch_t*)d_channels, (gpsb_aux*)d_aux, c->d_codes,
                                                                c->d_signal, c->ring_ms, ms0, n_ms, d_iq_log, d_nav_log,
                                                                d_results, d_prof, gate, d_first);
        int rc = check_launch(c, "k_track_run<profile>");
        unsigned long long h[16] = {};
        CU(cudaMemcpyAsync(h, d_prof, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        static unsigned long long tl[kLoopWarps * 4 * 8];                   // time line of channel 0, milliseconds 500..503
        CU(cudaMemcpyAsync(tl, d_prof + (size_t)n_ch * 16, sizeof tl, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(d_prof);
        if (n_ms > 504 && getenv("GPSB_LOOP_TIMELINE")) {
            unsigned long long t0 = ~0ull;
            for (unsigned long long v : tl) if (v && v < t0) t0 = v;
            fprintf(stderr, "[k_track_run time line, channel 0, ms %u..%u: clock64 - first stamp; points 0 loop top / 1 phase-2 compute / "
                            "2 redux / 3 red issued / 4 after barrier A / 5 offsets seen (control: sums loaded) / 6 phase 1 done "
                            "(code: offsets out, carrier: filters done) / 7 nco seen (control: done)]\n", ms0 + 500, ms0 + 503);
            for (int k = 0; k < 4; k++)
                for (int w = 0; w < kLoopWarps; w++) {
                    fprintf(stderr, "  ms+%d warp %2d:", k, w);
                    for (int p = 0; p < 8; p++) {
                        const unsigned long long v = tl[(w * 4 + k) * 8 + p];
                        if (v) fprintf(stderr, " %6llu", v - t0); else fprintf(stderr, "      -");
                    }
                    fprintf(stderr, "\n");
                }
        }
        const double n = n_ms ? (double)n_ms : 1.0;
        fprintf(stderr, "[k_track_run profile, channel 0, ticks per ms] workers: phase2+reduce %.0f (compute %.0f, redux %.0f, store %.0f), "
                        "A -> phase 1 done %.0f (phase 1 alone: plain warp %.0f, edge warp %.0f) | code thread %.0f | nav thread %.0f | "
                        "carrier thread %.0f (index 0: %.0f, %.0f%% float branch; others %.0f; waited at A %.0f) | loop total %.0f\n",
                h[0] / n, h[7] / n, h[8] / n, h[9] / n, h[1] / n, h[12] / n, h[13] / n, h[2] / n, h[4] / n, h[3] / n, h[10] / (n / 4),
                100.0 * h[11] / (n / 4), (h[3] - h[10]) / (n * 3 / 4), h[6] / n, h[5] / n);
        return rc;
    }
#ifdef GPSB_LOOP_EXPERIMENTS
    static const int experiment = getenv("GPSB_LOOP_EXPERIMENT") ? atoi(getenv("GPSB_LOOP_EXPERIMENT")) : 0;
#define GPSB_EXP_LAUNCH(E)                                                                                               \
    if (experiment == E) {                                                                                               \
        k_track_run<false, E><<<n_ch, kLoopThreads, 0, c->stream>>>((gps_ch_t*)d_channels, (gpsb_aux*)d_aux, c->d_codes, \
                                                                    c->d_signal, c->ring_ms, ms0, n_ms, d_iq_log,        \
                                                                    d_nav_log, d_results, nullptr, gate, d_first);       \
        return check_launch(c, "k_track_run<experiment>");                                                              \
    }
    GPSB_EXP_LAUNCH(1) GPSB_EXP_LAUNCH(2) GPSB_EXP_LAUNCH(3) GPSB_EXP_LAUNCH(4) GPSB_EXP_LAUNCH(5) GPSB_EXP_LAUNCH(7)
#endif
    const bool fixed = (flags & GPSB_LOOP_FIXED_SLOTS) != 0;     // no channel walks its slots: the build without the walk
    if (flags & GPSB_LOOP_STREAMING) {
        if (fixed)
            k_track_run<false, 0, true, false><<<n_ch, kLoopThreads, 0, c->stream>>>(
                (gps_ch_t*)d_channels, (gpsb_aux*)d_aux, c->d_codes, c->d_signal, c->ring_ms, ms0, n_ms, d_iq_log, d_nav_log,
                d_results, nullptr, gate, d_first);
        else
            k_track_run<false, 0, true><<<n_ch, kLoopThreads, 0, c->stream>>>(
                (gps_ch_t*)d_channels, (gpsb_aux*)d_aux, c->d_codes, c->d_signal, c->ring_ms, ms0, n_ms, d_iq_log, d_nav_log,
                d_results, nullptr, gate, d_first);
        return check_launch(c, "k_track_run<streaming>");
    }
    if (fixed)
        k_track_run<false, 0, false, false><<<n_ch, kLoopThreads, 0, c->stream>>>(
            (gps_ch_t*)d_channels, (gpsb_aux*)d_aux, c->d_codes, c->d_signal, c->ring_ms, ms0, n_ms, d_iq_log, d_nav_log,
            d_results, nullptr, gate, d_first);
    else
        k_track_run<false><<<n_ch, kLoopThreads, 0, c->stream>>>((gps_ch_t*)d_channels, (gpsb_aux*)d_aux, c->d_codes,
                                                                 c->d_signal, c->ring_ms, ms0, n_ms, d_iq_log, d_nav_log,
                                                                 d_results, nullptr, gate, d_first);
    return check_launch(c, "k_track_run");
}

int gpsb_track_loop_dev(gpsb_ctx* c, uint32_t n_ch, void* d_channels, void* d_aux, uint32_t ms0, uint32_t n_ms,
                        int16_t* d_iq_log, int8_t* d_nav_log, gpsb_loop_result* d_results)
{
    return track_loop_launch(c, n_ch, d_channels, d_aux, ms0, n_ms, d_iq_log, d_nav_log, d_results, 0u);
}

int gpsb_track_loop_dev_ex(gpsb_ctx* c, uint32_t n_ch, void* d_channels, void* d_aux, uint32_t ms0, uint32_t n_ms,
                           int16_t* d_iq_log, int8_t* d_nav_log, gpsb_loop_result* d_results, uint32_t flags)
{
    return track_loop_launch(c, n_ch, d_channels, d_aux, ms0, n_ms, d_iq_log, d_nav_log, d_results, flags);
}

/* ------------------------------------------------------------------ streaming ingest */
int gpsb_stream_reset(gpsb_ctx* c, uint32_t ms_valid_upto)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copy_stream));
    c->h_wm_ring[0] = ms_valid_upto;
    c->wm_next = 1;
    CU(cudaMemcpyAsync(c->d_watermark, &c->h_wm_ring[0], 4, cudaMemcpyHostToDevice, c->copy_stream));
    // the loop launched next on the context stream must see this value: order the two streams on the device, not on the host
    CU(cudaEventRecord(c->ev_reset, c->copy_stream));
    CU(cudaStreamWaitEvent(c->stream, c->ev_reset, 0));
    for (int i = 0; i < 256; i++) c->h_progress[i] = ms_valid_upto;
    c->h_progress[kAbortWord] = 0;
    return GPSB_OK;
}

int gpsb_stream_push(gpsb_ctx* c, uint32_t ms0, uint32_t n_ms, const uint8_t* packed)
{
    if (!c || !packed) return fail(GPSB_ERR_ARG, "gpsb_stream_push: null argument");
    if (n_ms == 0) return GPSB_OK;
    if (n_ms > c->ring_ms) return fail(GPSB_ERR_ARG, "n_ms %u exceeds ring capacity %u", n_ms, c->ring_ms);
    CU(cudaSetDevice(c->device));
    if (c->wm_next + 1 >= kWmSlots) {                    // the pinned source slots of earlier pushes are about to be reused
        CU(cudaStreamSynchronize(c->copy_stream));
        c->wm_next = 0;
    }
    const uint32_t f0 = ms0 % c->ring_ms;
    const uint32_t first = (f0 + n_ms <= c->ring_ms) ? n_ms : c->ring_ms - f0;
    CU(cudaMemcpy2DAsync((uint8_t*)c->d_signal + (size_t)f0 * GPSB_FRAME_BYTES, GPSB_FRAME_BYTES, packed, GPSB_MS_BYTES,
                         GPSB_MS_BYTES, first, cudaMemcpyHostToDevice, c->copy_stream));
    if (first < n_ms)
        CU(cudaMemcpy2DAsync((uint8_t*)c->d_signal, GPSB_FRAME_BYTES, packed + (size_t)first * GPSB_MS_BYTES, GPSB_MS_BYTES,
                             GPSB_MS_BYTES, n_ms - first, cudaMemcpyHostToDevice, c->copy_stream));
    // same stream: the frames have landed when the watermark moves
    uint32_t* slot = &c->h_wm_ring[c->wm_next++];
    *slot = ms0 + n_ms;
    CU(cudaMemcpyAsync(c->d_watermark, slot, 4, cudaMemcpyHostToDevice, c->copy_stream));
    return GPSB_OK;
}

int gpsb_stream_push_iq2(gpsb_ctx* c, uint32_t ms0, uint32_t n_ms, const uint8_t* samples)
{
    if (!c || !samples) return fail(GPSB_ERR_ARG, "gpsb_stream_push_iq2: null argument");
    if (n_ms == 0) return GPSB_OK;
    if (n_ms > c->ring_ms) return fail(GPSB_ERR_ARG, "n_ms %u exceeds ring capacity %u", n_ms, c->ring_ms);
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_ms * GPSB_MS_SAMPLES;
    if (bytes + 64 > c->iq2_cap) {                       // staging for the byte-per-sample container, grown on demand
        CU(cudaStreamSynchronize(c->copy_stream));
        if (c->d_iq2) cudaFree(c->d_iq2);
        c->d_iq2 = nullptr;
        c->iq2_cap = 0;
        CU(cudaMalloc(&c->d_iq2, bytes + 64));
        c->iq2_cap = bytes + 64;
    }
    if (c->wm_next + 1 >= kWmSlots) {
        CU(cudaStreamSynchronize(c->copy_stream));
        c->wm_next = 0;
    }
    // one stream: the next push's copy into the staging buffer waits for this push's pack kernel
    CU(cudaMemcpyAsync(c->d_iq2, samples, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    const uint32_t total = n_ms * kWords;
    k_pack_iq2<<<(total + 255) / 256, 256, 0, c->copy_stream>>>((const uint4*)c->d_iq2, c->d_signal, ms0 % c->ring_ms, c->ring_ms, n_ms);
    int rc = check_launch(c, "k_pack_iq2");
    if (rc) return rc;
    uint32_t* slot = &c->h_wm_ring[c->wm_next++];
    *slot = ms0 + n_ms;
    CU(cudaMemcpyAsync(c->d_watermark, slot, 4, cudaMemcpyHostToDevice, c->copy_stream));
    return GPSB_OK;
}

int gpsb_stream_wait(gpsb_ctx* c)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copy_stream));
    return GPSB_OK;
}

uint32_t gpsb_stream_progress(const gpsb_ctx* c, uint32_t n_ch)
{
    if (!c || n_ch == 0) return 0;
    uint32_t lo = ((volatile uint32_t*)c->h_progress)[0];
    for (uint32_t i = 1; i < n_ch && i < 256; i++) {
        const uint32_t v = ((volatile uint32_t*)c->h_progress)[i];
        if ((int32_t)(v - lo) < 0) lo = v;
    }
    return lo;
}

int gpsb_stream_abort(gpsb_ctx* c)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    ((volatile uint32_t*)c->h_progress)[kAbortWord] = 1u;
    return GPSB_OK;
}

int gpsb_stream_copies_pending(gpsb_ctx* c)
{
    if (!c) return 0;
    return cudaStreamQuery(c->copy_stream) == cudaErrorNotReady ? 1 : 0;
}

int gpsb_stream_loop_running(gpsb_ctx* c)
{
    if (!c || !c->loop_open) return 0;
    return cudaStreamQuery(c->stream) == cudaErrorNotReady ? 1 : 0;
}

uint32_t gpsb_stream_timeout_ms(const gpsb_ctx* c) { return c ? c->stream_timeout_ms : 0; }

int gpsb_stream_set_timeout_ms(gpsb_ctx* c, uint32_t ms)
{
    if (!c || ms == 0) return fail(GPSB_ERR_ARG, "gpsb_stream_set_timeout_ms: bad argument");
    c->stream_timeout_ms = ms;
    return GPSB_OK;
}

int gpsb_track_loop_begin(gpsb_ctx* c, uint32_t n_ch, void* channels, uint32_t channel_bytes, void* aux,
                          uint32_t aux_bytes, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log,
                          gpsb_loop_result* results, uint32_t flags)
{
    if (!c || !channels || !aux || !results) return fail(GPSB_ERR_ARG, "gpsb_track_loop: null argument");
    if (channel_bytes != sizeof(gps_ch_t) || aux_bytes != sizeof(gpsb_aux))
        return fail(GPSB_ERR_ARG, "record sizes %u/%u do not match this library's %zu/%zu", channel_bytes, aux_bytes,
                    sizeof(gps_ch_t), sizeof(gpsb_aux));
    if (n_ch == 0) return fail(GPSB_ERR_ARG, "gpsb_track_loop: no channels");
    const gps_ch_t* ch = (const gps_ch_t*)channels;
    for (uint32_t i = 0; i < n_ch; i++) {
        if (ch[i].prn >= c->max_sv) return fail(GPSB_ERR_ARG, "channel %u: prn %u out of range (max_sv %u)", i, ch[i].prn, c->max_sv);
        if (!c->code_set[ch[i].prn]) return fail(GPSB_ERR_STATE, "channel %u: no code set for slot %u", i, ch[i].prn);
    }
    if (c->loop_open) return fail(GPSB_ERR_STATE, "gpsb_track_loop_begin: the previous loop has not been ended (gpsb_track_loop_end)");
    pthread_mutex_lock(&c->call_lock);                   // held until gpsb_track_loop_end: the staging buffers are in use
    auto bail = [c](int rc) { pthread_mutex_unlock(&c->call_lock); return rc; };
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    auto& o = c->open_loop;
    bool pretrack = false;               // some channel is still in pre-track: k_pretrack_run goes ahead of the loop
    for (uint32_t i = 0; i < n_ch; i++)
        pretrack |= ch[i].tracking_data.state == GPS_NEED_PRE_TRACK || ch[i].tracking_data.state == GPS_PRE_TRACK_RUN;
    if (pretrack && (flags & GPSB_LOOP_STREAMING)) pretrack = false;      // a streamed run starts from tracking channels
    o.channels = channels; o.aux = aux; o.results = results; o.iq_log = iq_log; o.nav_log = nav_log;
    o.ch_b = (size_t)n_ch * sizeof(gps_ch_t);
    o.aux_b = (size_t)n_ch * sizeof(gpsb_aux);
    o.res_b = (size_t)n_ch * sizeof(gpsb_loop_result);
    o.iq_b = iq_log ? (size_t)n_ms * n_ch * 12 : 0;
    o.nav_b = nav_log ? (size_t)n_ms * n_ch : 0;
    o.o_aux = up(o.ch_b); o.o_res = o.o_aux + up(o.aux_b); o.o_iq = o.o_res + up(o.res_b); o.o_nav = o.o_iq + up(o.iq_b);
    const size_t o_first = o.o_nav + up(o.nav_b);
    const size_t total = o_first + up((size_t)n_ch * 4);
    int rc = ensure_stage(c, total);
    if (rc) return bail(rc);
    uint8_t* h = (uint8_t*)c->h_stage;
    uint8_t* d = (uint8_t*)c->d_stage;
    memcpy(h, channels, o.ch_b);
    memcpy(h + o.o_aux, aux, o.aux_b);
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d, h, o.o_aux + o.aux_b, cudaMemcpyHostToDevice, c->stream);   // records: one copy in
    if (e != cudaSuccess) return bail(fail(GPSB_ERR_CUDA, "record upload failed: %s", cudaGetErrorString(e)));
    // host records: the library sees for itself whether any channel walks its slots (core/gpsb_loop_core.h, lc_walk_*)
    {
        const gpsb_aux* ax = (const gpsb_aux*)aux;
        bool walks = false;
        for (uint32_t i = 0; i < n_ch; i++) walks |= ax[i].walk_enable || ax[i].slot_phase || ax[i].skip_len;
        flags = walks ? (flags & ~GPSB_LOOP_FIXED_SLOTS) : (flags | GPSB_LOOP_FIXED_SLOTS);
    }
    const uint32_t* d_first = nullptr;
    if (pretrack) {
        // rows of the logs a channel spends in pre-track stay blank (zero sums, no nav bit)
        if (o.iq_b) e = cudaMemsetAsync(d + o.o_iq, 0, o.iq_b, c->stream);
        if (e == cudaSuccess && o.nav_b) e = cudaMemsetAsync(d + o.o_nav, 0xFF, o.nav_b, c->stream);
        if (e != cudaSuccess) return bail(fail(GPSB_ERR_CUDA, "log clear failed: %s", cudaGetErrorString(e)));
        k_pretrack_run<<<n_ch, kSearchThreads, 0, c->stream>>>((gps_ch_t*)d, (gpsb_aux*)(d + o.o_aux), c->d_codes, c->d_signal,
                                                             c->ring_ms, ms0, n_ms, (uint32_t*)(d + o_first));
        rc = check_launch(c, "k_pretrack_run");
        if (rc) return bail(rc);
        d_first = (const uint32_t*)(d + o_first);
    }
    rc = track_loop_launch(c, n_ch, d, d + o.o_aux, ms0, n_ms, iq_log ? (int16_t*)(d + o.o_iq) : nullptr,
                           nav_log ? (int8_t*)(d + o.o_nav) : nullptr, (gpsb_loop_result*)(d + o.o_res), flags, d_first);
    if (rc) return bail(rc);
    const size_t back = (o.nav_b ? o.o_nav + o.nav_b : o.iq_b ? o.o_iq + o.iq_b : o.o_res + o.res_b);
    e = cudaMemcpyAsync(h, d, back, cudaMemcpyDeviceToHost, c->stream);                                       // records + logs: one copy out
    if (e != cudaSuccess) return bail(fail(GPSB_ERR_CUDA, "result download failed: %s", cudaGetErrorString(e)));
    c->loop_open = true;
    return GPSB_OK;
}

int gpsb_track_loop_end(gpsb_ctx* c)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    if (!c->loop_open) return fail(GPSB_ERR_STATE, "gpsb_track_loop_end without gpsb_track_loop_begin");
    c->loop_open = false;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    int rc = GPSB_OK;
    if (e != cudaSuccess) rc = fail(GPSB_ERR_CUDA, "k_track_run failed: %s", cudaGetErrorString(e));
    else {
        const auto& o = c->open_loop;
        const uint8_t* h = (const uint8_t*)c->h_stage;
        memcpy(o.channels, h, o.ch_b);
        memcpy(o.aux, h + o.o_aux, o.aux_b);
        memcpy(o.results, h + o.o_res, o.res_b);
        if (o.iq_log) memcpy(o.iq_log, h + o.o_iq, o.iq_b);
        if (o.nav_log) memcpy(o.nav_log, h + o.o_nav, o.nav_b);
    }
    pthread_mutex_unlock(&c->call_lock);
    return rc;
}

int gpsb_code_rounds(gpsb_ctx* c, uint32_t n_ch, void* channels, uint32_t channel_bytes, void* aux, uint32_t aux_bytes,
                     uint32_t ms0, uint32_t n_ms, uint32_t busy_mask, const uint8_t* mode, uint32_t* used)
{
    if (!c || !channels || !aux || !mode || !used) return fail(GPSB_ERR_ARG, "gpsb_code_rounds: null argument");
    if (channel_bytes != sizeof(gps_ch_t) || aux_bytes != sizeof(gpsb_aux))
        return fail(GPSB_ERR_ARG, "record sizes %u/%u do not match this library's %zu/%zu", channel_bytes, aux_bytes,
                    sizeof(gps_ch_t), sizeof(gpsb_aux));
    if (n_ch == 0) return fail(GPSB_ERR_ARG, "gpsb_code_rounds: no channels");
    if (n_ms > c->ring_ms) return fail(GPSB_ERR_ARG, "gpsb_code_rounds: %u snapshots exceed the ring (%u ms)", n_ms, c->ring_ms);
    const gps_ch_t* ch = (const gps_ch_t*)channels;
    for (uint32_t i = 0; i < n_ch; i++) {
        if (!mode[i]) continue;
        if (mode[i] > 2) return fail(GPSB_ERR_ARG, "channel %u: mode %u", i, mode[i]);
        if (ch[i].prn >= c->max_sv) return fail(GPSB_ERR_ARG, "channel %u: prn %u out of range (max_sv %u)", i, ch[i].prn, c->max_sv);
        if (!c->code_set[ch[i].prn]) return fail(GPSB_ERR_STATE, "channel %u: no code set for slot %u", i, ch[i].prn);
    }
    CallGuard guard(c);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t ch_b = (size_t)n_ch * sizeof(gps_ch_t), aux_b = (size_t)n_ch * sizeof(gpsb_aux);
    const size_t o_aux = up(ch_b), o_used = o_aux + up(aux_b), o_mode = o_used + up((size_t)n_ch * 4);
    const size_t total = o_mode + up(n_ch);
    int rc = ensure_stage(c, total);
    if (rc) return rc;
    uint8_t* h = (uint8_t*)c->h_stage;
    uint8_t* d = (uint8_t*)c->d_stage;
    memcpy(h, channels, ch_b);
    memcpy(h + o_aux, aux, aux_b);
    memcpy(h + o_mode, mode, n_ch);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(d, h, o_mode + n_ch, cudaMemcpyHostToDevice, c->stream));                  // records + modes: one copy in
    k_code_rounds_run<<<n_ch, kSearchThreads, 0, c->stream>>>((gps_ch_t*)d, (gpsb_aux*)(d + o_aux), c->d_codes, c->d_signal,
                                                            c->ring_ms, ms0, n_ms, busy_mask, d + o_mode, (uint32_t*)(d + o_used));
    rc = check_launch(c, "k_code_rounds_run");
    if (rc) return rc;
    CU(cudaMemcpyAsync(h, d, o_used + (size_t)n_ch * 4, cudaMemcpyDeviceToHost, c->stream));      // records + counts: one copy out
    CU(cudaStreamSynchronize(c->stream));
    memcpy(channels, h, ch_b);
    memcpy(aux, h + o_aux, aux_b);
    memcpy(used, h + o_used, (size_t)n_ch * 4);
    return GPSB_OK;
}

int gpsb_track_loop(gpsb_ctx* c, uint32_t n_ch, void* channels, uint32_t channel_bytes, void* aux,
                    uint32_t aux_bytes, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log,
                    gpsb_loop_result* results)
{
    if (c && n_ch == 0) return GPSB_OK;
    int rc = gpsb_track_loop_begin(c, n_ch, channels, channel_bytes, aux, aux_bytes, ms0, n_ms, iq_log, nav_log, results, 0u);
    if (rc) return rc;
    return gpsb_track_loop_end(c);
}

int gpsb_l0_loop_math(gpsb_ctx* c, int kind, int32_t ip_lo, uint32_t n_ip, float* out)
{
    if (!c || !out || (kind != 0 && kind != 1)) return fail(GPSB_ERR_ARG, "gpsb_l0_loop_math: bad argument");
    if (n_ip == 0) return GPSB_OK;
    if (ip_lo < -8184 || (int64_t)ip_lo + n_ip > 8185) return fail(GPSB_ERR_ARG, "ip range outside [-8184, 8184]");
    CallGuard guard(c);
    const size_t bytes = (size_t)n_ip * 16369u * sizeof(float);
    int rc = ensure_stage(c, bytes);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    k_l0_loop_math<<<148 * 8, 256, 0, c->stream>>>(kind, ip_lo, (int)n_ip, (float*)c->d_stage);
    rc = check_launch(c, "k_l0_loop_math");
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->h_stage, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(out, c->h_stage, bytes);
    return GPSB_OK;
}

int gpsb_search_dev(gpsb_ctx* c, uint32_t n, const gpsb_search_req* d_req, gpsb_search_res* d_res)
{
    if (!c || !d_req || !d_res) return fail(GPSB_ERR_ARG, "gpsb_search_dev: null argument");
    if (n == 0) return GPSB_OK;
    CU(cudaSetDevice(c->device));
    SweepParams sp = {};
    k_search<false><<<n, kSearchThreads, 0, c->stream>>>(d_req, sp, d_res, nullptr, nullptr, c->d_codes, c->d_signal,
                                                         c->ring_ms);
    return check_launch(c, "k_search");
}

// Requests whose window is at least this wide go through the dp4a kernel (which always evaluates all
// 2046 offsets at a cost of about 2046/8 direct offsets per satellite); narrower ones stay direct.
static const uint32_t kWideWindow = 384;

int gpsb_search(gpsb_ctx* c, uint32_t n, const gpsb_search_req* req, gpsb_search_res* res)
{
    if (!c || !req || !res) return fail(GPSB_ERR_ARG, "gpsb_search: null argument");
    if (n == 0) return GPSB_OK;
    int rc = check_search(c, n, req);
    if (rc) return rc;
    CallGuard guard(c);
    CU(cudaSetDevice(c->device));
    // staging layout: [direct requests][direct index map][groups][results by original index]
    const size_t req_b = (size_t)n * sizeof(gpsb_search_req), map_b = (size_t)n * 4;
    const size_t grp_b = (size_t)n * sizeof(AcqGroup), res_b = (size_t)n * sizeof(gpsb_search_res);
    const size_t map_off = (req_b + 255) & ~(size_t)255;
    const size_t grp_off = (map_off + map_b + 255) & ~(size_t)255;
    const size_t res_off = (grp_off + grp_b + 255) & ~(size_t)255;
    rc = ensure_stage(c, res_off + res_b);
    if (rc) return rc;
    uint8_t* h = (uint8_t*)c->h_stage;
    gpsb_search_req* h_req = (gpsb_search_req*)h;
    uint32_t* h_map = (uint32_t*)(h + map_off);
    AcqGroup* h_grp = (AcqGroup*)(h + grp_off);
    uint32_t n_direct = 0, n_grp = 0, max_in_group = 0;
    memset(h + res_off, 0, res_b);              // empty windows (start >= stop) report max 0, phase 0, avg 0
    for (uint32_t i = 0; i < n; i++) {
        const gpsb_search_req& r = req[i];
        if (r.start >= r.stop) continue;
        const bool wide = c->sweep_method != GPSB_SWEEP_DIRECT && (uint32_t)(r.stop - r.start) >= kWideWindow;
        if (!wide) {
            h_req[n_direct] = r;
            h_map[n_direct++] = i;
            continue;
        }
        AcqGroup* g = n_grp ? &h_grp[n_grp - 1] : nullptr;
        if (!g || g->n_sv >= 8 || g->ms_index != r.ms_index || g->acc0 != r.acc0 || g->step32 != r.step32 ||
            g->off_bits != r.off_bits) {
            g = &h_grp[n_grp++];
            memset(g, 0, sizeof *g);
            g->ms_index = r.ms_index;
            g->acc0 = r.acc0;
            g->step32 = r.step32;
            g->off_bits = r.off_bits;
        }
        g->sv_slot[g->n_sv] = r.sv_slot;
        g->start[g->n_sv] = r.start;
        g->stop[g->n_sv] = r.stop;
        g->res_index[g->n_sv] = i;
        g->n_sv++;
        if (g->n_sv > max_in_group) max_in_group = g->n_sv;
    }
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, res_off + res_b, cudaMemcpyHostToDevice, c->stream));
    uint8_t* d = (uint8_t*)c->d_stage;
    gpsb_search_res* d_res = (gpsb_search_res*)(d + res_off);
    if (n_direct) {
        SweepParams sp = {};
        k_search<false><<<n_direct, kSearchThreads, 0, c->stream>>>((const gpsb_search_req*)d, sp, d_res,
                                                                   (const uint32_t*)(d + map_off), nullptr,
                                                                   c->d_codes, c->d_signal, c->ring_ms);
        rc = check_launch(c, "k_search");
        if (rc) return rc;
    }
    if (n_grp) {
        SweepParams sp = {};
        const AcqGroup* dg = (const AcqGroup*)(d + grp_off);
        if (max_in_group > 4) rc = launch_acq<8, false>(c, dim3(n_grp), dg, sp, 0, d_res, n);
        else if (max_in_group > 1) rc = launch_acq<4, false>(c, dim3(n_grp), dg, sp, 0, d_res, n);
        else rc = launch_acq<1, false>(c, dim3(n_grp), dg, sp, 0, d_res, n);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(h + res_off, d + res_off, res_b, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(res, h + res_off, res_b);
    return GPSB_OK;
}

int gpsb_session_begin(gpsb_ctx* c, uint32_t n_slots)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    if (n_slots == 0 || n_slots > (uint32_t)kRtMaxCells) return fail(GPSB_ERR_ARG, "session of %u slots (1..%d)", n_slots, kRtMaxCells);
    if (c->session_slots) return fail(GPSB_ERR_STATE, "a tracking session is already open");
    if (getenv("GPSB_DISABLE_SESSION"))   // profilers replay kernels and cannot feed a resident one
        return fail(GPSB_ERR_STATE, "tracking sessions disabled by GPSB_DISABLE_SESSION");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));            // codes / frames uploaded so far are visible to the session
    for (uint32_t i = 0; i < n_slots; i++) {         // nothing pending: command tag == answered tag
        c->h_cmd[i].w[3] = c->h_cmd[i].w[7] = c->slot_seq[i];
        c->h_rsp[i].seq = c->slot_seq[i];
    }
    c->session_slots = n_slots;
    int rc = session_launch(c);
    if (rc) c->session_slots = 0;
    return rc;
}

int gpsb_session_end(gpsb_ctx* c)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    if (!c->session_slots) return GPSB_OK;
    for (uint32_t i = 0; i < c->session_slots; i++) {
        volatile uint32_t* w = c->h_cmd[i].w;
        w[3] = kRtStop;
        w[7] = kRtStop;
    }
    const uint32_t n = c->session_slots;
    c->session_slots = 0;
    CU(cudaStreamSynchronize(c->rt_stream));
    for (uint32_t i = 0; i < n; i++) c->h_cmd[i].w[3] = c->h_cmd[i].w[7] = c->slot_seq[i];
    return GPSB_OK;
}

/* Per-slot, thread-safe halves of a session exchange: each slot must be driven by one thread at a time, different
 * slots may be driven by different threads concurrently (gpsb_rx_track_run gives every worker its own channels). */
int gpsb_session_post(gpsb_ctx* c, uint32_t slot, const gpsb_epl_req* req, uint32_t* seq)
{
    if (!c || !seq) return fail(GPSB_ERR_ARG, "gpsb_session_post: null argument");
    if (slot >= c->session_slots) return fail(GPSB_ERR_STATE, "slot %u is not part of the open session", slot);
    if (req) {
        if (req->sv_slot >= c->max_sv || !c->code_set[req->sv_slot]) return fail(GPSB_ERR_STATE, "no code set for slot %u", req->sv_slot);
        if (req->off_e >= GPSB_OFFSETS || req->off_p >= GPSB_OFFSETS || req->off_l >= GPSB_OFFSETS || req->off_bits > 15)
            return fail(GPSB_ERR_ARG, "offset out of range");
    }
    *seq = post_slot(c, slot, req);
    return GPSB_OK;
}

int gpsb_session_wait(gpsb_ctx* c, uint32_t slot, uint32_t seq, int16_t out6[6])
{
    if (!c || !out6) return fail(GPSB_ERR_ARG, "gpsb_session_wait: null argument");
    if (slot >= c->session_slots) return fail(GPSB_ERR_STATE, "slot %u is not part of the open session", slot);
    return wait_slot(c, slot, seq, out6, c->rt_stream, true);
}

uint32_t gpsb_session_slots(const gpsb_ctx* c) { return c ? c->session_slots : 0; }

int gpsb_set_realtime(gpsb_ctx* c, int enabled)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    c->realtime = enabled ? 1 : 0;
    return GPSB_OK;
}

int gpsb_set_sweep_method(gpsb_ctx* c, int method)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    if (method != GPSB_SWEEP_DIRECT && method != GPSB_SWEEP_DP4A && method != GPSB_SWEEP_DP4A_FULL)
        return fail(GPSB_ERR_ARG, "unknown sweep method %d", method);
    c->sweep_method = method;
    return GPSB_OK;
}

int gpsb_search_iq(gpsb_ctx* c, const gpsb_search_req* req, int16_t* iq)
{
    if (!c || !req || !iq) return fail(GPSB_ERR_ARG, "gpsb_search_iq: null argument");
    int rc = check_search(c, 1, req);
    if (rc) return rc;
    if (req->start >= req->stop) return GPSB_OK;
    CU(cudaSetDevice(c->device));
    size_t n_off = (size_t)(req->stop - req->start);
    size_t res_off = 256, iq_off = 512, iq_b = n_off * 4;
    rc = ensure_stage(c, iq_off + iq_b);
    if (rc) return rc;
    memcpy(c->h_stage, req, sizeof *req);
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, sizeof *req, cudaMemcpyHostToDevice, c->stream));
    SweepParams sp = {};
    k_search<false><<<1, kSearchThreads, 0, c->stream>>>(
        (const gpsb_search_req*)c->d_stage, sp, (gpsb_search_res*)((uint8_t*)c->d_stage + res_off), nullptr,
        (int16_t*)((uint8_t*)c->d_stage + iq_off), c->d_codes, c->d_signal, c->ring_ms);
    rc = check_launch(c, "k_search(iq)");
    if (rc) return rc;
    CU(cudaMemcpyAsync((uint8_t*)c->h_stage + iq_off, (uint8_t*)c->d_stage + iq_off, iq_b,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(iq, (uint8_t*)c->h_stage + iq_off, iq_b);
    return GPSB_OK;
}

int gpsb_sweep_dev(gpsb_ctx* c, const uint32_t* d_sv_slots, uint32_t n_sv, const uint32_t* d_step32,
                   uint32_t n_bins, uint32_t ms0, uint32_t n_ms, uint32_t off_bits, gpsb_search_res* d_res)
{
    if (!c || !d_sv_slots || !d_step32 || !d_res) return fail(GPSB_ERR_ARG, "gpsb_sweep_dev: null argument");
    if (off_bits > 15) return fail(GPSB_ERR_ARG, "off_bits %u > 15", off_bits);
    uint64_t cells = (uint64_t)n_sv * n_bins * n_ms;
    if (cells == 0) return GPSB_OK;
    if (cells > 0x7FFFFFFFull) return fail(GPSB_ERR_ARG, "sweep of %llu cells is too large", (unsigned long long)cells);
    CU(cudaSetDevice(c->device));
    SweepParams sp = {d_sv_slots, d_step32, n_bins, ms0, n_ms, off_bits};
    if (c->sweep_method == GPSB_SWEEP_DIRECT) {
        k_search<true><<<(uint32_t)cells, kSearchThreads, 0, c->stream>>>(nullptr, sp, d_res, nullptr, nullptr,
                                                                        c->d_codes, c->d_signal, c->ring_ms);
        return check_launch(c, "k_search(sweep)");
    }
    // byte-popcount / dp4a search: per (bin, ms) x tile of up to 8 satellites one CTA per offset parity
    const uint32_t groups = n_bins * n_ms;
    if (n_sv > 4) return launch_acq<8, true>(c, dim3(groups, (n_sv + 7) / 8), nullptr, sp, n_sv, d_res, (size_t)cells);
    if (n_sv > 1) return launch_acq<4, true>(c, dim3(groups, 1), nullptr, sp, n_sv, d_res, (size_t)cells);
    return launch_acq<1, true>(c, dim3(groups, 1), nullptr, sp, n_sv, d_res, (size_t)cells);
}

int gpsb_sweep(gpsb_ctx* c, const uint32_t* sv_slots, uint32_t n_sv, const uint32_t* step32, uint32_t n_bins,
               uint32_t ms0, uint32_t n_ms, uint32_t off_bits, gpsb_search_res* res)
{
    if (!c || !sv_slots || !step32 || !res) return fail(GPSB_ERR_ARG, "gpsb_sweep: null argument");
    for (uint32_t i = 0; i < n_sv; i++) {
        if (sv_slots[i] >= c->max_sv) return fail(GPSB_ERR_ARG, "sv_slots[%u] = %u out of range", i, sv_slots[i]);
        if (!c->code_set[sv_slots[i]]) return fail(GPSB_ERR_STATE, "no code set for slot %u", sv_slots[i]);
    }
    size_t cells = (size_t)n_sv * n_bins * n_ms;
    if (cells == 0) return GPSB_OK;
    CU(cudaSetDevice(c->device));
    size_t sv_b = (size_t)n_sv * 4, st_b = (size_t)n_bins * 4;
    size_t st_off = (sv_b + 255) & ~(size_t)255;
    size_t res_off = (st_off + st_b + 255) & ~(size_t)255;
    size_t res_b = cells * sizeof(gpsb_search_res);
    int rc = ensure_stage(c, res_off + res_b);
    if (rc) return rc;
    memcpy(c->h_stage, sv_slots, sv_b);
    memcpy((uint8_t*)c->h_stage + st_off, step32, st_b);
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, st_off + st_b, cudaMemcpyHostToDevice, c->stream));
    rc = gpsb_sweep_dev(c, (const uint32_t*)c->d_stage, n_sv, (const uint32_t*)((uint8_t*)c->d_stage + st_off),
                        n_bins, ms0, n_ms, off_bits, (gpsb_search_res*)((uint8_t*)c->d_stage + res_off));
    if (rc) return rc;
    CU(cudaMemcpyAsync((uint8_t*)c->h_stage + res_off, (uint8_t*)c->d_stage + res_off, res_b,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(res, (uint8_t*)c->h_stage + res_off, res_b);
    return GPSB_OK;
}

/* ------------------------------------------------------------------ multi-GPU: sharded sweep + all-gather */
// NCCL is bound at run time (dlopen): a single-GPU user of the library needs no NCCL at all.  If the process has one
// loaded already (torch ships its own copy) that copy is used, so there is never a second NCCL in one process.
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, gpsb_nccl_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
pthread_mutex_t g_nccl_lock = PTHREAD_MUTEX_INITIALIZER;

int nccl_bind()
{
    pthread_mutex_lock(&g_nccl_lock);
    if (!g_nccl.lib) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // whatever the process already uses
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
            g_nccl.CommInitRank = (int (*)(void**, int, gpsb_nccl_id, int))dlsym(h, "ncclCommInitRank");
            g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
            g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
            g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
            if (g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather) g_nccl.lib = h;
        }
    }
    const bool ok = g_nccl.lib != nullptr;
    pthread_mutex_unlock(&g_nccl_lock);
    return ok ? GPSB_OK : fail(GPSB_ERR_STATE, "NCCL (libnccl.so.2) is not available: %s", dlerror() ? dlerror() : "symbols missing");
}
const char* nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"; }

// all[(r * n_local + k) * n_sv + v] (rank r's k-th cell group = group r + k * world) -> grid[(v * n_bins + b) * n_ms + m]
__global__ void k_unshard(const gpsb_search_res* __restrict__ all, gpsb_search_res* __restrict__ grid, uint32_t n_sv,
                          uint32_t n_bins, uint32_t n_ms, uint32_t world, uint32_t n_local)
{
    const uint32_t cells = n_sv * n_bins * n_ms;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
        const uint32_t g = i % (n_bins * n_ms), v = i / (n_bins * n_ms);
        grid[i] = all[((size_t)(g % world) * n_local + g / world) * n_sv + v];
    }
}
}  // namespace

int gpsb_comm_unique_id(gpsb_nccl_id* out)
{
    if (!out) return fail(GPSB_ERR_ARG, "gpsb_comm_unique_id: null argument");
    int rc = nccl_bind();
    if (rc) return rc;
    const int n = g_nccl.GetUniqueId(out);
    return n == 0 ? GPSB_OK : fail(GPSB_ERR_CUDA, "ncclGetUniqueId: %s", nccl_err(n));
}

int gpsb_comm_init(gpsb_ctx* c, int rank, int n_ranks, const gpsb_nccl_id* id)
{
    if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(GPSB_ERR_ARG, "gpsb_comm_init: bad argument");
    if (c->comm) return fail(GPSB_ERR_STATE, "the context already has a communicator");
    int rc = nccl_bind();
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    void* comm = nullptr;
    const int n = g_nccl.CommInitRank(&comm, n_ranks, *id, rank);
    if (n != 0) return fail(GPSB_ERR_CUDA, "ncclCommInitRank: %s", nccl_err(n));
    c->comm = comm;
    c->comm_rank = rank;
    c->comm_size = n_ranks;
    return GPSB_OK;
}

int gpsb_comm_destroy(gpsb_ctx* c)
{
    if (!c) return fail(GPSB_ERR_ARG, "null context");
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(c->comm);
        c->comm = nullptr;
    }
    c->comm_rank = 0;
    c->comm_size = 1;
    return GPSB_OK;
}

int gpsb_comm_rank(const gpsb_ctx* c) { return c ? c->comm_rank : 0; }
int gpsb_comm_size(const gpsb_ctx* c) { return c ? c->comm_size : 1; }

// Enqueue on the context stream: this rank's share of the (bin, ms) cell groups (group g belongs to rank g % size, so
// every rank keeps whole 8-satellite tiles), ncclAllGather of the dense blocks straight out of device memory, and the
// permutation into the (sv, bin, ms) grid.  No host copy anywhere; *d_grid_out = the grid in this context's memory.
int gpsb_sweep_gather_dev(gpsb_ctx* c, const uint32_t* d_sv_slots, uint32_t n_sv, const uint32_t* d_step32, uint32_t n_bins,
                          uint32_t ms0, uint32_t n_ms, uint32_t off_bits, gpsb_search_res** d_grid_out)
{
    if (!c || !d_sv_slots || !d_step32) return fail(GPSB_ERR_ARG, "gpsb_sweep_gather_dev: null argument");
    if (off_bits > 15) return fail(GPSB_ERR_ARG, "off_bits %u > 15", off_bits);
    if (c->sweep_method == GPSB_SWEEP_DIRECT) return fail(GPSB_ERR_STATE, "the sharded sweep runs the dp4a search only");
    const uint32_t world = (uint32_t)c->comm_size, rank = (uint32_t)c->comm_rank;
    if (world > 1 && !c->comm) return fail(GPSB_ERR_STATE, "no communicator: call gpsb_comm_init first");
    const uint32_t groups = n_bins * n_ms;
    const uint64_t cells = (uint64_t)n_sv * groups;
    if (cells == 0) return GPSB_OK;
    if (cells > 0x7FFFFFFFull) return fail(GPSB_ERR_ARG, "sweep of %llu cells is too large", (unsigned long long)cells);
    CU(cudaSetDevice(c->device));
    const uint32_t n_local = (groups + world - 1) / world;                 // block size, equal on every rank (padded)
    const uint32_t mine = groups > rank ? (groups - rank + world - 1) / world : 0;
    const size_t part = (size_t)n_local * n_sv;
    if (part > c->part_cap || cells > c->grid_cap) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->d_part) cudaFree(c->d_part);
        if (c->d_all) cudaFree(c->d_all);
        if (c->d_grid) cudaFree(c->d_grid);
        c->d_part = c->d_all = c->d_grid = nullptr;
        c->part_cap = c->grid_cap = 0;
        CU(cudaMalloc(&c->d_part, part * sizeof(gpsb_search_res)));
        CU(cudaMalloc(&c->d_all, part * world * sizeof(gpsb_search_res)));
        CU(cudaMalloc(&c->d_grid, (size_t)cells * sizeof(gpsb_search_res)));
        CU(cudaMemsetAsync(c->d_part, 0, part * sizeof(gpsb_search_res), c->stream));
        c->part_cap = part;
        c->grid_cap = (size_t)cells;
    }
    if (mine) {
        SweepParams sp = {d_sv_slots, d_step32, n_bins, ms0, n_ms, off_bits, rank, world - 1u, 1u};
        int rc;
        if (n_sv > 4) rc = launch_acq<8, true>(c, dim3(mine, (n_sv + 7) / 8), nullptr, sp, n_sv, c->d_part, part);
        else if (n_sv > 1) rc = launch_acq<4, true>(c, dim3(mine, 1), nullptr, sp, n_sv, c->d_part, part);
        else rc = launch_acq<1, true>(c, dim3(mine, 1), nullptr, sp, n_sv, c->d_part, part);
        if (rc) return rc;
    }
    const gpsb_search_res* all = c->d_part;
    if (world > 1) {
        const int n = g_nccl.AllGather(c->d_part, c->d_all, part * sizeof(gpsb_search_res), /* ncclChar */ 0, c->comm, c->stream);
        if (n != 0) return fail(GPSB_ERR_CUDA, "ncclAllGather: %s", nccl_err(n));
        all = c->d_all;
    }
    k_unshard<<<(unsigned)((cells + 255) / 256), 256, 0, c->stream>>>(all, c->d_grid, n_sv, n_bins, n_ms, world, n_local);
    int rc = check_launch(c, "k_unshard");
    if (rc) return rc;
    if (d_grid_out) *d_grid_out = c->d_grid;
    return GPSB_OK;
}

int gpsb_sweep_gather(gpsb_ctx* c, const uint32_t* sv_slots, uint32_t n_sv, const uint32_t* step32, uint32_t n_bins,
                      uint32_t ms0, uint32_t n_ms, uint32_t off_bits, gpsb_search_res* res)
{
    if (!c || !sv_slots || !step32 || !res) return fail(GPSB_ERR_ARG, "gpsb_sweep_gather: null argument");
    for (uint32_t i = 0; i < n_sv; i++) {
        if (sv_slots[i] >= c->max_sv) return fail(GPSB_ERR_ARG, "sv_slots[%u] = %u out of range", i, sv_slots[i]);
        if (!c->code_set[sv_slots[i]]) return fail(GPSB_ERR_STATE, "no code set for slot %u", sv_slots[i]);
    }
    const size_t cells = (size_t)n_sv * n_bins * n_ms;
    if (cells == 0) return GPSB_OK;
    CU(cudaSetDevice(c->device));
    const size_t sv_b = (size_t)n_sv * 4, st_b = (size_t)n_bins * 4;
    const size_t st_off = (sv_b + 255) & ~(size_t)255;
    const size_t res_off = (st_off + st_b + 255) & ~(size_t)255;
    const size_t res_b = cells * sizeof(gpsb_search_res);
    int rc = ensure_stage(c, res_off + res_b);
    if (rc) return rc;
    memcpy(c->h_stage, sv_slots, sv_b);
    memcpy((uint8_t*)c->h_stage + st_off, step32, st_b);
    CU(cudaMemcpyAsync(c->d_stage, c->h_stage, st_off + st_b, cudaMemcpyHostToDevice, c->stream));
    gpsb_search_res* d_grid = nullptr;
    rc = gpsb_sweep_gather_dev(c, (const uint32_t*)c->d_stage, n_sv, (const uint32_t*)((uint8_t*)c->d_stage + st_off), n_bins,
                               ms0, n_ms, off_bits, &d_grid);
    if (rc) return rc;
    CU(cudaMemcpyAsync((uint8_t*)c->h_stage + res_off, d_grid, res_b, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(res, (uint8_t*)c->h_stage + res_off, res_b);
    return GPSB_OK;
}

/* ------------------------------------------------------------------ level 0 */
int gpsb_l0_generate_prn_data2(gpsb_ctx* c, const uint8_t chips[GPSB_CHIPS], uint16_t* data, uint16_t offset_bits)
{
    if (!c || !chips || !data) return fail(GPSB_ERR_ARG, "gpsb_l0_generate_prn_data2: null argument");
    CU(cudaSetDevice(c->device));
    uint32_t* E = c->d_l0;
    uint32_t* R = c->d_l0 + kWords;
    CU(cudaMemcpyAsync(c->d_chips, chips, GPSB_CHIPS, cudaMemcpyHostToDevice, c->stream));
    k_expand_code<<<1, 256, 0, c->stream>>>(c->d_chips, E, nullptr);
    int rc = check_launch(c, "k_expand_code");
    if (rc) return rc;
    k_l0_replica<<<1, 256, 0, c->stream>>>(E, offset_bits, R);
    rc = check_launch(c, "k_l0_replica");
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->h_stage, R, kWords * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(data, c->h_stage, GPSB_MS_BYTES);  // 1023 words; word 1023 (spill) is the caller's
    return GPSB_OK;
}

int gpsb_l0_shift_to_zero_freq(gpsb_ctx* c, const uint8_t* signal_data, uint8_t* data_i, uint8_t* data_q,
                               uint32_t acc0, uint32_t step32, uint32_t* acc_out)
{
    if (!c || !signal_data || !data_i || !data_q) return fail(GPSB_ERR_ARG, "gpsb_l0_shift_to_zero_freq: null argument");
    CU(cudaSetDevice(c->device));
    uint32_t* S = c->d_l0;
    uint32_t* I = c->d_l0 + kWords;
    uint32_t* Q = c->d_l0 + 2 * kWords;
    memset(c->h_stage, 0, GPSB_FRAME_BYTES);
    memcpy(c->h_stage, signal_data, GPSB_MS_BYTES);
    CU(cudaMemcpyAsync(S, c->h_stage, GPSB_FRAME_BYTES, cudaMemcpyHostToDevice, c->stream));
    k_l0_mix<<<1, 256, 0, c->stream>>>(S, acc0, step32, I, Q);
    int rc = check_launch(c, "k_l0_mix");
    if (rc) return rc;
    uint8_t* h = (uint8_t*)c->h_stage + GPSB_FRAME_BYTES;
    CU(cudaMemcpyAsync(h, I, 2 * kWords * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(data_i, h, kMixWords * 4);                 // 2044 bytes, gps_misc.c:229
    memcpy(data_q, h + kWords * 4, kMixWords * 4);
    if (acc_out) *acc_out = acc0 + (uint32_t)kMixWords * step32;
    return GPSB_OK;
}

static int l0_search_impl(gpsb_ctx* c, const uint16_t* prn_p, const uint16_t* data_i, const uint16_t* data_q,
                          uint32_t start, uint32_t stop, gpsb_search_res* res, int16_t* iq_first)
{
    if (!c || !prn_p || !data_i || !data_q) return fail(GPSB_ERR_ARG, "level-0 correlator: null argument");
    if (stop > GPSB_OFFSETS) return fail(GPSB_ERR_ARG, "offset %u out of range", stop);
    CU(cudaSetDevice(c->device));
    uint8_t* h = (uint8_t*)c->h_stage;
    memset(h, 0, 3 * GPSB_FRAME_BYTES);
    memcpy(h, prn_p, GPSB_MS_BYTES);
    memcpy(h + GPSB_FRAME_BYTES, data_i, GPSB_MS_BYTES);
    memcpy(h + 2 * GPSB_FRAME_BYTES, data_q, GPSB_MS_BYTES);
    CU(cudaMemcpyAsync(c->d_l0, h, 3 * GPSB_FRAME_BYTES, cudaMemcpyHostToDevice, c->stream));
    gpsb_search_res* d_res = (gpsb_search_res*)(c->d_l0 + 3 * kWords);
    int16_t* d_iq = (int16_t*)(c->d_l0 + 3 * kWords + 4);
    // iq is only requested for single-offset windows here (fits the scratch)
    k_l0_search<<<1, kSearchThreads, 0, c->stream>>>(c->d_l0, c->d_l0 + kWords, c->d_l0 + 2 * kWords, start, stop,
                                                     d_res, iq_first ? d_iq : nullptr);
    int rc = check_launch(c, "k_l0_search");
    if (rc) return rc;
    uint8_t* hr = h + 3 * GPSB_FRAME_BYTES;
    CU(cudaMemcpyAsync(hr, d_res, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (res) memcpy(res, hr, sizeof *res);
    if (iq_first) memcpy(iq_first, hr + 16, 4);
    return GPSB_OK;
}

int gpsb_l0_correlation_iq(gpsb_ctx* c, const uint16_t* prn_p, const uint16_t* data_i, const uint16_t* data_q,
                           uint16_t offset, int16_t* res_i, int16_t* res_q)
{
    if (!res_i || !res_q) return fail(GPSB_ERR_ARG, "gpsb_l0_correlation_iq: null result pointer");
    if (offset >= GPSB_OFFSETS) return fail(GPSB_ERR_ARG, "offset %u out of range", offset);
    int16_t iq[2];
    int rc = l0_search_impl(c, prn_p, data_i, data_q, offset, offset + 1u, nullptr, iq);
    if (rc) return rc;
    *res_i = iq[0];
    *res_q = iq[1];
    return GPSB_OK;
}

int gpsb_l0_correlation8(gpsb_ctx* c, const uint16_t* prn_p, const uint16_t* data_i, const uint16_t* data_q,
                         uint16_t offset, int16_t* res)
{
    if (!res) return fail(GPSB_ERR_ARG, "gpsb_l0_correlation8: null result pointer");
    if (offset >= GPSB_OFFSETS) return fail(GPSB_ERR_ARG, "offset %u out of range", offset);
    gpsb_search_res r;
    int rc = l0_search_impl(c, prn_p, data_i, data_q, offset, offset + 1u, &r, nullptr);
    if (rc) return rc;
    *res = (int16_t)r.max;
    return GPSB_OK;
}

int gpsb_l0_correlation_search(gpsb_ctx* c, const uint16_t* prn_p, const uint16_t* data_i, const uint16_t* data_q,
                               uint16_t start_shift, uint16_t stop_shift, uint16_t* aver_val, uint16_t* phase,
                               uint16_t* max_val)
{
    gpsb_search_res r;
    int rc = l0_search_impl(c, prn_p, data_i, data_q, start_shift, stop_shift, &r, nullptr);
    if (rc) return rc;
    if (aver_val) *aver_val = r.avg;
    if (phase) *phase = r.phase;
    if (max_val) *max_val = r.max;
    return GPSB_OK;
}

}  // extern "C"
