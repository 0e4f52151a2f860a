/*
 * gpsb_epl_batch.cuh - k_epl_batch: open-loop integrate-and-dump over LARGE batches of cells, one WARP per cell.
 *
 * The closed loop (k_track_run) is serial per channel; a batch of cells whose NCO words and code offsets are all
 * known in advance - a recording replayed along a known trajectory, the reference's single-millisecond prompt
 * correlation (gps_correlation_iq, Firmware/project_main/GPS/gps_misc.c:128-145) repeated over a long recording,
 * SURVEY.md section 8(d) config 1 "batched" - has no such chain and is bound by how fast frames leave HBM
 * (2046 B per cell when every cell has its own millisecond) and by the POPC pipe (2 per word and arm).
 *
 *   frame     each lane fetches 4 x 16 bytes of the cell's frame straight into registers (coalesced 512-byte
 *             requests, L1 bypassed), one cell ahead of the one being correlated; nothing is staged in shared memory
 *   replica   the periodically extended replica stream of every (satellite slot, sub-byte shift) is resident in
 *             HBM/L2 (RXT, built when the code is set: per slot 16 shifts x 16 byte-displaced copies x 1040 words, two
 *             periods, 1 MB); the window of data word w for byte offset `off` is the word at byte 4w - off + 2046 of the
 *             copy displaced by that position's low four bits (no shift, no wrap to test for)
 *   carrier   closed form of gps_misc.c:229-239 per word; the quadrant patterns are one byte repeated (top byte of
 *             0x09999999 aside), so a pattern pair costs two PRMT byte broadcasts, shared by the arms
 *   edges     word 511 (never mixed; 16 or 8 replica bits against zero data) is special-cased in its unrolled slot;
 *             for odd offsets nine "edge" lanes then take back the three bytes the reference skips (gps_misc.c:59-89)
 *             - the bookkeeping of core/gpsb_epl_core.h in direct (I, Q) form
 *   reduce    REDUX per arm on packed I | Q << 16 sums, lane 0 stores the cell's int16 results
 *
 * kArms = 3: IE,QE,IP,QP,IL,QL per request (tracking.c:115-138); kArms = 1: the prompt arm only (I, Q).
 */
#pragma once

#include "../core/gpsb_epl_core.h"
#include "gpsb_kernels.cuh"

namespace gpsb {

constexpr int kRxtWords = 1040;         // two 2046-byte periods + the longest run a lane reads past them, rounded up to 16 bytes
constexpr int kRxtShifts = 16;
constexpr int kRxtCopies = 16;          // the stream displaced by 0..15 BYTES: a run starting at any byte is 16-byte aligned in one of them
constexpr int kBatchThreads = 256;
#ifndef GPSB_BATCH_CTAS
#define GPSB_BATCH_CTAS 2
#endif
constexpr int kBatchCtasPerSm = GPSB_BATCH_CTAS;

// RXT[slot][b][r][x] = bytes 4x+r .. 4x+r+3 (mod 2046) of the replica buffer gps_generate_prn_data2(b) produces
// (gps_misc.c:282-300: chip k at sample bits [16k+b, 16k+b+16), no wrap, the spill beyond 2046 bytes dropped).
// The sixteen copies r = 0..15 are the same stream displaced by r BYTES, so that a run starting at ANY byte of the stream
// starts 16-byte aligned in one of them: a lane fetches the replica windows of four data words with one 128-bit load and
// uses them as they are - no funnel shift per word, no fifth word from the neighbouring lane (round 2: the ALU pipe is
// what bounds the kernel, and the byte offset used to cost it one shift per word and arm).
__global__ void k_build_rxt(const uint32_t* __restrict__ E, uint32_t* __restrict__ rxt)
{
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < kRxtShifts * kRxtCopies * kRxtWords; i += blockDim.x * gridDim.x) {
        const uint32_t b = (uint32_t)(i / (kRxtCopies * kRxtWords));
        const int r = (i / kRxtWords) % kRxtCopies;
        const int x = i % kRxtWords;
        uint32_t v = 0;
        for (int j = 0; j < 4; j++) {
            const int byte = (4 * x + r + j) % (int)GPSB_MS_BYTES;
            const uint32_t w = ec_replica_word(E, byte >> 2, b);
            v |= ((w >> (8 * (byte & 3))) & 0xFFu) << (8 * j);
        }
        rxt[i] = v;
    }
}

// PRMT with the selector taken as it is (the __byte_perm intrinsic masks it first: two more instructions per word)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// cos / sin quadrant patterns of carrier phase ph = nco >> 30 (gps_misc.c:216-217) by byte broadcast: every pattern is
// one byte repeated, except that the top byte of 0x09999999 is 0x09 - bytes 0..2 come from one table word (selector
// nibbles ph), byte 3 from a second one (selector nibble 4 + ph)
__device__ __forceinline__ void quadrant_patterns(uint32_t nco, uint32_t& cp, uint32_t& sp)
{
    const uint32_t sel = (nco >> 30) * 0x1111u + 0x4000u;
    cp = prmt(0x3366CC99u, 0x3366CC09u, sel);     // cos[ph]
    sp = prmt(0x66CC9933u, 0x66CC0933u, sel);     // sin[ph] = cos[(ph + 3) & 3]
}

// The same pattern pair from a table in shared memory: 32 entries of (cos, sin) indexed by the top FIVE bits of the NCO
// word (every quadrant eight times), so the index is one shift; the address is formed by a multiply-add on the FMA pipe
// (`eight` holds 8 in a register the assembler cannot see through) and the pair arrives by one 8-byte load.  Per word
// that is 1 ALU + 1 FMA + 1 LSU instruction instead of 3 ALU + 1 FMA (shift, selector, two PRMT): the ALU pipe is what
// bounds k_epl_batch_tma<1>.  A warp has a table of its own (warps leave these kernels independently).
#ifndef GPSB_BATCH_LUT
#define GPSB_BATCH_LUT 1
#endif
__device__ __forceinline__ void quadrant_lut_fill(uint2* lut, int lane)
{
    uint32_t cp, sp;
    quadrant_patterns((uint32_t)lane << 27, cp, sp);
    lut[lane] = make_uint2(cp, sp);
    __syncwarp();
}
__device__ __forceinline__ void quadrant_patterns_lut(uint32_t nco, uint32_t lut_addr, uint32_t eight, uint32_t& cp, uint32_t& sp)
{
    uint32_t a;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(nco >> 27), "r"(eight), "r"(lut_addr));
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cp), "=r"(sp) : "r"(a));
}
__device__ __forceinline__ uint32_t batch_opaque_one()      // 1, from a special register (gridDim.y == 1 in these launches)
{
    uint32_t v;
    asm volatile("mov.u32 %0, %%nctaid.y;" : "=r"(v));
    return v;
}

// frame index of millisecond ms: the modulo only when the ring has wrapped (an integer modulo is ~20 instructions)
__device__ __forceinline__ uint32_t ring_frame(uint32_t ms, uint32_t ring_ms) { return ms < ring_ms ? ms : ms % ring_ms; }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// 32-bit replica window for data word w and byte offset off: bytes 4w - off (+ one period, so that the position is
// never negative) .. +3 of the periodic stream
__device__ __forceinline__ uint32_t rxt_window(const uint32_t* __restrict__ rx, int w, uint32_t off)
{
    const int p = 4 * w - (int)off + (int)GPSB_MS_BYTES;      // copy 0 of the table: the stream itself
    return __funnelshift_r(__ldg(rx + (p >> 2)), __ldg(rx + (p >> 2) + 1), ((uint32_t)p & 3u) * 8u);
}

// One cell by one warp: `f` holds the cell's frame, lane l the 16-byte groups l, l+32, l+64, l+96.
template <int kArms>
__device__ __forceinline__ void batch_cell(const gpsb_epl_req& rq, const uint4 (&f)[4], int lane, uint32_t c,
                                           int16_t* __restrict__ out, const uint32_t* __restrict__ rxt,
                                           const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t lut_addr)
{
    const uint32_t eight = batch_opaque_one() << 3;
    const uint32_t* __restrict__ rx = rxt + ((size_t)rq.sv_slot * kRxtShifts + (rq.off_bits & 15u)) * (kRxtCopies * kRxtWords);
    const uint32_t offs[3] = {kArms == 3 ? rq.off_e : rq.off_p, rq.off_p, rq.off_l};
    uint32_t acc[kArms];            // packed I | Q << 16 (a whole millisecond is at most 16368 per component)
    uint32_t acc_q[kArms];          // three arms: Q apart until the end (measured: the packed form costs them 3 %)
    // Replica run per arm: data word w meets the stream at byte 4w - off + 2046 (RXT spans two periods, so there is no
    // wrap to test for).  A lane's group of four data words is sixteen consecutive stream bytes starting at byte
    // p + 512 k: one aligned 128-bit load from the copy of the stream displaced by (p & 15) bytes.
    uint4 rr[kArms][4];
#pragma unroll
    for (int a = 0; a < kArms; a++) {
        acc[a] = acc_q[a] = 0u;
        const int p = 16 * lane - (int)offs[kArms == 3 ? a : 1] + (int)GPSB_MS_BYTES;
        const int copy = p & 15;
        const uint32_t* q = rx + copy * kRxtWords + ((p - copy) >> 2);        // 16-byte aligned
#pragma unroll
        for (int k = 0; k < 4; k++) rr[a][k] = __ldg(reinterpret_cast<const uint4*>(q + 128 * k));
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int w0 = 4 * (lane + 32 * k);
        const uint32_t s[4] = {f[k].x, f[k].y, f[k].z, f[k].w};
        uint32_t r[kArms][4];
#pragma unroll
        for (int a = 0; a < kArms; a++) {
            r[a][0] = rr[a][k].x; r[a][1] = rr[a][k].y; r[a][2] = rr[a][k].z; r[a][3] = rr[a][k].w;
        }
        // The ALU pipe is what bounds this kernel (70 % busy, profiles/k_epl_batch_tma1_r2.txt); multiply-adds issue on the
        // FMA pipe, so the NCO word of every data word and the Q half of the packed sum are formed by IMAD.
        uint32_t nco_run = rq.acc0 + (uint32_t)w0 * rq.step32;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t cp, sp;
            const uint32_t nco = kArms == 1 ? ((uint32_t)(w0 + j) * rq.step32 + rq.acc0) : nco_run;
            nco_run += rq.step32;
            if (GPSB_BATCH_LUT) quadrant_patterns_lut(nco, lut_addr, eight, cp, sp);
            else quadrant_patterns(nco, cp, sp);
#pragma unroll
            for (int a = 0; a < kArms; a++) {
                const uint32_t win = r[a][j];
                uint32_t xi = s[j] ^ cp ^ win, xq = s[j] ^ sp ^ win;
                if (k == 3 && j == 3) {
                    // word 511 (lane 31) is never mixed - its data is 0 - and only its bytes 2044 and, for even
                    // offsets, 2045 exist: I = Q = those replica bits (gps_misc.c:229, :59-89)
                    const uint32_t keep = (offs[kArms == 3 ? a : 1] & 1u) ? 0x000000FFu : 0x0000FFFFu;
                    xi = lane == 31 ? (win & keep) : xi;
                    xq = lane == 31 ? (win & keep) : xq;
                }
                acc[a] += (uint32_t)__popc(xi);
                if (kArms == 1) acc[a] = (uint32_t)__popc(xq) * 65536u + acc[a];
                else acc_q[a] += (uint32_t)__popc(xq);
            }
        }
    }
    if (kArms != 1) {
#pragma unroll
        for (int a = 0; a < kArms; a++) acc[a] += acc_q[a] << 16;
    }

    // Odd offsets 2k+1 skip the replica words whose data bytes are {2045, 0} and {off-2, off-1} (gps_misc.c:59-89).
    // Byte 2045 is handled above; edge lanes take back what the loop counted for the other three: role 0 byte 0,
    // roles 1 / 2 bytes off-2 / off-1 (offsets >= 3).
    bool any_odd = false;
#pragma unroll
    for (int a = 0; a < kArms; a++) any_odd |= (offs[kArms == 3 ? a : 1] & 1u) != 0u;
    if (any_odd && lane < 3 * kArms) {
        const int role = lane / kArms, a = lane % kArms;
        const uint32_t o = offs[kArms == 3 ? a : 1];
        const int d = role == 0 ? 0 : (int)o - (role == 1 ? 2 : 1);
        if ((o & 1u) && (role == 0 || o >= 3u)) {
            const int word = d >> 2;
            const uint32_t m = 0xFFu << (8 * (d & 3));
            const uint32_t win = rxt_window(rx, word, o);
            uint32_t v;
            if (word < kWords - 1) {
                const uint32_t sw = __ldg(signal + (size_t)ring_frame(rq.ms_index, ring_ms) * kWords + word);
                uint32_t cp, sp;
                quadrant_patterns(rq.acc0 + (uint32_t)word * rq.step32, cp, sp);
                v = (uint32_t)__popc((sw ^ cp ^ win) & m) + ((uint32_t)__popc((sw ^ sp ^ win) & m) << 16);
            } else {
                v = (uint32_t)__popc(win & m) * 0x00010001u;      // byte 2044 of offset 2045: counted above against zero data
            }
#pragma unroll
            for (int b = 0; b < kArms; b++) acc[b] -= (b == a) ? v : 0u;
        }
    }

    uint32_t tot[kArms];
#pragma unroll
    for (int a = 0; a < kArms; a++) tot[a] = __reduce_add_sync(0xFFFFFFFFu, acc[a]);
    if (lane == 0) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out + (size_t)c * 2 * kArms);
#pragma unroll
        for (int a = 0; a < kArms; a++) {
            const uint32_t i16 = (uint16_t)(int16_t)((int)(tot[a] & 0xFFFFu) - kHalfSum);
            const uint32_t q16 = (uint16_t)(int16_t)((int)(tot[a] >> 16) - kHalfSum);
            o32[a] = i16 | (q16 << 16);
        }
    }
}

template <int kArms>
__global__ void __launch_bounds__(kBatchThreads, kBatchCtasPerSm)
k_epl_batch(const gpsb_epl_req* __restrict__ reqs, int16_t* __restrict__ out, const uint32_t* __restrict__ rxt,
            const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t n)
{
    __shared__ uint2 lut[kBatchThreads / 32][32];
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * kBatchThreads) >> 5;
    uint32_t c = (blockIdx.x * kBatchThreads + threadIdx.x) >> 5;
    if (c >= n) return;
    quadrant_lut_fill(lut[threadIdx.x >> 5], lane);
    const uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(lut[threadIdx.x >> 5]);

    // Software pipeline, unrolled by two so that the two frame buffers swap roles without register moves: while cell i
    // is correlated, the frame of cell i+1 and the request of cell i+2 are in flight - neither the request fetch
    // (which the frame address depends on) nor the frame fetch is waited for.
    auto load_req = [&](uint32_t idx, const gpsb_epl_req& fallback) { return idx < n ? reqs[idx] : fallback; };
    auto load_frame = [&](const gpsb_epl_req& r, uint4 (&f)[4]) {
        const uint4* fp = reinterpret_cast<const uint4*>(signal + (size_t)ring_frame(r.ms_index, ring_ms) * kWords);
#pragma unroll
        for (int k = 0; k < 4; k++) f[k] = ldg_stream(fp + lane + 32 * k);
    };
    gpsb_epl_req rq_a = reqs[c];
    gpsb_epl_req rq_b = load_req(c + warps, rq_a);
    uint4 fa[4], fb[4];
    load_frame(rq_a, fa);
    for (;;) {
        load_frame(rq_b, fb);                                  // past the end: re-reads a frame that is in L2 anyway
        const gpsb_epl_req rq_c = load_req(c + 2 * warps, rq_b);
        batch_cell<kArms>(rq_a, fa, lane, c, out, rxt, signal, ring_ms, lut_addr);
        c += warps;
        if (c >= n) break;
        load_frame(rq_c, fa);
        rq_a = load_req(c + 2 * warps, rq_c);                  // request of the cell after next, into the free slot
        batch_cell<kArms>(rq_b, fb, lane, c, out, rxt, signal, ring_ms, lut_addr);
        c += warps;
        if (c >= n) break;
        rq_b = rq_a;                                           // roles for the next round: a = cell c, b = cell c + warps
        rq_a = rq_c;
    }
}


/* ------------------------------------------------------------------------------------------------------------
 * k_epl_batch_tma: the same cells, frames fed by the TMA engine into a shared-memory ring.
 *
 * The register-staged kernel above keeps ONE frame per warp in flight (16 registers per lane) and is limited by that:
 * ncu shows it waiting on the scoreboard of those loads with 23 % of the warp slots occupied (102 registers, 2 CTAs
 * per SM).  Here every warp owns a ring of kTmaStages 2048-byte frame buffers in shared memory; lane 0 issues one
 * cp.async.bulk per cell, kTmaStages cells ahead, each completing on its own mbarrier (complete_tx), so a warp has
 * kTmaStages frames in flight at no register cost and three CTAs fit an SM.  The consumer side waits on the stage's
 * mbarrier, pulls its four 16-byte groups out of shared memory (conflict free: consecutive lanes, consecutive
 * groups) and runs the very same batch_cell as above, so the two kernels cannot differ in their sums.
 * Requests: the frame address of a cell depends on its request, so lane 0 keeps the ms_index of the cell it will
 * fetch next one iteration ahead; the request of the cell being correlated is loaded one cell ahead by all lanes. */
#ifndef GPSB_BATCH_TMA_STAGES
#define GPSB_BATCH_TMA_STAGES 4
#endif
constexpr int kTmaStages = GPSB_BATCH_TMA_STAGES;
#ifndef GPSB_BATCH_TMA_CTAS
#define GPSB_BATCH_TMA_CTAS 3
#endif
constexpr int kTmaCtasPerSm = GPSB_BATCH_TMA_CTAS;
constexpr int kTmaWarps = kBatchThreads / 32;
struct BatchTmaSmem {
    uint4 frame[kTmaWarps][kTmaStages][GPSB_FRAME_BYTES / 16];     // 8 x 4 x 2 KB
    unsigned long long full[kTmaWarps][kTmaStages];
    uint2 lut[kTmaWarps][32];                                      // quadrant pattern pairs, one table per warp
};

template <int kArms>
__global__ void __launch_bounds__(kBatchThreads, kArms == 3 ? 2 : kTmaCtasPerSm)     // three arms need their 128 registers
k_epl_batch_tma(const gpsb_epl_req* __restrict__ reqs, int16_t* __restrict__ out, const uint32_t* __restrict__ rxt,
                const uint32_t* __restrict__ signal, uint32_t ring_ms, uint32_t n)
{
    extern __shared__ __align__(128) unsigned char batch_smem_raw[];
    BatchTmaSmem& sm = *reinterpret_cast<BatchTmaSmem*>(batch_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warps = (gridDim.x * kBatchThreads) >> 5;
    const uint32_t c0 = (blockIdx.x * kBatchThreads + threadIdx.x) >> 5;
    if (c0 >= n) return;                                           // whole warps leave; nothing below is CTA-wide
    const uint32_t mine = (n - c0 + warps - 1) / warps;            // cells of this warp: c0 + i * warps
    quadrant_lut_fill(sm.lut[warp], lane);
    const uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(sm.lut[warp]);

    auto sa = [](const void* p) { return (uint32_t)__cvta_generic_to_shared(p); };
    auto fetch = [&](uint32_t ms_index, int stage) {               // lane 0: one bulk copy of a whole frame
        const void* src = signal + (size_t)ring_frame(ms_index, ring_ms) * kWords;
        const uint32_t bar = sa(&sm.full[warp][stage]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)GPSB_FRAME_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sa(&sm.frame[warp][stage][0])),
                     "l"(src), "r"((uint32_t)GPSB_FRAME_BYTES), "r"(bar)
                     : "memory");
    };
    if (lane == 0) {
        for (int s = 0; s < kTmaStages; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(&sm.full[warp][s])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t next_ms = 0;                                          // lane 0: ms_index of cell `kTmaStages` ahead
    if (lane == 0) {
        for (uint32_t s = 0; s < (uint32_t)kTmaStages && s < mine; s++) fetch(reqs[c0 + s * warps].ms_index, (int)s);
        if ((uint32_t)kTmaStages < mine) next_ms = reqs[c0 + (uint32_t)kTmaStages * warps].ms_index;
    }
    gpsb_epl_req rq = reqs[c0];
    for (uint32_t i = 0; i < mine; i++) {
        const uint32_t c = c0 + i * warps;
        const int stage = (int)(i % (uint32_t)kTmaStages);
        const gpsb_epl_req rq_next = i + 1 < mine ? reqs[c + warps] : rq;          // in flight during this cell
        uint32_t ms_after = 0;                                      // lane 0: request word of the fetch after the next one
        if (lane == 0 && i + kTmaStages + 1 < mine) ms_after = reqs[c + (uint32_t)(kTmaStages + 1) * warps].ms_index;
        {   // the frame of this cell has landed
            const uint32_t bar = sa(&sm.full[warp][stage]), parity = (i / (uint32_t)kTmaStages) & 1u;
            asm volatile(
                "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar),
                "r"(parity)
                : "memory");
        }
        uint4 f[4];
#pragma unroll
        for (int k = 0; k < 4; k++) f[k] = sm.frame[warp][stage][lane + 32 * k];
        __syncwarp();                                               // every lane has its groups: the buffer is free again
        if (lane == 0 && i + kTmaStages < mine) fetch(next_ms, stage);
        next_ms = ms_after;
        batch_cell<kArms>(rq, f, lane, c, out, rxt, signal, ring_ms, lut_addr);
        rq = rq_next;
    }
}

}  // namespace gpsb
